// rigidbody/RigidBodySystem.h -- the reference's pipeline driver (include/rigidbody/RigidBodySystem.h:14-88)
// with the same public surface: body/joint setup, step(dt), solver selection by solverId/solverIter,
// callbacks. step() runs the whole pipeline on a B200 through the C ABI (include/slrbs_b200.h);
// the public host fields of bodies, joints and contacts are refreshed after every step.
#pragma once
#include "util/Types.h"

#include <functional>
#include <memory>
#include <unordered_map>
#include <vector>

class Contact;
class CollisionDetect;
class Joint;
class Solver;
class RigidBody;
struct slrbs_ctx;

typedef std::function<void(std::vector<RigidBody*>&)> PreStepFunc;
typedef std::function<void()> ResetFunc;

class RigidBodySystem {
public:
    RigidBodySystem();
    virtual ~RigidBodySystem();

    void step(float _dt);
    void clear();
    void addBody(RigidBody* _b);
    void addJoint(Joint* _j);

    const std::vector<RigidBody*>& getBodies() const { return m_bodies; }
    std::vector<RigidBody*>& getBodies() { return m_bodies; }
    const std::vector<Contact*>& getContacts() const;
    std::vector<Contact*>& getContacts();
    const std::vector<Joint*>& getJoints() const { return m_joints; }
    std::vector<Joint*>& getJoints() { return m_joints; }

    void setPreStepFunc(PreStepFunc _func) { m_preStepFunc = _func; }
    void setResetFunc(ResetFunc _func) { m_resetFunc = _func; }
    void setEnableCollisionDetection(bool _enableCollisions) { m_collisionsEnabled = _enableCollisions; }
    // EXTENSION (not in the reference): also generate sphere-plane, box-plane and box-box contacts, which the
    // reference's dispatch ignores (src/collision/CollisionDetect.cpp:45-73). Off by default.
    void setExtendedContacts(bool _on) { m_extContacts = _on; }
    // EXTENSION: replace the solver object behind solverId = id (0..3); the system takes ownership. The reference keeps
    // its solvers in a file-static array (RigidBodySystem.cpp:17,29-31), so there a user edits that file instead.
    void setSolver(int id, Solver* solver);

    int solverIter;
    int solverId;   // PGS = 0, Conj. Gradient = 1, Conj. Residual = 2; extension: block principal pivoting = 3

    // ---- device controls (additions; the reference is CPU-only) ----------------------------
    void setDevice(int cudaDevice) { m_device = cudaDevice; }
    // Mirror contacts (p, n, t, b, pene, lambda) into Contact objects after each step (default on).
    void setContactReadback(bool on) { m_contactReadback = on; }
    // Also mirror J0/J1/J0Minv/J1Minv of contacts and joints (default off; costs a download).
    void setRowReadback(bool on) { m_rowReadback = on; }
    slrbs_ctx* deviceContext() { return m_ctx; }
    float lastDt() const { return m_lastDt; }
    // stage access for the Solver / CollisionDetect seams
    void deviceSolve(int solverId, int iters, float h);
    void deviceDetect();
    void deviceContactJacobians(float h);

private:
    std::vector<RigidBody*> m_bodies;
    std::vector<Joint*> m_joints;
    std::unique_ptr<CollisionDetect> m_collisionDetect;
    bool m_collisionsEnabled;
    bool m_extContacts = false;
    PreStepFunc m_preStepFunc;
    ResetFunc m_resetFunc;
    // per-system solver objects (the reference keeps them in a file-static array shared by every
    // system, RigidBodySystem.cpp:17,29-31; kept per system here so batches do not alias)
    std::unique_ptr<Solver> m_solvers[4];

    // device mirror
    slrbs_ctx* m_ctx = nullptr;
    int m_device = 0;
    int m_contactCapacity = 0;
    bool m_topologyDirty = true;
    bool m_contactReadback = true;
    bool m_rowReadback = false;
    bool m_extActive = false;
    float m_lastDt = 0.01f;
    std::vector<float> m_paramSnapshot, m_stateSnapshot, m_lambdaSnapshot;
    mutable std::unordered_map<const RigidBody*, int> m_bodyIndex;   // body -> index, rebuilt when bodies were added / cleared
    mutable bool m_bodyIndexDirty = true;
    int bodyIndexOf(const RigidBody* b) const;
    bool hostSolve(Solver* solver, float dt);

    void ensureContext();
    void uploadAll();
    void uploadStateIfChanged();
    void packParams(std::vector<float>& out) const;
    void packState(std::vector<float>& out) const;
    void packLambda(std::vector<float>& out) const;
    void readBack(bool contacts);
    void mirrorContacts();
    void hostForceReset();
    void computeInertias();
    void check(int rc, const char* what) const;
};

// scalar * quaternion and quaternion + quaternion helpers of the reference (RigidBodySystem.h:78-88)
inline Eigen::Quaternionf operator*(float a, const Eigen::Quaternionf& q) { return Eigen::Quaternionf(a * q.w(), a * q.x(), a * q.y(), a * q.z()); }
inline Eigen::Quaternionf operator+(const Eigen::Quaternionf& a, const Eigen::Quaternionf& b)
{
    return Eigen::Quaternionf(a.w() + b.w(), a.x() + b.x(), a.y() + b.y(), a.z() + b.z());
}

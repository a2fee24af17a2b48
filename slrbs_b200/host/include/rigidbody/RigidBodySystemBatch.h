// rigidbody/RigidBodySystemBatch.h -- N independent copies of one RigidBodySystem stepped together on one B200.
//
// EXTENSION of the reference's surface (include/rigidbody/RigidBodySystem.h:21-56 describes ONE system whose step() is
// called once per frame and whose public fields are read back every frame). Batched use -- many environments built by
// the same Scenarios.h builder, RL style -- keeps the same knobs (solverIter, solverId, Contact::mu, collisions on/off,
// setPreStepFunc's forces, RigidBodySystemState's save / restore) but
//   * step(dt, n) is asynchronous and reads nothing back,
//   * state, external forces and resets are addressed by scene index range,
//   * host fields are refreshed only on request (getState / readScene).
// Every scene evolves exactly as a RigidBodySystem built the same way would: same kernels, same order (tests:
// tests/test_gpu_host_mirror.py::test_batch_matches_single_systems).
#pragma once
#include "rigidbody/RigidBodySystem.h"

#include <vector>

class RigidBodySystemBatch {
public:
    // nScenes copies of the prototype's bodies, joints and CURRENT state (public fields) on cuda:device;
    // contactsPerScene = 0 sizes the contact capacity as RigidBodySystem does (8 per body, at least 64)
    RigidBodySystemBatch(const RigidBodySystem& prototype, int nScenes, int device = 0, int contactsPerScene = 0);
    ~RigidBodySystemBatch();
    RigidBodySystemBatch(const RigidBodySystemBatch&) = delete;
    RigidBodySystemBatch& operator=(const RigidBodySystemBatch&) = delete;

    int numScenes() const { return m_scenes; }
    int numBodies() const { return m_nb; }
    int numJoints() const { return m_nj; }

    // the reference's public knobs, read at every step() like RigidBodySystem.cpp:181 reads them
    int solverIter = 10;
    int solverId = 0;                 // PGS = 0, Conj. Gradient = 1, Conj. Residual = 2; extension: pivoting = 3
    float mu = 0.8f;                  // Contact::mu (one value for the whole batch)
    void setEnableCollisionDetection(bool on) { m_collisions = on; }
    void setExtendedContacts(bool on) { m_ext = on; }

    // RigidBodySystem::step n times for every scene; returns at once (device resident, no readback)
    void step(float dt, int nSteps = 1);
    // wait for the device; throws if a scene produced more contacts than the capacity (contacts were dropped)
    void sync();

    // dense [nScenes][numBodies()][k] host arrays, scenes [scene0, scene0 + nScenes); any pointer may be null
    void setState(int scene0, int nScenes, const float* x3, const float* q4xyzw, const float* xdot3, const float* omega3);
    void getState(int scene0, int nScenes, float* x3, float* q4xyzw, float* xdot3, float* omega3);
    // what a PreStepFunc would leave in RigidBody::f / tau (RigidBodySystem.cpp:70-73): added to gravity, or absolute
    void setExternalForces(int scene0, int nScenes, const float* f3, const float* tau3, bool absolute = false);
    void clearExternalForces(int scene0, int nScenes);
    // RigidBodySystemState (RigidBodyState.cpp:48-72) for the batch: one snapshot, restored by scene index range
    void saveState();
    void resetScenes(int scene0, int nScenes);
    void getContactCounts(int scene0, int nScenes, int* counts);
    // one scene's x, q, xdot, omega (and joint lambdas) written into the public fields of a host system with the same bodies
    void readScene(int scene, RigidBodySystem& into);

    slrbs_ctx* deviceContext() { return m_ctx; }

private:
    slrbs_ctx* m_ctx = nullptr;
    int m_scenes = 0, m_nb = 0, m_nj = 0, m_device = 0;
    bool m_collisions = true, m_ext = false;
    void check(int rc, const char* what) const;
    void pushParams();
};

// joint/Joint.h -- the reference's constraint base class (include/joint/Joint.h:9-50): public
// body pointers, Jacobian blocks, phi, lambda, attachment frames. After RigidBodySystem::step()
// these fields hold what the device computed (lambda always; J blocks when row readback is on).
#pragma once
#include <Eigen/Dense>
#include "util/Types.h"

class RigidBody;

enum eConstraintType { kContact = 0, kSpherical, kHinge };

class Joint {
public:
    Joint(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Quaternionf& _q0,
          const Eigen::Vector3f& _r1, const Eigen::Quaternionf& _q1);
    Joint(RigidBody* _body0, RigidBody* _body1);
    virtual ~Joint() {}

    RigidBody* body0;
    RigidBody* body1;
    JBlock J0, J1, J0Minv, J1Minv;
    Eigen::VectorXf phi;
    Eigen::VectorXf lambda;
    unsigned int idx;
    unsigned int dim;
    Eigen::Vector3f r0, r1;
    Eigen::Quaternionf q0, q1;

    virtual eConstraintType getType() const = 0;
    // Host-side evaluation of the Jacobian from the bodies' current host fields. step() does this
    // on the device; this is for callers that inspect a joint between steps.
    virtual void computeJacobian() = 0;

protected:
    Joint();
    void sizeBlocks(int rows);
    void fillMinv();
};

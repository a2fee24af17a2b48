// rigidbody/RigidBodyState.h -- in-memory save / restore of x, xdot, q, omega
// (include/rigidbody/RigidBodyState.h, src/rigidbody/RigidBodyState.cpp:18-72).
#pragma once
#include <util/Types.h>
#include <vector>

class RigidBody;
class RigidBodySystem;

class RigidBodyState {
public:
    RigidBodyState() {}
    RigidBodyState(const RigidBody& _body) { save(_body); }
    void restore(RigidBody& _body);
    void save(const RigidBody& _body);
private:
    Eigen::Vector3f x, xdot;
    Eigen::Quaternionf q;
    Eigen::Vector3f omega;
};

class RigidBodySystemState {
public:
    RigidBodySystemState(const RigidBodySystem& _system) { save(_system); }
    void restore(RigidBodySystem& _system);
    void save(const RigidBodySystem& _system);
private:
    std::vector<RigidBodyState> rigidBodyStates;
};

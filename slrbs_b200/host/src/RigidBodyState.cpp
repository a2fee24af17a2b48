// RigidBodyState.cpp -- mirror of src/rigidbody/RigidBodyState.cpp:18-72.
#include "rigidbody/RigidBodyState.h"
#include "contact/Contact.h"
#include "rigidbody/RigidBody.h"
#include "rigidbody/RigidBodySystem.h"

#include <cassert>

void RigidBodyState::save(const RigidBody& b) { x = b.x; xdot = b.xdot; q = b.q; omega = b.omega; }

void RigidBodyState::restore(RigidBody& b)
{
    b.x = x; b.xdot = xdot; b.q = q; b.omega = omega;
    b.f.setZero();
    b.fc.setZero();
    b.contacts.clear();
}

void RigidBodySystemState::save(const RigidBodySystem& sys)
{
    const auto& bodies = sys.getBodies();
    rigidBodyStates.resize(bodies.size());
    for (size_t i = 0; i < bodies.size(); ++i) rigidBodyStates[i] = RigidBodyState(*bodies[i]);
}

void RigidBodySystemState::restore(RigidBodySystem& sys)
{
    const auto& bodies = sys.getBodies();
    assert(bodies.size() == rigidBodyStates.size());
    for (size_t i = 0; i < bodies.size() && i < rigidBodyStates.size(); ++i) rigidBodyStates[i].restore(*bodies[i]);
    // the reference clears the pointer vector without deleting (a leak, RigidBodyState.cpp:71); the
    // mirror owns its Contact objects, so delete them. Joint lambda is deliberately left alone.
    for (Contact* c : sys.getContacts()) delete c;
    sys.getContacts().clear();
}

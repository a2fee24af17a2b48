// collision/CollisionDetect.h -- the reference's collision front end (include/collision/CollisionDetect.h:14-58).
// The pair tests and narrow phase run on the device (slrbs_stage_detect); this class owns the
// host-side Contact mirror that RigidBodySystem::getContacts() exposes.
#pragma once
#include <Eigen/Dense>
#include <vector>

class Contact;
class RigidBody;
class RigidBodySystem;

class CollisionDetect {
public:
    CollisionDetect(RigidBodySystem* rigidBodySystem) : m_rigidBodySystem(rigidBodySystem) {}
    ~CollisionDetect() { clear(); }

    // Runs the device narrow phase on the system's current device state and mirrors the contacts.
    void detectCollisions();
    // Deletes the mirrored contacts and clears every body's contact list (CollisionDetect.cpp:87-100).
    void clear();
    // Runs the device Jacobian/row assembly for the current contacts (CollisionDetect.cpp:78-85).
    void computeContactJacobians();

    const std::vector<Contact*>& getContacts() const { return m_contacts; }
    std::vector<Contact*>& getContacts() { return m_contacts; }

private:
    friend class RigidBodySystem;
    RigidBodySystem* m_rigidBodySystem;
    std::vector<Contact*> m_contacts;
};

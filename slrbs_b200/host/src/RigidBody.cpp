// RigidBody.cpp -- host mirror of src/rigidbody/RigidBody.cpp: constructors (local inertia and its
// inverse, :27-28,65-66), updateInertiaMatrix (:78-89), addForceAtPos (:91-96), getVelocityAtPos (:98-102).
#include "rigidbody/RigidBody.h"
#include "contact/Contact.h"
#include "joint/Joint.h"
#include "polyscope/surface_mesh.h"
#include "util/MeshAssets.h"

int RigidBody::counter = 0;

void RigidBody::init(float _mass)
{
    fixed = false;
    mass = _mass;
    q = Eigen::Quaternionf(1, 0, 0, 0);
    Iinv.setZero();
    I.setZero();                       // uninitialised upstream (RigidBody.cpp:9-25); zero here
    Ibody = geometry->computeInertia(mass);
    IbodyInv = Ibody.inverse();
    mesh = polyscope::registerSurfaceMesh(std::to_string(RigidBody::counter), 0, 0);   // headless stub
    RigidBody::counter++;
}

RigidBody::RigidBody(float _mass, Geometry* _geometry, const std::string&) : geometry(_geometry), mesh(nullptr) { init(_mass); }
RigidBody::RigidBody(float _mass, Geometry* _geometry, const Mesh&) : geometry(_geometry), mesh(nullptr) { init(_mass); }

void RigidBody::updateInertiaMatrix()
{
    if (!fixed) {
        I = q * Ibody * q.inverse();
        Iinv = q * IbodyInv * q.inverse();
    } else {
        Iinv.setZero();
    }
}

void RigidBody::addForceAtPos(const Eigen::Vector3f& pos, const Eigen::Vector3f& force)
{
    const Eigen::Vector3f r = pos - x;
    f += force;
    tau += r.cross(force);
}

void RigidBody::getVelocityAtPos(const Eigen::Vector3f& pos, Eigen::Vector3f& vel)
{
    const Eigen::Vector3f r = pos - x;
    vel = xdot + r.cross(omega);       // the reference's sign (RigidBody.cpp:101)
}

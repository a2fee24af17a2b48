// rigidbody/RigidBody.h -- the reference's per-body record (include/rigidbody/RigidBody.h:21-60):
// same public fields, written directly by scene builders and callers between steps.
#pragma once
#include "collision/Geometry.h"
#include "util/Types.h"
#include <memory>
#include <string>
#include <vector>

class Contact;
class Joint;
struct Mesh;
namespace polyscope { class SurfaceMesh; }

class RigidBody {
public:
    RigidBody(float _mass, Geometry* _geometry, const std::string& meshFilename = "");
    RigidBody(float _mass, Geometry* _geometry, const Mesh& _mesh);

    static int counter;

    void updateInertiaMatrix();
    void addForceAtPos(const Eigen::Vector3f& pos, const Eigen::Vector3f& force);
    void getVelocityAtPos(const Eigen::Vector3f& pos, Eigen::Vector3f& vel);

    bool fixed;
    float mass;
    Eigen::Matrix3f I, Iinv;
    Eigen::Matrix3f Ibody, IbodyInv;
    Eigen::Vector3f x;
    Eigen::Quaternionf q;
    Eigen::Vector3f xdot, omega;
    Eigen::Vector3f f, tau;
    Eigen::Vector3f fc, tauc;

    std::unique_ptr<Geometry> geometry;
    std::vector<Contact*> contacts;
    std::vector<Joint*> joints;
    polyscope::SurfaceMesh* mesh;

private:
    void init(float _mass);
};

// util/MeshAssets.h -- shadows the reference's include/util/MeshAssets.h for oracle/_ref only (TEST
// INFRASTRUCTURE).  The reference's src/util/MeshAssets.cpp builds render meshes (and does not compile
// under GCC: std::sqrtf at :217); the step path never reads them, so the same signatures
// (include/util/MeshAssets.h:8-47) return empty meshes here.
#pragma once
#include <Eigen/Dense>
#include <cstddef>
#include <map>
#include <string>
#include "polyscope/surface_mesh.h"

struct Mesh {
    Mesh() : name(""), meshV(), meshF() {}
    std::string name;
    Eigen::MatrixXf meshV;
    Eigen::MatrixXi meshF;
};
typedef std::map<unsigned int, Mesh> MeshCache;
class MeshAssetRegistry {
public:
    static Mesh* loadObj(const std::string&) { return nullptr; }
    static void clear() {}
};
inline Mesh createCylinder(size_t, float, float) { return Mesh(); }
inline Mesh createPlane(const Eigen::Vector3f&, const Eigen::Vector3f&) { return Mesh(); }
inline Mesh createBox(const Eigen::Vector3f&) { return Mesh(); }
inline Mesh createSphere(float) { return Mesh(); }

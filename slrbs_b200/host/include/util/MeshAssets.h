// util/MeshAssets.h -- render meshes are out of the hot path (the collision code never reads them,
// reference src/util/MeshAssets.cpp is render-only). Only the signatures the scene builders call
// are kept (include/util/MeshAssets.h:44-47); they return an empty Mesh.
#pragma once
#include <Eigen/Dense>
#include <cstddef>
#include <string>

struct Mesh {
    std::string name;
    Eigen::MatrixXf meshV;
    Eigen::MatrixXi meshF;
};

inline Mesh createCylinder(size_t, float, float) { return Mesh(); }
inline Mesh createPlane(const Eigen::Vector3f&, const Eigen::Vector3f&) { return Mesh(); }
inline Mesh createBox(const Eigen::Vector3f&) { return Mesh(); }
inline Mesh createSphere(float) { return Mesh(); }

// polyscope/surface_mesh.h -- headless stand-in for the Polyscope v2.2.1 structure the reference
// registers for every body (src/rigidbody/RigidBody.cpp:38,69) and styles from its scene builders
// (include/rigidbody/Scenarios.h:37-38,62-66). The viewer is out of the hot path; these calls are no-ops.
#pragma once
#include <array>
#include <string>

namespace polyscope {

class SurfaceMesh {
public:
    SurfaceMesh* setSurfaceColor(std::array<float, 3>) { return this; }
    SurfaceMesh* setTransparency(float) { return this; }
    SurfaceMesh* setSmoothShade(bool) { return this; }
    SurfaceMesh* setEdgeWidth(float) { return this; }
    SurfaceMesh* setEnabled(bool) { return this; }
};

template <class V, class F>
inline SurfaceMesh* registerSurfaceMesh(const std::string&, const V&, const F&)
{
    static SurfaceMesh headless;
    return &headless;
}
inline void removeAllStructures() {}

}  // namespace polyscope

// polyscope/surface_mesh.h -- headless no-op stand-in (TEST INFRASTRUCTURE, oracle/_ref only) for the
// Polyscope v2.2.1 calls the reference makes from RigidBody's constructors (src/rigidbody/RigidBody.cpp:38-40,
// 69-71) and from its scene builders (include/rigidbody/Scenarios.h).  The viewer is not on the step path.
#pragma once
#include <array>
#include <iostream>
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
template <class V, class F> inline SurfaceMesh* registerSurfaceMesh(const std::string&, const V&, const F&)
{
    static SurfaceMesh headless;
    return &headless;
}
inline void removeAllStructures() {}
}  // namespace polyscope

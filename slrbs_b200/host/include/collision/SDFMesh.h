// collision/SDFMesh.h -- EXTENSION, no reference counterpart (the reference loads resources/bunny.obj
// for rendering only and has no SDF code): a closed triangle mesh as collision geometry, held as a signed-distance grid
// plus its vertices as sample points (csrc/sdf.cuh). Geometry type 4; contacts need RigidBodySystem::setExtendedContacts(true).
// The grid is built on the device (slrbs_sdf_build). Bodies share one SDFShape through a shared_ptr; a RigidBodySystem
// holds at most 16 distinct shapes. Inertia: that of the solid box around the vertices.
#pragma once
#include "collision/Geometry.h"
#include "util/OBJLoader.h"
#include "slrbs_b200.h"

#include <algorithm>
#include <cmath>
#include <memory>
#include <stdexcept>
#include <vector>

struct SDFShape {
    std::vector<float> verts;        // body frame, 3 per vertex
    std::vector<int32_t> tris;
    std::vector<float> grid;         // node (i,j,k) at origin + cell*(i,j,k): grid[(k*ny+j)*nx+i], negative inside
    int nx = 0, ny = 0, nz = 0;
    float origin[3] = { 0.f, 0.f, 0.f };
    float cell = 0.f;
    float bbmin[3], bbmax[3];
    mutable int shapeId = -1;        // slot in the owning system's device shape table (set by the system)

    // grid over the vertex bounding box plus `margin` cells on every side
    static std::shared_ptr<SDFShape> fromMesh(std::vector<float> v, std::vector<int32_t> t, float _cell, int margin = 3, int device = 0)
    {
        auto s = std::make_shared<SDFShape>();
        if (v.empty() || t.empty()) throw std::runtime_error("SDFShape: empty mesh");
        for (int k = 0; k < 3; ++k) { s->bbmin[k] = 1e30f; s->bbmax[k] = -1e30f; }
        for (size_t i = 0; i < v.size(); ++i) { const int k = (int)(i % 3); s->bbmin[k] = std::min(s->bbmin[k], v[i]); s->bbmax[k] = std::max(s->bbmax[k], v[i]); }
        s->cell = _cell;
        int dims[3];
        for (int k = 0; k < 3; ++k) {
            s->origin[k] = s->bbmin[k] - (float)margin * _cell;
            dims[k] = (int)std::ceil((s->bbmax[k] + (float)margin * _cell - s->origin[k]) / _cell) + 1;
        }
        s->nx = dims[0]; s->ny = dims[1]; s->nz = dims[2];
        s->grid.resize((size_t)s->nx * s->ny * s->nz);
        s->verts = std::move(v); s->tris = std::move(t);
        if (slrbs_sdf_build(device, (int)s->verts.size() / 3, s->verts.data(), (int)s->tris.size() / 3, s->tris.data(), s->nx, s->ny, s->nz,
                            s->origin, s->cell, s->grid.data()) < 0)
            throw std::runtime_error(std::string("slrbs_sdf_build: ") + slrbs_last_error());
        return s;
    }

    // OBJLoader::load, vertices scaled by `scale` and re-centred on the middle of their bounding box
    static std::shared_ptr<SDFShape> fromOBJ(const std::string& path, float scale, float _cell, int device = 0)
    {
        Eigen::MatrixXf V; Eigen::MatrixXi F;
        if (!OBJLoader::load(path, V, F)) throw std::runtime_error("SDFShape: cannot read " + path);
        std::vector<float> v((size_t)V.rows() * 3);
        std::vector<int32_t> t((size_t)F.rows() * 3);
        float lo[3] = { 1e30f, 1e30f, 1e30f }, hi[3] = { -1e30f, -1e30f, -1e30f };
        for (int i = 0; i < V.rows(); ++i) for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], V(i, k)); hi[k] = std::max(hi[k], V(i, k)); }
        for (int i = 0; i < V.rows(); ++i) for (int k = 0; k < 3; ++k) v[3 * i + k] = scale * (V(i, k) - 0.5f * (lo[k] + hi[k]));
        for (int i = 0; i < F.rows(); ++i) for (int k = 0; k < 3; ++k) t[3 * i + k] = F(i, k);
        return fromMesh(std::move(v), std::move(t), _cell, 3, device);
    }
};

class SDFMesh : public Geometry {
public:
    std::shared_ptr<SDFShape> shape;
    SDFMesh(const std::shared_ptr<SDFShape>& _shape) : shape(_shape) {}
    Eigen::Matrix3f computeInertia(float m) override
    {
        const float dx = shape->bbmax[0] - shape->bbmin[0], dy = shape->bbmax[1] - shape->bbmin[1], dz = shape->bbmax[2] - shape->bbmin[2];
        const float k = 1.0f / 12.0f;
        return diag(k * m * (dy * dy + dz * dz), k * m * (dx * dx + dz * dz), k * m * (dx * dx + dy * dy));
    }
    eGeometryType getType() const override { return kSDFMesh; }
    void packParams(float* g) const override { g[0] = (float)shape->shapeId; }
};

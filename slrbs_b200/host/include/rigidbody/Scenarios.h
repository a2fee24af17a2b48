// rigidbody/Scenarios.h -- the reference's five scene builders (include/rigidbody/Scenarios.h:18,84,112,153,178)
// with the same names and resulting systems (body order, joint order, initial state), written as small
// table-driven helpers, headless (no viewer styling). createSphereStack / createMarbleBoxN are
// additions used by the benchmarks: same primitives, more bodies.
#pragma once
#include "collision/SDFMesh.h"
#include "joint/Hinge.h"
#include "joint/Spherical.h"
#include "rigidbody/RigidBody.h"
#include "rigidbody/RigidBodySystem.h"
#include "polyscope/polyscope.h"
#include "util/MeshAssets.h"
#include "util/Types.h"

#include <Eigen/Dense>
#include <cstdlib>

class Scenarios {
    typedef Eigen::Vector3f V3;
    typedef Eigen::Quaternionf Quat;

    static RigidBody* sphere(RigidBodySystem& sys, float mass, float r, const V3& x)
    {
        RigidBody* b = new RigidBody(mass, new Sphere(r), createSphere(r));
        b->x = x;
        sys.addBody(b);
        return b;
    }
    static RigidBody* box(float mass, const V3& dim, const V3& x, bool fixed)
    {
        RigidBody* b = new RigidBody(mass, new Box(dim), createBox(dim));
        b->x = x;
        b->fixed = fixed;
        return b;
    }
    static RigidBody* groundPlane()
    {
        RigidBody* p = new RigidBody(1.0f, new Plane({ 0.0f, 0.0f, 0.0f }, { 0.0f, 1.0f, 0.0f }), "");
        p->x = { 0.0f, 0.0f, 0.0f };
        p->fixed = true;
        return p;
    }
    static void begin(RigidBodySystem& sys)
    {
        sys.clear();
        polyscope::removeAllStructures();
    }

public:
    // 9 x 9 x 2 marbles in an open box of four walls and a floor (Scenarios.h:18-80).
    static void createMarbleBox(RigidBodySystem& sys)
    {
        begin(sys);
        const float r = 0.5f;
        for (int i = 0; i < 9; ++i)
            for (int j = 0; j < 9; ++j)
                for (int layer = 0; layer < 2; ++layer)   // lower marble then upper marble of each column
                    sphere(sys, 1.0f, r, V3(-4.0f + (float)i * 1.0f, 2.0f + (float)layer, -4.0f + (float)j * 1.0f));
        const V3 side(0.4f, 4.0f, 10.0f), bottom(10.0f, 0.4f, 10.0f);
        const Quat quarter(Eigen::AngleAxisf(1.57f, V3(0, 1, 0)));
        RigidBody* w0 = box(1.0f, side, { 4.75f, 2.0f, 0.0f }, true);
        RigidBody* w1 = box(1.0f, side, { -4.75f, 2.0f, 0.0f }, true);
        RigidBody* w2 = box(1.0f, side, { 0.0f, 2.0f, 4.75f }, true);
        RigidBody* w3 = box(1.0f, side, { 0.0f, 2.0f, -4.75f }, true);
        RigidBody* fl = box(1.0f, bottom, { 0.0f, 0.0f, 0.0f }, true);
        w2->q = quarter;
        w3->q = quarter;
        for (RigidBody* b : { w0, w1, w2, w3, fl }) sys.addBody(b);
    }

    // One spinning sphere dropped on a fixed slab (Scenarios.h:84-108).
    static void createSphereOnBox(RigidBodySystem& sys)
    {
        begin(sys);
        RigidBody* s = sphere(sys, 1.0f, 0.5f, V3(0.0f, 4.0f, 0.0f));
        s->omega = V3(10.0f, 0.0f, 0.0f);
        sys.addBody(box(1.0f, V3(10.0f, 0.4f, 10.0f), V3(0.0f, 0.0f, 0.0f), true));
    }

    // Chain of 20 unit boxes on hinges, last link heavy and kicked sideways (Scenarios.h:112-148).
    static void createSwingingBoxes(RigidBodySystem& sys)
    {
        begin(sys);
        const int N = 20;
        const V3 dim({ 1.0f, 1.0f, 1.0f });
        const V3 dx(0.0f, 1.5f, 0.0f);
        RigidBody* parent = box(1.0f, dim, V3(0.0f, 1.5f * (float)N, 0.0f), true);
        sys.addBody(parent);
        for (int i = 0; i < N - 1; ++i) {
            const float mass = (i == N - 2) ? 100.0f : 1.0f;
            RigidBody* link = box(mass, dim, parent->x - dx, false);
            Joint* h = new Hinge(parent, link, -0.5f * dx, Quat::Identity(), 0.5f * dx, Quat::Identity());
            sys.addBody(link);
            sys.addJoint(h);
            parent = link;
        }
        parent->xdot = { 0.0f, 0.0f, 10.0f };
    }

    // Tilted cylinder falling on a plane (Scenarios.h:153-174).
    static void createCylinderOnPlane(RigidBodySystem& sys)
    {
        begin(sys);
        const float height = 2.0f, radius = 1.0f;
        RigidBody* cyl = new RigidBody(1.0f, new Cylinder(height, radius), createCylinder(16, radius, height));
        cyl->x = { 0.0f, 2.0f, 0.0f };
        cyl->q = Eigen::AngleAxisf(0.57f, V3(0.0f, 0.0f, 1.0f));
        sys.addBody(cyl);
        sys.addBody(groundPlane());
    }

    // Chassis + four cylinder wheels on hinges, rolling on a plane (Scenarios.h:178-247).
    static void createCarScene(RigidBodySystem& sys)
    {
        begin(sys);
        RigidBody* chassis = box(5.0f, V3(2.0f, 0.5f, 3.0f), V3(0.0f, 0.5f, 0.0f), false);
        chassis->xdot = { 0.0f, 0.0f, -10.0f };
        sys.addBody(chassis);
        const float wx[4] = { 1.0f, -1.0f, 1.0f, -1.0f };      // lf, rf, lr, rr
        const float wz[4] = { 1.5f, 1.5f, -1.5f, -1.5f };
        RigidBody* wheel[4];
        for (int k = 0; k < 4; ++k) {
            wheel[k] = new RigidBody(1.0f, new Cylinder(0.2f, 0.5f), createCylinder(16, 0.5f, 0.2f));
            wheel[k]->x = { wx[k], 0.5f, wz[k] };
            wheel[k]->q = Eigen::AngleAxisf(1.57079f, V3(0.0f, 0.0f, 1.0f));
            sys.addBody(wheel[k]);
        }
        sys.addBody(groundPlane());
        for (int k = 0; k < 4; ++k)
            sys.addJoint(new Hinge(chassis, wheel[k], { wx[k], 0.0f, wz[k] }, Quat::Identity(), { 0.0f, 0.0f, 0.0f },
                                   Quat(Eigen::AngleAxisf(1.57f, V3(0, 0, 1)))));
    }

    // ---- benchmark additions (not in the reference) -------------------------------------------
    // k spheres (r 0.5, m 1) stacked at y = 0.7 + i on createSphereOnBox's slab: spheres, then slab.
    static void createSphereStack(RigidBodySystem& sys, int k)
    {
        begin(sys);
        for (int i = 0; i < k; ++i) sphere(sys, 1.0f, 0.5f, V3(0.0f, 0.7f + (float)i * 1.0f, 0.0f));
        sys.addBody(box(1.0f, V3(10.0f, 0.4f, 10.0f), V3(0.0f, 0.0f, 0.0f), true));
    }

    // EXTENSION scene (BASELINE config 0 "box-stack on plane"; the reference has neither the scene nor box-box / box-plane
    // contacts): k unit boxes stacked on the ground plane, every other one turned 0.3 rad about y. Boxes bottom-up, then the plane.
    static void createBoxStack(RigidBodySystem& sys, int k)
    {
        begin(sys);
        for (int i = 0; i < k; ++i) {
            RigidBody* b = box(1.0f, V3(1.0f, 1.0f, 1.0f), V3(0.0f, 0.5f + (float)i * 1.0f, 0.0f), false);
            if (i & 1) b->q = Eigen::AngleAxisf(0.3f, V3(0, 1, 0));
            sys.addBody(b);
        }
        sys.addBody(groundPlane());
        sys.setExtendedContacts(true);
    }

    // EXTENSION scene (BASELINE config 3 "bunny pile"; the reference has no mesh collision): n copies of one closed mesh
    // (an OBJ file such as the reference's resources/bunny.obj, scaled so its bounding box is about one unit across), each
    // turned by a different angle, dropped in a column onto the ground plane. Meshes bottom-up, then the plane.
    static void createMeshPile(RigidBodySystem& sys, const std::string& objPath, int n, float scale = 1.0f, float cell = 0.06f)
    {
        begin(sys);
        std::shared_ptr<SDFShape> shape = SDFShape::fromOBJ(objPath, scale, cell);
        const float h = shape->bbmax[1] - shape->bbmin[1];
        for (int i = 0; i < n; ++i) {
            RigidBody* b = new RigidBody(1.0f, new SDFMesh(shape), "");
            b->x = V3(0.05f * (float)(i % 3 - 1), 0.6f * h + 1.1f * h * (float)i, 0.05f * (float)(i % 2));
            b->q = Quat(Eigen::AngleAxisf(0.7f * (float)i, V3(0, 1, 0)));
            sys.addBody(b);
        }
        sys.addBody(groundPlane());
        sys.setExtendedContacts(true);
    }

    // createMarbleBox's construction scaled to nx x nz columns of ny marbles, container sized to fit.
    static void createMarbleBoxN(RigidBodySystem& sys, int nx, int nz, int ny)
    {
        begin(sys);
        const float x0 = -0.5f * (float)(nx - 1), z0 = -0.5f * (float)(nz - 1);
        for (int i = 0; i < nx; ++i)
            for (int j = 0; j < nz; ++j)
                for (int k = 0; k < ny; ++k)
                    sphere(sys, 1.0f, 0.5f, V3(x0 + (float)i * 1.0f, 2.0f + (float)k * 1.0f, z0 + (float)j * 1.0f));
        const float ex = (float)nx + 1.0f, ez = (float)nz + 1.0f, hy = (float)ny + 2.0f;
        const float wallx = 0.5f * ex - 0.25f, wallz = 0.5f * ez - 0.25f;
        const Quat quarter(Eigen::AngleAxisf(1.57f, V3(0, 1, 0)));
        RigidBody* w0 = box(1.0f, V3(0.4f, hy, ez), V3(wallx, 0.5f * hy, 0.0f), true);
        RigidBody* w1 = box(1.0f, V3(0.4f, hy, ez), V3(-wallx, 0.5f * hy, 0.0f), true);
        RigidBody* w2 = box(1.0f, V3(0.4f, hy, ex), V3(0.0f, 0.5f * hy, wallz), true);
        RigidBody* w3 = box(1.0f, V3(0.4f, hy, ex), V3(0.0f, 0.5f * hy, -wallz), true);
        RigidBody* fl = box(1.0f, V3(ex, 0.4f, ez), V3(0.0f, 0.0f, 0.0f), true);
        w2->q = quarter;
        w3->q = quarter;
        for (RigidBody* b : { w0, w1, w2, w3, fl }) sys.addBody(b);
    }
};

// headless.cpp -- the reference's main loop without the viewer (main.cpp:8-16 starts SimViewer, whose draw callback
// builds a scene with Scenarios and calls RigidBodySystem::step(dt / subSteps) every frame, SimViewer.cpp:235-266).
// This program is written against the reference's OWN headers and class names; compiled against
// slrbs_b200/host/include it runs on the B200 through libslrbs_host.so / libslrbs_b200.so:
//
//     g++ -std=c++17 -Islrbs_b200/host/include -Iinclude examples/headless.cpp -Lslrbs_b200 -lslrbs_host -lslrbs_b200
//
// usage: slrbs_headless <marble_box|sphere_on_box|swinging_boxes|cylinder_on_plane|car> [steps] [solverId]
//        slrbs_headless mesh_pile <steps> <solverId> <file.obj> [n] [scale]      (extension: SDF-mesh bodies)
//        slrbs_headless batch <scene> <nScenes> <steps>                          (extension: RigidBodySystemBatch, N copies of a scene)
// prints "i x y z qx qy qz qw" for every body after the last step, then "contacts N"; with SLRBS_DUMP=<file> also writes
// every body's pose after every step (for offline rendering).
#include "rigidbody/RigidBodySystem.h"
#include "rigidbody/RigidBodySystemBatch.h"
#include "rigidbody/RigidBody.h"
#include "rigidbody/Scenarios.h"
#include "contact/Contact.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <string>

int main(int argc, char* argv[])
{
    std::string scene = argc > 1 ? argv[1] : "marble_box";
    int steps = argc > 2 ? std::atoi(argv[2]) : 100;
    // batch <scene> <nScenes> <steps>: the same scene builders, N copies stepped together without readback
    const bool batched = scene == "batch";
    const int nScenes = batched && argc > 3 ? std::atoi(argv[3]) : 1;
    if (batched) { scene = argc > 2 ? argv[2] : "car"; steps = argc > 4 ? std::atoi(argv[4]) : 100; }
    RigidBodySystem sys;
    if (scene == "marble_box") Scenarios::createMarbleBox(sys);
    else if (scene == "sphere_on_box") Scenarios::createSphereOnBox(sys);
    else if (scene == "swinging_boxes") Scenarios::createSwingingBoxes(sys);
    else if (scene == "cylinder_on_plane") Scenarios::createCylinderOnPlane(sys);
    else if (scene == "car") Scenarios::createCarScene(sys);
    else if (scene == "mesh_pile") {                            // extension: slrbs_headless mesh_pile <steps> <solverId> <file.obj> [n] [scale]
        if (argc < 5) { std::fprintf(stderr, "mesh_pile needs an OBJ file\n"); return 2; }
        Scenarios::createMeshPile(sys, argv[4], argc > 5 ? std::atoi(argv[5]) : 5, argc > 6 ? (float)std::atof(argv[6]) : 1.0f);
    }
    else { std::fprintf(stderr, "unknown scene %s\n", scene.c_str()); return 2; }
    if (batched) {
        RigidBodySystemBatch envs(sys, nScenes);
        envs.saveState();
        const auto b0 = std::chrono::steady_clock::now();
        envs.step(0.01f, steps);
        envs.sync();
        const double bms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - b0).count();
        envs.readScene(nScenes - 1, sys);                        // the last scene's state into the prototype's public fields
        int k = 0;
        for (RigidBody* b : sys.getBodies())
            std::printf("%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", k++, b->x[0], b->x[1], b->x[2], b->q.x(), b->q.y(), b->q.z(), b->q.w());
        std::fprintf(stderr, "%s x %d scenes: %d steps in %.1f ms (%.3g body-steps/s)\n", scene.c_str(), nScenes, steps, bms,
                     1e3 * (double)envs.numBodies() * nScenes * steps / bms);
        return 0;
    }
    sys.solverId = argc > 3 ? std::atoi(argv[3]) : 0;            // PGS = 0, CG = 1, CR = 2 (RigidBodySystem.h:56)
    sys.solverIter = 10;
    const float dt = 0.01f;                                      // SimViewer.cpp:113
    const auto t0 = std::chrono::steady_clock::now();
    // state dump for offline rendering: SLRBS_DUMP=<file> writes "step body x y z qx qy qz qw" after every step
    std::FILE* dump = std::getenv("SLRBS_DUMP") ? std::fopen(std::getenv("SLRBS_DUMP"), "w") : nullptr;
    for (int i = 0; i < steps; ++i) {
        sys.step(dt);
        if (dump) {
            int k = 0;
            for (RigidBody* b : sys.getBodies())
                std::fprintf(dump, "%d %d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", i, k++, b->x[0], b->x[1], b->x[2], b->q.x(), b->q.y(), b->q.z(), b->q.w());
        }
    }
    if (dump) std::fclose(dump);
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    int i = 0;
    for (RigidBody* b : sys.getBodies())
        std::printf("%d %.9g %.9g %.9g %.9g %.9g %.9g %.9g\n", i++, b->x[0], b->x[1], b->x[2], b->q.x(), b->q.y(), b->q.z(), b->q.w());
    std::printf("contacts %zu\n", sys.getContacts().size());
    std::fprintf(stderr, "%s: %d steps in %.1f ms\n", scene.c_str(), steps, ms);
    return 0;
}

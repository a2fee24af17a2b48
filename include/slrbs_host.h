/*
 * slrbs_host.h -- C entry points of libslrbs_host.so, the C++ mirror of the reference's class
 * surface (slrbs_b200/host: RigidBodySystem, RigidBody, Joint/Hinge/Spherical, Contact, Solver,
 * Scenarios). These wrappers let a non-C++ harness (bench.py, the tests) build the reference's
 * scenes through Scenarios.h (include/rigidbody/Scenarios.h:18,84,112,153,178), flatten them into
 * the dense arrays of include/slrbs_b200.h, and drive RigidBodySystem::step
 * (src/rigidbody/RigidBodySystem.cpp:52-121) exactly as a C++ caller would.
 */
#ifndef SLRBS_HOST_H
#define SLRBS_HOST_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct slrbs_host_system slrbs_host_system;

const char* slrbs_host_last_error(void);
/* scene: "marble_box", "sphere_on_box", "swinging_boxes", "cylinder_on_plane", "car" (the reference's
 * five) or "sphere_stack" (a = k), "marble_box_n" (a,b,c = nx,nz,ny), "empty". Needs no GPU. */
slrbs_host_system* slrbs_host_scene_create(const char* scene, int a, int b, int c);
void slrbs_host_scene_destroy(slrbs_host_system* s);
int  slrbs_host_num_bodies(const slrbs_host_system* s);
int  slrbs_host_num_joints(const slrbs_host_system* s);
int  slrbs_host_num_contacts(const slrbs_host_system* s);
/* the system flattened into the upload arrays of slrbs_upload_bodies / slrbs_upload_joints */
void slrbs_host_export_bodies(const slrbs_host_system* s, float* x3, float* q4, float* v3, float* w3, float* mass,
                              float* ibody9, float* ibodyinv9, int32_t* fixed, int32_t* geom_type, float* geom6);
void slrbs_host_export_joints(const slrbs_host_system* s, int32_t* type, int32_t* body0, int32_t* body1,
                              float* r0, float* q0, float* r1, float* q1, float* lambda5);
/* public fields solverId / solverIter, Contact::mu, setEnableCollisionDetection */
void slrbs_host_set_solver(slrbs_host_system* s, int solver_id, int solver_iter, float mu, int collisions);
/* a constant extra force on one body applied through setPreStepFunc + addForceAtPos / direct f, tau writes */
void slrbs_host_set_prestep_force(slrbs_host_system* s, int body, const float* f3, const float* tau3);
void slrbs_host_set_state(slrbs_host_system* s, const float* x3, const float* q4, const float* v3, const float* w3);
/* RigidBodySystem::step n times on cuda:device; 0 on success, -1 with slrbs_host_last_error() otherwise */
int  slrbs_host_step(slrbs_host_system* s, float dt, int n, int device);
void slrbs_host_get_state(const slrbs_host_system* s, float* x3, float* q4, float* v3, float* w3);
void slrbs_host_get_contacts(const slrbs_host_system* s, int32_t* body0, int32_t* body1, float* p3, float* n3,
                             float* phi, float* lambda3);
void slrbs_host_get_joint_lambda(const slrbs_host_system* s, float* lambda5);
/* RigidBodySystemState save / restore (src/rigidbody/RigidBodyState.cpp:48-72) */
void slrbs_host_save(slrbs_host_system* s);
void slrbs_host_restore(slrbs_host_system* s);

/* EXTENSION: install a host-side Gauss-Seidel solver written against the reference's Solver.h (only solve(h) overridden) as
 * solver `solver_id`; step() then runs the host-solver path (rows mirrored to the host objects, lambdas sent back) */
void slrbs_host_use_host_solver(slrbs_host_system* s, int solver_id);

/* RigidBodySystemBatch (slrbs_b200/host/include/rigidbody/RigidBodySystemBatch.h): n_scenes copies of a system's bodies, joints
 * and current state in one device context; step is asynchronous and reads nothing back. 0 / -1 + slrbs_host_last_error(). */
typedef struct slrbs_host_batch slrbs_host_batch;
slrbs_host_batch* slrbs_host_batch_create(const slrbs_host_system* prototype, int n_scenes, int device, int contacts_per_scene);
void slrbs_host_batch_destroy(slrbs_host_batch* b);
int  slrbs_host_batch_set_solver(slrbs_host_batch* b, int solver_id, int solver_iter, float mu, int collisions);
int  slrbs_host_batch_step(slrbs_host_batch* b, float dt, int n);
int  slrbs_host_batch_sync(slrbs_host_batch* b);
int  slrbs_host_batch_set_state(slrbs_host_batch* b, int scene0, int n_scenes, const float* x3, const float* q4, const float* v3, const float* w3);
int  slrbs_host_batch_get_state(slrbs_host_batch* b, int scene0, int n_scenes, float* x3, float* q4, float* v3, float* w3);
int  slrbs_host_batch_set_forces(slrbs_host_batch* b, int scene0, int n_scenes, const float* f3, const float* tau3, int absolute);
int  slrbs_host_batch_save(slrbs_host_batch* b);
int  slrbs_host_batch_reset(slrbs_host_batch* b, int scene0, int n_scenes);
int  slrbs_host_batch_contact_counts(slrbs_host_batch* b, int scene0, int n_scenes, int32_t* counts);
int  slrbs_host_batch_read_scene(slrbs_host_batch* b, int scene, slrbs_host_system* into);

#ifdef __cplusplus
}
#endif
#endif

// ref_capi.cpp -- C driver around the REFERENCE'S OWN classes (TEST INFRASTRUCTURE, oracle/_ref only).
//
// oracle/Makefile compiles this file together with the reference's unmodified sources, taken where
// they lie under /root/reference (src/rigidbody/*.cpp, src/collision/CollisionDetect.cpp,
// src/contact/Contact.cpp, src/joint/*.cpp, src/solvers/*.cpp; include/rigidbody/Scenarios.h is
// included below), against oracle/ref_shim/Eigen/Dense (API shim, see its header) and no-op Polyscope /
// MeshAssets headers.  It exports the same orc_* entry points as oracle/slrbs_oracle.h (minus the
// per-stage ones: RigidBodySystem::step is monolithic and its collaborators are private), so
// tests/oracle_lib.py can drive either library and tests/test_ref_vs_oracle.py can pin the C
// restatement against the reference's own code, scene by scene.
//
// Nothing here is product code; nothing under slrbs_b200/ links or loads it.
#include "rigidbody/Scenarios.h"
#include "rigidbody/RigidBodyState.h"
#include "collision/CollisionDetect.h"
#include "contact/Contact.h"

#include <chrono>
#include <cstring>
#include <sstream>

namespace {

struct Pending { int body; bool at_pos; Eigen::Vector3f a, b; };

}  // namespace

struct orc_system {
    RigidBodySystem* sys;
    RigidBodySystemState* saved;
    bool collisions;
    float mu;                          // Contact::mu is process-wide in the reference; set before each step
    std::vector<Pending> pending;      // forces a pre-step callback adds (RigidBodySystem.cpp:70-73)
};

namespace {

// The reference keeps its solvers in a file-static array that every RigidBodySystem constructor
// overwrites (src/rigidbody/RigidBodySystem.cpp:17,29-31): the solvers always serve the system built
// last.  Tests keep several systems alive, so before a system steps it is re-homed into a freshly
// constructed RigidBodySystem (bodies and joints move across through the public accessors).
orc_system* g_latest = nullptr;

void install_hook(orc_system* s)
{
    s->sys->setPreStepFunc([s](std::vector<RigidBody*>& bodies) {
        for (const Pending& p : s->pending) {
            RigidBody* b = bodies[(size_t)p.body];
            if (p.at_pos) b->addForceAtPos(p.a, p.b);
            else { b->f += p.a; b->tau += p.b; }
        }
        s->pending.clear();
    });
}

void rehome(orc_system* s)
{
    if (g_latest == s) return;
    RigidBodySystem* n = new RigidBodySystem();
    n->getBodies().swap(s->sys->getBodies());
    n->getJoints().swap(s->sys->getJoints());
    n->solverIter = s->sys->solverIter;
    n->solverId = s->sys->solverId;
    n->setEnableCollisionDetection(s->collisions);
    delete s->sys;                      // deletes its (stale) contacts; bodies/joints have moved
    s->sys = n;
    install_hook(s);
    g_latest = s;
}

int body_index(const RigidBodySystem* sys, const RigidBody* b)
{
    const auto& bodies = sys->getBodies();
    for (size_t i = 0; i < bodies.size(); ++i) if (bodies[i] == b) return (int)i;
    return -1;
}

void put3(float* dst, int i, const Eigen::Vector3f& v) { if (dst) { dst[3 * i] = v(0); dst[3 * i + 1] = v(1); dst[3 * i + 2] = v(2); } }
void putq(float* dst, int i, const Eigen::Quaternionf& q) { if (dst) { dst[4 * i] = q.x(); dst[4 * i + 1] = q.y(); dst[4 * i + 2] = q.z(); dst[4 * i + 3] = q.w(); } }
void putm(float* dst, int i, const Eigen::Matrix3f& m) { if (dst) for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) dst[9 * i + 3 * r + c] = m(r, c); }
void put_rows(float* dst, int stride, int i, const JBlock& J, int maxrows)
{
    if (!dst) return;
    for (int r = 0; r < maxrows; ++r) for (int k = 0; k < 6; ++k)
        dst[stride * i + 6 * r + k] = (r < J.rows()) ? J(r, k) : 0.f;
}

struct Silence {                         // the scene builders print a banner to std::cout
    std::streambuf* old; std::ostringstream sink;
    Silence() : old(std::cout.rdbuf(sink.rdbuf())) {}
    ~Silence() { std::cout.rdbuf(old); }
};

RigidBody* add_sphere_at(RigidBodySystem& sys, float x, float y, float z)
{
    RigidBody* b = new RigidBody(1.0f, new Sphere(0.5f), createSphere(0.5f));
    b->x = Eigen::Vector3f(x, y, z);
    sys.addBody(b);
    return b;
}

RigidBody* add_fixed_box(RigidBodySystem& sys, const Eigen::Vector3f& dim, const Eigen::Vector3f& x, bool turned)
{
    RigidBody* b = new RigidBody(1.0f, new Box(dim), createBox(dim));
    b->fixed = true;
    b->x = x;
    if (turned) b->q = Eigen::AngleAxisf(1.57f, Eigen::Vector3f(0, 1, 0));
    sys.addBody(b);
    return b;
}

}  // namespace

extern "C" {

const char* orc_impl_name(void) { return "reference sources (sheldona/slrbs) + Eigen API shim"; }

orc_system* orc_create(void)
{
    orc_system* s = new orc_system();
    s->sys = new RigidBodySystem();
    s->saved = nullptr;
    s->collisions = true;
    s->mu = 0.8f;
    g_latest = s;
    install_hook(s);
    return s;
}

void orc_destroy(orc_system* s)
{
    if (!s) return;
    if (g_latest == s) g_latest = nullptr;
    delete s->saved;
    delete s->sys;
    delete s;
}

void orc_clear(orc_system* s) { s->sys->clear(); delete s->saved; s->saved = nullptr; }

int orc_add_body(orc_system* s, float mass, int geom_type, const float* g, int fixed,
                 const float* x3, const float* q, const float* v3, const float* w3)
{
    Geometry* geom = nullptr;
    switch (geom_type) {
    case kSphere:   geom = new Sphere(g[0]); break;
    case kBox:      geom = new Box(Eigen::Vector3f(g[0], g[1], g[2])); break;
    case kPlane:    geom = new Plane(Eigen::Vector3f(g[0], g[1], g[2]), Eigen::Vector3f(g[3], g[4], g[5])); break;
    case kCylinder: geom = new Cylinder(g[0], g[1]); break;
    default: return -1;
    }
    RigidBody* b = new RigidBody(mass, geom, "");
    b->fixed = fixed != 0;
    if (x3) b->x = Eigen::Vector3f(x3[0], x3[1], x3[2]);
    if (q) b->q = Eigen::Quaternionf(q[3], q[0], q[1], q[2]);
    if (v3) b->xdot = Eigen::Vector3f(v3[0], v3[1], v3[2]);
    if (w3) b->omega = Eigen::Vector3f(w3[0], w3[1], w3[2]);
    s->sys->addBody(b);
    return (int)s->sys->getBodies().size() - 1;
}

int orc_add_hinge(orc_system* s, int b0, int b1, const float* r0, const float* q0, const float* r1, const float* q1)
{
    auto& B = s->sys->getBodies();
    s->sys->addJoint(new Hinge(B[(size_t)b0], B[(size_t)b1], Eigen::Vector3f(r0[0], r0[1], r0[2]),
                               Eigen::Quaternionf(q0[3], q0[0], q0[1], q0[2]), Eigen::Vector3f(r1[0], r1[1], r1[2]),
                               Eigen::Quaternionf(q1[3], q1[0], q1[1], q1[2])));
    return (int)s->sys->getJoints().size() - 1;
}

int orc_add_spherical(orc_system* s, int b0, int b1, const float* r0, const float* r1)
{
    auto& B = s->sys->getBodies();
    s->sys->addJoint(new Spherical(B[(size_t)b0], B[(size_t)b1], Eigen::Vector3f(r0[0], r0[1], r0[2]),
                                   Eigen::Vector3f(r1[0], r1[1], r1[2])));
    return (int)s->sys->getJoints().size() - 1;
}

// Contact::mu is one process-wide static in the reference (src/contact/Contact.cpp:4)
void orc_set_params(orc_system* s, float mu, int solver_id, int solver_iter, int collisions)
{
    s->mu = mu;
    s->sys->solverId = solver_id;
    s->sys->solverIter = solver_iter;
    s->collisions = collisions != 0;
    s->sys->setEnableCollisionDetection(s->collisions);
}

void orc_step(orc_system* s, float dt) { rehome(s); Contact::mu = s->mu; s->sys->step(dt); }

void orc_add_force_at_pos(orc_system* s, int body, const float* p, const float* f)
{
    s->pending.push_back({ body, true, Eigen::Vector3f(p[0], p[1], p[2]), Eigen::Vector3f(f[0], f[1], f[2]) });
}
void orc_add_force(orc_system* s, int body, const float* f, const float* tau)
{
    s->pending.push_back({ body, false, Eigen::Vector3f(f[0], f[1], f[2]), Eigen::Vector3f(tau[0], tau[1], tau[2]) });
}

int orc_num_bodies(const orc_system* s) { return (int)s->sys->getBodies().size(); }
int orc_num_joints(const orc_system* s) { return (int)s->sys->getJoints().size(); }
int orc_num_contacts(const orc_system* s) { return (int)s->sys->getContacts().size(); }
int orc_num_joint_rows(const orc_system* s) { int n = 0; for (Joint* j : s->sys->getJoints()) n += (int)j->dim; return n; }

void orc_get_state(const orc_system* s, float* x3, float* q, float* v3, float* w3)
{
    const auto& B = s->sys->getBodies();
    for (size_t i = 0; i < B.size(); ++i) { put3(x3, (int)i, B[i]->x); putq(q, (int)i, B[i]->q); put3(v3, (int)i, B[i]->xdot); put3(w3, (int)i, B[i]->omega); }
}

void orc_set_state(orc_system* s, const float* x3, const float* q, const float* v, const float* w)
{
    auto& B = s->sys->getBodies();
    for (size_t i = 0; i < B.size(); ++i) {
        if (x3) B[i]->x = Eigen::Vector3f(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
        if (q) B[i]->q = Eigen::Quaternionf(q[4 * i + 3], q[4 * i], q[4 * i + 1], q[4 * i + 2]);
        if (v) B[i]->xdot = Eigen::Vector3f(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        if (w) B[i]->omega = Eigen::Vector3f(w[3 * i], w[3 * i + 1], w[3 * i + 2]);
    }
}

void orc_get_body_params(const orc_system* s, float* mass, float* ibody9, float* ibodyinv9, int* fixed, int* geom_type, float* g)
{
    const auto& B = s->sys->getBodies();
    for (size_t i = 0; i < B.size(); ++i) {
        const RigidBody* b = B[i];
        if (mass) mass[i] = b->mass;
        putm(ibody9, (int)i, b->Ibody); putm(ibodyinv9, (int)i, b->IbodyInv);
        if (fixed) fixed[i] = b->fixed ? 1 : 0;
        const eGeometryType t = b->geometry->getType();
        if (geom_type) geom_type[i] = (int)t;
        if (g) {
            float* o = g + 6 * i;
            for (int k = 0; k < 6; ++k) o[k] = 0.f;
            if (t == kSphere) o[0] = dynamic_cast<Sphere*>(b->geometry.get())->radius;
            else if (t == kBox) { const Box* bx = dynamic_cast<Box*>(b->geometry.get()); o[0] = bx->dim[0]; o[1] = bx->dim[1]; o[2] = bx->dim[2]; }
            else if (t == kPlane) { const Plane* pl = dynamic_cast<Plane*>(b->geometry.get()); for (int k = 0; k < 3; ++k) { o[k] = pl->p[k]; o[3 + k] = pl->n[k]; } }
            else if (t == kCylinder) { const Cylinder* cy = dynamic_cast<Cylinder*>(b->geometry.get()); o[0] = cy->height; o[1] = cy->radius; }
        }
    }
}

void orc_get_body_dyn(const orc_system* s, float* f3, float* tau3, float* fc3, float* tauc3, float* I9, float* Iinv9)
{
    const auto& B = s->sys->getBodies();
    for (size_t i = 0; i < B.size(); ++i) {
        put3(f3, (int)i, B[i]->f); put3(tau3, (int)i, B[i]->tau); put3(fc3, (int)i, B[i]->fc); put3(tauc3, (int)i, B[i]->tauc);
        putm(I9, (int)i, B[i]->I); putm(Iinv9, (int)i, B[i]->Iinv);
    }
}

void orc_get_contacts(const orc_system* s, int* body0, int* body1, float* p3, float* n3, float* t3, float* b3, float* phi, float* lambda3)
{
    const auto& Cs = s->sys->getContacts();
    for (size_t i = 0; i < Cs.size(); ++i) {
        const Contact* c = Cs[i];
        if (body0) body0[i] = body_index(s->sys, c->body0);
        if (body1) body1[i] = body_index(s->sys, c->body1);
        put3(p3, (int)i, c->p); put3(n3, (int)i, c->n); put3(t3, (int)i, c->t); put3(b3, (int)i, c->b);
        if (phi) phi[i] = c->pene;
        if (lambda3) for (int k = 0; k < 3; ++k) lambda3[3 * i + k] = c->lambda(k);
    }
}

void orc_get_contact_rows(const orc_system* s, float* J0, float* J1, float* J0M, float* J1M)
{
    const auto& Cs = s->sys->getContacts();
    for (size_t i = 0; i < Cs.size(); ++i) {
        put_rows(J0, 18, (int)i, Cs[i]->J0, 3); put_rows(J1, 18, (int)i, Cs[i]->J1, 3);
        put_rows(J0M, 18, (int)i, Cs[i]->J0Minv, 3); put_rows(J1M, 18, (int)i, Cs[i]->J1Minv, 3);
    }
}

void orc_get_joints(const orc_system* s, int* type, int* body0, int* body1, float* r0, float* q0, float* r1, float* q1)
{
    const auto& Js = s->sys->getJoints();
    for (size_t i = 0; i < Js.size(); ++i) {
        const Joint* j = Js[i];
        if (type) type[i] = (int)j->getType();
        if (body0) body0[i] = body_index(s->sys, j->body0);
        if (body1) body1[i] = body_index(s->sys, j->body1);
        put3(r0, (int)i, j->r0); put3(r1, (int)i, j->r1); putq(q0, (int)i, j->q0); putq(q1, (int)i, j->q1);
    }
}

void orc_get_joint_rows(const orc_system* s, float* J0, float* J1, float* J0M, float* J1M, float* phi5, float* lambda5)
{
    const auto& Js = s->sys->getJoints();
    for (size_t i = 0; i < Js.size(); ++i) {
        const Joint* j = Js[i];
        put_rows(J0, 30, (int)i, j->J0, 5); put_rows(J1, 30, (int)i, j->J1, 5);
        put_rows(J0M, 30, (int)i, j->J0Minv, 5); put_rows(J1M, 30, (int)i, j->J1Minv, 5);
        for (int r = 0; r < 5; ++r) {
            if (phi5) phi5[5 * i + r] = r < (int)j->dim ? j->phi(r) : 0.f;
            if (lambda5) lambda5[5 * i + r] = r < (int)j->dim ? j->lambda(r) : 0.f;
        }
    }
}

void orc_set_joint_lambda(orc_system* s, const float* lambda5)
{
    auto& Js = s->sys->getJoints();
    for (size_t i = 0; i < Js.size(); ++i) for (int r = 0; r < (int)Js[i]->dim; ++r) Js[i]->lambda(r) = lambda5[5 * i + r];
}

// RigidBodySystemState (src/rigidbody/RigidBodyState.cpp:42-72)
void orc_save_state(orc_system* s)
{
    if (!s->saved) s->saved = new RigidBodySystemState(*s->sys);
    else s->saved->save(*s->sys);
}
void orc_restore_state(orc_system* s) { if (s->saved) s->saved->restore(*s->sys); }

void orc_scene_marble_box(orc_system* s)        { Silence q; Scenarios::createMarbleBox(*s->sys); }
void orc_scene_sphere_on_box(orc_system* s)     { Silence q; Scenarios::createSphereOnBox(*s->sys); }
void orc_scene_swinging_boxes(orc_system* s)    { Silence q; Scenarios::createSwingingBoxes(*s->sys); }
void orc_scene_cylinder_on_plane(orc_system* s) { Silence q; Scenarios::createCylinderOnPlane(*s->sys); }
void orc_scene_car(orc_system* s)               { Silence q; Scenarios::createCarScene(*s->sys); }

// synthetic scenes of SURVEY.md 8(d), same construction as oracle/slrbs_oracle.c, built here from the
// reference's own constructors
void orc_scene_sphere_stack(orc_system* s, int k)
{
    s->sys->clear();
    for (int i = 0; i < k; ++i) add_sphere_at(*s->sys, 0.0f, 0.7f + (float)i * 1.0f, 0.0f);
    add_fixed_box(*s->sys, Eigen::Vector3f(10.0f, 0.4f, 10.0f), Eigen::Vector3f(0, 0, 0), false);
}

void orc_scene_marble_box_n(orc_system* s, int nx, int nz, int ny)
{
    s->sys->clear();
    const float x0 = -0.5f * (float)(nx - 1), z0 = -0.5f * (float)(nz - 1);
    for (int i = 0; i < nx; ++i) for (int j = 0; j < nz; ++j) for (int k = 0; k < ny; ++k)
        add_sphere_at(*s->sys, x0 + (float)i * 1.0f, 2.0f + (float)k * 1.0f, z0 + (float)j * 1.0f);
    const float ex = (float)nx + 1.0f, ez = (float)nz + 1.0f, hy = (float)ny + 2.0f;
    const float wallx = 0.5f * ex - 0.25f, wallz = 0.5f * ez - 0.25f;
    add_fixed_box(*s->sys, Eigen::Vector3f(0.4f, hy, ez), Eigen::Vector3f(wallx, 0.5f * hy, 0.0f), false);
    add_fixed_box(*s->sys, Eigen::Vector3f(0.4f, hy, ez), Eigen::Vector3f(-wallx, 0.5f * hy, 0.0f), false);
    add_fixed_box(*s->sys, Eigen::Vector3f(0.4f, hy, ex), Eigen::Vector3f(0.0f, 0.5f * hy, wallz), true);
    add_fixed_box(*s->sys, Eigen::Vector3f(0.4f, hy, ex), Eigen::Vector3f(0.0f, 0.5f * hy, -wallz), true);
    add_fixed_box(*s->sys, Eigen::Vector3f(ex, 0.4f, ez), Eigen::Vector3f(0.0f, 0.0f, 0.0f), false);
}

// timed like the reference's viewer loop (src/viewer/SimViewer.cpp:237-266: chrono around step())
double orc_time_steps(orc_system* s, float dt, int n)
{
    rehome(s);
    Contact::mu = s->mu;
    const auto t0 = std::chrono::steady_clock::now();
    for (int i = 0; i < n; ++i) s->sys->step(dt);
    const auto t1 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"

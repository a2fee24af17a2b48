// host_capi.cpp -- C wrappers of include/slrbs_host.h around the C++ mirror classes.
#include "slrbs_host.h"

#include "contact/Contact.h"
#include "polyscope/polyscope.h"
#include "rigidbody/RigidBodyState.h"
#include "rigidbody/RigidBodySystemBatch.h"
#include "solvers/Solver.h"
#include "rigidbody/Scenarios.h"

#include <algorithm>
#include <cstring>
#include <vector>
#include <memory>
#include <string>

struct slrbs_host_batch { std::unique_ptr<RigidBodySystemBatch> batch; };

struct slrbs_host_system {
    RigidBodySystem sys;
    std::unique_ptr<RigidBodySystemState> saved;
    int forceBody = -1;
    Eigen::Vector3f force, torque;
};

namespace {
thread_local std::string g_err;
int bodyIndex(const RigidBodySystem& sys, const RigidBody* b)
{
    const auto& bodies = sys.getBodies();
    for (size_t i = 0; i < bodies.size(); ++i) if (bodies[i] == b) return (int)i;
    return -1;
}
}

extern "C" {

const char* slrbs_host_last_error(void) { return g_err.c_str(); }

slrbs_host_system* slrbs_host_scene_create(const char* scene, int a, int b, int c)
{
    const std::string name(scene ? scene : "");
    std::unique_ptr<slrbs_host_system> h(new slrbs_host_system());
    if (name == "marble_box") Scenarios::createMarbleBox(h->sys);
    else if (name == "sphere_on_box") Scenarios::createSphereOnBox(h->sys);
    else if (name == "swinging_boxes") Scenarios::createSwingingBoxes(h->sys);
    else if (name == "cylinder_on_plane") Scenarios::createCylinderOnPlane(h->sys);
    else if (name == "car") Scenarios::createCarScene(h->sys);
    else if (name == "sphere_stack") Scenarios::createSphereStack(h->sys, a);
    else if (name == "marble_box_n") Scenarios::createMarbleBoxN(h->sys, a, b, c);
    else if (name == "box_stack") Scenarios::createBoxStack(h->sys, a);
    else if (name == "empty") {}
    else { g_err = "unknown scene: " + name; return nullptr; }
    return h.release();
}

void slrbs_host_scene_destroy(slrbs_host_system* s) { delete s; }
int slrbs_host_num_bodies(const slrbs_host_system* s) { return (int)s->sys.getBodies().size(); }
int slrbs_host_num_joints(const slrbs_host_system* s) { return (int)s->sys.getJoints().size(); }
int slrbs_host_num_contacts(const slrbs_host_system* s) { return (int)s->sys.getContacts().size(); }

void slrbs_host_get_state(const slrbs_host_system* s, float* x3, float* q4, float* v3, float* w3)
{
    const auto& bodies = s->sys.getBodies();
    for (size_t i = 0; i < bodies.size(); ++i) {
        const RigidBody* b = bodies[i];
        for (int k = 0; k < 3; ++k) {
            if (x3) x3[3 * i + k] = b->x[k];
            if (v3) v3[3 * i + k] = b->xdot[k];
            if (w3) w3[3 * i + k] = b->omega[k];
        }
        if (q4) { q4[4 * i] = b->q.x(); q4[4 * i + 1] = b->q.y(); q4[4 * i + 2] = b->q.z(); q4[4 * i + 3] = b->q.w(); }
    }
}

void slrbs_host_set_state(slrbs_host_system* s, const float* x3, const float* q4, const float* v3, const float* w3)
{
    auto& bodies = s->sys.getBodies();
    for (size_t i = 0; i < bodies.size(); ++i) {
        RigidBody* b = bodies[i];
        if (x3) b->x = Eigen::Vector3f(x3[3 * i], x3[3 * i + 1], x3[3 * i + 2]);
        if (q4) b->q = Eigen::Quaternionf(q4[4 * i + 3], q4[4 * i], q4[4 * i + 1], q4[4 * i + 2]);
        if (v3) b->xdot = Eigen::Vector3f(v3[3 * i], v3[3 * i + 1], v3[3 * i + 2]);
        if (w3) b->omega = Eigen::Vector3f(w3[3 * i], w3[3 * i + 1], w3[3 * i + 2]);
    }
}

void slrbs_host_export_bodies(const slrbs_host_system* s, float* x3, float* q4, float* v3, float* w3, float* mass,
                              float* ibody9, float* ibodyinv9, int32_t* fixed, int32_t* geom_type, float* geom6)
{
    slrbs_host_get_state(s, x3, q4, v3, w3);
    const auto& bodies = s->sys.getBodies();
    for (size_t i = 0; i < bodies.size(); ++i) {
        const RigidBody* b = bodies[i];
        if (mass) mass[i] = b->mass;
        if (fixed) fixed[i] = b->fixed ? 1 : 0;
        if (geom_type) geom_type[i] = (int32_t)b->geometry->getType();
        if (geom6) { for (int k = 0; k < 6; ++k) geom6[6 * i + k] = 0.f; b->geometry->packParams(geom6 + 6 * i); }
        if (ibody9) std::memcpy(ibody9 + 9 * i, b->Ibody.data(), 9 * sizeof(float));
        if (ibodyinv9) std::memcpy(ibodyinv9 + 9 * i, b->IbodyInv.data(), 9 * sizeof(float));
    }
}

void slrbs_host_export_joints(const slrbs_host_system* s, int32_t* type, int32_t* body0, int32_t* body1,
                              float* r0, float* q0, float* r1, float* q1, float* lambda5)
{
    const auto& joints = s->sys.getJoints();
    for (size_t i = 0; i < joints.size(); ++i) {
        const Joint* j = joints[i];
        if (type) type[i] = (int32_t)j->getType();
        if (body0) body0[i] = bodyIndex(s->sys, j->body0);
        if (body1) body1[i] = bodyIndex(s->sys, j->body1);
        for (int k = 0; k < 3; ++k) { if (r0) r0[3 * i + k] = j->r0[k]; if (r1) r1[3 * i + k] = j->r1[k]; }
        if (q0) { q0[4 * i] = j->q0.x(); q0[4 * i + 1] = j->q0.y(); q0[4 * i + 2] = j->q0.z(); q0[4 * i + 3] = j->q0.w(); }
        if (q1) { q1[4 * i] = j->q1.x(); q1[4 * i + 1] = j->q1.y(); q1[4 * i + 2] = j->q1.z(); q1[4 * i + 3] = j->q1.w(); }
        if (lambda5) for (int k = 0; k < 5; ++k) lambda5[5 * i + k] = k < (int)j->dim ? j->lambda(k) : 0.f;
    }
}

void slrbs_host_set_solver(slrbs_host_system* s, int solver_id, int solver_iter, float mu, int collisions)
{
    s->sys.solverId = solver_id;
    s->sys.solverIter = solver_iter;
    Contact::mu = mu;
    s->sys.setEnableCollisionDetection(collisions != 0);
}

void slrbs_host_set_prestep_force(slrbs_host_system* s, int body, const float* f3, const float* tau3)
{
    if (body < 0 || (!f3 && !tau3)) { s->forceBody = -1; s->sys.setPreStepFunc(nullptr); return; }
    s->forceBody = body;
    s->force = f3 ? Eigen::Vector3f(f3[0], f3[1], f3[2]) : Eigen::Vector3f(0, 0, 0);
    s->torque = tau3 ? Eigen::Vector3f(tau3[0], tau3[1], tau3[2]) : Eigen::Vector3f(0, 0, 0);
    slrbs_host_system* self = s;
    s->sys.setPreStepFunc([self](std::vector<RigidBody*>& bodies) {
        RigidBody* b = bodies[(size_t)self->forceBody];
        b->f += self->force;
        b->tau += self->torque;
    });
}

int slrbs_host_step(slrbs_host_system* s, float dt, int n, int device)
{
    try {
        s->sys.setDevice(device);
        for (int i = 0; i < n; ++i) s->sys.step(dt);
        return 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}

void slrbs_host_get_contacts(const slrbs_host_system* s, int32_t* body0, int32_t* body1, float* p3, float* n3, float* phi, float* lambda3)
{
    const auto& cs = s->sys.getContacts();
    for (size_t i = 0; i < cs.size(); ++i) {
        const Contact* c = cs[i];
        if (body0) body0[i] = bodyIndex(s->sys, c->body0);
        if (body1) body1[i] = bodyIndex(s->sys, c->body1);
        for (int k = 0; k < 3; ++k) {
            if (p3) p3[3 * i + k] = c->p[k];
            if (n3) n3[3 * i + k] = c->n[k];
            if (lambda3) lambda3[3 * i + k] = c->lambda(k);
        }
        if (phi) phi[i] = c->pene;
    }
}

void slrbs_host_get_joint_lambda(const slrbs_host_system* s, float* lambda5)
{
    slrbs_host_export_joints(s, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, lambda5);
}

void slrbs_host_save(slrbs_host_system* s)
{
    if (!s->saved) s->saved.reset(new RigidBodySystemState(s->sys));
    else s->saved->save(s->sys);
}

void slrbs_host_restore(slrbs_host_system* s)
{
    if (s->saved) s->saved->restore(s->sys);
}

// ---- a Solver subclass written against the reference's header: only solve(h) is overridden, so it runs on the host -----------
// Projected Gauss-Seidel over the mirrored constraint objects, in the reference's row order (joints, then contacts) with one
// velocity-impulse accumulator per body -- the same maths as SolverBoxPGS.cpp:126-250 written from the public fields a user
// solver sees (J0, J1, J0Minv, J1Minv, phi, lambda, body->f / tau / xdot / omega / fixed). Used by the tests to exercise the
// host-solver path of RigidBodySystem::step.
namespace {
class HostGaussSeidel : public Solver {
public:
    explicit HostGaussSeidel(RigidBodySystem* sys) : Solver(sys) {}
    void solve(float h) override
    {
        auto& bodies = m_rigidBodySystem->getBodies();
        std::vector<Joint*> rows(m_rigidBodySystem->getJoints().begin(), m_rigidBodySystem->getJoints().end());
        const size_t nJoints = rows.size();
        for (Contact* c : m_rigidBodySystem->getContacts()) { c->lambda.setZero(); rows.push_back(c); }
        std::vector<int> index(bodies.size());
        std::vector<std::vector<float>> w(bodies.size(), std::vector<float>(6, 0.f));      // sum of J^T lambda per body
        auto slot = [&](const RigidBody* b) { for (size_t i = 0; i < bodies.size(); ++i) if (bodies[i] == b) return (int)i; return -1; };
        struct Row { int b0, b1; std::vector<float> A, rhs; };
        std::vector<Row> R(rows.size());
        for (size_t i = 0; i < rows.size(); ++i) {
            Joint* j = rows[i];
            const int d = (int)j->dim;
            Row& r = R[i];
            r.b0 = j->body0->fixed ? -1 : slot(j->body0); r.b1 = j->body1->fixed ? -1 : slot(j->body1);
            r.A.assign((size_t)d * d, 0.f); r.rhs.assign(d, 0.f);
            for (int a = 0; a < d; ++a) {
                float b = -(0.3f / h) * j->phi(a);
                const RigidBody* bs[2] = { j->body0, j->body1 };
                const JBlock* Js[2] = { &j->J0, &j->J1 };
                const JBlock* Ms[2] = { &j->J0Minv, &j->J1Minv };
                for (int s = 0; s < 2; ++s) {
                    if (bs[s]->fixed) continue;
                    for (int k = 0; k < 3; ++k) {
                        b -= h * (*Ms[s])(a, k) * bs[s]->f[k] + h * (*Ms[s])(a, 3 + k) * bs[s]->tau[k];
                        b -= (*Js[s])(a, k) * bs[s]->xdot[k] + (*Js[s])(a, 3 + k) * bs[s]->omega[k];
                    }
                    for (int c = 0; c < d; ++c) for (int k = 0; k < 6; ++k) r.A[(size_t)a * d + c] += (*Ms[s])(a, k) * (*Js[s])(c, k);
                }
                r.rhs[a] = b;
                if (i < nJoints) r.A[(size_t)a * d + a] += 1e-5f;
            }
            for (int a = 0; a < d; ++a) {                        // warm start: joints keep last step's impulses
                if (r.b0 >= 0) for (int k = 0; k < 6; ++k) w[r.b0][k] += j->J0(a, k) * j->lambda(a);
                if (r.b1 >= 0) for (int k = 0; k < 6; ++k) w[r.b1][k] += j->J1(a, k) * j->lambda(a);
            }
        }
        for (int it = 0; it < m_maxIter; ++it)
            for (size_t i = 0; i < rows.size(); ++i) {
                Joint* j = rows[i];
                const int d = (int)j->dim;
                const Row& r = R[i];
                std::vector<float> x(r.rhs);
                for (int a = 0; a < d; ++a) {                    // x = b - sum over the OTHER constraints of the two bodies
                    for (int k = 0; k < 6; ++k) {
                        if (r.b0 >= 0) { float own = 0.f; for (int c = 0; c < d; ++c) own += j->J0(c, k) * j->lambda(c); x[a] -= j->J0Minv(a, k) * (w[r.b0][k] - own); }
                        if (r.b1 >= 0) { float own = 0.f; for (int c = 0; c < d; ++c) own += j->J1(c, k) * j->lambda(c); x[a] -= j->J1Minv(a, k) * (w[r.b1][k] - own); }
                    }
                }
                std::vector<float> lam(d);
                if (i >= nJoints) {                              // contact: normal >= 0, tangents inside the friction box
                    const float* A = r.A.data();
                    lam[0] = std::max(0.f, (x[0] - A[1] * j->lambda(1) - A[2] * j->lambda(2)) / (A[0] + 1e-3f));
                    const float lim = Contact::mu * lam[0];
                    lam[1] = std::min(lim, std::max(-lim, (x[1] - A[3] * lam[0] - A[5] * j->lambda(2)) / A[4]));
                    lam[2] = std::min(lim, std::max(-lim, (x[2] - A[6] * lam[0] - A[7] * lam[1]) / A[8]));
                } else {                                         // joint: dense solve of the d x d block (Gaussian elimination)
                    std::vector<float> M(r.A), y(x);
                    for (int c = 0; c < d; ++c) for (int a = c + 1; a < d; ++a) {
                        const float m = M[(size_t)a * d + c] / M[(size_t)c * d + c];
                        for (int k = c; k < d; ++k) M[(size_t)a * d + k] -= m * M[(size_t)c * d + k];
                        y[a] -= m * y[c];
                    }
                    for (int a = d - 1; a >= 0; --a) {
                        float sum = y[a];
                        for (int k = a + 1; k < d; ++k) sum -= M[(size_t)a * d + k] * lam[k];
                        lam[a] = sum / M[(size_t)a * d + a];
                    }
                }
                for (int a = 0; a < d; ++a) {
                    const float dl = lam[a] - j->lambda(a);
                    if (r.b0 >= 0) for (int k = 0; k < 6; ++k) w[r.b0][k] += j->J0(a, k) * dl;
                    if (r.b1 >= 0) for (int k = 0; k < 6; ++k) w[r.b1][k] += j->J1(a, k) * dl;
                    j->lambda(a) = lam[a];
                }
            }
    }
};
}  // namespace

/* replaces the solver behind solver_id (0..3) by a host-side Gauss-Seidel written against the reference's Solver.h
 * (deviceSolverId() not overridden): RigidBodySystem::step then runs the host-solver path */
void slrbs_host_use_host_solver(slrbs_host_system* s, int solver_id)
{
    s->sys.setSolver(solver_id, new HostGaussSeidel(&s->sys));
}

/* ---- RigidBodySystemBatch ------------------------------------------------------------------------------------------- */
slrbs_host_batch* slrbs_host_batch_create(const slrbs_host_system* proto, int n_scenes, int device, int contacts_per_scene)
{
    try {
        std::unique_ptr<slrbs_host_batch> h(new slrbs_host_batch());
        h->batch.reset(new RigidBodySystemBatch(proto->sys, n_scenes, device, contacts_per_scene));
        return h.release();
    } catch (const std::exception& e) { g_err = e.what(); return nullptr; }
}
void slrbs_host_batch_destroy(slrbs_host_batch* b) { delete b; }
#define BATCH_TRY(stmt) try { stmt; return 0; } catch (const std::exception& e) { g_err = e.what(); return -1; }
int slrbs_host_batch_set_solver(slrbs_host_batch* b, int solver_id, int solver_iter, float mu, int collisions)
{
    b->batch->solverId = solver_id; b->batch->solverIter = solver_iter; b->batch->mu = mu;
    b->batch->setEnableCollisionDetection(collisions != 0);
    return 0;
}
int slrbs_host_batch_step(slrbs_host_batch* b, float dt, int n) { BATCH_TRY(b->batch->step(dt, n)) }
int slrbs_host_batch_sync(slrbs_host_batch* b) { BATCH_TRY(b->batch->sync()) }
int slrbs_host_batch_set_state(slrbs_host_batch* b, int scene0, int ns, const float* x3, const float* q4, const float* v3, const float* w3)
{ BATCH_TRY(b->batch->setState(scene0, ns, x3, q4, v3, w3)) }
int slrbs_host_batch_get_state(slrbs_host_batch* b, int scene0, int ns, float* x3, float* q4, float* v3, float* w3)
{ BATCH_TRY(b->batch->getState(scene0, ns, x3, q4, v3, w3)) }
int slrbs_host_batch_set_forces(slrbs_host_batch* b, int scene0, int ns, const float* f3, const float* tau3, int absolute)
{ BATCH_TRY(if (f3 || tau3) b->batch->setExternalForces(scene0, ns, f3, tau3, absolute != 0); else b->batch->clearExternalForces(scene0, ns)) }
int slrbs_host_batch_save(slrbs_host_batch* b) { BATCH_TRY(b->batch->saveState()) }
int slrbs_host_batch_reset(slrbs_host_batch* b, int scene0, int ns) { BATCH_TRY(b->batch->resetScenes(scene0, ns)) }
int slrbs_host_batch_contact_counts(slrbs_host_batch* b, int scene0, int ns, int32_t* counts) { BATCH_TRY(b->batch->getContactCounts(scene0, ns, counts)) }
int slrbs_host_batch_read_scene(slrbs_host_batch* b, int scene, slrbs_host_system* into) { BATCH_TRY(b->batch->readScene(scene, into->sys)) }
#undef BATCH_TRY

}  // extern "C"

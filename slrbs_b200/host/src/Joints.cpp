// Joints.cpp -- host mirror of src/joint/{Joint,Hinge,Spherical}.cpp and src/contact/Contact.cpp.
// The per-step Jacobians are computed on the device; computeJacobian() here evaluates the same
// rows from the host fields for callers that inspect a constraint between steps.
#include "contact/Contact.h"
#include "joint/Hinge.h"
#include "joint/Spherical.h"
#include "rigidbody/RigidBody.h"

float Contact::mu = 0.8f;

Joint::Joint() : body0(nullptr), body1(nullptr), idx(0), dim(0), q0(Eigen::Quaternionf::Identity()), q1(Eigen::Quaternionf::Identity()) {}
Joint::Joint(RigidBody* _body0, RigidBody* _body1)
    : body0(_body0), body1(_body1), idx(0), dim(0), q0(Eigen::Quaternionf::Identity()), q1(Eigen::Quaternionf::Identity()) {}
Joint::Joint(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Quaternionf& _q0,
             const Eigen::Vector3f& _r1, const Eigen::Quaternionf& _q1)
    : body0(_body0), body1(_body1), idx(0), dim(0), r0(_r0), r1(_r1), q0(_q0), q1(_q1) {}

void Joint::sizeBlocks(int rows)
{
    dim = (unsigned)rows;
    J0.setZero(rows, 6); J1.setZero(rows, 6); J0Minv.setZero(rows, 6); J1Minv.setZero(rows, 6);
    phi.setZero(rows); lambda.setZero(rows);
}

// J M^-1: linear block scaled by 1/mass, angular block times the world inverse inertia
void Joint::fillMinv()
{
    const RigidBody* bodies[2] = { body0, body1 };
    JBlock* Js[2] = { &J0, &J1 };
    JBlock* Ms[2] = { &J0Minv, &J1Minv };
    for (int s = 0; s < 2; ++s) {
        const float im = 1.0f / bodies[s]->mass;
        const Eigen::Matrix3f& Ii = bodies[s]->Iinv;
        for (unsigned r = 0; r < dim; ++r) {
            for (int c = 0; c < 3; ++c) (*Ms[s])(r, c) = im * (*Js[s])(r, c);
            for (int c = 0; c < 3; ++c)
                (*Ms[s])(r, 3 + c) = (*Js[s])(r, 3) * Ii(0, c) + ((*Js[s])(r, 4) * Ii(1, c) + (*Js[s])(r, 5) * Ii(2, c));
        }
    }
}

static void putRow(JBlock& J, int r, const Eigen::Vector3f& lin, const Eigen::Vector3f& ang)
{
    for (int c = 0; c < 3; ++c) { J(r, c) = lin[c]; J(r, 3 + c) = ang[c]; }
}
// rows 0..2 of a point constraint: [ +-I | hat(a) ]
static void putPointRows(JBlock& J, float sign, const Eigen::Vector3f& a)
{
    putRow(J, 0, Eigen::Vector3f(sign, 0, 0), Eigen::Vector3f(0.f, -a[2], a[1]));
    putRow(J, 1, Eigen::Vector3f(0, sign, 0), Eigen::Vector3f(a[2], 0.f, -a[0]));
    putRow(J, 2, Eigen::Vector3f(0, 0, sign), Eigen::Vector3f(-a[1], a[0], 0.f));
}

Hinge::Hinge(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Quaternionf& _q0,
             const Eigen::Vector3f& _r1, const Eigen::Quaternionf& _q1)
    : Joint(_body0, _body1, _r0, _q0, _r1, _q1)
{
    sizeBlocks(5);
}

void Hinge::computeJacobian()
{
    const Eigen::Vector3f rr0 = body0->q * r0, rr1 = body1->q * r1;
    const Eigen::Vector3f nn = body0->q * (q0 * Eigen::Vector3f(1, 0, 0));
    const Eigen::Vector3f uu = body1->q * (q1 * Eigen::Vector3f(0, 1, 0));
    const Eigen::Vector3f vv = body1->q * (q1 * Eigen::Vector3f(0, 0, 1));
    const Eigen::Vector3f nxu = nn.cross(uu), nxv = nn.cross(vv);
    const Eigen::Vector3f e = ((body0->x + rr0) - body1->x) - rr1;
    for (int k = 0; k < 3; ++k) phi(k) = e[k];
    phi(3) = nn.dot(uu);
    phi(4) = nn.dot(vv);
    putPointRows(J0, 1.f, -rr0);
    putPointRows(J1, -1.f, rr1);
    const Eigen::Vector3f zero(0, 0, 0);
    putRow(J0, 3, zero, nxu); putRow(J0, 4, zero, nxv);
    putRow(J1, 3, zero, -nxu); putRow(J1, 4, zero, -nxv);
    fillMinv();
}

Spherical::Spherical(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Vector3f& _r1)
    : Joint(_body0, _body1, _r0, Eigen::Quaternionf::Identity(), _r1, Eigen::Quaternionf::Identity())
{
    sizeBlocks(3);
}

void Spherical::computeJacobian()
{
    const Eigen::Vector3f rr0 = body0->q * r0, rr1 = body1->q * r1;
    const Eigen::Vector3f e = ((body0->x + rr0) - body1->x) - rr1;
    for (int k = 0; k < 3; ++k) phi(k) = e[k];
    putPointRows(J0, 1.f, -rr0);
    putPointRows(J1, -1.f, rr1);
    fillMinv();
}

Contact::Contact(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _p, const Eigen::Vector3f& _n, float _phi)
    : Joint(_body0, _body1), p(_p), n(_n), pene(_phi)
{
    sizeBlocks(3);
    phi(0) = _phi;
    body0->contacts.push_back(this);
    body1->contacts.push_back(this);
}

void Contact::computeContactFrame()
{
    t = n.cross(Eigen::Vector3f(1, 0, 0));
    if (t.norm() < 1e-5f) t = n.cross(Eigen::Vector3f(0, 1, 0));
    t.normalize();
    b = n.cross(t);
    b.normalize();
}

void Contact::computeJacobian()
{
    const Eigen::Vector3f rr0 = p - body0->x, rr1 = p - body1->x;
    sizeBlocks(3);
    phi(0) = pene;
    const Eigen::Vector3f d[3] = { n, t, b };
    for (int k = 0; k < 3; ++k) {
        putRow(J0, k, d[k], rr0.cross(d[k]));
        putRow(J1, k, -d[k], -(rr1.cross(d[k])));
    }
    fillMinv();
}

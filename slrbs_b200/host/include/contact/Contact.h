// contact/Contact.h -- contact constraint with box friction (include/contact/Contact.h:11-48).
// Contacts are produced by the device narrow phase each step and mirrored into these objects.
#pragma once
#include "joint/Joint.h"

class Contact : public Joint {
public:
    static float mu;            // process-wide friction coefficient, as in the reference (Contact.cpp:4)

    Contact(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _p, const Eigen::Vector3f& _n, float _phi);
    eConstraintType getType() const override { return kContact; }

    Eigen::Vector3f p, n, t, b;
    float pene;

    void computeJacobian() override;
    void computeContactFrame();
};

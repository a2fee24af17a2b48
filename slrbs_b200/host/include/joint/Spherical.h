// joint/Spherical.h -- 3-row ball joint (include/joint/Spherical.h, src/joint/Spherical.cpp:22-52)
#pragma once
#include "joint/Joint.h"

class Spherical : public Joint {
public:
    Spherical(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Vector3f& _r1);
    eConstraintType getType() const override { return kSpherical; }
    void computeJacobian() override;
};

// joint/Hinge.h -- 5-row hinge (include/joint/Hinge.h, src/joint/Hinge.cpp:21-63)
#pragma once
#include "joint/Joint.h"

class Hinge : public Joint {
public:
    Hinge(RigidBody* _body0, RigidBody* _body1, const Eigen::Vector3f& _r0, const Eigen::Quaternionf& _q0,
          const Eigen::Vector3f& _r1, const Eigen::Quaternionf& _q1);
    eConstraintType getType() const override { return kHinge; }
    void computeJacobian() override;
};

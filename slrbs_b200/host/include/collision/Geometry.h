// collision/Geometry.h -- the reference's geometry hierarchy (include/collision/Geometry.h): same
// class names, public fields, constructors and type ids; inertia formulas per :38-43,63-70,91-102,120-124.
#pragma once
#include "util/Types.h"

enum eGeometryType { kSphere, kBox, kPlane, kCylinder, kSDFMesh /* extension, collision/SDFMesh.h */ };

class Geometry {
public:
    virtual ~Geometry() {}
    virtual Eigen::Matrix3f computeInertia(float _mass) = 0;
    virtual eGeometryType getType() const = 0;
    // six floats in the C ABI's geom6 convention (include/slrbs_b200.h)
    virtual void packParams(float* g6) const = 0;
protected:
    static Eigen::Matrix3f diag(float a, float b, float c)
    {
        Eigen::Matrix3f I;
        I(0, 0) = a; I(1, 1) = b; I(2, 2) = c;
        return I;
    }
};

class Sphere : public Geometry {
public:
    float radius;
    Sphere(float _radius) : radius(_radius) {}
    Eigen::Matrix3f computeInertia(float m) override { const float v = (2.0f / 5.0f) * m * radius * radius; return diag(v, v, v); }
    eGeometryType getType() const override { return kSphere; }
    void packParams(float* g) const override { g[0] = radius; }
};

class Box : public Geometry {
public:
    Eigen::Vector3f dim;
    Box(const Eigen::Vector3f& _dim) : dim(_dim) {}
    Eigen::Matrix3f computeInertia(float m) override
    {
        const float k = 1.0f / 12.0f;
        return diag(k * m * (dim[1] * dim[1] + dim[2] * dim[2]), k * m * (dim[0] * dim[0] + dim[2] * dim[2]),
                    k * m * (dim[0] * dim[0] + dim[1] * dim[1]));
    }
    eGeometryType getType() const override { return kBox; }
    void packParams(float* g) const override { g[0] = dim[0]; g[1] = dim[1]; g[2] = dim[2]; }
};

class Cylinder : public Geometry {
public:
    float height, radius;        // axis = local y
    Cylinder(float _height, float _radius) : height(_height), radius(_radius) {}
    Eigen::Matrix3f computeInertia(float m) override
    {
        const float s = 1.0f / 12.0f;
        const float h2 = height * height, r2 = radius * radius;
        const float lat = (3.0f * r2 + h2);
        return diag(s * m * lat, 0.5f * m * r2, s * m * lat);
    }
    eGeometryType getType() const override { return kCylinder; }
    void packParams(float* g) const override { g[0] = height; g[1] = radius; }
};

class Plane : public Geometry {
public:
    Eigen::Vector3f p, n;
    Plane(const Eigen::Vector3f& _p, const Eigen::Vector3f& _n) : p(_p), n(_n) {}
    Eigen::Matrix3f computeInertia(float) override { return Eigen::Matrix3f::Zero(); }   // static bodies only
    eGeometryType getType() const override { return kPlane; }
    void packParams(float* g) const override { for (int i = 0; i < 3; ++i) { g[i] = p[i]; g[3 + i] = n[i]; } }
};

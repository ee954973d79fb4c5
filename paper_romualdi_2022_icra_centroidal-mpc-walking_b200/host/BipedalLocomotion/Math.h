// Math.h -- minimal stand-ins for the Eigen / manif / BLF::Math types that appear in the reference's calls into
// BipedalLocomotion::ReducedModelControllers::CentroidalMPC (src/centroidal-mpc-walking/src/CentroidalMPCBlock.cpp:405-407,
// 579; src/WholeBodyQPBlock.cpp:824-828): Eigen::Vector3d, manif::SE3d (pose with translation() / rotation() / quat()),
// BLF Math::Wrenchd (force() / torque()).  Eigen, manif and BLF are not available in this build environment; the names and
// accessors are kept so that the reference's call sites read the same.
#pragma once

#include <array>
#include <cmath>

namespace Eigen {
struct Vector3d {
    double v[3] = {0.0, 0.0, 0.0};
    Vector3d() = default;
    Vector3d(double x, double y, double z) : v{x, y, z} {}
    double& operator()(int i) { return v[i]; }
    double operator()(int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    double operator[](int i) const { return v[i]; }
    double* data() { return v; }
    const double* data() const { return v; }
    static Vector3d Zero() { return Vector3d(); }
    void setZero() { v[0] = v[1] = v[2] = 0.0; }
    Vector3d operator+(const Vector3d& o) const { return {v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]}; }
    Vector3d operator-(const Vector3d& o) const { return {v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]}; }
    Vector3d operator*(double s) const { return {v[0] * s, v[1] * s, v[2] * s}; }
    Vector3d& operator+=(const Vector3d& o) { for (int i = 0; i < 3; ++i) v[i] += o.v[i]; return *this; }
    Vector3d& operator/=(double s) { for (int i = 0; i < 3; ++i) v[i] /= s; return *this; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    Vector3d cross(const Vector3d& o) const
    { return {v[1] * o.v[2] - v[2] * o.v[1], v[2] * o.v[0] - v[0] * o.v[2], v[0] * o.v[1] - v[1] * o.v[0]}; }
};
// 3 x 3, column major like Eigen's default (and like vec(R) in the MPC parameter vector)
struct Matrix3d {
    double m[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    double& operator()(int r, int c) { return m[3 * c + r]; }
    double operator()(int r, int c) const { return m[3 * c + r]; }
    const double* data() const { return m; }
    static Matrix3d Identity() { return Matrix3d(); }
    Vector3d operator*(const Vector3d& x) const
    {
        Vector3d y;
        for (int r = 0; r < 3; ++r) y.v[r] = (*this)(r, 0) * x.v[0] + (*this)(r, 1) * x.v[1] + (*this)(r, 2) * x.v[2];
        return y;
    }
    Matrix3d transpose() const
    {
        Matrix3d t;
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) t(r, c) = (*this)(c, r);
        return t;
    }
};
struct Quaterniond {  // w, x, y, z
    double w = 1, x = 0, y = 0, z = 0;
    Quaterniond() = default;
    Quaterniond(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
    Matrix3d toRotationMatrix() const
    {
        const double n = std::sqrt(w * w + x * x + y * y + z * z);
        const double a = w / n, b = x / n, c = y / n, d = z / n;
        Matrix3d R;
        R(0, 0) = 1 - 2 * (c * c + d * d); R(0, 1) = 2 * (b * c - a * d); R(0, 2) = 2 * (b * d + a * c);
        R(1, 0) = 2 * (b * c + a * d); R(1, 1) = 1 - 2 * (b * b + d * d); R(1, 2) = 2 * (c * d - a * b);
        R(2, 0) = 2 * (b * d - a * c); R(2, 1) = 2 * (c * d + a * b); R(2, 2) = 1 - 2 * (b * b + c * c);
        return R;
    }
};
}  // namespace Eigen

namespace manif {
// pose of a contact: translation + rotation (the subset of manif::SE3d the MPC path touches)
class SE3d {
public:
    SE3d() = default;
    SE3d(const Eigen::Vector3d& t, const Eigen::Quaterniond& q) : m_t(t), m_R(q.toRotationMatrix()) {}
    SE3d(const Eigen::Vector3d& t, const Eigen::Matrix3d& R) : m_t(t), m_R(R) {}
    static SE3d Identity() { return SE3d(); }
    const Eigen::Vector3d& translation() const { return m_t; }
    void translation(const Eigen::Vector3d& t) { m_t = t; }
    const Eigen::Matrix3d& rotation() const { return m_R; }
    void rotation(const Eigen::Matrix3d& R) { m_R = R; }
    Eigen::Vector3d act(const Eigen::Vector3d& p) const { return m_R * p + m_t; }
    static SE3d fromYaw(const Eigen::Vector3d& t, double yaw)
    {
        Eigen::Matrix3d R;
        R(0, 0) = std::cos(yaw); R(0, 1) = -std::sin(yaw); R(1, 0) = std::sin(yaw); R(1, 1) = std::cos(yaw);
        return SE3d(t, R);
    }
private:
    Eigen::Vector3d m_t;
    Eigen::Matrix3d m_R;
};
}  // namespace manif

namespace BipedalLocomotion {
namespace Math {
// BLF Math::Wrenchd: 6-vector (force, torque)
class Wrenchd {
public:
    Eigen::Vector3d& force() { return m_f; }
    const Eigen::Vector3d& force() const { return m_f; }
    Eigen::Vector3d& torque() { return m_t; }
    const Eigen::Vector3d& torque() const { return m_t; }
    static Wrenchd Zero() { return Wrenchd(); }
    void setZero() { m_f.setZero(); m_t.setZero(); }
private:
    Eigen::Vector3d m_f, m_t;
};
}  // namespace Math
}  // namespace BipedalLocomotion

// Basic numeric types of the host surface (mirrors src/lib/MeshFEM/Types.hh:8 --
// Real is double everywhere -- without Eigen, which is not available offline).
#ifndef MESHFEM_B200_TYPES_HH
#define MESHFEM_B200_TYPES_HH
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <initializer_list>
#include <iosfwd>
#include <stdexcept>
#include <string>
#include <vector>

typedef double Real;

// Fixed-size column vector with the handful of Eigen operations the path uses.
template <size_t N>
struct VectorND {
    std::array<Real, N> v{};
    VectorND() { v.fill(0.0); }
    VectorND(std::initializer_list<Real> l) {
        v.fill(0.0);
        size_t i = 0;
        for (Real x : l) { if (i < N) v[i++] = x; }
    }
    static VectorND Zero() { return VectorND(); }
    static constexpr size_t size() { return N; }
    Real &operator[](size_t i) { return v[i]; }
    Real operator[](size_t i) const { return v[i]; }
    VectorND &operator+=(const VectorND &b) { for (size_t i = 0; i < N; ++i) v[i] += b[i]; return *this; }
    VectorND &operator-=(const VectorND &b) { for (size_t i = 0; i < N; ++i) v[i] -= b[i]; return *this; }
    VectorND &operator*=(Real s) { for (size_t i = 0; i < N; ++i) v[i] *= s; return *this; }
    VectorND &operator/=(Real s) { for (size_t i = 0; i < N; ++i) v[i] /= s; return *this; }
    friend VectorND operator+(VectorND a, const VectorND &b) { return a += b; }
    friend VectorND operator-(VectorND a, const VectorND &b) { return a -= b; }
    friend VectorND operator*(Real s, VectorND a) { return a *= s; }
    friend VectorND operator*(VectorND a, Real s) { return a *= s; }
    friend VectorND operator/(VectorND a, Real s) { return a /= s; }
    VectorND operator-() const { VectorND r; for (size_t i = 0; i < N; ++i) r[i] = -v[i]; return r; }
    Real dot(const VectorND &b) const { Real s = 0; for (size_t i = 0; i < N; ++i) s += v[i] * b[i]; return s; }
    Real squaredNorm() const { return dot(*this); }
    Real norm() const { return std::sqrt(squaredNorm()); }
    VectorND cwiseMin(const VectorND &b) const { VectorND r; for (size_t i = 0; i < N; ++i) r[i] = v[i] < b[i] ? v[i] : b[i]; return r; }
    VectorND cwiseMax(const VectorND &b) const { VectorND r; for (size_t i = 0; i < N; ++i) r[i] = v[i] > b[i] ? v[i] : b[i]; return r; }
    bool operator==(const VectorND &b) const { return v == b.v; }
};

typedef VectorND<2> Vector2D;
typedef VectorND<3> Vector3D;
typedef Vector2D Point2D;
typedef Vector3D Point3D;
template <size_t N> using PointND = VectorND<N>;
template <size_t N> using IVectorND = std::array<int, N>;

inline Vector3D cross(const Vector3D &a, const Vector3D &b) {
    return Vector3D{a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
}

// truncateFrom3D<VectorND<N>> (Types.hh): drop trailing components
template <class Vec>
inline Vec truncateFrom3D(const Vector3D &p) {
    Vec r;
    for (size_t i = 0; i < Vec::size(); ++i) r[i] = p[i];
    return r;
}
template <size_t N>
inline Vector3D padTo3D(const VectorND<N> &p) {
    Vector3D r;
    for (size_t i = 0; i < N; ++i) r[i] = p[i];
    return r;
}
#endif

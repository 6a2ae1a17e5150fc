// Regions used by boundary conditions (mirrors the used part of Geometry.hh:150-300):
// BBox with inclusive containment (Geometry.hh:276-279) and box% interpolation (:259-262).
#ifndef MESHFEM_B200_GEOMETRY_HH
#define MESHFEM_B200_GEOMETRY_HH
#include <MeshFEM/Types.hh>

#include <algorithm>

template <class _Vector>
struct Region {
    typedef _Vector Vector;
    Vector minCorner, maxCorner;
    virtual bool containsPoint(const Vector &p) const = 0;
    Vector dimensions() const { return maxCorner - minCorner; }
    virtual ~Region() {}
};

template <class _Vector>
struct BBox : public Region<_Vector> {
    typedef _Vector Vector;
    BBox() {}
    BBox(const Vector &mn, const Vector &mx) { this->minCorner = mn; this->maxCorner = mx; }
    void unionPoint(const Vector &p) {
        this->minCorner = this->minCorner.cwiseMin(p);
        this->maxCorner = this->maxCorner.cwiseMax(p);
    }
    Vector interpolatePoint(const Vector &v) const {
        Vector r;
        for (size_t i = 0; i < Vector::size(); ++i)
            r[i] = this->minCorner[i] + v[i] * (this->maxCorner[i] - this->minCorner[i]);
        return r;
    }
    Vector center() const { return 0.5 * (this->minCorner + this->maxCorner); }
    bool containsPoint(const Vector &p) const override {
        for (size_t i = 0; i < Vector::size(); ++i)
            if (!(p[i] >= this->minCorner[i]) || !(p[i] <= this->maxCorner[i])) return false;
        return true;
    }
    Real volume() const {
        Real r = 1.0;
        for (size_t i = 0; i < Vector::size(); ++i) r *= this->maxCorner[i] - this->minCorner[i];
        return r;
    }
};

// Unordered index tuples (Geometry.hh:393-447) for element-indexed boundary conditions.
struct UnorderedTriplet {
    int v[3];
    UnorderedTriplet(int v0, int v1, int v2) {
        v[0] = std::min(v0, std::min(v1, v2));
        v[2] = std::max(v0, std::max(v1, v2));
        v[1] = v0 ^ v1 ^ v2 ^ v[0] ^ v[2];
    }
    bool operator<(const UnorderedTriplet &b) const {
        if (v[0] != b.v[0]) return v[0] < b.v[0];
        if (v[1] != b.v[1]) return v[1] < b.v[1];
        return v[2] < b.v[2];
    }
    bool operator==(const UnorderedTriplet &b) const { return v[0] == b.v[0] && v[1] == b.v[1] && v[2] == b.v[2]; }
};
#endif

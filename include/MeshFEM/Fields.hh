// Flat field containers (mirrors the used part of Fields.hh: VectorField is N x domainSize
// column-major, i.e. flat index N*i + c, Fields.hh:46-50).
#ifndef MESHFEM_B200_FIELDS_HH
#define MESHFEM_B200_FIELDS_HH
#include <MeshFEM/SymmetricMatrix.hh>
#include <MeshFEM/Types.hh>

enum class DomainType { PER_ELEMENT, PER_NODE, GUESS, ANY };
enum FieldType { FIELD_SCALAR, FIELD_VECTOR, FIELD_MATRIX };

template <typename _Real>
class ScalarField {
public:
    ScalarField() {}
    explicit ScalarField(size_t n) : m_v(n, 0) {}
    size_t domainSize() const { return m_v.size(); }
    size_t dim() const { return 1; }
    size_t N() const { return 1; }
    FieldType fieldType() const { return FIELD_SCALAR; }
    _Real &operator[](size_t i) { return m_v[i]; }
    _Real operator[](size_t i) const { return m_v[i]; }
    std::array<_Real, 1> operator()(size_t i) const { return {m_v[i]}; }
    std::vector<_Real> &data() { return m_v; }
    const std::vector<_Real> &data() const { return m_v; }

private:
    std::vector<_Real> m_v;
};

template <typename _Real, size_t _N>
class VectorField {
public:
    VectorField() {}
    explicit VectorField(size_t n) : m_v(n * _N, 0) {}
    size_t domainSize() const { return m_v.size() / _N; }
    size_t dim() const { return _N; }
    size_t N() const { return _N; }
    FieldType fieldType() const { return FIELD_VECTOR; }
    void resizeDomain(size_t n) { m_v.assign(n * _N, 0); }
    void clear() { std::fill(m_v.begin(), m_v.end(), 0); }
    // flat scalar access x[N*i + c]
    _Real &operator[](size_t k) { return m_v[k]; }
    _Real operator[](size_t k) const { return m_v[k]; }
    size_t size() const { return m_v.size(); }
    VectorND<_N> operator()(size_t i) const {
        VectorND<_N> r;
        for (size_t c = 0; c < _N; ++c) r[c] = m_v[_N * i + c];
        return r;
    }
    void set(size_t i, const VectorND<_N> &v) { for (size_t c = 0; c < _N; ++c) m_v[_N * i + c] = v[c]; }
    void add(size_t i, const VectorND<_N> &v) { for (size_t c = 0; c < _N; ++c) m_v[_N * i + c] += v[c]; }
    std::vector<_Real> &data() { return m_v; }
    const std::vector<_Real> &data() const { return m_v; }
    VectorField &operator*=(_Real s) { for (auto &x : m_v) x *= s; return *this; }
    _Real maxMag() const {
        _Real m = 0;
        for (size_t i = 0; i < domainSize(); ++i) m = std::max(m, (*this)(i).norm());
        return m;
    }

private:
    std::vector<_Real> m_v;
};

template <typename _Real, size_t _N>
class SymmetricMatrixField {
public:
    static constexpr size_t F = flatLen(_N);
    SymmetricMatrixField() {}
    explicit SymmetricMatrixField(size_t n) : m_v(n * F, 0) {}
    size_t domainSize() const { return m_v.size() / F; }
    size_t dim() const { return F; }
    size_t N() const { return _N; }
    FieldType fieldType() const { return FIELD_MATRIX; }
    SymmetricMatrixValue<_Real, _N> operator()(size_t i) const {
        SymmetricMatrixValue<_Real, _N> r;
        for (size_t k = 0; k < F; ++k) r[k] = m_v[F * i + k];
        return r;
    }
    std::vector<_Real> &data() { return m_v; }
    const std::vector<_Real> &data() const { return m_v; }

private:
    std::vector<_Real> m_v;
};
// Per-element field of symmetric-matrix interpolants: nodesPerElem nodal values per element, the form in
// which the reference writes full-degree strain/stress ($ElementNodeData, MSHFieldWriter.hh:262-306).
template <typename _Real, size_t _N>
class SymmetricMatrixInterpolantField {
public:
    static constexpr size_t F = flatLen(_N);
    SymmetricMatrixInterpolantField() {}
    SymmetricMatrixInterpolantField(size_t numElements, size_t nodesPerElem) : m_npe(nodesPerElem), m_v(numElements * nodesPerElem * F, 0) {}
    size_t domainSize() const { return m_npe ? m_v.size() / (m_npe * F) : 0; }
    size_t nodesPerElement() const { return m_npe; }
    size_t N() const { return _N; }
    _Real &operator()(size_t e, size_t n, size_t k) { return m_v[(e * m_npe + n) * F + k]; }
    _Real operator()(size_t e, size_t n, size_t k) const { return m_v[(e * m_npe + n) * F + k]; }
    const std::vector<_Real> &data() const { return m_v; }

private:
    size_t m_npe = 0;
    std::vector<_Real> m_v;
};
#endif

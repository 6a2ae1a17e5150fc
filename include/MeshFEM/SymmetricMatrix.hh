// Flattened symmetric N x N matrix value (mirrors the used part of SymmetricMatrix.hh).
#ifndef MESHFEM_B200_SYMMETRICMATRIX_HH
#define MESHFEM_B200_SYMMETRICMATRIX_HH
#include <MeshFEM/Flattening.hh>
#include <MeshFEM/Types.hh>

template <typename _Real, size_t t_N>
class SymmetricMatrixValue {
public:
    static constexpr size_t N = t_N;
    static constexpr size_t flatSize() { return flatLen(t_N); }
    static constexpr size_t size() { return t_N; }
    SymmetricMatrixValue() { m_data.fill(0); }
    _Real &operator[](size_t i) { return m_data[i]; }
    _Real operator[](size_t i) const { return m_data[i]; }
    _Real &operator()(size_t i, size_t j) { return m_data[flattenIndices<t_N>(i, j)]; }
    _Real operator()(size_t i, size_t j) const { return m_data[flattenIndices<t_N>(i, j)]; }
    void clear() { m_data.fill(0); }
    // e_ij = .5 (e_i e_j^T + e_j e_i^T): 1 on diagonal entries, 0.5 on shear entries (SymmetricMatrix.hh:407-413)
    static SymmetricMatrixValue CanonicalBasis(size_t i) {
        if (i >= flatSize()) throw std::runtime_error("Illegal basis element number.");
        SymmetricMatrixValue e;
        e[i] = (i < t_N) ? 1.0 : 0.5;
        return e;
    }
    SymmetricMatrixValue operator-() const { SymmetricMatrixValue r; for (size_t i = 0; i < flatSize(); ++i) r[i] = -m_data[i]; return r; }
    SymmetricMatrixValue &operator+=(const SymmetricMatrixValue &b) { for (size_t i = 0; i < flatSize(); ++i) m_data[i] += b[i]; return *this; }
    SymmetricMatrixValue &operator*=(_Real s) { for (auto &x : m_data) x *= s; return *this; }
    // single contraction with a vector (SymmetricMatrix.hh:150-160)
    VectorND<t_N> contract(const VectorND<t_N> &v) const {
        VectorND<t_N> r;
        for (size_t i = 0; i < t_N; ++i) for (size_t j = 0; j < t_N; ++j) r[i] += (*this)(i, j) * v[j];
        return r;
    }
    _Real doubleContract(const SymmetricMatrixValue &b) const {
        _Real s = 0;
        for (size_t i = 0; i < t_N; ++i) for (size_t j = 0; j < t_N; ++j) s += (*this)(i, j) * b(i, j);
        return s;
    }
    const std::array<_Real, flatLen(t_N)> &flattened() const { return m_data; }

private:
    std::array<_Real, flatLen(t_N)> m_data;
};
#endif

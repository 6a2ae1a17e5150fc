// Rank-4 elasticity tensor stored as the symmetric flattened "D" matrix (mirrors the used part
// of ElasticityTensor.hh: setIsotropic :100-134, setOrthotropic3D/2D :136-164, operator() :274-277,
// inverse :315-323, doubleContract :435-447).  Small dense inverse by Gauss-Jordan (Eigen is not
// available offline).
#ifndef MESHFEM_B200_ELASTICITYTENSOR_HH
#define MESHFEM_B200_ELASTICITYTENSOR_HH
#include <MeshFEM/SymmetricMatrix.hh>

#include <array>
#include <cmath>
#include <iomanip>
#include <ostream>
#include <vector>

namespace tensor_detail {
template <size_t F>
inline void invertInPlace(Real (&A)[F][F]) {
    Real I[F][F];
    for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) I[i][j] = (i == j);
    for (size_t c = 0; c < F; ++c) {
        size_t piv = c;
        for (size_t r = c + 1; r < F; ++r) if (std::abs(A[r][c]) > std::abs(A[piv][c])) piv = r;
        if (A[piv][c] == 0.0) throw std::runtime_error("Singular tensor matrix");
        if (piv != c) for (size_t j = 0; j < F; ++j) { std::swap(A[piv][j], A[c][j]); std::swap(I[piv][j], I[c][j]); }
        const Real d = A[c][c];
        for (size_t j = 0; j < F; ++j) { A[c][j] /= d; I[c][j] /= d; }
        for (size_t r = 0; r < F; ++r) {
            if (r == c) continue;
            const Real f = A[r][c];
            if (f == 0.0) continue;
            for (size_t j = 0; j < F; ++j) { A[r][j] -= f * A[c][j]; I[r][j] -= f * I[c][j]; }
        }
    }
    for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) A[i][j] = I[i][j];
}
}  // namespace tensor_detail

// Rank-4 tensor with the minor symmetries only, stored as the full flattened F x F matrix
// (ElasticityTensor<Real, Dim, false> of the reference): results of tensor : tensor contractions.
template <typename _Real, size_t _Dim>
struct MinorSymmetricTensor {
    static constexpr size_t F = flatLen(_Dim);
    _Real d[F][F];
    MinorSymmetricTensor() { for (auto &r : d) for (auto &x : r) x = 0; }
    _Real operator()(size_t i, size_t j, size_t k, size_t l) const { return d[flattenIndices<_Dim>(i, j)][flattenIndices<_Dim>(k, l)]; }
    _Real D(size_t i, size_t j) const { return d[i][j]; }
    _Real &D(size_t i, size_t j) { return d[i][j]; }
    // F(A : B) = F(A) S F(B), S = shear doubler (ElasticityTensor.hh:483-495)
    template <class Other>
    MinorSymmetricTensor doubleContract(const Other &B) const {
        MinorSymmetricTensor r;
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) {
            _Real s = 0;
            for (size_t k = 0; k < F; ++k) s += d[i][k] * (k >= _Dim ? 2.0 : 1.0) * B.D(k, j);
            r.d[i][j] = s;
        }
        return r;
    }
    // Mathematica array syntax (:613-633)
    void writeUnflattened(std::ostream &os) const {
        os << "{";
        for (size_t i = 0; i < _Dim; ++i) {
            os << "{";
            for (size_t j = 0; j < _Dim; ++j) {
                os << "{";
                for (size_t k = 0; k < _Dim; ++k) {
                    os << "{";
                    for (size_t l = 0; l < _Dim; ++l) { os << (*this)(i, j, k, l); if (l < _Dim - 1) os << ", "; }
                    os << ((k < _Dim - 1) ? "}, " : "}");
                }
                os << ((j < _Dim - 1) ? "}, " : "}");
            }
            os << ((i < _Dim - 1) ? "}, " : "}");
        }
        os << "}";
    }
};

template <typename _Real, size_t _Dim>
class ElasticityTensor {
public:
    static constexpr size_t Dim = _Dim;
    static constexpr size_t F = flatLen(_Dim);
    typedef SymmetricMatrixValue<_Real, _Dim> SMatrix;

    ElasticityTensor() { clear(); }
    ElasticityTensor(_Real E, _Real nu) { setIsotropic(E, nu); }
    void clear() { for (auto &r : m_d) for (auto &x : r) x = 0; }
    void setIdentity() { setIsotropicLame(0, 0.5); }

    void setIsotropic(_Real E, _Real nu) {
        _Real lambda = (nu * E) / ((1.0 + nu) * (1.0 - 2.0 * nu));
        const _Real mu = E / (2.0 + 2.0 * nu);
        if (_Dim == 2) lambda = (nu * E) / (1.0 - nu * nu);   // plane stress (:111-112)
        setIsotropicLame(lambda, mu);
    }
    void setIsotropicLame(_Real lambda, _Real mu) {
        clear();
        for (size_t i = 0; i < _Dim; ++i) for (size_t j = 0; j < _Dim; ++j) m_d[i][j] = lambda + (i == j ? 2 * mu : 0.0);
        for (size_t i = _Dim; i < F; ++i) m_d[i][i] = mu;
    }
    void setOrthotropic3D(_Real Ex, _Real Ey, _Real Ez, _Real nuYX, _Real nuZX, _Real nuZY, _Real muYZ, _Real muZX,
                          _Real muXY) {
        if (_Dim != 3) throw std::runtime_error("setOrthotropic3D call on non-3D tensor");
        clear();
        m_d[0][0] = 1.0 / Ex; m_d[0][1] = -nuYX / Ey; m_d[0][2] = -nuZX / Ez;
        m_d[1][1] = 1.0 / Ey; m_d[1][2] = -nuZY / Ez;
        m_d[2][2] = 1.0 / Ez;
        m_d[3][3] = 1.0 / muYZ; m_d[4][4] = 1.0 / muZX; m_d[5][5] = 1.0 / muXY;
        m_symmetrizeFromUpper();
        tensor_detail::invertInPlace<F>(m_d);
    }
    void setOrthotropic2D(_Real Ex, _Real Ey, _Real nuYX, _Real muXY) {
        if (_Dim != 2) throw std::runtime_error("setOrthotropic2D call on non-2D tensor");
        clear();
        m_d[0][0] = 1.0 / Ex; m_d[0][1] = -nuYX / Ey;
        m_d[1][1] = 1.0 / Ey;
        m_d[2][2] = 1.0 / muXY;
        m_symmetrizeFromUpper();
        tensor_detail::invertInPlace<F>(m_d);
    }

    _Real operator()(size_t i, size_t j, size_t k, size_t l) const { return D(flattenIndices<_Dim>(i, j), flattenIndices<_Dim>(k, l)); }
    _Real D(size_t i, size_t j) const { return (i <= j) ? m_d[i][j] : m_d[j][i]; }
    _Real &D(size_t i, size_t j) { return (i <= j) ? m_d[i][j] : m_d[j][i]; }
    // full symmetric flattened matrix, row-major (what mfem_b200_set_material_* takes)
    void getFlat(_Real *out) const { for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) out[i * F + j] = D(i, j); }
    void setFlat(const _Real *in) { for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) m_d[i][j] = in[i * F + j]; m_symmetrizeFromUpper(); }

    // D * shearDoubled(strain)  (:435-447)
    SMatrix doubleContract(const SMatrix &in) const {
        SMatrix out;
        for (size_t i = 0; i < F; ++i) {
            _Real s = 0;
            for (size_t j = 0; j < F; ++j) s += D(i, j) * (j >= _Dim ? 2.0 : 1.0) * in[j];
            out[i] = s;
        }
        return out;
    }
    // A : B for two rank-4 tensors; the result has no major symmetry in general (:483-495)
    template <class Other>
    MinorSymmetricTensor<_Real, _Dim> doubleContractTensor(const Other &B) const {
        MinorSymmetricTensor<_Real, _Dim> r;
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) {
            _Real s = 0;
            for (size_t k = 0; k < F; ++k) s += D(i, k) * (k >= _Dim ? 2.0 : 1.0) * B.D(k, j);
            r.d[i][j] = s;
        }
        return r;
    }
    // E^-1 with E : E^-1 = identity: invert D, then halve shear rows and columns (:315-323)
    ElasticityTensor inverse() const {
        ElasticityTensor r;
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) r.m_d[i][j] = D(i, j);
        tensor_detail::invertInPlace<F>(r.m_d);
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) {
            if (i >= _Dim) r.m_d[i][j] *= 0.5;
            if (j >= _Dim) r.m_d[i][j] *= 0.5;
        }
        return r;
    }
    ElasticityTensor &operator+=(const ElasticityTensor &b) { for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) m_d[i][j] += b.m_d[i][j]; return *this; }
    ElasticityTensor &operator*=(_Real s) { for (auto &r : m_d) for (auto &x : r) x *= s; return *this; }
    ElasticityTensor &operator/=(_Real s) { return (*this) *= (1.0 / s); }
    friend ElasticityTensor operator*(ElasticityTensor a, _Real s) { return a *= s; }
    // row i of D viewed as a flattened symmetric matrix (DRowAsSymMatrix)
    void addToRow(size_t i, const SMatrix &m) { for (size_t j = 0; j < F; ++j) m_d[i][j] += m[j]; }
    void symmetrizeFromFull() { for (size_t i = 0; i < F; ++i) for (size_t j = i + 1; j < F; ++j) m_d[j][i] = m_d[i][j]; }
    // Orthotropic parameters of the tensor, read off its inverse (ElasticityTensor.hh:193-225 of the reference;
    // shear entries of the flattened compliance are 1/(4 mu)).
    void getOrthotropic3D(_Real &Ex, _Real &Ey, _Real &Ez, _Real &nuYX, _Real &nuZX, _Real &nuZY, _Real &muYZ, _Real &muZX,
                          _Real &muXY) const {
        if (_Dim != 3) throw std::runtime_error("getOrthotropic3D call on non-3D tensor");
        const ElasticityTensor S = inverse();
        Ex = 1.0 / S.D(0, 0), Ey = 1.0 / S.D(1, 1), Ez = 1.0 / S.D(2, 2);
        nuYX = -S.D(0, 1) * Ey, nuZX = -S.D(0, 2) * Ez, nuZY = -S.D(1, 2) * Ez;
        muYZ = 0.25 / S.D(3, 3), muZX = 0.25 / S.D(4, 4), muXY = 0.25 / S.D(5, 5);
    }
    void getOrthotropic2D(_Real &Ex, _Real &Ey, _Real &nuYX, _Real &muXY) const {
        if (_Dim != 2) throw std::runtime_error("getOrthotropic2D call on non-2D tensor");
        const ElasticityTensor S = inverse();
        Ex = 1.0 / S.D(0, 0), Ey = 1.0 / S.D(1, 1);
        nuYX = -S.D(0, 1) * Ey;
        muXY = 0.25 / S.D(2, 2);
    }
    // mu_avg / mu_iso(E_avg, nu_avg)  (:251-268)
    _Real anisotropy() const {
        _Real muAvg, EAvg, nuAvg;
        if (_Dim == 2) {
            _Real Ex, Ey, nuYX, muXY;
            getOrthotropic2D(Ex, Ey, nuYX, muXY);
            EAvg = (Ex + Ey) / 2.0, nuAvg = nuYX, muAvg = muXY;
        } else {
            _Real Ex, Ey, Ez, nuYX, nuZX, nuZY, muYZ, muZX, muXY;
            getOrthotropic3D(Ex, Ey, Ez, nuYX, nuZX, nuZY, muYZ, muZX, muXY);
            EAvg = (Ex + Ey + Ez) / 3.0, nuAvg = (nuYX + nuZX + nuZY) / 3.0, muAvg = (muYZ + muZX + muXY) / 3.0;
        }
        return muAvg / (EAvg / (2 * (1 + nuAvg)));
    }
    // Eigenstrains E : s = lambda s (:555-579): ordinary symmetric eigenproblem of
    // D^(1/2) F(E) D^(1/2) (D = shear doubler), eigenvalues ascending, strains[k] = D^(-1/2) q_k.
    // Cyclic Jacobi on the F x F matrix (Eigen's SelfAdjointEigenSolver in the reference).
    struct EigenDecomposition {
        std::array<_Real, F> lambdas;
        std::array<std::array<_Real, F>, F> strains;   // strains[k][component]
    };
    EigenDecomposition computeEigenstrains() const {
        _Real A[F][F], Q[F][F];
        const _Real rt2 = std::sqrt(2.0);
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) {
            A[i][j] = D(i, j) * (i >= _Dim ? rt2 : 1.0) * (j >= _Dim ? rt2 : 1.0);
            Q[i][j] = (i == j);
        }
        for (int sweep = 0; sweep < 64; ++sweep) {
            _Real off = 0, diag = 0;
            for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) (i == j ? diag : off) += A[i][j] * A[i][j];
            if (off <= 1e-32 * diag) break;
            for (size_t p = 0; p < F; ++p) for (size_t q = p + 1; q < F; ++q) {
                if (A[p][q] == 0.0) continue;
                const _Real theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
                const _Real t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                const _Real c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (size_t k = 0; k < F; ++k) { const _Real akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
                for (size_t k = 0; k < F; ++k) { const _Real apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
                for (size_t k = 0; k < F; ++k) { const _Real qkp = Q[k][p], qkq = Q[k][q]; Q[k][p] = c * qkp - s * qkq; Q[k][q] = s * qkp + c * qkq; }
            }
        }
        std::array<size_t, F> order;
        for (size_t i = 0; i < F; ++i) order[i] = i;
        for (size_t i = 0; i < F; ++i) for (size_t j = i + 1; j < F; ++j) if (A[order[j]][order[j]] < A[order[i]][order[i]]) std::swap(order[i], order[j]);
        EigenDecomposition r;
        for (size_t k = 0; k < F; ++k) {
            r.lambdas[k] = A[order[k]][order[k]];
            for (size_t i = 0; i < F; ++i) r.strains[k][i] = Q[i][order[k]] / (i >= _Dim ? rt2 : 1.0);
        }
        return r;
    }
    // sum_ijkl a_ijkl b_ijkl (:498-506): every flattened shear index stands for two index pairs
    _Real quadrupleContract(const ElasticityTensor &b) const {
        _Real s = 0;
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) s += (i >= _Dim ? 2.0 : 1.0) * (j >= _Dim ? 2.0 : 1.0) * D(i, j) * b.D(i, j);
        return s;
    }
    _Real frobeniusNormSq() const { return quadrupleContract(*this); }        // :508
    // Change of coordinates E'_ijkl = E_pqrs R_ip R_jq R_kr R_ls (:515-541); R row-major _Dim x _Dim, any invertible
    // matrix (DeformedCells_cli uses the deformation jacobian and its inverse).
    ElasticityTensor transform(const _Real (&R)[_Dim][_Dim]) const {
        ElasticityTensor result;
        for (size_t i = 0; i < _Dim; ++i) for (size_t j = i; j < _Dim; ++j)
            for (size_t k = 0; k < _Dim; ++k) for (size_t l = k; l < _Dim; ++l) {
                const size_t ij = flattenIndices<_Dim>(i, j), kl = flattenIndices<_Dim>(k, l);
                if (ij > kl) continue;
                _Real comp = 0;
                for (size_t p = 0; p < _Dim; ++p) for (size_t q = 0; q < _Dim; ++q)
                    for (size_t r = 0; r < _Dim; ++r) for (size_t t = 0; t < _Dim; ++t)
                        comp += (*this)(p, q, r, t) * R[i][p] * R[j][q] * R[k][r] * R[l][t];
                result.m_d[ij][kl] = comp;
            }
        result.m_symmetrizeFromUpper();
        return result;
    }
    // row-major F x F coefficients (:636-654) and the orthotropic parameter list
    // (2D: Ex Ey nuYX muXY -- the reference's own comment says "nuXY" for the 4th, the value is muXY; 3D: Ex Ey Ez nuYX nuZX nuZY muYZ muZX muXY; :168-184)
    std::vector<_Real> getCoefficients() const {
        std::vector<_Real> c;
        for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) c.push_back(D(i, j));
        return c;
    }
    std::vector<_Real> getOrthotropicParameters() const {
        std::vector<_Real> m(_Dim == 2 ? 4 : 9);
        if (_Dim == 2) getOrthotropic2D(m[0], m[1], m[2], m[3]);
        else getOrthotropic3D(m[0], m[1], m[2], m[3], m[4], m[5], m[6], m[7], m[8]);
        return m;
    }
    void printOrthotropic(std::ostream &os) const {                           // :236-249
        const auto m = getOrthotropicParameters();
        for (size_t i = 0; i < m.size(); ++i) os << (i ? "\t" : "") << m[i];
        os << std::endl;
    }

    // Isotropic-equivalent moduli read off a compliance-like inverse (PeriodicHomogenization_cli.cc:126-171)
    friend std::ostream &operator<<(std::ostream &os, const ElasticityTensor &E) {
        for (size_t i = 0; i < F; ++i) {
            for (size_t j = 0; j < F; ++j) os << (j ? "\t" : "") << E.D(i, j);
            os << std::endl;
        }
        return os;
    }

private:
    _Real m_d[F][F];
    void m_symmetrizeFromUpper() { for (size_t i = 0; i < F; ++i) for (size_t j = i + 1; j < F; ++j) m_d[j][i] = m_d[i][j]; }
};
#endif

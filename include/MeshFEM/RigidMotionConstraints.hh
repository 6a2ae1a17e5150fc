// Lagrange-multiplier rows of the elasticity system and how this path solves them.
//
// The reference poses  [K C^T; C 0] [u; l] = [f; d]  whenever assembleConstrainedSystem emits
// constraint rows (LinearElasticity.hh:1201-1249: the no-rigid-motion rows of
// m_appendInfinitesimalRotationMatrix :1530-1568 and m_appendTranslationMatrix :1571-1593) and hands
// the indefinite matrix to UMFPACK (SparseMatrices.hh:2332-2348, 2580-2595).  The device solver here is
// a PCG, i.e. SPD/SPSD only, so the saddle point is resolved on the host around it:
//
//   every row of C is a rigid-mode functional and the rows exist precisely because K_ff (K with the
//   fixed variables removed) is singular with a null space Z spanned by the rigid modes the fixed
//   variables leave free.  With  Z^T K_ff = 0  and  W = C_f Z  (m x m, invertible iff the reference's
//   saddle-point matrix is):
//       W^T l          = Z^T f                      (multipliers: the unbalanced part of the load)
//       K_ff u_p       = f - C^T l - K_fc u_c       (consistent SPSD system -> PCG from x0 = 0)
//       W a            = d - C u_p ,  u = u_p + Z a (fix the rigid part so that C u = d)
//   The fixed variables' columns of C go to the right-hand side exactly as fixVariables does for the
//   reference's augmented matrix (SparseMatrices.hh:2389-2500).
//
// Z is found, not assumed: among the candidate rigid modes (translations; infinitesimal rotations when
// nodes and DoFs coincide) the combinations that vanish on every fixed variable.  More rows than free
// modes (down to none: Dirichlet conditions that already make K_ff definite, with no_rigid_motion on top) is
// the same saddle point with a Schur complement for the surplus rows -- m - k extra SPSD solves, see solve();
// fewer rows than free modes is singular for the reference too, and we throw.
#ifndef MESHFEM_B200_RIGIDMOTIONCONSTRAINTS_HH
#define MESHFEM_B200_RIGIDMOTIONCONSTRAINTS_HH
#include <algorithm>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

namespace RigidMotionConstraints {
typedef double Real;
typedef std::vector<Real> Vec;

struct Rows {                       // dense m x n (m <= 6); rhs[m]
    std::vector<Vec> rows;
    Vec rhs;
    size_t m() const { return rows.size(); }
    void clear() { rows.clear(); rhs.clear(); }
};

// cyclic Jacobi eigen-decomposition of a small symmetric matrix (row-major n x n, destroyed);
// eigenvectors in the COLUMNS of V
inline void symmetricEigen(size_t n, std::vector<Real> &A, Vec &evals, std::vector<Real> &V) {
    V.assign(n * n, 0.0);
    for (size_t i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 64; ++sweep) {
        Real off = 0.0, diag = 0.0;
        for (size_t i = 0; i < n; ++i)
            for (size_t j = 0; j < n; ++j) (i == j ? diag : off) += A[i * n + j] * A[i * n + j];
        if (off <= 1e-30 * std::max(diag, Real(1e-300))) break;
        for (size_t p = 0; p + 1 < n; ++p)
            for (size_t q = p + 1; q < n; ++q) {
                const Real apq = A[p * n + q];
                if (apq == 0.0) continue;
                const Real theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const Real t = (theta >= 0 ? 1.0 : -1.0) / (std::abs(theta) + std::sqrt(theta * theta + 1.0));
                const Real c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (size_t k = 0; k < n; ++k) {           // A <- A J
                    const Real akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - s * akq;
                    A[k * n + q] = s * akp + c * akq;
                }
                for (size_t k = 0; k < n; ++k) {           // A <- J^T A
                    const Real apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - s * aqk;
                    A[q * n + k] = s * apk + c * aqk;
                }
                for (size_t k = 0; k < n; ++k) {
                    const Real vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - s * vkq;
                    V[k * n + q] = s * vkp + c * vkq;
                }
            }
    }
    evals.resize(n);
    for (size_t i = 0; i < n; ++i) evals[i] = A[i * n + i];
}

// solve the small dense system M x = b (row-major m x m) by LU with partial pivoting
inline Vec denseSolve(size_t m, std::vector<Real> M, Vec b, const char *what) {
    Real scale = 0.0;
    for (Real v : M) scale = std::max(scale, std::abs(v));
    for (size_t k = 0; k < m; ++k) {
        size_t piv = k;
        for (size_t i = k + 1; i < m; ++i) if (std::abs(M[i * m + k]) > std::abs(M[piv * m + k])) piv = i;
        if (std::abs(M[piv * m + k]) <= 1e-12 * scale || scale == 0.0)
            throw std::runtime_error(std::string("Singular constraint system: ") + what);
        if (piv != k) { for (size_t j = 0; j < m; ++j) std::swap(M[k * m + j], M[piv * m + j]); std::swap(b[k], b[piv]); }
        for (size_t i = k + 1; i < m; ++i) {
            const Real l = M[i * m + k] / M[k * m + k];
            if (l == 0.0) continue;
            for (size_t j = k; j < m; ++j) M[i * m + j] -= l * M[k * m + j];
            b[i] -= l * b[k];
        }
    }
    Vec x(m);
    for (size_t kk = m; kk-- > 0;) {
        Real s = b[kk];
        for (size_t j = kk + 1; j < m; ++j) s -= M[kk * m + j] * x[j];
        x[kk] = s / M[kk * m + kk];
    }
    return x;
}

inline Real dot(const Vec &a, const Vec &b) {
    long double s = 0.0L;
    for (size_t i = 0; i < a.size(); ++i) s += (long double)a[i] * b[i];
    return (Real)s;
}

// The rigid modes (columns, each of length n) that vanish on every fixed variable: combinations of the
// candidates `modes`.  Candidates are scaled to unit max-norm first so that the rank decision does not
// depend on the units of the mesh.
inline std::vector<Vec> freeRigidModes(size_t n, std::vector<Vec> modes, const std::vector<size_t> &fixedVars) {
    std::vector<Vec> kept;
    for (auto &b : modes) {
        if (b.size() != n) throw std::runtime_error("freeRigidModes: bad mode size");
        Real mx = 0.0;
        for (Real v : b) mx = std::max(mx, std::abs(v));
        if (mx == 0.0) continue;                           // e.g. a rotation of a mesh collapsed onto its axis
        for (Real &v : b) v /= mx;
        kept.push_back(std::move(b));
    }
    modes.swap(kept);
    const size_t k = modes.size();
    if (k == 0) return {};
    std::vector<Real> G(k * k, 0.0);
    for (size_t v : fixedVars) {
        if (v >= n) throw std::runtime_error("freeRigidModes: fixed variable out of range");
        for (size_t i = 0; i < k; ++i)
            for (size_t j = 0; j < k; ++j) G[i * k + j] += modes[i][v] * modes[j][v];
    }
    Real trace = 0.0;
    for (size_t i = 0; i < k; ++i) trace += G[i * k + i];
    Vec evals;
    std::vector<Real> V;
    symmetricEigen(k, G, evals, V);
    std::vector<Vec> Z;
    for (size_t j = 0; j < k; ++j) {
        if (std::abs(evals[j]) > 1e-10 * std::max(trace, Real(1.0))) continue;
        Vec z(n, 0.0);
        for (size_t i = 0; i < k; ++i) {
            const Real w = V[i * k + j];
            if (w == 0.0) continue;
            for (size_t q = 0; q < n; ++q) z[q] += w * modes[i][q];
        }
        for (size_t v : fixedVars) z[v] = 0.0;
        Z.push_back(std::move(z));
    }
    return Z;
}

// Solve the saddle-point systems for the right-hand sides fs (see the header comment).
//   solveSPSD(rhs) -> u: solves K_ff u_f = rhs_f - K_fc u_c for every right-hand side and returns the full
//   vectors with the fixed values in place (the device PCG; K_ff may be singular, the systems are consistent).
//
// General form (k = free rigid modes, m = rows, k <= m): with W = C_f Z (m x k, full column rank) the multipliers are
//   l = l0 + Nn mu,   l0 = W (W^T W)^-1 Z^T f  (the part the null space of K_ff dictates),  Nn = null(W^T)  (m x (m-k)),
// and with  K_ff u_p = f - C^T l0 - K_fc u_c  and  K_ff y_j = C_f^T Nn_j  (m-k more consistent solves, shared by all
// right-hand sides)
//   u = u_p - sum_j mu_j y_j + Z a,     [ -C y_1 .. -C y_(m-k) | W ] (mu; a) = d - C u_p .
// k = m (the rows remove exactly the null space -- every configuration the no-rigid-motion rows were made for): Nn is
// empty and this is the three-line recipe of the header.  k = 0 (Dirichlet conditions already make K_ff definite, rows
// on top: the reference solves that saddle point too): the classical Schur complement S = C K_ff^-1 C^T.
template <class SolveFn>
std::vector<Vec> solve(size_t n, const Rows &C, const std::vector<size_t> &fixedVars, const std::vector<Vec> &candidateModes,
                       const std::vector<Vec> &fs, SolveFn &&solveSPSD, std::vector<Vec> *multipliers = nullptr) {
    const size_t m = C.m();
    if (C.rhs.size() != m) throw std::runtime_error("Bad constraint rows");
    for (const auto &r : C.rows) if (r.size() != n) throw std::runtime_error("Bad constraint rows");
    for (const auto &f : fs) if (f.size() != n) throw std::runtime_error("Bad RHS");
    std::vector<Vec> Z = freeRigidModes(n, candidateModes, fixedVars);
    const size_t k = Z.size();
    if (k > m)
        throw std::runtime_error("Unsupported constrained system: " + std::to_string(m) + " Lagrange-multiplier row(s) but the fixed variables leave " +
                                 std::to_string(k) + " rigid mode(s) free (the rows must remove the null space of the stiffness matrix)");
    std::vector<uint8_t> isFixed(n, 0);
    for (size_t v : fixedVars) isFixed[v] = 1;
    // W = C_f Z  (m x k)
    std::vector<Real> W(m * k);
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < k; ++j) {
            long double s = 0.0L;
            for (size_t q = 0; q < n; ++q) if (!isFixed[q]) s += (long double)C.rows[i][q] * Z[j][q];
            W[i * k + j] = (Real)s;
        }
    // Nn = null(W^T): eigenvectors of W W^T (m x m) with vanishing eigenvalue; exactly m - k of them iff W has full column rank
    const size_t e = m - k;
    std::vector<Vec> Nn;
    if (e > 0) {
        std::vector<Real> G(m * m, 0.0), V;
        Real trace = 0.0;
        for (size_t i = 0; i < m; ++i)
            for (size_t j = 0; j < m; ++j) {
                Real s = 0.0;
                for (size_t t = 0; t < k; ++t) s += W[i * k + t] * W[j * k + t];
                G[i * m + j] = s;
                if (i == j) trace += s;
            }
        Vec evals;
        symmetricEigen(m, G, evals, V);
        for (size_t j = 0; j < m; ++j) {
            if (std::abs(evals[j]) > 1e-10 * std::max(trace, Real(1e-300)) && k > 0) continue;
            Vec v(m);
            for (size_t i = 0; i < m; ++i) v[i] = V[i * m + j];
            Nn.push_back(std::move(v));
        }
        if (Nn.size() != e) throw std::runtime_error("Singular constraint system: constraint rows do not control the free rigid modes");
    }
    // W^T W (k x k) for the least-norm multipliers
    std::vector<Real> WtW(k * k, 0.0);
    for (size_t a = 0; a < k; ++a)
        for (size_t b = 0; b < k; ++b)
            for (size_t i = 0; i < m; ++i) WtW[a * k + b] += W[i * k + a] * W[i * k + b];
    std::vector<Vec> lambdas, rhs;
    for (const auto &f : fs) {
        Vec l(m, 0.0);
        if (k > 0) {
            Vec ztf(k);
            for (size_t j = 0; j < k; ++j) ztf[j] = dot(Z[j], f);      // Z vanishes on the fixed variables
            if (k == m) {                                               // square W: W^T l = Z^T f directly
                std::vector<Real> Wt(m * m);
                for (size_t i = 0; i < m; ++i)
                    for (size_t j = 0; j < m; ++j) Wt[j * m + i] = W[i * m + j];
                l = denseSolve(m, Wt, ztf, "constraint rows do not control the free rigid modes");
            } else {
                const Vec c = denseSolve(k, WtW, ztf, "constraint rows do not control the free rigid modes");
                for (size_t i = 0; i < m; ++i)
                    for (size_t j = 0; j < k; ++j) l[i] += W[i * k + j] * c[j];
            }
        }
        Vec b = f;
        for (size_t i = 0; i < m; ++i)
            if (l[i] != 0.0) for (size_t q = 0; q < n; ++q) b[q] -= l[i] * C.rows[i][q];
        lambdas.push_back(std::move(l));
        rhs.push_back(std::move(b));
    }
    // the extra systems K_ff y_j = C_f^T Nn_j ride on the first right-hand side: y_j = u(rhs_0 + C^T Nn_j) - u(rhs_0)
    // (the solver keeps the fixed values in place; they cancel in the difference)
    const size_t nf = fs.size();
    if (e > 0 && nf == 0) return {};
    for (size_t j = 0; j < e; ++j) {
        Vec b = rhs[0];
        for (size_t i = 0; i < m; ++i)
            if (Nn[j][i] != 0.0) for (size_t q = 0; q < n; ++q) b[q] += Nn[j][i] * C.rows[i][q];
        rhs.push_back(std::move(b));
    }
    std::vector<Vec> us = solveSPSD(rhs);
    if (us.size() != nf + e) throw std::runtime_error("constrained solve: solver returned the wrong number of solutions");
    for (auto &u : us)
        if (u.size() != n) throw std::runtime_error("constrained solve: solver returned a vector of the wrong size");
    std::vector<Vec> Y(e);
    for (size_t j = 0; j < e; ++j) {
        Y[j] = us[nf + j];
        for (size_t q = 0; q < n; ++q) Y[j][q] -= us[0][q];
    }
    us.resize(nf);
    // M = [ -C y_1 .. -C y_e | W ]  (m x m)
    std::vector<Real> M(m * m, 0.0);
    for (size_t i = 0; i < m; ++i) {
        for (size_t j = 0; j < e; ++j) M[i * m + j] = -dot(C.rows[i], Y[j]);
        for (size_t j = 0; j < k; ++j) M[i * m + e + j] = W[i * k + j];
    }
    for (size_t r = 0; r < nf; ++r) {
        Vec &u = us[r];
        Vec d(m);
        for (size_t i = 0; i < m; ++i) d[i] = C.rhs[i] - dot(C.rows[i], u);     // fixed columns included: C u = d
        const Vec x = denseSolve(m, M, d, "constraint rows do not control the free rigid modes");
        for (size_t j = 0; j < e; ++j) {
            if (x[j] == 0.0) continue;
            for (size_t q = 0; q < n; ++q) u[q] -= x[j] * Y[j][q];
            for (size_t i = 0; i < m; ++i) lambdas[r][i] += x[j] * Nn[j][i];
        }
        for (size_t j = 0; j < k; ++j)
            if (x[e + j] != 0.0) for (size_t q = 0; q < n; ++q) u[q] += x[e + j] * Z[j][q];
    }
    if (multipliers) *multipliers = lambdas;
    return us;
}
}  // namespace RigidMotionConstraints
#endif

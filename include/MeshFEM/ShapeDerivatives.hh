// Discrete shape derivatives of the elasticity operators under vertex perturbations delta_p (mirrors the
// functionality of LinearElasticity.hh:232-331, 1286-1373 and PeriodicHomogenization.hh:383-563 of the reference).
// The straight-sided elements follow their vertices and nodal values are transported (Lagrangian derivative).
// With the piecewise-linear velocity dp_h = sum_k delta_p_k lambda_k the geometric rules are
//     delta grad phi_i = -(grad dp_h)^T grad phi_i ,      delta vol = vol div dp_h
// (EmbeddedElement.hh:269-278, 338-372), so that instead of the reference's per-element 30x30 delta-stiffness
// matrix each element contributes through the displacement gradient at the quadrature points:
//   (delta K u)_i = int [ div dp sigma(u) + C : delta eps(u) ] grad phi_i + sigma(u) delta grad phi_i ,
//   delta eps(u) = -sym(grad u grad dp_h).
// All integrands are polynomials of degree 2 (Deg - 1) and are integrated exactly by the reference's rules
// (GaussQuadrature.hh:115-127, 283-295).  Host loops over the elements: these are the geometry-sensitivity
// post-processing steps of an optimisation loop, O(elements); the solves they feed stay on the device.
#ifndef MESHFEM_B200_SHAPEDERIVATIVES_HH
#define MESHFEM_B200_SHAPEDERIVATIVES_HH
#include <MeshFEM/ElasticityTensor.hh>
#include <MeshFEM/Fields.hh>
#include <MeshFEM/Simplex.hh>

#include <array>
#include <vector>

namespace ShapeDerivatives {

// quadrature of degree 2 (Deg - 1) on the K-simplex: barycentric points and weights summing to one
template <size_t K, size_t Deg>
struct ElementQuadrature {
    static constexpr size_t numPoints = (Deg == 1) ? 1 : K + 1;
    Real lambda[numPoints][K + 1];
    Real weight[numPoints];
    ElementQuadrature() {
        if (Deg == 1) {
            for (size_t v = 0; v <= K; ++v) lambda[0][v] = 1.0 / (K + 1);
            weight[0] = 1.0;
        } else {
            const Real c0 = (K == 3) ? 0.58541019662496845446 : 2.0 / 3.0, c1 = (K == 3) ? 0.13819660112501051518 : 1.0 / 6.0;
            for (size_t q = 0; q < numPoints; ++q) {
                for (size_t v = 0; v <= K; ++v) lambda[q][v] = (v == q) ? c0 : c1;
                weight[q] = 1.0 / (K + 1);
            }
        }
    }
};

// grad phi_i at the barycentric point lam (EmbeddedElement.hh:315-332): Deg 1 grad lambda_i; Deg 2 vertex functions
// (4 lam_i - 1) grad lambda_i and edge functions 4 (lam_e grad lambda_s + lam_s grad lambda_e)
template <size_t K, size_t Deg>
inline void gradPhis(const Real (&g)[K + 1][K], const Real *lam, Real (&gphi)[Simplex::numNodes(K, Deg)][K]) {
    constexpr size_t npe = Simplex::numNodes(K, Deg);
    for (size_t i = 0; i < npe; ++i) {
        if (Deg == 1) { for (size_t r = 0; r < K; ++r) gphi[i][r] = g[i][r]; }
        else if (i <= K) { for (size_t r = 0; r < K; ++r) gphi[i][r] = (4.0 * lam[i] - 1.0) * g[i][r]; }
        else {
            const size_t s = Simplex::edgeStartNode(i - (K + 1)), e = Simplex::edgeEndNode(i - (K + 1));
            for (size_t r = 0; r < K; ++r) gphi[i][r] = 4.0 * (lam[e] * g[s][r] + lam[s] * g[e][r]);
        }
    }
}

// grad dp_h of one element: G[a][r] = sum_k delta_p_k[a] grad lambda_k[r]
template <class Mesh, class VField>
inline void velocityGradient(const Mesh &mesh, size_t e, const Real (&g)[Mesh::K + 1][Mesh::K], const VField &deltaP, Real (&G)[Mesh::K][Mesh::K]) {
    constexpr size_t K = Mesh::K;
    for (size_t a = 0; a < K; ++a) for (size_t r = 0; r < K; ++r) G[a][r] = 0.0;
    for (size_t k = 0; k <= K; ++k) {
        const auto dp = deltaP(mesh.elementVertex(e, k));
        for (size_t a = 0; a < K; ++a) for (size_t r = 0; r < K; ++r) G[a][r] += dp[a] * g[k][r];
    }
}

template <size_t N>
inline SymmetricMatrixValue<Real, N> symmetrized(const Real (&A)[N][N]) {
    SymmetricMatrixValue<Real, N> s;
    for (size_t c = 0; c < N; ++c) for (size_t r = c; r < N; ++r) s(c, r) = 0.5 * (A[c][r] + A[r][c]);
    return s;
}

// One-form over the vertex positions with tensor values: (v)[c] is the derivative with respect to component c of
// vertex v (OneForm<ETensor, N> of the reference)
template <class ETensor, size_t N>
struct OneForm {
    std::vector<std::array<ETensor, N>> data;
    explicit OneForm(size_t numVertices = 0) : data(numVertices) {}
    std::array<ETensor, N> &operator()(size_t v) { return data[v]; }
    const std::array<ETensor, N> &operator()(size_t v) const { return data[v]; }
    size_t domainSize() const { return data.size(); }
    // apply to a per-vertex perturbation field
    template <class VField>
    ETensor operator[](const VField &deltaP) const {
        if (deltaP.domainSize() != data.size()) throw std::runtime_error("OneForm: per-vertex field expected");
        ETensor r;
        for (size_t v = 0; v < data.size(); ++v) {
            const auto dp = deltaP(v);
            for (size_t c = 0; c < N; ++c) { if (dp[c] == 0.0) continue; ETensor t = data[v][c]; t *= dp[c]; r += t; }
        }
        return r;
    }
};

}  // namespace ShapeDerivatives
#endif

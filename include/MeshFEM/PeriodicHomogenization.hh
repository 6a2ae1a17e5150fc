// Periodic homogenization (mirrors PeriodicHomogenization.hh:34-186 of the reference): the
// flatLen(N) cell problems share one assembled K (one GPU assembly, then flatLen(N) PCG solves
// against constantStrainLoad(-e_ij)), followed by the homogenized tensor in displacement form
// (boundary integral) or stress form (element averages).
#ifndef MESHFEM_B200_PERIODICHOMOGENIZATION_HH
#define MESHFEM_B200_PERIODICHOMOGENIZATION_HH
#include <MeshFEM/LinearElasticity.hh>

namespace PeriodicHomogenization {

template <class _Sim>
void solveCellProblems(std::vector<typename _Sim::VField> &w_ij, _Sim &sim, Real cellEpsilon = 1e-7,
                       bool ignorePeriodicMismatch = false, std::unique_ptr<PeriodicCondition<_Sim::N>> pc = nullptr) {
    typedef typename _Sim::VField VField;
    typedef typename _Sim::SMatrix SMatrix;
    constexpr size_t numStrains = SMatrix::flatSize();
    sim.applyPeriodicConditions(cellEpsilon, ignorePeriodicMismatch, std::move(pc));
    sim.applyNoRigidMotionConstraint();
    sim.setUsePinNoRigidTranslationConstraint(true);
    w_ij.reserve(numStrains), w_ij.clear();
    // the reference back-solves once per strain with one factorisation; here the numStrains loads go to
    // the device together and are solved by one batched PCG (one matrix stream for all of them)
    std::vector<VField> loads;
    for (size_t i = 0; i < numStrains; ++i) {
        BENCHMARK_START_TIMER("Constant Strain Load");
        loads.push_back(sim.constantStrainLoad(-SMatrix::CanonicalBasis(i)));
        BENCHMARK_STOP_TIMER("Constant Strain Load");
    }
    w_ij = sim.solve(loads);
}

template <class _Sim>
std::vector<typename _Sim::VField> solveCellProblems(_Sim &sim, Real cellEpsilon = 1e-7) {
    std::vector<typename _Sim::VField> w_ij;
    solveCellProblems(w_ij, sim, cellEpsilon);
    return w_ij;
}

// Eh_ijkl = 1/|Y| int_omega [E : strain(w_ij)]_kl + E_ijkl dV  (:72-100)
template <class _Sim>
typename _Sim::ETensor homogenizedElasticityTensor(const std::vector<typename _Sim::VField> &w_ij, const _Sim &sim,
                                                   Real baseCellVolume = 0.0) {
    const auto &mesh = sim.mesh();
    if (baseCellVolume == 0.0) baseCellVolume = mesh.boundingBox().volume();
    typedef typename _Sim::SMatrix SMatrix;
    typename _Sim::ETensor Eh;
    std::vector<typename _Sim::SMField> strains;
    for (const auto &w : w_ij) strains.push_back(sim.averageStrainField(w));
    for (size_t e = 0; e < mesh.numElements(); ++e) {
        typename _Sim::ETensor Econtrib;
        const auto &E = sim.elementTensor(e);
        for (size_t i = 0; i < w_ij.size(); ++i) Econtrib.addToRow(i, E.doubleContract(strains[i](e)));
        Econtrib += E;
        Econtrib *= mesh.elementVolume(e);
        Eh += Econtrib;
    }
    Eh /= baseCellVolume;
    (void)sizeof(SMatrix);
    return Eh;
}

// Displacement (boundary-integral) form, constant base tensor (:146-186)
template <class _Sim>
typename _Sim::ETensor homogenizedElasticityTensorDisplacementForm(const std::vector<typename _Sim::VField> &w_ij, const _Sim &sim,
                                                                   Real baseCellVolume = 0.0) {
    const auto &mesh = sim.mesh();
    typedef typename _Sim::Mesh Mesh;
    if (baseCellVolume == 0.0) baseCellVolume = mesh.boundingBox().volume();
    using SMatrix = typename _Sim::SMatrix;
    constexpr size_t N = _Sim::N, K = _Sim::K, Deg = _Sim::Degree;
    const typename _Sim::ETensor &EBase = sim.elementTensor(0);
    typename _Sim::ETensor Eh;
    // integrated boundary shape functions (Functions.hh:247-274)
    constexpr size_t npbe = Mesh::nodesPerBoundaryElement;
    Real wts[npbe];
    if (Deg == 1) for (size_t n = 0; n < npbe; ++n) wts[n] = 1.0 / K;
    else if (K == 3) { for (size_t n = 0; n < 3; ++n) { wts[n] = 0.0; wts[3 + n] = 1.0 / 3.0; } }
    else { wts[0] = wts[1] = 1.0 / 6.0; wts[2] = 4.0 / 6.0; }
    for (size_t be = 0; be < mesh.numBoundaryElements(); ++be) {
        const auto n = mesh.boundaryElementNormal(be);
        for (size_t i = 0; i < w_ij.size(); ++i) {
            VectorND<N> w_int;
            for (size_t ni = 0; ni < npbe; ++ni)
                w_int += (wts[ni] * mesh.boundaryElementVolume(be)) * w_ij[i](mesh.boundaryElementVolumeNode(be, ni));
            SMatrix nw_pq;
            for (size_t p = 0; p < N; ++p) for (size_t q = p; q < N; ++q) nw_pq(p, q) = 0.5 * (w_int[p] * n[q] + w_int[q] * n[p]);
            Eh.addToRow(i, EBase.doubleContract(nw_pq));
        }
    }
    Eh += EBase * mesh.volume();
    Eh /= baseCellVolume;
    return Eh;
}

// Macroscopic-strain-to-microscopic-strain tensors (:188-210): G_ijkl = [avg_e strain(w^kl) + e^kl]_ij, one
// (minor-symmetric only) tensor per element; column kl of the flattened matrix is that average strain.
template <class _Sim>
std::vector<MinorSymmetricTensor<Real, _Sim::N>> macroStrainToMicroStrainTensors(const std::vector<typename _Sim::VField> &w, const _Sim &sim) {
    constexpr size_t N = _Sim::N, F = flatLen(N);
    const size_t numElems = sim.mesh().numElements();
    std::vector<MinorSymmetricTensor<Real, N>> G(numElems);
    for (size_t ij = 0; ij < w.size(); ++ij) {
        const auto strain = sim.averageStrainField(w[ij]);
        const auto eij = _Sim::SMatrix::CanonicalBasis(ij);
        for (size_t e = 0; e < numElems; ++e) {
            const auto se = strain(e);
            for (size_t r = 0; r < F; ++r) G[e].d[r][ij] = se[r] + eij[r];
        }
    }
    return G;
}

// Exact discrete differential of the homogenized tensor with respect to the vertex positions (:383-478):
// dCh(v)[c] is the change of Ch per unit motion of vertex v along axis c, at fixed fluctuation displacements
// (they are stationary points of the cell energy, so this is the total derivative);
//   dCh_ijkl = 1/|Y| int [ div dp  eps_ij : C : eps_kl - sigma_kl : (grad w_ij grad dp) - sigma_ij : (grad w_kl grad dp) ],
// eps_ij = e_ij + strain(w_ij).  |Y| (the bounding box) is not differentiated, as in the reference.
template <class _Sim>
ShapeDerivatives::OneForm<typename _Sim::ETensor, _Sim::N>
homogenizedElasticityTensorDiscreteDifferential(const std::vector<typename _Sim::VField> &w, const _Sim &sim) {
    namespace SD = ShapeDerivatives;
    constexpr size_t N = _Sim::N, K = _Sim::K, Deg = _Sim::Degree, F = flatLen(N);
    typedef typename _Sim::Mesh Mesh;
    typedef typename _Sim::SMatrix SMatrix;
    constexpr size_t npe = Mesh::nodesPerElement;
    const auto &mesh = sim.mesh();
    if (w.size() != F) throw std::runtime_error("homogenizedElasticityTensorDiscreteDifferential: flatLen(N) fluctuation fields expected");
    SD::OneForm<typename _Sim::ETensor, N> dCh(mesh.numVertices());
    const SD::ElementQuadrature<K, Deg> quad;
    const Real invCell = 1.0 / mesh.boundingBox().volume();
    for (size_t e = 0; e < mesh.numElements(); ++e) {
        Real g[K + 1][K], gphi[npe][K];
        mesh.elementGradLambda(e, g);
        const Real vol = mesh.elementVolume(e);
        const auto &E = sim.elementTensor(e);
        // per (ij <= kl): the scalar energy and the N x N matrix A = grad w_ij^T sigma_kl + grad w_kl^T sigma_ij, integrated
        Real energy[F][F] = {}, A[F][F][N][N] = {};
        for (size_t q = 0; q < quad.numPoints; ++q) {
            SD::gradPhis<K, Deg>(g, quad.lambda[q], gphi);
            Real gw[F][N][N] = {};
            SMatrix eps[F], sig[F];
            for (size_t ij = 0; ij < F; ++ij) {
                for (size_t i = 0; i < npe; ++i) {
                    const auto wi = w[ij](mesh.elementNode(e, i));
                    for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) gw[ij][c][r] += wi[c] * gphi[i][r];
                }
                eps[ij] = SD::symmetrized<N>(gw[ij]);
                eps[ij] += SMatrix::CanonicalBasis(ij);
                sig[ij] = E.doubleContract(eps[ij]);
            }
            const Real wq = quad.weight[q] * vol;
            for (size_t ij = 0; ij < F; ++ij)
                for (size_t kl = ij; kl < F; ++kl) {
                    energy[ij][kl] += wq * eps[ij].doubleContract(sig[kl]);
                    for (size_t c = 0; c < N; ++c) for (size_t b = 0; b < N; ++b) {
                        Real acc = 0.0;
                        for (size_t a = 0; a < N; ++a) acc += gw[ij][a][c] * sig[kl](a, b) + gw[kl][a][c] * sig[ij](a, b);
                        A[ij][kl][c][b] += wq * acc;
                    }
                }
        }
        for (size_t v = 0; v <= K; ++v) {
            auto &out = dCh(mesh.elementVertex(e, v));
            for (size_t c = 0; c < N; ++c)
                for (size_t ij = 0; ij < F; ++ij)
                    for (size_t kl = ij; kl < F; ++kl) {
                        Real val = energy[ij][kl] * g[v][c];
                        for (size_t b = 0; b < N; ++b) val -= A[ij][kl][c][b] * g[v][b];
                        out[c].D(ij, kl) += invCell * val;
                    }
        }
    }
    return dCh;
}

// Change in the homogenized tensor under the per-vertex perturbation delta_p.  The reference integrates its
// continuous boundary shape derivative against the normal velocity (:480-510); here the exact discrete
// differential above is applied, which is what a finite difference of the discrete Ch converges to.
template <class _Sim>
typename _Sim::ETensor deltaHomogenizedElasticityTensor(const _Sim &sim, const std::vector<typename _Sim::VField> &w,
                                                        const typename _Sim::VField &delta_p) {
    return homogenizedElasticityTensorDiscreteDifferential(w, sim)[delta_p];
}

// Change in the fluctuation displacements under delta_p (:520-540): K dw_ij = delta load(-e_ij) - (delta K) w_ij
// with the constraints of the cell problems; the solves run on the device (one batched PCG).
template <class _Sim>
std::vector<typename _Sim::VField> deltaFluctuationDisplacements(const _Sim &sim, const std::vector<typename _Sim::VField> &w,
                                                                 const typename _Sim::VField &delta_p) {
    typedef typename _Sim::VField VField;
    typedef typename _Sim::SMatrix SMatrix;
    std::vector<VField> rhs;
    for (size_t ij = 0; ij < w.size(); ++ij) {
        VField r = sim.deltaConstantStrainLoad(-SMatrix::CanonicalBasis(ij), delta_p);
        const VField dKw = sim.applyDeltaStiffnessMatrix(w[ij], delta_p);
        for (size_t k = 0; k < r.data().size(); ++k) r.data()[k] -= dKw.data()[k];
        rhs.push_back(r);
    }
    return sim.solve(rhs);
}

// Change in the macro-to-micro strain tensors under delta_p (:542-560): column kl = delta avg strain(w_kl)
template <class _Sim>
std::vector<MinorSymmetricTensor<Real, _Sim::N>> deltaMacroStrainToMicroStrainTensors(const _Sim &sim, const std::vector<typename _Sim::VField> &w,
                                                                                      const std::vector<typename _Sim::VField> &delta_w,
                                                                                      const typename _Sim::VField &delta_p) {
    constexpr size_t F = flatLen(_Sim::N);
    const size_t numElems = sim.mesh().numElements();
    std::vector<MinorSymmetricTensor<Real, _Sim::N>> deltaG(numElems);
    for (size_t kl = 0; kl < w.size(); ++kl) {
        const auto dwe = sim.deltaAverageStrainField(w[kl], delta_w[kl], delta_p);
        for (size_t e = 0; e < numElems; ++e) {
            const auto s = dwe(e);
            for (size_t r = 0; r < F; ++r) deltaG[e].d[r][kl] = s[r];
        }
    }
    return deltaG;
}

}  // namespace PeriodicHomogenization
#endif

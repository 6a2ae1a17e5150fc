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

}  // namespace PeriodicHomogenization
#endif

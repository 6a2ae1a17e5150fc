// Homogenization of the orthotropic base cell (mirrors OrthotropicHomogenization.hh:42-240 of the
// reference): for a microstructure with reflective symmetry about the coordinate planes only the
// positive orthant of the cell is meshed; the cell problems then need no periodic DoFs but
// 1 + (flatLen(N) - N) different sets of fixed variables on the same stiffness matrix
//   stretch problems: the normal displacement component vanishes on every min/max cell face,
//   shear problem s:  the complementary components vanish there,
// and the tensor of the full cell follows from the orthant's by summing the sign-flipped reflections.
//
// The reference builds one SPSDSystem (one CHOLMOD factorisation) per fixed-variable set and returns
// them; here K is assembled once on the device and the fixed-variable mask + block-Jacobi
// preconditioner are swapped between the 1 + (flatLen(N) - N) groups of solves, so nothing is returned.
#ifndef MESHFEM_B200_ORTHOTROPICHOMOGENIZATION_HH
#define MESHFEM_B200_ORTHOTROPICHOMOGENIZATION_HH
#include <MeshFEM/PeriodicHomogenization.hh>

#include <bitset>

namespace PeriodicHomogenization {
namespace Orthotropic {

template <class _Sim>
void solveCellProblems(std::vector<typename _Sim::VField> &w_ij, _Sim &sim, Real cellEpsilon = 1e-7) {
    constexpr size_t N = _Sim::N;
    typedef typename _Sim::VField VField;
    sim.removePeriodicConditions();
    sim.removeNoRigidMotionConstraint();
    {   // the reference asserts that no other constraint is active (:57-62)
        std::vector<size_t> fv;
        std::vector<Real> fx;
        sim.getFixedVariables(fv, fx, true);
        if (!fv.empty()) throw std::runtime_error("Constraints unexpected.");
    }
    const auto &mesh = sim.mesh();
    const auto &cell = mesh.boundingBox();
    typedef PeriodicBoundaryMatcher::FaceMembership<N> FM;
    std::vector<FM> fm;
    fm.reserve(mesh.numBoundaryNodes());
    for (size_t bn = 0; bn < mesh.numBoundaryNodes(); ++bn)
        fm.emplace_back(mesh.nodePosition(mesh.volumeNodeForBoundaryNode(bn)), cell, cellEpsilon);

    BENCHMARK_START_TIMER("Constant Strain Load");
    std::vector<VField> l;
    for (size_t ij = 0; ij < flatLen(N); ++ij) l.push_back(sim.constantStrainLoad(-_Sim::SMatrix::CanonicalBasis(ij)));
    BENCHMARK_STOP_TIMER("Constant Strain Load");

    w_ij.clear();
    w_ij.reserve(flatLen(N));
    std::vector<size_t> fixedVars;
    // stretch system: N loads (:77-88)
    for (size_t bn = 0; bn < mesh.numBoundaryNodes(); ++bn)
        for (size_t c = 0; c < N; ++c)
            if (fm[bn].onMinOrMaxFace(c)) fixedVars.push_back(N * (size_t)mesh.volumeNodeForBoundaryNode(bn) + c);
    {
        std::vector<VField> loads(l.begin(), l.begin() + N);
        for (auto &w : sim.solveWithFixedVariables(fixedVars, std::vector<Real>(fixedVars.size(), 0.0), loads)) w_ij.push_back(std::move(w));
    }
    // shear systems: one load each (:90-119)
    for (size_t s = 0; s < flatLen(N) - N; ++s) {
        std::vector<bool> fixVar(N * mesh.numNodes(), false);
        for (size_t bn = 0; bn < mesh.numBoundaryNodes(); ++bn) {
            const size_t ni = (size_t)mesh.volumeNodeForBoundaryNode(bn);
            for (size_t c = 0; c < N; ++c) {
                if (!fm[bn].onMinOrMaxFace(c)) continue;
                if (N == 3) {
                    fixVar.at(N * ni + s) = true;
                    if (c != s) fixVar.at(N * ni + (N - (c + s))) = true;
                } else {
                    fixVar.at(N * ni + (c == 0)) = true;
                }
            }
        }
        fixedVars.clear();
        for (size_t i = 0; i < fixVar.size(); ++i) if (fixVar[i]) fixedVars.push_back(i);
        std::vector<VField> loads(1, l[N + s]);
        w_ij.push_back(std::move(sim.solveWithFixedVariables(fixedVars, std::vector<Real>(fixedVars.size(), 0.0), loads)[0]));
    }
}

constexpr inline size_t numReflectedCells(size_t N) { return size_t(1) << N; }

// sign of fluctuation displacement ij in reflected copy r of the orthant (:149-163)
template <size_t N>
Real fluctuationDisplacementSign(size_t ij, size_t r) {
    if (ij < N) return 1.0;
    std::bitset<N> isReflected(r);
    if (N == 3) isReflected.reset(ij - N);
    return (isReflected.count() == 1) ? -1.0 : 1.0;
}

template <size_t N>
ElasticityTensor<Real, N> homogenizedTensorFromOrthoCellQuantity(const ElasticityTensor<Real, N> &EhO) {
    ElasticityTensor<Real, N> Eh;
    for (size_t r = 0; r < numReflectedCells(N); ++r)
        for (size_t kl = 0; kl < flatLen(N); ++kl) {
            const Real s_kl = fluctuationDisplacementSign<N>(kl, r);
            for (size_t ij = 0; ij <= kl; ++ij) Eh.D(ij, kl) += fluctuationDisplacementSign<N>(ij, r) * s_kl * EhO.D(ij, kl);
        }
    Eh *= 1.0 / numReflectedCells(N);
    Eh.symmetrizeFromFull();
    return Eh;
}

template <class _Sim>
typename _Sim::ETensor homogenizedElasticityTensorDisplacementForm(const std::vector<typename _Sim::VField> &w_ij, const _Sim &sim,
                                                                   Real baseCellVolume = 0.0) {
    return homogenizedTensorFromOrthoCellQuantity(PeriodicHomogenization::homogenizedElasticityTensorDisplacementForm(w_ij, sim, baseCellVolume));
}
template <class _Sim>
typename _Sim::ETensor homogenizedElasticityTensor(const std::vector<typename _Sim::VField> &w_ij, const _Sim &sim, Real baseCellVolume = 0.0) {
    return homogenizedTensorFromOrthoCellQuantity(PeriodicHomogenization::homogenizedElasticityTensor(w_ij, sim, baseCellVolume));
}

}  // namespace Orthotropic
}  // namespace PeriodicHomogenization
#endif

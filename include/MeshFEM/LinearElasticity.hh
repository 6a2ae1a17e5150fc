// LinearElasticity::Simulator -- the reference's operator facade (LinearElasticity.hh:434-1659)
// kept name-for-name for the assemble-and-solve path, with the hot path re-routed through the
// C ABI of libmfem_b200 (include/mfem_b200.h):
//   constructor negative-volume check, m_assembleStiffnessMatrix, SPSDSystem fixVariables/solve,
//   constantStrainLoad, averageStrainField/averageStressField, applyStiffnessMatrix
// run on the GPU; boundary-condition bookkeeping (O(boundary) work) stays on the host exactly
// as in the reference: applyBoundaryConditions :881-1027, applyTranslationPins :1095-1111,
// analyzeDirichletPosedness :1169-1190, assembleConstrainedSystem :1201-1249,
// m_getDirichletVarsAndValues :1469-1518, m_pinNode :1595-1618, neumannLoad :703-717.
//
// Systems with Lagrange-multiplier rows (no_rigid_motion without periodicity, or an unconstrained
// translation without the pin option) are saddle-point systems the reference hands to UMFPACK; here the
// rows are resolved on the host around the device PCG (RigidMotionConstraints.hh: multipliers from the
// free rigid modes, a consistent semi-definite solve, then the rigid part fixed so that C u = d).
#ifndef MESHFEM_B200_LINEARELASTICITY_HH
#define MESHFEM_B200_LINEARELASTICITY_HH
#include <MeshFEM/BoundaryConditions.hh>
#include <MeshFEM/FEMMesh.hh>
#include <MeshFEM/Fields.hh>
#include <MeshFEM/GlobalBenchmark.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/RigidMotionConstraints.hh>
#include <MeshFEM/ShapeDerivatives.hh>
#include <MeshFEM/SparseMatrices.hh>

#include <iostream>
#include <memory>

namespace LinearElasticity {

template <size_t _N>
struct ETensorStoreGetter {
    typedef ElasticityTensor<Real, _N> ETensor;
    ETensorStoreGetter(const ETensor &E) : m_E(E) {}
    ETensorStoreGetter() : m_E(1, 0) {}
    const ETensor &operator()() const { return m_E; }
    ETensor &operator()() { return m_E; }
private:
    ETensor m_E;
};

template <size_t _K, size_t _Deg> using Mesh = FEMMesh<_K, _Deg, VectorND<_K>>;

template <class _Mesh>
class Simulator {
public:
    typedef _Mesh Mesh;
    static constexpr size_t N = _Mesh::K;
    static constexpr size_t K = _Mesh::K;
    static constexpr size_t Degree = _Mesh::Deg;
    typedef VectorND<N> Point;
    typedef ScalarField<Real> SField;
    typedef VectorField<Real, N> VField;
    typedef ElasticityTensor<Real, N> ETensor;
    typedef SymmetricMatrixValue<Real, N> SMatrix;
    typedef SymmetricMatrixField<Real, N> SMField;
    typedef TripletMatrix<Triplet<Real>> TMatrix;

    struct BoundaryNodeData {
        ComponentMask dirichletComponents;
        Point dirichletDisplacement;
        size_t dirichletRegionIdx = 0;
        bool hasDirichlet() const { return dirichletComponents.hasAny(N); }
        void setDirichlet(ComponentMask mask, const Point &val) {       // :392-406
            for (size_t c = 0; c < N; ++c) {
                if (!mask.has(c)) continue;
                if (!dirichletComponents.has(c)) { dirichletComponents.set(c); dirichletDisplacement[c] = val[c]; }
                else if (std::abs(dirichletDisplacement[c] - val[c]) > 1e-10)
                    throw std::runtime_error("Conflicting dirichlet displacements.");
            }
        }
        void setDirichletRegion(size_t idx) {
            if (dirichletRegionIdx != 0 && dirichletRegionIdx != idx)
                std::cerr << "WARNING: region traction currently unsupported for vertices belonging to multiple regions" << std::endl;
            dirichletRegionIdx = idx;
        }
    };
    struct BoundaryElementData {
        Point neumannTraction;
        bool isInternal = false;
    };

    // device >= 0: CUDA device of the handle.  device < 0: host-only mode (no handle is created;
    // boundary-condition bookkeeping, loads and fixed-variable lists work, everything that needs
    // the GPU throws) -- used by the CPU unit tests of the host logic.
    template <class Elements, class Vertices>
    Simulator(const Elements &elems, const Vertices &vertices, int device = 0)
        : m_useRigidMotionConstraint(false), m_useNRTPinConstraint(false), m_hostOnly(device < 0), m_mesh(elems, vertices) {
        m_system.setDevice(device);
        m_bnData.resize(m_mesh.numBoundaryNodes());
        m_beData.resize(m_mesh.numBoundaryElements());
        m_uploadMesh();          // throws "Mesh has negatively oriented elements." (:465-472)
        setMaterial(Materials::Constant<N>().getTensor());
    }

    const _Mesh &mesh() const { return m_mesh; }
    _Mesh &mesh() { return m_mesh; }

    // ---- material (element(i)->configure(store) in Simulate_cli.cc:104-175)
    void setMaterial(const ETensor &E) {
        m_E = E;
        m_perElementE.clear();
        if (m_hostOnly) return;
        Real D[ETensor::F * ETensor::F];
        E.getFlat(D);
        mfemCheck(h(), mfem_b200_set_material_constant(h(), D));
        m_system.clear();
    }
    void setPerElementMaterial(const std::vector<ETensor> &Es) {
        if (Es.size() != m_mesh.numElements()) throw std::runtime_error("Material parameter fields of incorrect size.");
        m_perElementE = Es;
        if (m_hostOnly) return;
        std::vector<Real> D(Es.size() * ETensor::F * ETensor::F);
        for (size_t e = 0; e < Es.size(); ++e) Es[e].getFlat(&D[e * ETensor::F * ETensor::F]);
        mfemCheck(h(), mfem_b200_set_material_per_element(h(), D.data()));
        m_system.clear();
    }
    const ETensor &elementTensor(size_t e) const { return m_perElementE.empty() ? m_E : m_perElementE[e]; }

    // ---- solve (:479-487, 657)
    VField solve(const VField &f) const {
        if (!m_system.isSet()) m_buildConstrainedSystem();
        if (m_constraintRows.m() > 0) return m_solveWithConstraintRows(std::vector<VField>(1, f))[0];
        BENCHMARK_START_TIMER_SECTION("Elasticity Solve");
        std::vector<Real> x;
        m_system.solve(f.data(), x);
        BENCHMARK_STOP_TIMER_SECTION("Elasticity Solve");
        return dofToNodeField(x);
    }
    VField solve() const { return solve(neumannLoad()); }
    // Solve against the assembled K with an explicit set of fixed variables instead of the ones the
    // boundary conditions imply (OrthotropicHomogenization.hh:64-119 builds one SPSDSystem per set; here
    // the matrix stays on the device and only the mask + preconditioner change).  The cached
    // constrained system is invalidated afterwards.
    std::vector<VField> solveWithFixedVariables(const std::vector<size_t> &fixedVars, const std::vector<Real> &fixedVarValues,
                                                const std::vector<VField> &fs) const {
        mfemCheck(h(), mfem_b200_clear_fixed_variables(h()));
        m_system.setAssembled(N * numDoFs());
        m_system.fixVariables(fixedVars, fixedVarValues);
        BENCHMARK_START_TIMER_SECTION("Elasticity Solve");
        std::vector<std::vector<Real>> rhs, xs;
        for (const auto &f : fs) rhs.push_back(f.data());
        m_system.solveMultiple(rhs, xs);
        BENCHMARK_STOP_TIMER_SECTION("Elasticity Solve");
        std::vector<VField> result;
        for (const auto &x : xs) result.push_back(dofToNodeField(x));
        m_system.clear();
        return result;
    }
    // all right-hand sides against one assembled system (batched on the device when there are flatLen(N))
    std::vector<VField> solve(const std::vector<VField> &fs) const {
        if (!m_system.isSet()) m_buildConstrainedSystem();
        if (m_constraintRows.m() > 0) return m_solveWithConstraintRows(fs);
        BENCHMARK_START_TIMER_SECTION("Elasticity Solve");
        std::vector<std::vector<Real>> rhs, xs;
        for (const auto &f : fs) rhs.push_back(f.data());
        m_system.solveMultiple(rhs, xs);
        BENCHMARK_STOP_TIMER_SECTION("Elasticity Solve");
        std::vector<VField> result;
        for (const auto &x : xs) result.push_back(dofToNodeField(x));
        return result;
    }
    void setSolverTolerance(double rtol, int maxIters = 200000) { m_system.setTolerance(rtol, maxIters); }
    const mfem_b200_solve_info &lastSolveInfo() const { return m_system.lastSolveInfo(); }
    // multipliers of the last solve with Lagrange rows, one vector per right-hand side
    const std::vector<std::vector<Real>> &lastLagrangeMultipliers() const { return m_lastMultipliers; }

    // ---- fields
    SMField averageStrainField(const VField &u) const {      // :528-537
        SMField s(m_mesh.numElements());
        mfemCheck(h(), mfem_b200_avg_strain_stress(h(), u.data().data(), s.data().data(), nullptr));
        return s;
    }
    SMField averageStressField(const VField &u) const {      // :540-549
        SMField s(m_mesh.numElements());
        mfemCheck(h(), mfem_b200_avg_strain_stress(h(), u.data().data(), nullptr, s.data().data()));
        return s;
    }
    // Full-degree strain / stress (Element::strain :99-116, Simulator::strainField / stressField): the strain
    // of a degree-Deg displacement is a degree-(Deg-1) interpolant -- one value per element for Deg 1, a
    // linear interpolant given by its values at the K+1 vertices for Deg 2.  Returned UPSAMPLED to the
    // element's full node set (vertex values, edge nodes = mean of the end points) as Simulate_cli writes it
    // (Simulate_cli.cc:208-224).  Host loop over the elements (post-processing, O(elements)).
    typedef SymmetricMatrixInterpolantField<Real, N> SMInterpField;
    SMInterpField strainField(const VField &u) const { return m_strainOrStressField(u, false); }
    SMInterpField stressField(const VField &u) const { return m_strainOrStressField(u, true); }

    template <class _SymMat>
    VField constantStrainLoad(const _SymMat &strain) const {  // :551-562
        VField load(numDoFs());
        mfemCheck(h(), mfem_b200_const_strain_load(h(), strain.flattened().data(), load.data().data()));
        return load;
    }
    VField applyStiffnessMatrix(const VField &u) const {      // :801-823
        VField load(m_mesh.numNodes());
        mfemCheck(h(), mfem_b200_apply_K(h(), u.data().data(), load.data().data()));
        return load;
    }
    template <class _Vec>
    VField dofToNodeField(const _Vec &x) const {              // :665-677
        VField f(m_mesh.numNodes());
        for (size_t i = 0; i < m_mesh.numNodes(); ++i) {
            const size_t d = DoF(i);
            for (size_t c = 0; c < N; ++c) f[N * i + c] = x[N * d + c];
        }
        return f;
    }
    VField nodeToVertexField(const VField &x) const {
        VField f(m_mesh.numVertices());
        for (size_t i = 0; i < m_mesh.numVertices(); ++i) for (size_t c = 0; c < N; ++c) f[N * i + c] = x[N * i + c];
        return f;
    }

    // Neumann load on the DoFs (:703-717) with BoundaryElement::nodalNeumannLoad (:341-347):
    // int phi_n over a boundary element = A/K per vertex (deg 1); deg 2: faces 0 at vertices and
    // A/3 at edge nodes, edges L/6, L/6, 4L/6 (Functions.hh:247-274).
    VField neumannLoad() const {
        VField load(numDoFs());
        constexpr size_t npbe = _Mesh::nodesPerBoundaryElement;
        Real w[npbe];
        if (Degree == 1) for (size_t n = 0; n < npbe; ++n) w[n] = 1.0 / K;
        else if (K == 3) { for (size_t n = 0; n < 3; ++n) { w[n] = 0.0; w[3 + n] = 1.0 / 3.0; } }
        else { w[0] = w[1] = 1.0 / 6.0; w[2] = 4.0 / 6.0; }
        for (size_t be = 0; be < m_mesh.numBoundaryElements(); ++be)
            for (size_t n = 0; n < npbe; ++n)
                load.add(DoF(m_mesh.boundaryElementVolumeNode(be, n)), (w[n] * m_mesh.boundaryElementVolume(be)) * m_beData[be].neumannTraction);
        for (const auto &ndf : m_nodalDeltaFunctionForces) load.add(DoF(ndf.first), ndf.second);
        return load;
    }

    // ---- discrete shape derivatives (ShapeDerivatives.hh; reference :1286-1373)
    // (delta K) u: change in the force of the fixed per-NODE field u under the per-vertex perturbation deltaP,
    // returned per DoF (:1301-1330)
    VField applyDeltaStiffnessMatrix(const VField &u, const VField &deltaP) const {
        namespace SD = ShapeDerivatives;
        if (u.domainSize() != m_mesh.numNodes()) throw std::runtime_error("applyDeltaStiffnessMatrix: per-node displacement expected");
        if (deltaP.domainSize() != m_mesh.numVertices()) throw std::runtime_error("applyDeltaStiffnessMatrix: per-vertex perturbation expected");
        if (!m_hostOnly) {        // device element kernel (csrc/shape.cu k_apply_delta_K); the loop below is the host-only mirror
            VField load(numDoFs());
            mfemCheck(h(), mfem_b200_apply_delta_K(h(), u.data().data(), deltaP.data().data(), (int64_t)m_mesh.numVertices(), load.data().data()));
            return load;
        }
        constexpr size_t npe = _Mesh::nodesPerElement;
        const SD::ElementQuadrature<K, Degree> quad;
        VField load(numDoFs());
        for (size_t e = 0; e < m_mesh.numElements(); ++e) {
            Real g[K + 1][K], G[K][K], gphi[npe][K];
            m_mesh.elementGradLambda(e, g);
            SD::velocityGradient(m_mesh, e, g, deltaP, G);
            Real div = 0.0;
            for (size_t a = 0; a < K; ++a) div += G[a][a];
            const Real vol = m_mesh.elementVolume(e);
            const ETensor &E = elementTensor(e);
            Point f[npe];
            for (size_t q = 0; q < quad.numPoints; ++q) {
                SD::gradPhis<K, Degree>(g, quad.lambda[q], gphi);
                Real gu[N][N] = {}, dgu[N][N] = {};
                for (size_t i = 0; i < npe; ++i) {
                    const auto ui = u(m_mesh.elementNode(e, i));
                    for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) gu[c][r] += ui[c] * gphi[i][r];
                }
                for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) for (size_t m = 0; m < K; ++m) dgu[c][r] -= gu[c][m] * G[m][r];
                const SMatrix sig = E.doubleContract(SD::symmetrized<N>(gu)), dsig = E.doubleContract(SD::symmetrized<N>(dgu));
                const Real wq = quad.weight[q] * vol;
                for (size_t i = 0; i < npe; ++i) {
                    Real dgphi[K];
                    for (size_t r = 0; r < K; ++r) { dgphi[r] = 0.0; for (size_t m = 0; m < K; ++m) dgphi[r] -= G[m][r] * gphi[i][m]; }
                    for (size_t c = 0; c < N; ++c) {
                        Real acc = 0.0;
                        for (size_t r = 0; r < K; ++r) acc += (div * sig(c, r) + dsig(c, r)) * gphi[i][r] + sig(c, r) * dgphi[r];
                        f[i][c] += wq * acc;
                    }
                }
            }
            for (size_t i = 0; i < npe; ++i) load.add(DoF(m_mesh.elementNode(e, i)), f[i]);
        }
        return load;
    }
    // change in constantStrainLoad(cstrain) under deltaP (:1333-1348)
    template <class _SymMat>
    VField deltaConstantStrainLoad(const _SymMat &cstrain, const VField &deltaP) const {
        namespace SD = ShapeDerivatives;
        if (deltaP.domainSize() != m_mesh.numVertices()) throw std::runtime_error("deltaConstantStrainLoad: per-vertex perturbation expected");
        if (!m_hostOnly) {        // device element kernel (csrc/shape.cu k_delta_const_strain_load)
            VField dload(numDoFs());
            Real eps[flatLen(N)];
            for (size_t kf = 0; kf < flatLen(N); ++kf) eps[kf] = cstrain[kf];
            mfemCheck(h(), mfem_b200_delta_const_strain_load(h(), eps, deltaP.data().data(), (int64_t)m_mesh.numVertices(), dload.data().data()));
            return dload;
        }
        constexpr size_t npe = _Mesh::nodesPerElement;
        VField dload(numDoFs());
        Real centroid[K + 1];
        for (size_t v = 0; v <= K; ++v) centroid[v] = 1.0 / (K + 1);
        for (size_t e = 0; e < m_mesh.numElements(); ++e) {
            Real g[K + 1][K], G[K][K], gavg[npe][K];
            m_mesh.elementGradLambda(e, g);
            SD::velocityGradient(m_mesh, e, g, deltaP, G);
            // grad phi_i is (at most) linear: its element average is its centroid value
            SD::gradPhis<K, Degree>(g, centroid, gavg);
            Real div = 0.0;
            for (size_t a = 0; a < K; ++a) div += G[a][a];
            const SMatrix s = elementTensor(e).doubleContract(cstrain);
            const Real vol = m_mesh.elementVolume(e);
            for (size_t i = 0; i < npe; ++i) {
                Point l;
                for (size_t c = 0; c < N; ++c)
                    for (size_t r = 0; r < K; ++r) {
                        Real dg = 0.0;
                        for (size_t m = 0; m < K; ++m) dg -= G[m][r] * gavg[i][m];
                        l[c] += vol * s(c, r) * (div * gavg[i][r] + dg);
                    }
                dload.add(DoF(m_mesh.elementNode(e, i)), l);
            }
        }
        return dload;
    }
    // change in the element-averaged strain: avg (delta strain)(u) + avg strain(deltaU) (:1365-1375)
    SMField deltaAverageStrainField(const VField &u, const VField &deltaU, const VField &deltaP) const {
        namespace SD = ShapeDerivatives;
        if (u.domainSize() != m_mesh.numNodes() || deltaU.domainSize() != m_mesh.numNodes()) throw std::runtime_error("deltaAverageStrainField: per-node fields expected");
        if (deltaP.domainSize() != m_mesh.numVertices()) throw std::runtime_error("deltaAverageStrainField: per-vertex perturbation expected");
        if (!m_hostOnly) {        // device element kernel (csrc/shape.cu k_delta_avg_strain)
            SMField ds(m_mesh.numElements());
            mfemCheck(h(), mfem_b200_delta_avg_strain(h(), u.data().data(), deltaU.data().data(), deltaP.data().data(),
                                                      (int64_t)m_mesh.numVertices(), ds.data().data()));
            return ds;
        }
        constexpr size_t npe = _Mesh::nodesPerElement;
        SMField ds(m_mesh.numElements());
        Real centroid[K + 1];
        for (size_t v = 0; v <= K; ++v) centroid[v] = 1.0 / (K + 1);
        for (size_t e = 0; e < m_mesh.numElements(); ++e) {
            Real g[K + 1][K], G[K][K], gavg[npe][K];
            m_mesh.elementGradLambda(e, g);
            SD::velocityGradient(m_mesh, e, g, deltaP, G);
            SD::gradPhis<K, Degree>(g, centroid, gavg);
            Real gu[N][N] = {}, total[N][N] = {};
            for (size_t i = 0; i < npe; ++i) {
                const auto ui = u(m_mesh.elementNode(e, i)), dui = deltaU(m_mesh.elementNode(e, i));
                for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) { gu[c][r] += ui[c] * gavg[i][r]; total[c][r] += dui[c] * gavg[i][r]; }
            }
            for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) for (size_t m = 0; m < K; ++m) total[c][r] -= gu[c][m] * G[m][r];
            const SMatrix eps = SD::symmetrized<N>(total);
            for (size_t kf = 0; kf < flatLen(N); ++kf) ds.data()[flatLen(N) * e + kf] = eps[kf];
        }
        return ds;
    }

    bool usingReducedDoFs() const { return m_dofForNode.size() == m_mesh.numNodes(); }
    size_t numDoFs() const { return usingReducedDoFs() ? m_numDoFs : m_mesh.numNodes(); }
    size_t DoF(size_t node) const { return usingReducedDoFs() ? m_dofForNode[node] : node; }

    // ---- periodic conditions (:845-854, 874-879)
    void applyPeriodicConditions(Real epsilon = 1e-7, bool ignoreMismatch = false, std::unique_ptr<PeriodicCondition<N>> pc = nullptr) {
        m_system.clear();
        if (!pc) pc.reset(new PeriodicCondition<N>(m_mesh, epsilon, ignoreMismatch));
        m_dofForNode = pc->periodicDoFsForNodes();
        m_numDoFs = pc->numPeriodicDoFs();
        for (size_t i = 0; i < m_mesh.numBoundaryElements(); ++i) m_beData[i].isInternal = pc->isPeriodicBE(i);
        m_uploadMesh();
    }
    // (re-)embed the mesh elements (:1279-1284): vertex positions change, connectivity, periodic DoF
    // identification and boundary conditions stay; the device copy of the mesh is refreshed.
    template <typename Vertices>
    void updateMeshNodePositions(const Vertices &vertices) {
        m_mesh.setNodePositions(vertices);
        m_system.clear();
        m_uploadMesh();
    }
    void removePeriodicConditions() {
        m_system.clear();
        m_dofForNode.clear();
        for (auto &be : m_beData) be.isInternal = false;
        m_uploadMesh();
    }
    bool isInternalBoundaryElement(size_t be) const { return m_beData[be].isInternal; }

    // ---- boundary conditions (:881-1027)
    void applyBoundaryConditions(const std::vector<CondPtr<N>> &conds) {
        ExpressionEnvironment env;
        const auto &mbb = m_mesh.boundingBox();
        env.setVectorValue("mesh_size_", mbb.dimensions());
        env.setVectorValue("mesh_min_", mbb.minCorner);
        env.setVectorValue("mesh_max_", mbb.maxCorner);
        size_t dirichletRegionIdx = 0;
        if (conds.size() > 0) m_system.clear();
        for (const auto &cond : conds) {
            env.setVectorValue("region_size_", cond->region->dimensions());
            env.setVectorValue("region_min_", cond->region->minCorner);
            env.setVectorValue("region_max_", cond->region->maxCorner);
            if (auto nc = dynamic_cast<const NeumannCondition<N> *>(cond.get())) {
                Real regionArea = 0.0;
                std::vector<size_t> region;
                for (size_t be = 0; be < m_mesh.numBoundaryElements(); ++be) {
                    Point center;
                    for (size_t c = 0; c < K; ++c) center += m_mesh.nodePosition(m_mesh.boundaryElementVolumeVertex(be, c));
                    center /= Real(K);
                    if (nc->containsPoint(center)) {
                        env.setXYZ(center);
                        regionArea += m_mesh.boundaryElementVolume(be);
                        region.push_back(be);
                        if (nc->type == NeumannType::Pressure) m_beData[be].neumannTraction = -nc->pressure(env) * m_mesh.boundaryElementNormal(be);
                        else m_beData[be].neumannTraction = nc->traction(env);
                    }
                }
                if (region.size() == 0) throw std::runtime_error("Neumann region unmatched");
                if (nc->type == NeumannType::Force)
                    for (size_t bei : region) m_beData[bei].neumannTraction /= regionArea;
            } else if (dynamic_cast<const TargetCondition<N> *>(cond.get()) || dynamic_cast<const TargetNodesCondition<N> *>(cond.get())) {
                std::cerr << "WARNING: ignoring target boundary conditions." << std::endl;
            } else if (auto dc = dynamic_cast<const DirichletCondition<N> *>(cond.get())) {
                ++dirichletRegionIdx;
                for (size_t bn = 0; bn < m_mesh.numBoundaryNodes(); ++bn) {
                    const Point p = m_mesh.nodePosition(m_mesh.volumeNodeForBoundaryNode(bn));
                    if (dc->containsPoint(p)) {
                        env.setXYZ(p);
                        m_bnData[bn].setDirichlet(dc->componentMask, dc->displacement(env));
                        m_bnData[bn].setDirichletRegion(dirichletRegionIdx);
                    }
                }
            } else if (auto dec = dynamic_cast<const DirichletElementsCondition<N> *>(cond.get())) {
                ++dirichletRegionIdx;
                for (size_t be = 0; be < m_mesh.numBoundaryElements(); ++be) {
                    IVectorND<N> idx;
                    for (size_t c = 0; c < K; ++c) idx[c] = m_mesh.boundaryElementVolumeVertex(be, c);
                    if (dec->containsElement(idx)) {
                        for (size_t n = 0; n < _Mesh::nodesPerBoundaryElement; ++n) {
                            const int vn = m_mesh.boundaryElementVolumeNode(be, n);
                            env.setXYZ(m_mesh.nodePosition(vn));
                            auto &bnd = m_bnData[m_mesh.boundaryNodeForVolumeNode(vn)];
                            bnd.setDirichlet(dec->componentMask, dec->displacement(env));
                            bnd.setDirichletRegion(dirichletRegionIdx);
                        }
                    }
                }
            } else if (auto nec = dynamic_cast<const NeumannElementsCondition<N> *>(cond.get())) {
                size_t numSet = 0;
                Real regionArea = 0.0;
                std::vector<size_t> forceRegion;
                for (size_t be = 0; be < m_mesh.numBoundaryElements(); ++be) {
                    UnorderedTriplet elem(m_mesh.boundaryElementVolumeVertex(be, 0), m_mesh.boundaryElementVolumeVertex(be, 1),
                                          (N == 3) ? m_mesh.boundaryElementVolumeVertex(be, 2) : 0);
                    if (nec->hasValueForElement(elem)) {
                        const auto &val = nec->getValue(elem);
                        if (val.type == NeumannType::Pressure) m_beData[be].neumannTraction = -val.pressure() * m_mesh.boundaryElementNormal(be);
                        else if (val.type == NeumannType::Traction) m_beData[be].neumannTraction = val.traction();
                        else { m_beData[be].neumannTraction = val.force(); regionArea += m_mesh.boundaryElementVolume(be); forceRegion.push_back(be); }
                        ++numSet;
                    }
                }
                if (numSet != nec->numElements()) throw std::runtime_error("Some element boundary conditions weren't matched.");
                for (size_t bei : forceRegion) m_beData[bei].neumannTraction /= regionArea;
            } else if (auto dnc = dynamic_cast<const DirichletNodesCondition<N> *>(cond.get())) {
                std::cerr << "WARNING: dirichlet region index currently not set for DirichletNodesCondition; region force printout will be inaccurate." << std::endl;
                for (size_t i = 0; i < dnc->indices.size(); ++i) {
                    const size_t ni = dnc->indices[i];
                    const int bn = ni < m_mesh.numNodes() ? m_mesh.boundaryNodeForVolumeNode(ni) : -1;
                    if (bn < 0) throw std::runtime_error("Condition applied to non-boundary node " + std::to_string(ni));
                    m_bnData[bn].setDirichlet(dnc->componentMask, dnc->displacements[i]);
                }
            } else if (auto fc = dynamic_cast<const DeltaForceCondition<N> *>(cond.get())) {
                for (size_t n = 0; n < m_mesh.numNodes(); ++n) {
                    const Point p = m_mesh.nodePosition(n);
                    if (fc->containsPoint(p)) { env.setXYZ(p); m_nodalDeltaFunctionForces.emplace_back(n, fc->force(env)); }
                }
            } else if (auto fnc = dynamic_cast<const DeltaForceNodesCondition<N> *>(cond.get())) {
                for (size_t i = 0; i < fnc->indices.size(); ++i) {
                    const size_t ni = fnc->indices[i];
                    if (ni >= m_mesh.numNodes()) throw std::runtime_error("DeltaForceNodesCondition node index out of bounds: " + std::to_string(ni));
                    m_nodalDeltaFunctionForces.emplace_back(ni, fnc->forces[i]);
                }
            } else throw std::runtime_error("Illegal BC type");
        }
    }

    void removeDirichletConditions() {
        int removeCount = 0;
        for (auto &bn : m_bnData) if (bn.hasDirichlet()) { bn.dirichletComponents.clear(); ++removeCount; }
        if (removeCount > 0) m_system.clear();
    }
    void removeNeumanConditions() { for (auto &be : m_beData) be.neumannTraction = Point::Zero(); }
    void removeAllBoundaryConditions() { removeNeumanConditions(); removeDirichletConditions(); }

    void applyNoRigidMotionConstraint() {                     // :1052-1059
        if (!m_useRigidMotionConstraint || m_rigidMotionConstraintRHS.size() != 0) {
            m_rigidMotionConstraintRHS.clear();
            m_system.clear();
            m_useRigidMotionConstraint = true;
        }
    }
    void setUsePinNoRigidTranslationConstraint(bool use) { if (use != m_useNRTPinConstraint) m_system.clear(); m_useNRTPinConstraint = use; }
    // match the rigid motion of the per-DoF field u: the no-rigid-motion rows with right-hand side R u (:1069-1076)
    void applyRigidMotionConstraint(const VField &u) {
        applyNoRigidMotionConstraint();
        m_system.clear();
        getRigidInnerProduct(u, m_rigidMotionConstraintRHS);
    }
    // R u for the rigid-mode matrix R = [rotation rows; translation rows] (:1114-1126, m_assembleRigidModeMatrix :1522-1528)
    void getRigidInnerProduct(const VField &u, std::vector<Real> &innerProduct) const {
        if (u.domainSize() != numDoFs()) throw std::runtime_error("getRigidInnerProduct: per-DoF field expected");
        ConstraintRows R;
        m_appendInfinitesimalRotationRows(R);
        m_appendTranslationRows(R);
        innerProduct.assign(R.m(), 0.0);
        for (size_t i = 0; i < R.m(); ++i) innerProduct[i] = RigidMotionConstraints::dot(R.rows[i], u.data());
    }
    // v -= sum_i (R_i . v) R_i / |R_i|^2 for the rigid-mode rows (:1128-1164; rows orthogonal, not normalised)
    void projectOutRigidComponent(VField &v, const std::vector<bool> &dofMask = std::vector<bool>()) const {
        if (v.domainSize() != numDoFs()) throw std::runtime_error("projectOutRigidComponent: per-DoF field expected");
        const bool hasDofMask = dofMask.size() == numDoFs();
        ConstraintRows R;
        m_appendInfinitesimalRotationRows(R);
        m_appendTranslationRows(R);
        std::vector<Real> rowSqNorms(R.m(), 0.0), innerProduct(R.m(), 0.0);
        auto &vd = v.data();
        for (size_t i = 0; i < R.m(); ++i)
            for (size_t j = 0; j < vd.size(); ++j) {
                if (hasDofMask && dofMask[j / N]) continue;
                rowSqNorms[i] += R.rows[i][j] * R.rows[i][j];
                innerProduct[i] += R.rows[i][j] * vd[j];
            }
        for (size_t i = 0; i < R.m(); ++i) {
            if (rowSqNorms[i] == 0.0) continue;
            for (size_t j = 0; j < vd.size(); ++j) {
                if (hasDofMask && dofMask[j / N]) continue;
                vd[j] -= innerProduct[i] * R.rows[i][j] / rowSqNorms[i];
            }
        }
    }
    void removeNoRigidMotionConstraint() { if (m_useRigidMotionConstraint) { m_system.clear(); m_useRigidMotionConstraint = false; } }

    void applyPeriodicPairDirichletConditions(std::vector<PeriodicPairDirichletCondition<N>> &pps) {   // :1087-1093
        for (auto &pp : pps) {
            std::pair<size_t, size_t> p = pp.pair(m_mesh);
            m_bnData[p.first].setDirichlet(pp.component(), Point::Zero());
            m_bnData[p.second].setDirichlet(pp.component(), Point::Zero());
        }
        if (!pps.empty()) m_system.clear();
    }

    void applyTranslationPins(const ComponentMask &c) {      // :1095-1111
        for (size_t d = 0; d < N; ++d) {
            if (!c.has(d)) continue;
            size_t bnMin = 0;
            for (size_t bn = 0; bn < m_mesh.numBoundaryNodes(); ++bn)
                if (m_mesh.nodePosition(m_mesh.volumeNodeForBoundaryNode(bn))[d] < m_mesh.nodePosition(m_mesh.volumeNodeForBoundaryNode(bnMin))[d]) bnMin = bn;
            ComponentMask dmask;
            dmask.set(d);
            m_bnData[bnMin].setDirichlet(dmask, Point::Zero());
            m_system.clear();
        }
    }

    void analyzeDirichletPosedness(ComponentMask &needsTranslations, ComponentMask &needsRotations) const {   // :1169-1190
        needsTranslations.set();
        size_t totalConstrained = 0;
        for (const auto &bn : m_bnData)
            for (size_t c = 0; c < N; ++c)
                if (bn.dirichletComponents.has(c)) { ++totalConstrained; needsTranslations.clear(c); }
        needsRotations.clear();
        if (totalConstrained == 0) needsRotations.set();
        else if (needsTranslations.hasAny(N) || (totalConstrained < ((N == 2) ? 3 : 6))) {
            std::cerr << "WARNING: analysis of partial Dirichlet rotational posedness not yet implemented!" << std::endl;
            std::cerr << "Unconstrained translation components: " << needsTranslations.componentString() << std::endl;
        }
    }

    // The constraint half of assembleConstrainedSystem (:1201-1249): which scalar variables are fixed and to
    // what, and the Lagrange-multiplier rows C with their right-hand side.
    typedef RigidMotionConstraints::Rows ConstraintRows;
    void assembleConstraints(std::vector<size_t> &fixedVars, std::vector<Real> &fixedVarValues, ConstraintRows &C,
                             bool allowIllPosed = false) const {
        fixedVars.clear(), fixedVarValues.clear();
        C.clear();
        if (m_useRigidMotionConstraint) {
            m_appendInfinitesimalRotationRows(C);              // NO RIGID ROTATIONS
            if (m_useNRTPinConstraint) m_pinNode(fixedVars, fixedVarValues);
            else m_appendTranslationRows(C);
            // rigid-motion = 0 unless a right-hand side was supplied (applyRigidMotionConstraint)
            C.rhs = m_rigidMotionConstraintRHS;
            if (C.rhs.size() == 0) C.rhs.assign(C.m(), 0.0);
            if (C.rhs.size() != C.m()) throw std::runtime_error("Invalid rigid motion RHS");
        } else if (!allowIllPosed) {
            ComponentMask needsTranslations, needsRotations;
            analyzeDirichletPosedness(needsTranslations, needsRotations);
            if (needsTranslations.hasAny(N)) {
                if (m_useNRTPinConstraint) m_pinNode(fixedVars, fixedVarValues, needsTranslations);
                else { m_appendTranslationRows(C, needsTranslations); C.rhs.assign(needsTranslations.count(N), 0.0); }
            }
            if (needsRotations.hasAny(N)) throw std::runtime_error("Unimplemented");
        }
        m_getDirichletVarsAndValues(fixedVars, fixedVarValues);
    }
    void getFixedVariables(std::vector<size_t> &fixedVars, std::vector<Real> &fixedVarValues, bool allowIllPosed = false) const {
        ConstraintRows C;
        assembleConstraints(fixedVars, fixedVarValues, C, allowIllPosed);
    }

    void reportRegionSurfaceForces(const VField &u) const {  // :1251-1270
        VField f = applyStiffnessMatrix(u);
        std::vector<Point> forces;
        for (size_t bni = 0; bni < m_mesh.numBoundaryNodes(); ++bni) {
            const size_t ri = m_bnData[bni].dirichletRegionIdx;
            if (ri + 1 > forces.size()) forces.resize(ri + 1, Point::Zero());
            forces[ri] += f(m_mesh.volumeNodeForBoundaryNode(bni));
        }
        for (size_t i = 0; i < forces.size(); ++i) {
            std::cout << "region " << i << " force:";
            for (size_t j = 0; j < N; ++j) std::cout << "\t" << forces[i][j];
            std::cout << std::endl;
        }
    }

    void dumpSystem(const std::string &path) const {         // :1272-1277
        if (!m_system.isSet()) m_buildConstrainedSystem();
        m_system.sumAndDumpUpper(path);
    }

    // Build *upper triangle* of the stiffness matrix as triplets (:1406-1466) -- assembled on the
    // device, exported through the ABI (parity checks and --dumpMatrix).
    void m_assembleStiffnessMatrix(TMatrix &Ktrip) const {
        mfemCheck(h(), mfem_b200_assemble(h()));
        int64_t nb = 0, nnzb = 0;
        mfemCheck(h(), mfem_b200_get_bsr_sizes(h(), &nb, &nnzb));
        std::vector<int64_t> rp((size_t)nb + 1);
        std::vector<int32_t> ci((size_t)nnzb);
        std::vector<Real> v((size_t)nnzb * N * N);
        mfemCheck(h(), mfem_b200_get_bsr(h(), rp.data(), ci.data(), v.data()));
        Ktrip.init(N * numDoFs(), N * numDoFs());
        for (int64_t bi = 0; bi < nb; ++bi)
            for (int64_t k = rp[bi]; k < rp[bi + 1]; ++k)
                for (size_t r = 0; r < N; ++r)
                    for (size_t c = 0; c < N; ++c) {
                        const size_t row = N * bi + r, col = N * ci[k] + c;
                        if (row <= col) Ktrip.addNZ(row, col, v[(size_t)k * N * N + r * N + c]);
                    }
    }

    mfem_b200_handle deviceHandle() const { return h(); }

private:
    mfem_b200_handle h() const {
        if (m_hostOnly) throw std::runtime_error("Simulator was constructed in host-only mode (device < 0): no GPU operations available");
        return m_system.handle();
    }

    void m_uploadMesh() {
        if (m_hostOnly) return;
        std::vector<int64_t> dof;
        if (usingReducedDoFs()) dof.assign(m_dofForNode.begin(), m_dofForNode.end());
        const int st = mfem_b200_set_mesh(h(), (int)N, (int)Degree, (int64_t)m_mesh.numNodes(), m_mesh.nodePositions().data(),
                                          (int64_t)m_mesh.numElements(), m_mesh.elementNodes().data(),
                                          dof.empty() ? nullptr : dof.data(), (int64_t)numDoFs());
        if (st == MFEM_B200_ERR_NEG_VOLUME) {
            std::cerr << mfem_b200_last_error(h()) << std::endl;
            throw std::runtime_error("Mesh has negatively oriented elements.\nCorrect with: mesh_convert --reorientNegativeElements.");
        }
        mfemCheck(h(), st);
        if (m_haveMaterialOnDevice) {
            if (m_perElementE.empty()) setMaterial(m_E); else setPerElementMaterial(std::vector<ETensor>(m_perElementE));
        }
        m_haveMaterialOnDevice = true;
    }

    void m_buildConstrainedSystem() const {                  // :1377-1404
        std::vector<size_t> fixedVars;
        std::vector<Real> fixedVarValues;
        BENCHMARK_START_TIMER("Assemble System");
        assembleConstraints(fixedVars, fixedVarValues, m_constraintRows);
        m_systemFixedVars = fixedVars;
        BENCHMARK_STOP_TIMER("Assemble System");
        mfemCheck(h(), mfem_b200_clear_fixed_variables(h()));
        m_system.setAssembled(N * numDoFs());
        BENCHMARK_START_TIMER_SECTION("Fix Variables");
        m_system.fixVariables(fixedVars, fixedVarValues);
        BENCHMARK_STOP_TIMER_SECTION("Fix Variables");
        m_system.setEconomyMode(true);
    }

    // ---- Lagrange rows, dense over the N * numDoFs() variables
    static constexpr size_t numRotModes = (N == 3) ? 3 : 1;
    // no-rigid-rotation rows (:1530-1568).  Periodic conditions pin the rotations, so the rows are skipped then.
    void m_appendInfinitesimalRotationRows(ConstraintRows &R) const {
        if ((N == 2) && (numDoFs() < m_mesh.numNodes())) return;
        if (numDoFs() + 1 < m_mesh.numNodes()) return;
        if (numDoFs() < m_mesh.numNodes()) throw std::runtime_error("Single pair periodic BC unsupported in 3D.");
        const size_t old = R.rows.size(), nn = m_mesh.numNodes();
        R.rows.resize(old + numRotModes, std::vector<Real>(N * numDoFs(), 0.0));
        for (size_t k = 0; k < nn; ++k) {
            const Point x = m_mesh.nodePosition(k);
            if (N == 3) {
                R.rows[old    ][N * k + 1] = -x[2]; R.rows[old    ][N * k + 2] =  x[1];    // x axis: (0, -z, y)
                R.rows[old + 1][N * k    ] =  x[2]; R.rows[old + 1][N * k + 2] = -x[0];    // y axis: (z, 0, -x)
                R.rows[old + 2][N * k    ] = -x[1]; R.rows[old + 2][N * k + 1] =  x[0];    // z axis: (-y, x, 0)
            } else {
                R.rows[old][N * k] = -x[1]; R.rows[old][N * k + 1] = x[0];                  // "z axis": (-y, x)
            }
        }
    }
    // no-rigid-translation rows on the (possibly periodic) DoFs (:1571-1593)
    void m_appendTranslationRows(ConstraintRows &T, const ComponentMask &components = ComponentMask("xyz")) const {
        for (size_t c = 0; c < N; ++c) {
            if (!components.has(c)) continue;
            std::vector<Real> row(N * numDoFs(), 0.0);
            for (size_t i = 0; i < numDoFs(); ++i) row[N * i + c] = 1.0;
            T.rows.push_back(std::move(row));
        }
    }
public:
    // candidates for the null space of K on the DoFs: translations always, infinitesimal rotations when
    // nodes and DoFs coincide (periodic identification is not rotation invariant)
    std::vector<std::vector<Real>> candidateRigidModes() const {
        ConstraintRows B;
        m_appendTranslationRows(B);
        if (numDoFs() == m_mesh.numNodes()) m_appendInfinitesimalRotationRows(B);
        return B.rows;
    }
private:
    std::vector<VField> m_solveWithConstraintRows(const std::vector<VField> &fs) const {
        BENCHMARK_START_TIMER_SECTION("Elasticity Solve");
        std::vector<std::vector<Real>> rhs;
        for (const auto &f : fs) rhs.push_back(f.data());
        // the system under the rows is singular (its rigid modes are free): block-Jacobi PCG handles a consistent
        // semi-definite system, the aggregation coarse space (whose E = Z'KZ would be singular too) stays off
        m_system.setOption("coarse_aggregates", 0);
        auto us = RigidMotionConstraints::solve(N * numDoFs(), m_constraintRows, m_systemFixedVars, candidateRigidModes(), rhs,
            [&](const std::vector<std::vector<Real>> &bs) {
                std::vector<std::vector<Real>> xs;
                if (bs.size() == 1) { xs.resize(1); m_system.solve(bs[0], xs[0]); }
                else m_system.solveMultiple(bs, xs);
                return xs;
            }, &m_lastMultipliers);
        BENCHMARK_STOP_TIMER_SECTION("Elasticity Solve");
        std::vector<VField> result;
        for (const auto &x : us) result.push_back(dofToNodeField(x));
        return result;
    }

    void m_getDirichletVarsAndValues(std::vector<size_t> &dirichletVars, std::vector<Real> &dirichletValues) const {   // :1469-1518
        std::vector<Point> constraintDisplacements;
        std::vector<size_t> constraintDoFs;
        std::vector<ComponentMask> constraintComponents;
        std::vector<int> constraintIndex(numDoFs(), -1);
        for (size_t i = 0; i < m_mesh.numBoundaryNodes(); ++i) {
            const auto &bn = m_bnData[i];
            if (!bn.hasDirichlet()) continue;
            const size_t dof = DoF(m_mesh.volumeNodeForBoundaryNode(i));
            if (constraintIndex[dof] < 0) {
                constraintIndex[dof] = (int)constraintDoFs.size();
                constraintDoFs.push_back(dof);
                constraintDisplacements.push_back(bn.dirichletDisplacement);
                constraintComponents.push_back(bn.dirichletComponents);
            } else {
                std::cerr << "WARNING: Dirichlet condition on periodic boundary applies to all identified nodes." << std::endl;
                const auto diff = bn.dirichletDisplacement - constraintDisplacements[constraintIndex[dof]];
                const bool cdiffer = (bn.dirichletComponents != constraintComponents[constraintIndex[dof]]);
                if ((diff.norm() > 1e-10) || cdiffer) throw std::runtime_error("Mismatched Dirichlet constraint on periodic DoF");
            }
        }
        for (size_t i = 0; i < constraintDoFs.size(); ++i)
            for (size_t c = 0; c < N; ++c) {
                if (!constraintComponents[i].has(c)) continue;
                dirichletVars.push_back(N * constraintDoFs[i] + c);
                dirichletValues.push_back(constraintDisplacements[i][c]);
            }
    }

    SMInterpField m_strainOrStressField(const VField &u, bool stress) const {
        constexpr size_t npe = _Mesh::nodesPerElement, F = flatLen(N);
        if (u.domainSize() != m_mesh.numNodes()) throw std::runtime_error("strainField: per-node displacement expected");
        SMInterpField out(m_mesh.numElements(), npe);
        for (size_t e = 0; e < m_mesh.numElements(); ++e) {
            Real g[K + 1][K];
            m_mesh.elementGradLambda(e, g);
            SMatrix atVertex[K + 1];
            const size_t nv = (Degree == 1) ? 1 : K + 1;            // evaluation points of the interpolant
            for (size_t v = 0; v < nv; ++v) {
                Real grad[N][N] = {};                               // grad[c][r] = d u_c / d x_r at vertex v
                for (size_t i = 0; i < npe; ++i) {
                    Real gphi[K] = {};
                    if (Degree == 1) { for (size_t r = 0; r < K; ++r) gphi[r] = g[i][r]; }
                    else if (i <= K) { const Real w = (i == v) ? 3.0 : -1.0; for (size_t r = 0; r < K; ++r) gphi[r] = w * g[i][r]; }
                    else {                                          // edge node (s, e): 4 (x_e grad l_s + x_s grad l_e)
                        const size_t k = i - (K + 1);
                        const size_t es = Simplex::edgeStartNode(k), ee = Simplex::edgeEndNode(k);
                        if (v == es) for (size_t r = 0; r < K; ++r) gphi[r] = 4.0 * g[ee][r];
                        else if (v == ee) for (size_t r = 0; r < K; ++r) gphi[r] = 4.0 * g[es][r];
                    }
                    const auto ui = u(m_mesh.elementNode(e, i));
                    for (size_t c = 0; c < N; ++c) for (size_t r = 0; r < K; ++r) grad[c][r] += ui[c] * gphi[r];
                }
                SMatrix eps;
                for (size_t c = 0; c < N; ++c) for (size_t r = c; r < N; ++r) eps(c, r) = 0.5 * (grad[c][r] + grad[r][c]);
                atVertex[v] = stress ? elementTensor(e).doubleContract(eps) : eps;
            }
            for (size_t n = 0; n < npe; ++n)
                for (size_t kf = 0; kf < F; ++kf) {
                    Real val;
                    if (Degree == 1) val = atVertex[0][kf];
                    else if (n <= K) val = atVertex[n][kf];
                    else { const size_t k = n - (K + 1); val = 0.5 * (atVertex[Simplex::edgeStartNode(k)][kf] + atVertex[Simplex::edgeEndNode(k)][kf]); }
                    out(e, n, kf) = val;
                }
        }
        return out;
    }

    void m_pinNode(std::vector<size_t> &fixedVars, std::vector<Real> &fixedVarValues,
                   const ComponentMask &components = ComponentMask("xyz")) const {   // :1595-1618
        size_t nodeToPin = m_mesh.numNodes();
        for (size_t i = 0; i < m_mesh.numNodes(); ++i)
            if (m_mesh.boundaryNodeForVolumeNode(i) < 0) { nodeToPin = i; break; }
        if (nodeToPin == m_mesh.numNodes()) nodeToPin = 0;
        for (size_t d = 0; d < N; ++d)
            if (components.has(d)) { fixedVars.push_back(N * DoF(nodeToPin) + d); fixedVarValues.push_back(0.0); }
    }

    size_t m_numDoFs = 0;
    std::vector<size_t> m_dofForNode;
    bool m_useRigidMotionConstraint, m_useNRTPinConstraint, m_hostOnly;
    std::vector<std::pair<size_t, Point>> m_nodalDeltaFunctionForces;
    std::vector<BoundaryNodeData> m_bnData;
    std::vector<BoundaryElementData> m_beData;
    ETensor m_E;
    std::vector<ETensor> m_perElementE;
    bool m_haveMaterialOnDevice = false;

protected:
    mutable SPSDSystem<Real> m_system;
    mutable ConstraintRows m_constraintRows;                 // Lagrange rows of the cached system
    mutable std::vector<size_t> m_systemFixedVars;
    mutable std::vector<std::vector<Real>> m_lastMultipliers;
    std::vector<Real> m_rigidMotionConstraintRHS;
    _Mesh m_mesh;
};

}  // namespace LinearElasticity
#endif

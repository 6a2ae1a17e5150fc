// C exports of host-side helpers for the Python tests and bench.py: synthetic `grid -t`
// meshes (src/bin/tools/grid.cc:115-137 of the reference) and FEMMesh construction.  These are
// input generators / host logic, not part of the GPU ABI (include/mfem_b200.h).
#include <MeshFEM/FEMMesh.hh>
#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/PeriodicHomogenization.hh>
#include <MeshFEM/MSHFieldParser.hh>
#include <MeshFEM/TensorProjection.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/Partition.hh>
#include <MeshFEM/MeshIO.hh>
#include <MeshFEM/filters/gen_grid.hh>
#include <MeshFEM/filters/hex_tet_subdiv.hh>
#include <MeshFEM/filters/quad_tri_subdiv.hh>

#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>

namespace {

struct HostMesh {
    int dim = 0, deg = 0;
    std::vector<MeshIO::IOVertex> vertices;
    std::vector<MeshIO::IOElement> elements;
    // FEMMesh flat data
    std::vector<double> nodes;
    std::vector<int32_t> elemNodes, bdryElemNodes, bdryElemVerts, bdryNodes;
    std::vector<double> bdryVol, bdryNormal;
    double bbmin[3] = {0, 0, 0}, bbmax[3] = {0, 0, 0};
    size_t nV = 0, nbe = 0;
    std::string err;
};

thread_local std::string g_err;

template <class Mesh>
void fillFrom(HostMesh &hm, const Mesh &m, int deg) {
    constexpr size_t K = Mesh::K;
    hm.deg = deg;
    hm.nodes = m.nodePositions();
    hm.elemNodes = m.elementNodes();
    hm.bdryElemNodes = m.boundaryElementNodes();
    hm.bdryElemVerts = m.boundaryElementVertices();
    hm.nV = m.numVertices();
    hm.nbe = m.numBoundaryElements();
    hm.bdryNodes.resize(m.numBoundaryNodes());
    for (size_t i = 0; i < hm.bdryNodes.size(); ++i) hm.bdryNodes[i] = m.volumeNodeForBoundaryNode(i);
    hm.bdryVol.resize(hm.nbe);
    hm.bdryNormal.resize(hm.nbe * K);
    for (size_t be = 0; be < hm.nbe; ++be) {
        hm.bdryVol[be] = m.boundaryElementVolume(be);
        for (size_t c = 0; c < K; ++c) hm.bdryNormal[be * K + c] = m.boundaryElementNormal(be)[c];
    }
    for (size_t c = 0; c < K; ++c) { hm.bbmin[c] = m.boundingBox().minCorner[c]; hm.bbmax[c] = m.boundingBox().maxCorner[c]; }
}

template <size_t K, size_t Deg>
void fill(HostMesh &hm) {
    if (hm.deg == (int)Deg && !hm.nodes.empty()) return;      // flat FEMMesh data of this degree already cached
    FEMMesh<K, Deg> m(hm.elements, hm.vertices);
    fillFrom(hm, m, (int)Deg);
}

}  // namespace

template <class T> struct TypeTag { typedef T type; };

extern "C" {

const char *mfemhost_last_error() { return g_err.c_str(); }

// `grid sx x sy [x sz] -t [-m min -M max]`: returns an opaque mesh holding vertices+simplices.
void *mfemhost_grid(int ndim, const int64_t *sizes, const double *minCorner, const double *maxCorner) {
    try {
        auto *hm = new HostMesh();
        std::vector<size_t> sz(sizes, sizes + ndim);
        std::vector<MeshIO::IOVertex> gv;
        std::vector<MeshIO::IOElement> ge;
        gen_grid(sz, gv, ge);
        if (minCorner && maxCorner) {
            Point3D scale, mn;
            for (int i = 0; i < 3; ++i) { mn[i] = i < ndim ? minCorner[i] : 0.0; scale[i] = i < ndim ? (maxCorner[i] - minCorner[i]) / sizes[i] : 0.0; }
            for (auto &v : gv) for (int i = 0; i < 3; ++i) v.point[i] = scale[i] * v.point[i] + mn[i];
        }
        std::vector<size_t> cellIdx;
        if (ndim == 2) quad_tri_subdiv(gv, ge, hm->vertices, hm->elements, cellIdx);
        else hex_tet_subdiv(gv, ge, hm->vertices, hm->elements, cellIdx);
        hm->dim = ndim;
        return hm;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

// Perforated periodic cell of BASELINE config 4: n^ndim voxels on [0,1]^ndim with the centred hole^ndim block
// removed BEFORE the symmetric simplex subdivision; unused vertices are dropped.
void *mfemhost_perforated_cell(int ndim, int64_t n, int64_t hole) {
    try {
        if (ndim != 2 && ndim != 3) throw std::runtime_error("perforated_cell: ndim must be 2 or 3");
        if (n < 1 || hole < 0 || hole >= n || ((n - hole) % 2) != 0) throw std::runtime_error("perforated_cell: need 0 <= hole < n and n - hole even");
        auto *hm = new HostMesh();
        std::vector<size_t> sz((size_t)ndim, (size_t)n);
        std::vector<MeshIO::IOVertex> gv;
        std::vector<MeshIO::IOElement> ge, kept;
        gen_grid(sz, gv, ge);
        for (auto &v : gv) for (int i = 0; i < ndim; ++i) v.point[i] /= (Real)n;
        const int64_t lo = (n - hole) / 2, hi = lo + hole;       // removed index range [lo, hi) per axis
        // gen_grid element order: slices (z) outermost, then rows (y), then columns (x)
        size_t e = 0;
        for (int64_t s = 0; s < (ndim == 3 ? n : 1); ++s)
            for (int64_t r = 0; r < n; ++r)
                for (int64_t c = 0; c < n; ++c, ++e) {
                    const bool inHole = c >= lo && c < hi && r >= lo && r < hi && (ndim == 2 || (s >= lo && s < hi));
                    if (!inHole) kept.push_back(ge[e]);
                }
        std::vector<MeshIO::IOVertex> sv;
        std::vector<MeshIO::IOElement> se;
        std::vector<size_t> cellIdx;
        if (ndim == 2) quad_tri_subdiv(gv, kept, sv, se, cellIdx);
        else hex_tet_subdiv(gv, kept, sv, se, cellIdx);
        std::vector<int64_t> remap(sv.size(), -1);
        for (const auto &el : se) for (size_t c = 0; c < el.size(); ++c) remap[el[c]] = 0;
        int64_t next = 0;
        for (size_t v = 0; v < sv.size(); ++v) if (remap[v] == 0) { remap[v] = next++; hm->vertices.push_back(sv[v]); }
        hm->elements = se;
        for (auto &el : hm->elements) for (size_t c = 0; c < el.size(); ++c) el[c] = (size_t)remap[el[c]];
        hm->dim = ndim;
        return hm;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void *mfemhost_load_mesh(const char *path, int *dimOut) {
    try {
        auto *hm = new HostMesh();
        auto type = MeshIO::load(path, hm->vertices, hm->elements);
        if (type == MeshIO::MESH_TET) hm->dim = 3;
        else if (type == MeshIO::MESH_TRI) hm->dim = 2;
        else { delete hm; throw std::runtime_error("Mesh must be pure triangle or tet."); }
        if (dimOut) *dimOut = hm->dim;
        return hm;
    } catch (const std::exception &e) { g_err = e.what(); return nullptr; }
}

void *mfemhost_from_arrays(int dim, int64_t nV, const double *V3, int64_t nE, const int64_t *E) {
    auto *hm = new HostMesh();
    hm->dim = dim;
    hm->vertices.resize(nV);
    for (int64_t i = 0; i < nV; ++i) hm->vertices[i].set(V3[3 * i], V3[3 * i + 1], V3[3 * i + 2]);
    hm->elements.assign(nE, MeshIO::IOElement(dim + 1));
    for (int64_t e = 0; e < nE; ++e) for (int c = 0; c <= dim; ++c) hm->elements[e][c] = (size_t)E[e * (dim + 1) + c];
    return hm;
}

void mfemhost_free(void *m) { delete static_cast<HostMesh *>(m); }

int mfemhost_raw_sizes(void *m, int64_t *nV, int64_t *nE) {
    auto *hm = static_cast<HostMesh *>(m);
    *nV = (int64_t)hm->vertices.size(); *nE = (int64_t)hm->elements.size();
    return 0;
}
int mfemhost_raw_copy(void *m, double *V3, int64_t *E) {
    auto *hm = static_cast<HostMesh *>(m);
    for (size_t i = 0; i < hm->vertices.size(); ++i) for (int c = 0; c < 3; ++c) V3[3 * i + c] = hm->vertices[i][c];
    const size_t n = hm->dim + 1;
    for (size_t e = 0; e < hm->elements.size(); ++e) for (size_t c = 0; c < n; ++c) E[e * n + c] = (int64_t)hm->elements[e][c];
    return 0;
}

// Build FEMMesh<dim,deg>; sizes: [numNodes, numElements, numBoundaryElements, numBoundaryNodes, numVertices]
int mfemhost_build_femmesh(void *m, int deg, int64_t *sizes5) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        if (hm->dim == 3 && deg == 1) fill<3, 1>(*hm);
        else if (hm->dim == 3 && deg == 2) fill<3, 2>(*hm);
        else if (hm->dim == 2 && deg == 1) fill<2, 1>(*hm);
        else if (hm->dim == 2 && deg == 2) fill<2, 2>(*hm);
        else throw std::runtime_error("bad dim/deg");
        sizes5[0] = (int64_t)(hm->nodes.size() / hm->dim);
        sizes5[1] = (int64_t)hm->elements.size();
        sizes5[2] = (int64_t)hm->nbe;
        sizes5[3] = (int64_t)hm->bdryNodes.size();
        sizes5[4] = (int64_t)hm->nV;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int mfemhost_femmesh_copy(void *m, double *nodes, int32_t *elemNodes, int32_t *bdryElemNodes, int32_t *bdryElemVerts,
                          int32_t *bdryNodes, double *bdryVol, double *bdryNormal, double *bbox6) {
    auto *hm = static_cast<HostMesh *>(m);
    auto cp = [](auto &v, auto *dst) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(hm->nodes, nodes); cp(hm->elemNodes, elemNodes); cp(hm->bdryElemNodes, bdryElemNodes);
    cp(hm->bdryElemVerts, bdryElemVerts); cp(hm->bdryNodes, bdryNodes); cp(hm->bdryVol, bdryVol);
    cp(hm->bdryNormal, bdryNormal);
    if (bbox6) for (int c = 0; c < 3; ++c) { bbox6[c] = hm->bbmin[c]; bbox6[3 + c] = hm->bbmax[c]; }
    return 0;
}

}  // extern "C"

// Host-side boundary-condition bookkeeping of LinearElasticity::Simulator (host-only mode, no GPU):
// parse a .bc JSON text, apply it, and return the fixed variables + values and the Neumann load.
// periodic != 0 additionally applies PeriodicCondition first (cell problems) with the pin constraint.
namespace {
struct BCResult {
    std::vector<int64_t> fixedVars, dofForNode;
    std::vector<double> fixedVals, load;
    std::vector<uint8_t> internalBE;
    int64_t numDoFs = 0;
    RigidMotionConstraints::Rows rows;                     // Lagrange rows of the configuration (usually none)
    std::vector<std::vector<double>> rigidModes;           // candidate null-space modes on the DoFs
};
thread_local BCResult g_bc;

template <size_t K, size_t Deg>
void runBC(HostMesh &hm, const char *bcJson, int periodic) {
    typedef LinearElasticity::Simulator<LinearElasticity::Mesh<K, Deg>> Sim;
    Sim sim(hm.elements, hm.vertices, -1);
    fillFrom(hm, sim.mesh(), (int)Deg);        // the FEMMesh is built once; femmesh() reuses it
    if (periodic == 1 || periodic == 2) {
        sim.applyPeriodicConditions(1e-7);
        sim.applyNoRigidMotionConstraint();
        sim.setUsePinNoRigidTranslationConstraint(periodic != 2);   // 2: translation rows instead of the pinned node
    } else if (periodic == 4) sim.setUsePinNoRigidTranslationConstraint(true);   // 4: pin option without periodicity
    if (bcJson && bcJson[0]) {
        std::istringstream is(bcJson);
        bool noRigidMotion;
        std::vector<PeriodicPairDirichletCondition<K>> pps;
        ComponentMask pin;
        auto conds = readBoundaryConditions<K>(is, sim.mesh().boundingBox(), noRigidMotion, pps, pin);
        sim.applyTranslationPins(pin);
        sim.applyBoundaryConditions(conds);
        sim.applyPeriodicPairDirichletConditions(pps);
        if (noRigidMotion) sim.applyNoRigidMotionConstraint();
    }
    std::vector<size_t> fv;
    std::vector<Real> fx;
    sim.assembleConstraints(fv, fx, g_bc.rows);
    g_bc.rigidModes.clear();
    if (g_bc.rows.m() > 0) g_bc.rigidModes = sim.candidateRigidModes();
    g_bc.fixedVars.assign(fv.begin(), fv.end());
    g_bc.fixedVals = fx;
    g_bc.load = sim.neumannLoad().data();
    g_bc.numDoFs = (int64_t)sim.numDoFs();
    g_bc.dofForNode.resize(sim.mesh().numNodes());
    for (size_t n = 0; n < sim.mesh().numNodes(); ++n) g_bc.dofForNode[n] = (int64_t)sim.DoF(n);
    g_bc.internalBE.resize(sim.mesh().numBoundaryElements());
    for (size_t be = 0; be < g_bc.internalBE.size(); ++be) g_bc.internalBE[be] = sim.isInternalBoundaryElement(be);
}
}  // namespace

extern "C" {

// sizes3: [numFixed, numDoFs, numBoundaryElements]
int mfemhost_apply_bc(void *m, int deg, const char *bcJson, int periodic, int64_t *sizes3) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        if (hm->dim == 3 && deg == 1) runBC<3, 1>(*hm, bcJson, periodic);
        else if (hm->dim == 3 && deg == 2) runBC<3, 2>(*hm, bcJson, periodic);
        else if (hm->dim == 2 && deg == 1) runBC<2, 1>(*hm, bcJson, periodic);
        else if (hm->dim == 2 && deg == 2) runBC<2, 2>(*hm, bcJson, periodic);
        else throw std::runtime_error("bad dim/deg");
        sizes3[0] = (int64_t)g_bc.fixedVars.size(); sizes3[1] = g_bc.numDoFs; sizes3[2] = (int64_t)g_bc.internalBE.size();
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int mfemhost_bc_copy(int64_t *fixedVars, double *fixedVals, double *load, int64_t *dofForNode, uint8_t *internalBE) {
    auto cp = [](auto &v, auto *dst) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(g_bc.fixedVars, fixedVars); cp(g_bc.fixedVals, fixedVals); cp(g_bc.load, load); cp(g_bc.dofForNode, dofForNode);
    cp(g_bc.internalBE, internalBE);
    return 0;
}

// Lagrange-multiplier rows of the last mfemhost_apply_bc: counts[0] = rows, counts[1] = candidate rigid modes,
// counts[2] = variables per row; rows / rhs / modes are copied when non-null.
int mfemhost_bc_rows(int64_t *counts3, double *rows, double *rhs, double *modes) {
    const size_t n = g_bc.load.size();                     // N * numDoFs
    counts3[0] = (int64_t)g_bc.rows.m(); counts3[1] = (int64_t)g_bc.rigidModes.size(); counts3[2] = (int64_t)n;
    for (size_t i = 0; rows && i < g_bc.rows.m(); ++i) std::memcpy(rows + i * n, g_bc.rows.rows[i].data(), n * sizeof(double));
    if (rhs && g_bc.rows.m()) std::memcpy(rhs, g_bc.rows.rhs.data(), g_bc.rows.m() * sizeof(double));
    for (size_t i = 0; modes && i < g_bc.rigidModes.size(); ++i) std::memcpy(modes + i * n, g_bc.rigidModes[i].data(), n * sizeof(double));
    return 0;
}

// RigidMotionConstraints::solve with the SPSD solve supplied by the caller (the Simulator passes the device
// PCG; the CPU tests of the host algebra pass their own): cb(nrhs, rhs[nrhs*n], u[nrhs*n]) returns 0 on success
// and must solve K_ff u_f = rhs_f - K_fc u_c with the fixed values in place.
typedef int (*mfemhost_spsd_solve_cb)(int nrhs, const double *rhs, double *u);
int mfemhost_constrained_solve(int64_t n, int m, const double *rows, const double *rowsRhs, int64_t nFixed, const int64_t *fixedVars,
                               int nModes, const double *modes, int nrhs, const double *fs, double *us, double *multipliers,
                               mfemhost_spsd_solve_cb cb) {
    try {
        namespace RMC = RigidMotionConstraints;
        RMC::Rows C;
        for (int i = 0; i < m; ++i) C.rows.emplace_back(rows + (size_t)i * n, rows + (size_t)(i + 1) * n);
        C.rhs.assign(rowsRhs, rowsRhs + m);
        std::vector<size_t> fv(fixedVars, fixedVars + nFixed);
        std::vector<RMC::Vec> B, F;
        for (int i = 0; i < nModes; ++i) B.emplace_back(modes + (size_t)i * n, modes + (size_t)(i + 1) * n);
        for (int k = 0; k < nrhs; ++k) F.emplace_back(fs + (size_t)k * n, fs + (size_t)(k + 1) * n);
        std::vector<RMC::Vec> lambdas;
        auto U = RMC::solve((size_t)n, C, fv, B, F, [&](const std::vector<RMC::Vec> &bs) {
            std::vector<double> flat(bs.size() * (size_t)n), out(bs.size() * (size_t)n);
            for (size_t k = 0; k < bs.size(); ++k) std::copy(bs[k].begin(), bs[k].end(), flat.begin() + k * n);
            if (cb((int)bs.size(), flat.data(), out.data()) != 0) throw std::runtime_error("constrained solve: the SPSD solve callback failed");
            std::vector<RMC::Vec> xs;
            for (size_t k = 0; k < bs.size(); ++k) xs.emplace_back(out.begin() + k * n, out.begin() + (k + 1) * n);
            return xs;
        }, &lambdas);
        for (int k = 0; k < nrhs; ++k) {
            std::copy(U[k].begin(), U[k].end(), us + (size_t)k * n);
            if (multipliers) std::copy(lambdas[k].begin(), lambdas[k].end(), multipliers + (size_t)k * m);
        }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// .material text -> flattened tensor (flat x flat row-major); returns 0 or -1
int mfemhost_material(int dim, const char *jsonText, double *Dout, char *roundTripJson, int roundTripCap) {
    try {
        auto j = mjson::json::parse(std::string(jsonText));
        std::string rt;
        if (dim == 3) { Materials::Constant<3> mat; mat.setFromJson(j); mat.getTensor().getFlat(Dout); rt = mat.getJson().dump(); }
        else { Materials::Constant<2> mat; mat.setFromJson(j); mat.getTensor().getFlat(Dout); rt = mat.getJson().dump(); }
        if (roundTripJson && roundTripCap > 0) { std::strncpy(roundTripJson, rt.c_str(), roundTripCap - 1); roundTripJson[roundTripCap - 1] = 0; }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// tinyexpr-style expression evaluation with variables x, y, z
int mfemhost_eval_expr(const char *expr, double x, double y, double z, double *out) {
    try {
        ExpressionEnvironment env;
        env.setValue("x", x); env.setValue("y", y); env.setValue("z", z);
        *out = Expression(expr).eval(env);
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// MSHFieldParser: number of entries of the named field (kind 0 scalar, 1 vector, 2 symmetric matrix;
// domain 0 per-element, 1 per-node, 2 any) and, when `out` is non-null, its values (flattened).
int mfemhost_msh_field(int dim, const char *path, const char *name, int kind, int domain, double *out, int64_t cap,
                       int64_t *nEntries, int *actualDomain) {
    try {
        const DomainType want = domain == 0 ? DomainType::PER_ELEMENT : (domain == 1 ? DomainType::PER_NODE : DomainType::ANY);
        DomainType got = want;
        auto run = [&](auto &parser) {
            const std::vector<Real> *data = nullptr;
            size_t n = 0;
            if (kind == 0) { const auto &f = parser.scalarField(name, want, got); data = &f.data(); n = f.domainSize(); }
            else if (kind == 1) { const auto &f = parser.vectorField(name, want, got); data = &f.data(); n = f.domainSize(); }
            else { const auto &f = parser.symmetricMatrixField(name, want); data = &f.data(); n = f.domainSize(); }
            *nEntries = (int64_t)n;
            if (out) {
                if ((int64_t)data->size() > cap) throw std::runtime_error("output buffer too small");
                std::memcpy(out, data->data(), data->size() * sizeof(Real));
            }
        };
        if (dim == 3) { MSHFieldParser<3> p(path); run(p); } else { MSHFieldParser<2> p(path); run(p); }
        if (actualDomain) *actualDomain = got == DomainType::PER_ELEMENT ? 0 : 1;
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// ElasticityTensor analysis used by PeriodicHomogenization_cli: eigenstrains (ascending eigenvalues,
// strains[k][component]), compliance D(inverse()), orthotropic parameters, anisotropy.
int mfemhost_closest_isotropic(int dim, const double *Dflat, double *out) {
    try {
        if (dim == 3) { ElasticityTensor<Real, 3> E; E.setFlat(Dflat); closestIsotropicTensor(E).getFlat(out); }
        else { ElasticityTensor<Real, 2> E; E.setFlat(Dflat); closestIsotropicTensor(E).getFlat(out); }
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int mfemhost_tensor_analysis(int dim, const double *Dflat, double *lambdas, double *strains, double *compliance,
                             double *ortho, double *anisotropy) {
    try {
        auto run = [&](auto E) {
            constexpr size_t F = decltype(E)::F;
            E.setFlat(Dflat);
            const auto eig = E.computeEigenstrains();
            for (size_t k = 0; k < F; ++k) { lambdas[k] = eig.lambdas[k]; for (size_t i = 0; i < F; ++i) strains[k * F + i] = eig.strains[k][i]; }
            E.inverse().getFlat(compliance);
            *anisotropy = E.anisotropy();
            if (decltype(E)::Dim == 3) E.getOrthotropic3D(ortho[0], ortho[1], ortho[2], ortho[3], ortho[4], ortho[5], ortho[6], ortho[7], ortho[8]);
            else E.getOrthotropic2D(ortho[0], ortho[1], ortho[2], ortho[3]);
        };
        if (dim == 3) run(ElasticityTensor<Real, 3>()); else run(ElasticityTensor<Real, 2>());
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// Simulator::strainField / stressField (full-degree, upsampled to the element's nodes) in host-only mode:
// out[numElements][nodesPerElem][flat]; with path != NULL also written as $ElementNodeData "strain"/"stress"
// next to the $NodeData "u" (MSHFieldWriter, full-degree output as Simulate_cli -D).
int mfemhost_strain_field(void *m, int deg, const double *uNodes, int stress, const double *Dflat, double *out,
                          const char *path, int binary) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        auto run = [&](auto simTag) {
            typedef typename decltype(simTag)::type Sim;
            Sim sim(hm->elements, hm->vertices, -1);
            typename Sim::ETensor E;
            E.setFlat(Dflat);
            sim.setMaterial(E);
            typename Sim::VField u(sim.mesh().numNodes());
            std::copy(uNodes, uNodes + u.size(), u.data().begin());
            const auto f = stress ? sim.stressField(u) : sim.strainField(u);
            std::copy(f.data().begin(), f.data().end(), out);
            if (path) {
                MSHFieldWriter writer(path, sim.mesh(), false, MeshIO::MESH_GUESS, binary != 0);
                writer.addField("u", u, DomainType::PER_NODE);
                writer.addField(stress ? "stress" : "strain", f, DomainType::PER_ELEMENT);
            }
        };
        if (hm->dim == 3 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 1>>>());
        else if (hm->dim == 3 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 2>>>());
        else if (hm->dim == 2 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 1>>>());
        else if (hm->dim == 2 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 2>>>());
        else throw std::runtime_error("bad dim/deg");
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// Host half of DeformedCells_cli --homogenize: periodic conditions on the undeformed cell, nodes moved by
// x -> J (x - centre) (J row-major dim x dim, NULL = identity), then homogenizedElasticityTensorDisplacementForm of
// the given fluctuation displacements w[flat][numNodes][dim] over |bbox| det J.  Also returns the neumann-free
// constant-strain bookkeeping the CLI relies on: dofForNode is unchanged by the deformation (checked).
int mfemhost_deformed_displacement_form(void *m, int deg, const double *Dflat, const double *J, const double *w, double *EhOut,
                                        double *deformedNodes) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        auto run = [&](auto simTag) {
            typedef typename decltype(simTag)::type Sim;
            constexpr size_t N = Sim::N, F = flatLen(N);
            Sim sim(hm->elements, hm->vertices, -1);
            typename Sim::ETensor E;
            E.setFlat(Dflat);
            sim.setMaterial(E);
            const auto bbox = sim.mesh().boundingBox();
            const auto center = bbox.center();
            sim.applyPeriodicConditions();
            sim.applyNoRigidMotionConstraint();
            sim.setUsePinNoRigidTranslationConstraint(true);
            std::vector<size_t> dofBefore(sim.mesh().numNodes());
            for (size_t n = 0; n < dofBefore.size(); ++n) dofBefore[n] = sim.DoF(n);
            Real det = 1.0;
            if (J) {
                std::vector<MeshIO::IOVertex> deformed;
                for (size_t v = 0; v < sim.mesh().numVertices(); ++v) {
                    MeshIO::IOVertex d;
                    const auto p = sim.mesh().nodePosition(v);
                    for (size_t i = 0; i < N; ++i) { Real acc = 0.0; for (size_t j = 0; j < N; ++j) acc += J[i * N + j] * (p[j] - center[j]); d[i] = acc; }
                    deformed.push_back(d);
                }
                sim.updateMeshNodePositions(deformed);
                det = (N == 2) ? J[0] * J[3] - J[1] * J[2]
                               : J[0] * (J[4] * J[8 % (N * N)] - J[5] * J[7 % (N * N)]) - J[1] * (J[3] * J[8 % (N * N)] - J[5] * J[6 % (N * N)]) +
                                 J[2] * (J[3] * J[7 % (N * N)] - J[4] * J[6 % (N * N)]);
            }
            for (size_t n = 0; n < dofBefore.size(); ++n) if (sim.DoF(n) != dofBefore[n]) throw std::runtime_error("periodic DoFs changed under deformation");
            std::vector<typename Sim::VField> w_ij;
            const size_t nn = sim.mesh().numNodes();
            for (size_t i = 0; i < F; ++i) {
                typename Sim::VField wi(nn);
                std::copy(w + i * nn * N, w + (i + 1) * nn * N, wi.data().begin());
                w_ij.push_back(wi);
            }
            PeriodicHomogenization::homogenizedElasticityTensorDisplacementForm(w_ij, sim, bbox.volume() * det).getFlat(EhOut);
            if (deformedNodes) std::copy(sim.mesh().nodePositions().begin(), sim.mesh().nodePositions().end(), deformedNodes);
        };
        if (hm->dim == 3 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 1>>>());
        else if (hm->dim == 3 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 2>>>());
        else if (hm->dim == 2 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 1>>>());
        else if (hm->dim == 2 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 2>>>());
        else throw std::runtime_error("bad dim/deg");
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// PeriodicCondition variants (BoundaryConditions.hh:457-610 of the reference): geometric matching with optional
// mismatch tolerance and non-periodic axes, or identified pairs read from a file (pcFile != NULL).
// out: dofForNode[numNodes], isPeriodicBE[numBoundaryElements]; returns numDoFs through numDofs.
int mfemhost_periodic_condition(void *m, int deg, double eps, int ignoreMismatch, int nIgnoreDims, const int64_t *ignoreDims,
                                const char *pcFile, int64_t *dofForNode, uint8_t *isPeriodicBE, int64_t *numDofs) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        auto run = [&](auto meshTag) {
            typedef typename decltype(meshTag)::type Mesh;
            Mesh mesh(hm->elements, hm->vertices);
            std::unique_ptr<PeriodicCondition<Mesh::K>> pc;
            if (pcFile) pc.reset(new PeriodicCondition<Mesh::K>(mesh, std::string(pcFile)));
            else pc.reset(new PeriodicCondition<Mesh::K>(mesh, eps, ignoreMismatch != 0, std::vector<size_t>(ignoreDims, ignoreDims + nIgnoreDims)));
            const auto &d = pc->periodicDoFsForNodes();
            for (size_t n = 0; n < d.size(); ++n) dofForNode[n] = (int64_t)d[n];
            for (size_t be = 0; be < mesh.numBoundaryElements(); ++be) isPeriodicBE[be] = pc->isPeriodicBE(be);
            *numDofs = (int64_t)pc->numPeriodicDoFs();
        };
        if (hm->dim == 3 && deg == 1) run(TypeTag<LinearElasticity::Mesh<3, 1>>());
        else if (hm->dim == 3 && deg == 2) run(TypeTag<LinearElasticity::Mesh<3, 2>>());
        else if (hm->dim == 2 && deg == 1) run(TypeTag<LinearElasticity::Mesh<2, 1>>());
        else if (hm->dim == 2 && deg == 2) run(TypeTag<LinearElasticity::Mesh<2, 2>>());
        else throw std::runtime_error("bad dim/deg");
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// E : G : S for flattened tensors (what PeriodicHomogenization_cli --m2mstress writes per element), through the
// same MinorSymmetricTensor / doubleContractTensor code; text = writeUnflattened of the result.
int mfemhost_m2m_tensor(int dim, const double *Eflat, const double *Gflat, const double *Sflat, double *out, char *text, int textCap) {
    try {
        auto run = [&](auto E) {
            constexpr size_t N = decltype(E)::Dim, F = flatLen(N);
            decltype(E) S;
            E.setFlat(Eflat); S.setFlat(Sflat);
            MinorSymmetricTensor<Real, N> G;
            for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) G.d[i][j] = Gflat[i * F + j];
            const auto M = E.doubleContractTensor(G.doubleContract(S));
            for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) out[i * F + j] = M.d[i][j];
            std::ostringstream os;
            os << std::setprecision(16);
            M.writeUnflattened(os);
            if (text && textCap > 0) { std::strncpy(text, os.str().c_str(), textCap - 1); text[textCap - 1] = 0; }
        };
        if (dim == 3) run(ElasticityTensor<Real, 3>()); else run(ElasticityTensor<Real, 2>());
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

// Discrete shape derivatives on the host-only Simulator (ShapeDerivatives.hh): for the per-vertex perturbation deltaP
//   dKu[numDoFs*N]      = applyDeltaStiffnessMatrix(u, deltaP)
//   dload[numDoFs*N]    = deltaConstantStrainLoad(strain, deltaP)
//   dstrain[numElems*F] = deltaAverageStrainField(u, du, deltaP)
//   dCh[numVertices*N*F*F] = homogenizedElasticityTensorDiscreteDifferential(w)   (when w != NULL)
// periodic != 0 applies the periodic conditions first (loads are per DoF).
int mfemhost_shape_derivatives(void *m, int deg, int periodic, const double *Dflat, const double *u, const double *du, const double *strain,
                               const double *deltaP, const double *w, double *dKu, double *dload, double *dstrain, double *dCh) {
    auto *hm = static_cast<HostMesh *>(m);
    try {
        auto run = [&](auto simTag) {
            typedef typename decltype(simTag)::type Sim;
            constexpr size_t N = Sim::N, F = flatLen(N);
            Sim sim(hm->elements, hm->vertices, -1);
            typename Sim::ETensor E;
            E.setFlat(Dflat);
            sim.setMaterial(E);
            if (periodic) sim.applyPeriodicConditions();
            const size_t nn = sim.mesh().numNodes(), nv = sim.mesh().numVertices();
            typename Sim::VField U(nn), DU(nn), DP(nv);
            std::copy(u, u + nn * N, U.data().begin());
            std::copy(du, du + nn * N, DU.data().begin());
            std::copy(deltaP, deltaP + nv * N, DP.data().begin());
            typename Sim::SMatrix eps;
            for (size_t k = 0; k < F; ++k) eps[k] = strain[k];
            const auto a = sim.applyDeltaStiffnessMatrix(U, DP);
            std::copy(a.data().begin(), a.data().end(), dKu);
            const auto b = sim.deltaConstantStrainLoad(eps, DP);
            std::copy(b.data().begin(), b.data().end(), dload);
            const auto c = sim.deltaAverageStrainField(U, DU, DP);
            std::copy(c.data().begin(), c.data().end(), dstrain);
            if (w && dCh) {
                std::vector<typename Sim::VField> w_ij;
                for (size_t i = 0; i < F; ++i) {
                    typename Sim::VField wi(nn);
                    std::copy(w + i * nn * N, w + (i + 1) * nn * N, wi.data().begin());
                    w_ij.push_back(wi);
                }
                const auto form = PeriodicHomogenization::homogenizedElasticityTensorDiscreteDifferential(w_ij, sim);
                for (size_t v = 0; v < nv; ++v) for (size_t cc = 0; cc < N; ++cc) form(v)[cc].getFlat(dCh + (v * N + cc) * F * F);
                // the one-form applied to deltaP must agree with the contraction of its entries
                typename Sim::ETensor applied = PeriodicHomogenization::deltaHomogenizedElasticityTensor(sim, w_ij, DP), manual;
                for (size_t v = 0; v < nv; ++v) for (size_t cc = 0; cc < N; ++cc) { auto t = form(v)[cc]; t *= DP(v)[cc]; manual += t; }
                for (size_t i = 0; i < F; ++i) for (size_t j = i; j < F; ++j)
                    if (std::abs(applied.D(i, j) - manual.D(i, j)) > 1e-12 * (1.0 + std::abs(manual.D(i, j)))) throw std::runtime_error("OneForm application mismatch");
            }
        };
        if (hm->dim == 3 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 1>>>());
        else if (hm->dim == 3 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<3, 2>>>());
        else if (hm->dim == 2 && deg == 1) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 1>>>());
        else if (hm->dim == 2 && deg == 2) run(TypeTag<LinearElasticity::Simulator<LinearElasticity::Mesh<2, 2>>>());
        else throw std::runtime_error("bad dim/deg");
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}

int mfemhost_save_mesh(void *m, const char *path) {
    auto *hm = static_cast<HostMesh *>(m);
    try { MeshIO::save(path, hm->vertices, hm->elements); return 0; }
    catch (const std::exception &e) { g_err = e.what(); return -1; }
}

}  // extern "C"

// ---- element partitioning (include/MeshFEM/Partition.hh) for the multi-GPU bench and tests
namespace { thread_local Partition::LocalPart g_part; }
static int g_partitioner = -1;          // -1: from the environment
extern "C" {
int mfemhost_set_partitioner(int method) { g_partitioner = method; return 0; }
// sizes5: [nLocalElems, nLocalNodes, nNeighbors, nSharedTotal, nOwned]
int mfemhost_partition(int dim, int64_t nNodes, const double *nodes, int64_t nElems, int npe, const int32_t *elemNodes,
                       int nParts, int rank, const int64_t *dofForNode, int64_t nDofs, int64_t *sizes6) {
    try {
        // 0 = slabs (default), 1 = recursive coordinate bisection; mfemhost_set_partitioner or MESHFEM_PARTITIONER=rcb
        int method = g_partitioner;
        if (method < 0) {
            const char *env = std::getenv("MESHFEM_PARTITIONER");
            method = (env && std::string(env) == "rcb") ? 1 : 0;
        }
        auto part = method == 1 ? Partition::rcbPartition(dim, nNodes, nodes, nElems, npe, elemNodes, nParts)
                                : Partition::slabPartition(dim, nNodes, nodes, nElems, npe, elemNodes, nParts);
        g_part = Partition::extractPart(rank, nParts, nNodes, nElems, npe, elemNodes, part, dofForNode, nDofs);
        sizes6[0] = (int64_t)g_part.elems.size(); sizes6[1] = (int64_t)g_part.nodes.size();
        sizes6[2] = (int64_t)g_part.neighborRanks.size(); sizes6[3] = (int64_t)g_part.sharedLocal.size();
        int64_t owned = 0; for (auto o : g_part.owned) owned += o;
        sizes6[4] = owned;
        sizes6[5] = (int64_t)g_part.dofs.size();
        return 0;
    } catch (const std::exception &e) { g_err = e.what(); return -1; }
}
int mfemhost_partition_copy(int64_t *elems, int64_t *nodes, int32_t *elemNodesLocal, uint8_t *owned, int32_t *neighborRanks,
                            int64_t *neighborOffsets, int32_t *sharedLocal, int64_t *dofs, int64_t *dofForNodeLocal) {
    auto cp = [](auto &v, auto *dst) { if (dst && !v.empty()) std::memcpy(dst, v.data(), v.size() * sizeof(v[0])); };
    cp(g_part.elems, elems); cp(g_part.nodes, nodes); cp(g_part.elemNodes, elemNodesLocal); cp(g_part.owned, owned);
    cp(g_part.neighborRanks, neighborRanks); cp(g_part.neighborOffsets, neighborOffsets); cp(g_part.sharedLocal, sharedLocal);
    cp(g_part.dofs, dofs); cp(g_part.dofForNode, dofForNodeLocal);
    return 0;
}
}  // extern "C"

// pybind11 modules `tensors`, `sparse_matrices` and `periodic_homogenization` -- the reference's second
// public operator surface (src/python_bindings/tensors.cc:20-127, sparse_matrices.cc:30-66,
// periodic_homogenization.cc:36-172), bound over this repository's host classes so that the
// assemble-and-solve work runs on the GPU through libmfem_b200.
//
// One translation unit, three modules (MODULE_NAME selects which PYBIND11_MODULE is emitted; the build
// compiles it three times).  Differences from the reference, all at the edges:
//  * no Eigen here: matrices/vectors cross the boundary as numpy arrays;
//  * the reference passes `mesh` objects of its `mesh` module; `homogenize` / `probe` take
//    (vertices [nv x 3|2], elements [ne x (N+1)], degree) instead;
//  * SPSDSystem takes the block size of K (variables ordered blockDim*DoF + component) and exposes the
//    PCG controls (setTolerance); C / C_rhs constraint rows are not supported (SPD path only);
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/OrthotropicHomogenization.hh>
#include <MeshFEM/PeriodicHomogenization.hh>
#include <MeshFEM/SparseMatrices.hh>

#include <sstream>

namespace py = pybind11;
using NpArr = py::array_t<double, py::array::c_style | py::array::forcecast>;
using NpIdx = py::array_t<int64_t, py::array::c_style | py::array::forcecast>;

// ---------------------------------------------------------------------------------------------
#if defined(BIND_TENSORS) || defined(BIND_HOMOGENIZATION)
template <size_t N>
static NpArr tensorD(const ElasticityTensor<Real, N> &E) {
    constexpr size_t F = flatLen(N);
    NpArr D({F, F});
    auto d = D.mutable_unchecked<2>();
    for (size_t i = 0; i < F; ++i) for (size_t j = 0; j < F; ++j) d(i, j) = E.D(i, j);
    return D;
}
template <size_t N>
static ElasticityTensor<Real, N> tensorFromArray(const NpArr &D) {
    constexpr size_t F = flatLen(N);
    if (D.ndim() != 2 || (size_t)D.shape(0) != F || (size_t)D.shape(1) != F) throw std::runtime_error("expected a flat x flat matrix");
    ElasticityTensor<Real, N> E;
    E.setFlat(D.data());
    return E;
}
#endif

#ifdef BIND_TENSORS
struct ETensorEigenDecomposition {
    NpArr eigenstrains;   // flat x flat, COLUMN k = k-th eigenstrain (as the reference's `strains`)
    NpArr eigenvalues;    // ascending
};

template <size_t N>
static void bindTensors(py::module &m) {
    typedef ElasticityTensor<Real, N> ETensor;
    constexpr size_t F = flatLen(N);
    const std::string name = "ElasticityTensor" + std::to_string(N) + "D";
    auto cls = py::class_<ETensor>(m, name.c_str())
        .def(py::init<>())
        .def(py::init([](const std::string &materialFile) { return Materials::Constant<N>(materialFile).getTensor(); }), py::arg("material_file"))
        .def(py::init<Real, Real>(), py::arg("E"), py::arg("nu"))
        .def(py::init([](const NpArr &D) { return tensorFromArray<N>(D); }), py::arg("D"))
        .def("setIsotropic", &ETensor::setIsotropic, py::arg("E"), py::arg("nu"))
        .def("setIdentity", &ETensor::setIdentity)
        .def("getOrthotropicParameters", [](const ETensor &E) {
            std::vector<Real> p(N == 3 ? 9 : 4);
            if constexpr (N == 3) E.getOrthotropic3D(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
            else E.getOrthotropic2D(p[0], p[1], p[2], p[3]);
            return p; })
        .def("anisotropy", &ETensor::anisotropy)
        .def("__call__", [](const ETensor &E, size_t i, size_t j, size_t k, size_t l) {
            if ((i >= N) || (j >= N) || (k >= N) || (l >= N)) throw std::runtime_error("Index out of bounds");
            return E(i, j, k, l); })
        .def_property_readonly("D", [](const ETensor &E) { return tensorD<N>(E); })
        .def("doubleContract", [](const ETensor &E, const NpArr &smat) {
            // flattened symmetric matrix (flat) or a field of them (n x flat)
            auto one = [&](const double *in, double *out) {
                typename ETensor::SMatrix s;
                for (size_t i = 0; i < F; ++i) s[i] = in[i];
                const auto r = E.doubleContract(s);
                for (size_t i = 0; i < F; ++i) out[i] = r[i];
            };
            if (smat.ndim() == 1 && (size_t)smat.shape(0) == F) { NpArr out(F); one(smat.data(), out.mutable_data()); return out; }
            if (smat.ndim() == 2 && (size_t)smat.shape(1) == F) {
                NpArr out({(size_t)smat.shape(0), F});
                for (py::ssize_t i = 0; i < smat.shape(0); ++i) one(smat.data() + i * F, out.mutable_data() + i * F);
                return out;
            }
            throw std::runtime_error("doubleContract: expected flat or n x flat values"); }, py::arg("smat"))
        .def("computeEigenstrains", [](const ETensor &E) {
            const auto eig = E.computeEigenstrains();
            ETensorEigenDecomposition r;
            r.eigenstrains = NpArr({F, F});
            r.eigenvalues = NpArr(F);
            auto s = r.eigenstrains.mutable_unchecked<2>();
            for (size_t k = 0; k < F; ++k) {
                r.eigenvalues.mutable_data()[k] = eig.lambdas[k];
                for (size_t i = 0; i < F; ++i) s(i, k) = eig.strains[k][i];
            }
            return r; })
        .def("inverse", &ETensor::inverse)
        .def("pseudoinverse", &ETensor::inverse)
        .def("frobeniusNormSq", &ETensor::frobeniusNormSq)
        .def("quadrupleContract", [](const ETensor &E, const ETensor &other) { return E.quadrupleContract(other); }, py::arg("E"))
        .def("transform", [](const ETensor &E, const NpArr &R) {
            if (R.ndim() != 2 || (size_t)R.shape(0) != N || (size_t)R.shape(1) != N) throw std::runtime_error("transform: R must be N x N");
            Real M[N][N];
            auto r = R.template unchecked<2>();
            for (size_t i = 0; i < N; ++i) for (size_t j = 0; j < N; ++j) M[i][j] = r(i, j);
            return E.transform(M); }, py::arg("R"), "Apply a *orthogonal* change of coordinates to this tensor")
        .def("__sub__", [](const ETensor &a, const ETensor &b) { ETensor r = b; r *= -1.0; r += a; return r; })
        .def("__repr__", [](const ETensor &E) {
            std::stringstream ss;
            ss << N << "D elasticity tensor with orthotropic moduli:";
            std::vector<Real> p(N == 3 ? 9 : 4);
            if constexpr (N == 3) E.getOrthotropic3D(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], p[8]);
            else E.getOrthotropic2D(p[0], p[1], p[2], p[3]);
            for (Real v : p) ss << " " << v;
            return ss.str(); });
    if constexpr (N == 3)
        cls.def("setOrthotropic", &ETensor::setOrthotropic3D, py::arg("Ex"), py::arg("Ey"), py::arg("Ez"), py::arg("nuYX"),
                py::arg("nuZX"), py::arg("nuZY"), py::arg("muYZ"), py::arg("myZX"), py::arg("muXY"));
    else
        cls.def("setOrthotropic", &ETensor::setOrthotropic2D, py::arg("Ex"), py::arg("Ey"), py::arg("nuYX"), py::arg("muXY"));
}

PYBIND11_MODULE(tensors, m) {
    m.doc() = "Tensors used for elasticity simulations (ElasticityTensor2D / ElasticityTensor3D)";
    py::module detail = m.def_submodule("detail");
    py::class_<ETensorEigenDecomposition>(detail, "ETensorEigenDecomposition")
        .def_readonly("eigenstrains", &ETensorEigenDecomposition::eigenstrains)
        .def_readonly("eigenvalues", &ETensorEigenDecomposition::eigenvalues);
    bindTensors<2>(m);
    bindTensors<3>(m);
}
#endif

// ---------------------------------------------------------------------------------------------
#ifdef BIND_SPARSE_MATRICES
typedef TripletMatrix<Triplet<Real>> TMatrix;

PYBIND11_MODULE(sparse_matrices, m) {
    m.doc() = "Triplet matrices and the (GPU) SPSD system solver";
    py::class_<TMatrix>(m, "TripletMatrix")
        .def(py::init<size_t, size_t>(), py::arg("m") = 0, py::arg("n") = 0)
        .def_readwrite("m", &TMatrix::m)
        .def_readwrite("n", &TMatrix::n)
        .def("nnz", &TMatrix::nnz)
        .def("reserve", &TMatrix::reserve)
        .def("addNZ", &TMatrix::addNZ, py::arg("i"), py::arg("j"), py::arg("v"))
        .def("addNZs", [](TMatrix &A, const NpIdx &i, const NpIdx &j, const NpArr &v) {
            if (i.size() != j.size() || i.size() != v.size()) throw std::runtime_error("addNZs: size mismatch");
            for (py::ssize_t k = 0; k < i.size(); ++k) A.addNZ((size_t)i.data()[k], (size_t)j.data()[k], v.data()[k]); },
            "Append many triplets at once (numpy arrays)")
        .def("sumRepeated", &TMatrix::sumRepeated, "Compress the matrix by summing together all the entries with the same row, column index")
        .def("apply", [](const TMatrix &A, const NpArr &x) {
            if ((size_t)x.size() != A.n) throw std::runtime_error("apply: size mismatch");
            NpArr y(A.m);
            double *yp = y.mutable_data();
            std::fill(yp, yp + A.m, 0.0);
            for (const auto &t : A.nz) yp[t.i] += t.v * x.data()[t.j];
            return y; }, "Apply the sparse matrix to a vector")
        .def("triplets", [](const TMatrix &A) {
            NpIdx i(A.nnz()), j(A.nnz());
            NpArr v(A.nnz());
            for (size_t k = 0; k < A.nnz(); ++k) { i.mutable_data()[k] = (int64_t)A.nz[k].i; j.mutable_data()[k] = (int64_t)A.nz[k].j; v.mutable_data()[k] = A.nz[k].v; }
            return py::make_tuple(i, j, v); })
        .def("compressedColumn", [](const TMatrix &A) {
            py::object csc = py::module::import("scipy.sparse").attr("csc_matrix");
            NpIdx i(A.nnz()), j(A.nnz());
            NpArr v(A.nnz());
            for (size_t k = 0; k < A.nnz(); ++k) { i.mutable_data()[k] = (int64_t)A.nz[k].i; j.mutable_data()[k] = (int64_t)A.nz[k].j; v.mutable_data()[k] = A.nz[k].v; }
            return csc(py::make_tuple(v, py::make_tuple(i, j)), py::arg("shape") = py::make_tuple(A.m, A.n)); })
        .def("dumpBinary", &TMatrix::dumpBinary)
        .def("readBinary", &TMatrix::readBinary);

    typedef SPSDSystem<Real> Sys;
    py::class_<Sys>(m, "SPSDSystem", "A (constrained) SPSD system that can be solved for several different right-hand sides.")
        .def(py::init([](const TMatrix &K, int blockDim, int device) { return new Sys(K, blockDim, device); }), py::arg("K"),
             py::arg("blockDim") = 3, py::arg("device") = 0)
        .def("fixVariables", [](Sys &s, const std::vector<size_t> &vars, const std::vector<double> &vals) { s.fixVariables(vars, vals); },
             py::arg("fixedVars"), py::arg("fixedVarValues") = std::vector<double>())
        .def("setTolerance", &Sys::setTolerance, py::arg("rtol"), py::arg("maxIters") = 200000)
        .def("solve", [](Sys &s, const NpArr &b) {
            std::vector<double> f(b.data(), b.data() + b.size()), u;
            s.solve(f, u);
            NpArr out(u.size());
            std::copy(u.begin(), u.end(), out.mutable_data());
            return out; })
        .def("lastSolveInfo", [](const Sys &s) {
            const auto &i = s.lastSolveInfo();
            py::dict d;
            d["iterations"] = i.iterations; d["converged"] = (bool)i.converged; d["rel_residual"] = i.rel_residual; d["seconds"] = i.seconds;
            return d; });
}
#endif

// ---------------------------------------------------------------------------------------------
#ifdef BIND_HOMOGENIZATION
struct HomogenizationResult {
    NpArr Ch;                          // flat x flat
    std::vector<NpArr> w_ij;           // numNodes x N each
    std::vector<NpArr> strain_w_ij;    // numElements x flat each
    int N = 0;
};

static void toIO(const NpArr &V, const NpIdx &F, std::vector<MeshIO::IOVertex> &verts, std::vector<MeshIO::IOElement> &elems) {
    if (V.ndim() != 2 || (V.shape(1) != 2 && V.shape(1) != 3)) throw std::runtime_error("vertices must be nv x 2 or nv x 3");
    if (F.ndim() != 2 || (F.shape(1) != 3 && F.shape(1) != 4)) throw std::runtime_error("elements must be ne x 3 (triangles) or ne x 4 (tets)");
    for (py::ssize_t i = 0; i < V.shape(0); ++i)
        verts.emplace_back(V.at(i, 0), V.at(i, 1), V.shape(1) == 3 ? V.at(i, 2) : 0.0);
    for (py::ssize_t e = 0; e < F.shape(0); ++e) {
        elems.emplace_back((size_t)F.shape(1));
        for (py::ssize_t c = 0; c < F.shape(1); ++c) elems.back()[(size_t)c] = (size_t)F.at(e, c);
    }
}

template <size_t N, size_t Deg>
static HomogenizationResult runHomogenization(const std::vector<MeshIO::IOVertex> &verts, const std::vector<MeshIO::IOElement> &elems,
                                              const NpArr &Cbase, bool orthotropicCell, const std::string &manualPeriodicVerticesFile, bool center,
                                              bool ignoreMismatch, int device, double rtol) {
    typedef LinearElasticity::Simulator<LinearElasticity::Mesh<N, Deg>> Sim;
    constexpr size_t F = flatLen(N);
    Sim sim(elems, verts, device);
    sim.setMaterial(tensorFromArray<N>(Cbase));
    sim.setSolverTolerance(rtol);
    std::vector<typename Sim::VField> w_ij;
    HomogenizationResult r;
    r.N = (int)N;
    if (orthotropicCell) {
        PeriodicHomogenization::Orthotropic::solveCellProblems(w_ij, sim);
        r.Ch = tensorD<N>(PeriodicHomogenization::Orthotropic::homogenizedElasticityTensorDisplacementForm(w_ij, sim));
    } else {
        std::unique_ptr<PeriodicCondition<N>> pc;
        if (!manualPeriodicVerticesFile.empty()) pc.reset(new PeriodicCondition<N>(sim.mesh(), manualPeriodicVerticesFile));
        PeriodicHomogenization::solveCellProblems(w_ij, sim, 1e-7, ignoreMismatch, std::move(pc));
        r.Ch = tensorD<N>(PeriodicHomogenization::homogenizedElasticityTensorDisplacementForm(w_ij, sim));
    }
    if (center)
        for (auto &w : w_ij) {
            VectorND<N> total;
            for (size_t i = 0; i < w.domainSize(); ++i) total += w(i);
            total *= 1.0 / w.domainSize();
            for (size_t i = 0; i < w.domainSize(); ++i) w.set(i, w(i) - total);
        }
    for (const auto &w : w_ij) {
        NpArr a({w.domainSize(), N});
        std::copy(w.data().begin(), w.data().end(), a.mutable_data());
        r.w_ij.push_back(a);
        const auto s = sim.averageStrainField(w);
        NpArr sa({s.domainSize(), F});
        std::copy(s.data().begin(), s.data().end(), sa.mutable_data());
        r.strain_w_ij.push_back(sa);
    }
    return r;
}

static HomogenizationResult homogenize(const NpArr &V, const NpIdx &F, const NpArr &Cbase, int degree, bool orthotropicCell,
                                       const std::string &manualPeriodicVerticesFile, bool center, bool ignoreMismatch,
                                       int device, double rtol) {
    std::vector<MeshIO::IOVertex> verts;
    std::vector<MeshIO::IOElement> elems;
    toIO(V, F, verts, elems);
    const bool tet = F.shape(1) == 4;
    if (degree != 1 && degree != 2) throw std::runtime_error("degree must be 1 or 2");
    if (tet) return degree == 2 ? runHomogenization<3, 2>(verts, elems, Cbase, orthotropicCell, manualPeriodicVerticesFile, center, ignoreMismatch, device, rtol)
                                : runHomogenization<3, 1>(verts, elems, Cbase, orthotropicCell, manualPeriodicVerticesFile, center, ignoreMismatch, device, rtol);
    return degree == 2 ? runHomogenization<2, 2>(verts, elems, Cbase, orthotropicCell, manualPeriodicVerticesFile, center, ignoreMismatch, device, rtol)
                       : runHomogenization<2, 1>(verts, elems, Cbase, orthotropicCell, manualPeriodicVerticesFile, center, ignoreMismatch, device, rtol);
}

// getProbeResult (periodic_homogenization.cc:92-143): displacement and strain of the cell under a macroscopic
// strain: u = sum_i doubler_i e_i w_i, translated so that each displacement component averages to zero over
// the boundary nodes on the matching minimum face, plus the linear term e x.
template <size_t N, size_t Deg>
static py::tuple probeImpl(const std::vector<MeshIO::IOVertex> &verts, const std::vector<MeshIO::IOElement> &elems,
                           const HomogenizationResult &hr, const NpArr &macroStrain) {
    constexpr size_t F = flatLen(N);
    if ((size_t)macroStrain.size() != F) throw std::runtime_error("macroStrain must be a flattened symmetric matrix");
    FEMMesh<N, Deg, VectorND<N>> mesh(elems, verts);
    const size_t nn = mesh.numNodes(), ne = mesh.numElements();
    if (hr.w_ij.size() != F || (size_t)hr.w_ij[0].shape(0) != nn) throw std::runtime_error("homogenization result does not match the mesh");
    NpArr u({nn, N}), strain({ne, F});
    std::fill(u.mutable_data(), u.mutable_data() + nn * N, 0.0);
    std::fill(strain.mutable_data(), strain.mutable_data() + ne * F, 0.0);
    const double *e = macroStrain.data();
    for (size_t i = 0; i < F; ++i) {
        const double c = (i < N ? 1.0 : 2.0) * e[i];
        for (size_t k = 0; k < nn * N; ++k) u.mutable_data()[k] += c * hr.w_ij[i].data()[k];
        for (size_t k = 0; k < ne * F; ++k) strain.mutable_data()[k] += c * hr.strain_w_ij[i].data()[k];
    }
    const auto &bbox = mesh.boundingBox();
    VectorND<N> translation, numAveraged;
    for (size_t bn = 0; bn < mesh.numBoundaryNodes(); ++bn) {
        const size_t n = (size_t)mesh.volumeNodeForBoundaryNode(bn);
        const auto p = mesh.nodePosition(n);
        for (size_t d = 0; d < N; ++d)
            if (std::abs(p[d] - bbox.minCorner[d]) < 1e-9) { translation[d] += u.data()[n * N + d]; numAveraged[d] += 1.0; }
    }
    SymmetricMatrixValue<Real, N> E;
    for (size_t i = 0; i < F; ++i) E[i] = e[i];
    for (size_t n = 0; n < nn; ++n) {
        const auto lin = E.contract(mesh.nodePosition(n));
        for (size_t d = 0; d < N; ++d) u.mutable_data()[n * N + d] += lin[d] - translation[d] / numAveraged[d];
    }
    for (size_t el = 0; el < ne; ++el) for (size_t i = 0; i < F; ++i) strain.mutable_data()[el * F + i] += e[i];
    return py::make_tuple(u, strain);
}

static py::tuple probe(const NpArr &V, const NpIdx &F, int degree, const HomogenizationResult &hr, const NpArr &macroStrain) {
    std::vector<MeshIO::IOVertex> verts;
    std::vector<MeshIO::IOElement> elems;
    toIO(V, F, verts, elems);
    const bool tet = F.shape(1) == 4;
    if (tet) return degree == 2 ? probeImpl<3, 2>(verts, elems, hr, macroStrain) : probeImpl<3, 1>(verts, elems, hr, macroStrain);
    return degree == 2 ? probeImpl<2, 2>(verts, elems, hr, macroStrain) : probeImpl<2, 1>(verts, elems, hr, macroStrain);
}

// Shape sensitivity of the homogenized tensor (not in the reference's binding; its C++ callers use
// PeriodicHomogenization.hh:383-563): cell problems on the GPU, then the exact discrete differential dCh[v, c] on the
// host and -- when a per-vertex perturbation is given -- the change of the fluctuation displacements (one more
// batched solve on the GPU) and of Ch.
template <size_t N, size_t Deg>
static py::dict shapeDerivativeImpl(const std::vector<MeshIO::IOVertex> &verts, const std::vector<MeshIO::IOElement> &elems,
                                    const NpArr &Cbase, const py::object &deltaP, int device, double rtol) {
    typedef LinearElasticity::Simulator<LinearElasticity::Mesh<N, Deg>> Sim;
    constexpr size_t F = flatLen(N);
    Sim sim(elems, verts, device);
    sim.setMaterial(tensorFromArray<N>(Cbase));
    sim.setSolverTolerance(rtol);
    std::vector<typename Sim::VField> w_ij;
    PeriodicHomogenization::solveCellProblems(w_ij, sim);
    py::dict out;
    out["Ch"] = tensorD<N>(PeriodicHomogenization::homogenizedElasticityTensorDisplacementForm(w_ij, sim));
    const auto form = PeriodicHomogenization::homogenizedElasticityTensorDiscreteDifferential(w_ij, sim);
    const size_t nv = sim.mesh().numVertices(), nn = sim.mesh().numNodes();
    NpArr dCh({nv, N, F, F});
    for (size_t v = 0; v < nv; ++v) for (size_t c = 0; c < N; ++c) form(v)[c].getFlat(dCh.mutable_data() + (v * N + c) * F * F);
    out["dCh"] = dCh;
    NpArr w({F, nn, N});
    for (size_t i = 0; i < F; ++i) std::copy(w_ij[i].data().begin(), w_ij[i].data().end(), w.mutable_data() + i * nn * N);
    out["w_ij"] = w;
    if (!deltaP.is_none()) {
        const NpArr dp = deltaP.cast<NpArr>();
        if (dp.ndim() != 2 || (size_t)dp.shape(0) != nv || (size_t)dp.shape(1) != N) throw std::runtime_error("deltaP must be numVertices x N");
        typename Sim::VField DP(nv);
        std::copy(dp.data(), dp.data() + nv * N, DP.data().begin());
        const auto dw = PeriodicHomogenization::deltaFluctuationDisplacements(sim, w_ij, DP);
        NpArr dwa({F, nn, N});
        for (size_t i = 0; i < F; ++i) std::copy(dw[i].data().begin(), dw[i].data().end(), dwa.mutable_data() + i * nn * N);
        out["delta_w_ij"] = dwa;
        out["delta_Ch"] = tensorD<N>(form[DP]);
    }
    return out;
}

static py::dict shapeDerivative(const NpArr &V, const NpIdx &F, const NpArr &Cbase, int degree, const py::object &deltaP, int device, double rtol) {
    std::vector<MeshIO::IOVertex> verts;
    std::vector<MeshIO::IOElement> elems;
    toIO(V, F, verts, elems);
    const bool tet = F.shape(1) == 4;
    if (degree != 1 && degree != 2) throw std::runtime_error("degree must be 1 or 2");
    if (tet) return degree == 2 ? shapeDerivativeImpl<3, 2>(verts, elems, Cbase, deltaP, device, rtol) : shapeDerivativeImpl<3, 1>(verts, elems, Cbase, deltaP, device, rtol);
    return degree == 2 ? shapeDerivativeImpl<2, 2>(verts, elems, Cbase, deltaP, device, rtol) : shapeDerivativeImpl<2, 1>(verts, elems, Cbase, deltaP, device, rtol);
}

PYBIND11_MODULE(periodic_homogenization, m) {
    m.doc() = "Periodic homogenization of a base cell on the GPU";
    py::module detail = m.def_submodule("detail");
    py::class_<HomogenizationResult>(detail, "HomogenizationResult")
        .def_readonly("Ch", &HomogenizationResult::Ch)
        .def_readonly("w_ij", &HomogenizationResult::w_ij)
        .def_readonly("strain_w_ij", &HomogenizationResult::strain_w_ij);
    m.def("homogenize", &homogenize, py::arg("vertices"), py::arg("elements"), py::arg("Cbase"), py::arg("degree") = 2,
          py::arg("orthotropicCell") = false, py::arg("manualPeriodicVerticesFile") = std::string(),
          py::arg("centerFluctuationDisplacements") = true, py::arg("ignorePeriodicMismatch") = false, py::arg("device") = 0,
          py::arg("rtol") = 1e-10);
    m.def("probe", &probe, py::arg("vertices"), py::arg("elements"), py::arg("degree"), py::arg("homogenizationResult"),
          py::arg("macroStrain"));
    m.def("shapeDerivative", &shapeDerivative, py::arg("vertices"), py::arg("elements"), py::arg("Cbase"), py::arg("degree") = 2,
          py::arg("deltaP") = py::none(), py::arg("device") = 0, py::arg("rtol") = 1e-10,
          "Ch, its exact discrete differential dCh[v, c] with respect to the vertex positions, the fluctuation displacements and, for a "
          "per-vertex perturbation deltaP, delta_w_ij and delta_Ch");
}
#endif

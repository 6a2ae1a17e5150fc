// TripletMatrix (COO, the reference's dump/interchange format) and SPSDSystem, the solver
// policy seam of the reference (SparseMatrices.hh:2321-2716), re-pointed at the GPU: where the
// reference instantiates SPSDSystem<Real, UmfpackFactorizer, CholmodFactorizer>, this class
// forwards setConstrained/fixVariables/solve to the C ABI (include/mfem_b200.h) -- the matrix
// lives in HBM as block-CSR and the "factorizer" is the block-Jacobi PCG.
#ifndef MESHFEM_B200_SPARSEMATRICES_HH
#define MESHFEM_B200_SPARSEMATRICES_HH
#include <MeshFEM/GlobalBenchmark.hh>
#include <MeshFEM/Types.hh>
#include <mfem_b200.h>

#include <algorithm>
#include <fstream>
#include <stdexcept>
#include <vector>

template <typename _Real>
struct Triplet {
    size_t i, j;
    _Real v;
    Triplet(size_t ii, size_t jj, _Real vv) : i(ii), j(jj), v(vv) {}
    Triplet() : i(0), j(0), v(0) {}
};

template <class _Triplet = Triplet<Real>>
struct TripletMatrix {
    size_t m = 0, n = 0;
    std::vector<_Triplet> nz;
    TripletMatrix(size_t mm = 0, size_t nn = 0) : m(mm), n(nn) {}
    void init(size_t mm = 0, size_t nn = 0) { m = mm; n = nn; nz.clear(); }
    void reserve(size_t nnz) { nz.reserve(nnz); }
    size_t nnz() const { return nz.size(); }
    void addNZ(size_t i, size_t j, Real v) { nz.emplace_back(i, j, v); }
    // sort column-major, sum duplicates, prune v*v <= 0 (SparseMatrices.hh:280-374)
    void sumRepeated() {
        std::stable_sort(nz.begin(), nz.end(), [](const _Triplet &a, const _Triplet &b) { return a.j != b.j ? a.j < b.j : a.i < b.i; });
        size_t out = 0;
        for (size_t k = 0; k < nz.size();) {
            _Triplet t = nz[k++];
            while (k < nz.size() && nz[k].i == t.i && nz[k].j == t.j) t.v += nz[k++].v;
            if (t.v * t.v > 0) nz[out++] = t;
        }
        nz.resize(out);
    }
    // uint64 nnz | uint64 rows[] | uint64 cols[] | double vals[]  (SparseMatrices.hh:623-645)
    void dumpBinary(const std::string &path) const {
        std::ofstream os(path, std::ios::binary);
        if (!os.is_open()) throw std::runtime_error("Couldn't open output matrix file " + path);
        const uint64_t cnt = nz.size();
        os.write((const char *)&cnt, 8);
        for (const auto &t : nz) { uint64_t v = t.i; os.write((const char *)&v, 8); }
        for (const auto &t : nz) { uint64_t v = t.j; os.write((const char *)&v, 8); }
        for (const auto &t : nz) { double v = t.v; os.write((const char *)&v, 8); }
    }
    void readBinary(const std::string &path) {
        std::ifstream is(path, std::ios::binary);
        if (!is.is_open()) throw std::runtime_error("Couldn't open input matrix file " + path);
        uint64_t cnt;
        is.read((char *)&cnt, 8);
        std::vector<uint64_t> r(cnt), c(cnt);
        std::vector<double> v(cnt);
        is.read((char *)r.data(), 8 * cnt); is.read((char *)c.data(), 8 * cnt); is.read((char *)v.data(), 8 * cnt);
        nz.clear();
        for (uint64_t k = 0; k < cnt; ++k) { nz.emplace_back(r[k], c[k], v[k]); m = std::max<size_t>(m, r[k] + 1); n = std::max<size_t>(n, c[k] + 1); }
    }
};

inline void mfemCheck(mfem_b200_handle h, int status) {
    if (status != MFEM_B200_OK) throw std::runtime_error(mfem_b200_last_error(h));
}

// Solver policy facade over one mfem_b200 handle.
template <typename _Real>
class SPSDSystem {
public:
    SPSDSystem() {}
    // The reference's constructor (SparseMatrices.hh:2325-2348, set(K)): a symmetric positive
    // (semi-)definite matrix assembled by the caller, upper triangle in triplet form, variables ordered
    // blockDim*DoF + component.  The triplets are summed and laid out on the device; fixVariables / solve
    // then work as for a mesh-assembled system.
    template <class _TMatrix>
    explicit SPSDSystem(const _TMatrix &K, int blockDim = 3, int device = 0) : m_device(device) { set(K, blockDim); }
    template <class _TMatrix>
    void set(const _TMatrix &K, int blockDim = 3) {
        if (K.m != K.n) throw std::runtime_error("SPSDSystem: K must be square");
        std::vector<int64_t> I(K.nnz()), J(K.nnz());
        std::vector<_Real> V(K.nnz());
        bool upper = true;
        for (size_t k = 0; k < K.nnz(); ++k) {
            I[k] = (int64_t)K.nz[k].i; J[k] = (int64_t)K.nz[k].j; V[k] = K.nz[k].v;
            upper = upper && K.nz[k].i <= K.nz[k].j;
        }
        mfemCheck(handle(), mfem_b200_set_matrix_triplets(handle(), blockDim, (int64_t)K.m, (int64_t)K.nnz(), I.data(), J.data(),
                                                          V.data(), upper ? 1 : 0));
        m_numVars = K.m;
        m_isSet = true;
    }
    SPSDSystem(const SPSDSystem &) = delete;
    SPSDSystem &operator=(const SPSDSystem &) = delete;
    ~SPSDSystem() { if (m_handle) mfem_b200_destroy(m_handle); }

    mfem_b200_handle handle() {
        if (!m_handle) {
            mfem_b200_handle h = nullptr;
            if (mfem_b200_create(m_device, &h) != MFEM_B200_OK) throw std::runtime_error(mfem_b200_last_error(nullptr));
            m_handle = h;
        }
        return m_handle;
    }
    void setDevice(int d) { m_device = d; }
    bool isSet() const { return m_isSet; }
    void clear() {
        m_isSet = false;
        if (m_handle) mfem_b200_clear_fixed_variables(m_handle);
    }
    // "setConstrained": the stiffness matrix is assembled on the device from the mesh + material
    // previously given to the handle; C (Lagrange rows) must be empty on this path.
    void setAssembled(size_t numVars) {
        BENCHMARK_START_TIMER_SECTION("Set System");
        mfemCheck(handle(), mfem_b200_assemble(handle()));
        BENCHMARK_ADD_DEVICE_SECONDS("Assemble System (device)", mfem_b200_get_timer(handle(), "Assemble System"));
        BENCHMARK_ADD_DEVICE_SECONDS("Pattern (device)", mfem_b200_get_timer(handle(), "Pattern"));
        m_numVars = numVars;
        m_isSet = true;
        BENCHMARK_STOP_TIMER_SECTION("Set System");
    }
    void fixVariables(const std::vector<size_t> &fixedVars, const std::vector<_Real> &fixedVarValues = std::vector<_Real>()) {
        BENCHMARK_SCOPED_TIMER_SECTION timer("fixVariables");
        if (fixedVars.empty()) return;
        if (!fixedVarValues.empty() && fixedVarValues.size() != fixedVars.size()) throw std::runtime_error("Incorrect number of fixedVarValues");
        std::vector<int64_t> v(fixedVars.begin(), fixedVars.end());
        mfemCheck(handle(), mfem_b200_fix_variables(handle(), (int64_t)v.size(), v.data(), fixedVarValues.empty() ? nullptr : fixedVarValues.data()));
    }
    template <class _Vec, class _SolnVec>
    void solve(const _Vec &f, _SolnVec &u) {
        if (!isSet()) throw std::runtime_error("No system to solve");
        if (f.size() != m_numVars) throw std::runtime_error("Bad RHS");
        u.resize(f.size());
        mfem_b200_solve_info info;
        mfemCheck(handle(), mfem_b200_solve(handle(), 1, f.data(), u.data(), m_rtol, m_maxIters, &info));
        m_lastInfo = info;
        BENCHMARK_ADD_DEVICE_SECONDS("PCG (device)", info.seconds);
    }
    // Several right-hand sides against the same system -- the reference factorises once and back-solves
    // per right-hand side; here flatLen(N) right-hand sides (the cell problems) run as ONE batched PCG
    // whose SpMM streams the matrix once for all of them (csrc/solver_multi.inl).
    template <class _Vec, class _SolnVec>
    void solveMultiple(const std::vector<_Vec> &fs, std::vector<_SolnVec> &us) {
        if (!isSet()) throw std::runtime_error("No system to solve");
        const size_t nrhs = fs.size();
        std::vector<_Real> f(nrhs * m_numVars), u(nrhs * m_numVars);
        for (size_t k = 0; k < nrhs; ++k) {
            if (fs[k].size() != m_numVars) throw std::runtime_error("Bad RHS");
            std::copy(fs[k].begin(), fs[k].end(), f.begin() + k * m_numVars);
        }
        std::vector<mfem_b200_solve_info> info(nrhs);
        mfemCheck(handle(), mfem_b200_solve(handle(), (int)nrhs, f.data(), u.data(), m_rtol, m_maxIters, info.data()));
        us.resize(nrhs);
        double seconds = 0;
        for (size_t k = 0; k < nrhs; ++k) {
            us[k].assign(u.begin() + k * m_numVars, u.begin() + (k + 1) * m_numVars);
            seconds += info[k].seconds;
        }
        m_lastInfo = info.back();
        BENCHMARK_ADD_DEVICE_SECONDS("PCG (device)", seconds);
    }
    void setTolerance(double rtol, int maxIters) { m_rtol = rtol; m_maxIters = maxIters; }
    // solver options of the C ABI (include/mfem_b200.h: "coarse_aggregates", "coarse_fine_nodes", ...)
    void setOption(const char *name, long long value) { mfemCheck(handle(), mfem_b200_set_option(handle(), name, (int64_t)value)); }
    const mfem_b200_solve_info &lastSolveInfo() const { return m_lastInfo; }
    void setEconomyMode(bool) {}
    void sumAndDumpUpper(const std::string &path) { mfemCheck(handle(), mfem_b200_dump_upper_triplets(handle(), path.c_str())); }

private:
    mfem_b200_handle m_handle = nullptr;
    int m_device = 0;
    bool m_isSet = false;
    size_t m_numVars = 0;
    double m_rtol = 1e-10;
    int m_maxIters = 200000;
    mfem_b200_solve_info m_lastInfo{};
};
#endif

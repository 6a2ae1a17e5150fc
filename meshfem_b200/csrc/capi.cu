// C ABI of libmfem_b200 (include/mfem_b200.h): argument checking, host<->device staging,
// exception -> status translation.  No compute lives here.
#include <cstring>
#include <fstream>

#include <algorithm>

#include "core.cuh"

using namespace mfem;

namespace mfem {
void comm_destroy(mfem_b200_ctx *c);   // comm.cu
}

#define API_BEGIN(h)                                   \
    if (!(h)) return MFEM_B200_ERR_INVALID;            \
    try {                                              \
        MFEM_CUDA(cudaSetDevice((h)->device));
#define API_END(h)                                     \
        return MFEM_B200_OK;                           \
    } catch (const CudaError &e) {                     \
        (h)->err = e.what();                           \
        return e.status;                               \
    } catch (const std::exception &e) {                \
        (h)->err = e.what();                           \
        return MFEM_B200_ERR_INVALID;                  \
    }

static thread_local std::string g_createError;

// ---------------------------------------------------------------------------------------------------------
// caching pool behind DevBuf (core.cuh)
#include <mutex>
#include <unordered_map>
namespace mfem {
namespace {
struct PoolBlock { void *p; size_t bytes; int device; };
std::mutex g_poolMutex;
std::vector<PoolBlock> g_poolFree;                              // cached blocks
std::unordered_map<void *, PoolBlock> g_poolLive;               // blocks handed out
bool pool_enabled() {
    static const bool on = [] { const char *e = getenv("MFEM_B200_POOL"); return !(e && e[0] == '0'); }();
    return on;
}
void pool_release_locked(int device /* -1: all */) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (size_t k = 0; k < g_poolFree.size();) {
        if (device < 0 || g_poolFree[k].device == device) {
            cudaSetDevice(g_poolFree[k].device);
            cudaFree(g_poolFree[k].p);
            g_poolFree[k] = g_poolFree.back();
            g_poolFree.pop_back();
        } else ++k;
    }
    cudaSetDevice(cur);
}
}  // namespace

void *pool_alloc(size_t bytes) {
    int dev = 0;
    MFEM_CUDA(cudaGetDevice(&dev));
    const size_t want = (bytes + 511) & ~size_t(511);
    std::lock_guard<std::mutex> lock(g_poolMutex);
    if (pool_enabled()) {
        // best fit among the cached blocks of this device: at least `want`, at most 12.5 % + 1 MiB larger
        size_t best = g_poolFree.size();
        for (size_t k = 0; k < g_poolFree.size(); ++k) {
            const PoolBlock &b = g_poolFree[k];
            if (b.device != dev || b.bytes < want || b.bytes > want + want / 8 + (1u << 20)) continue;
            if (best == g_poolFree.size() || b.bytes < g_poolFree[best].bytes) best = k;
        }
        if (best != g_poolFree.size()) {
            PoolBlock b = g_poolFree[best];
            g_poolFree[best] = g_poolFree.back();
            g_poolFree.pop_back();
            g_poolLive[b.p] = b;
            return b.p;
        }
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, want);
    if (e != cudaSuccess) {                      // out of memory: give the cached blocks back and try once more
        cudaGetLastError();
        pool_release_locked(dev);
        e = cudaMalloc(&p, want);
    }
    if (e != cudaSuccess)
        throw CudaError(MFEM_B200_ERR_CUDA, std::string("cudaMalloc of ") + std::to_string(want) + " bytes failed: " + cudaGetErrorString(e));
    g_poolLive[p] = PoolBlock{p, want, dev};
    return p;
}

void pool_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> lock(g_poolMutex);
    auto it = g_poolLive.find(p);
    if (it == g_poolLive.end()) { cudaFree(p); return; }
    PoolBlock b = it->second;
    g_poolLive.erase(it);
    if (pool_enabled()) {
        // cudaFree synchronises the device, and code that drops a temporary while its last kernel is still in flight
        // relies on that; keep the guarantee (a block may be handed to a handle on another stream next)
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != b.device) cudaSetDevice(b.device);
        cudaDeviceSynchronize();
        if (cur != b.device) cudaSetDevice(cur);
        g_poolFree.push_back(b);
    } else {
        int cur = 0;
        cudaGetDevice(&cur);
        if (cur != b.device) cudaSetDevice(b.device);
        cudaFree(b.p);
        if (cur != b.device) cudaSetDevice(cur);
    }
}

void pool_release_all() {
    std::lock_guard<std::mutex> lock(g_poolMutex);
    pool_release_locked(-1);
}
}  // namespace mfem

extern "C" {

int mfem_b200_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return MFEM_B200_ERR_CUDA;
    return n;
}

int mfem_b200_create(int device, mfem_b200_handle *out) {
    if (!out) return MFEM_B200_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        g_createError = std::string("no usable CUDA device (") + cudaGetErrorString(e) +
                        "); this library has no CPU fallback";
        return MFEM_B200_ERR_CUDA;
    }
    if (device < 0 || device >= n) { g_createError = "device index out of range"; return MFEM_B200_ERR_INVALID; }
    if (cudaSetDevice(device) != cudaSuccess) { g_createError = "cudaSetDevice failed"; return MFEM_B200_ERR_CUDA; }
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, device);
    if (prop.major < 10) {
        g_createError = "device compute capability " + std::to_string(prop.major) + "." + std::to_string(prop.minor) +
                        " < 10.0: this library is built for sm_100a only";
        return MFEM_B200_ERR_CUDA;
    }
    auto *c = new mfem_b200_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        g_createError = "cudaStreamCreate failed";
        return MFEM_B200_ERR_CUDA;
    }
    *out = c;
    return MFEM_B200_OK;
}

int mfem_b200_destroy(mfem_b200_handle h) {
    if (!h) return MFEM_B200_ERR_INVALID;
    cudaSetDevice(h->device);
    comm_destroy(h);
    free_work_multi(h);
    free_coarse_space(h);
    if (h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    delete h;
    return MFEM_B200_OK;
}

const char *mfem_b200_last_error(mfem_b200_handle h) {
    if (!h) return g_createError.c_str();
    return h->err.c_str();
}

int mfem_b200_set_option(mfem_b200_handle h, const char *name, int64_t value) {
    API_BEGIN(h)
    MFEM_REQUIRE(name, MFEM_B200_ERR_INVALID, "null option name");
    const std::string n(name);
    if (n == "reorder") {
        MFEM_REQUIRE(h->nElems == 0, MFEM_B200_ERR_INVALID, "option 'reorder' must be set before set_mesh");
        h->opt_reorder = value != 0;
    } else if (n == "assembly") {
        MFEM_REQUIRE(value >= 0 && value <= 2, MFEM_B200_ERR_INVALID, "assembly must be 0 (block-owner), 1 (coloured) or 2 (owner-gather)");
        h->opt_assembly = (int)value;
    } else if (n == "spmm_kernel") {
        h->opt_spmm_kernel = (int)value;
    } else if (n == "batch_rhs") {
        h->opt_batch_rhs = value != 0;
    } else if (n == "graph") {
        h->opt_graph = value != 0;
    } else if (n == "spmv_kernel") {
        MFEM_REQUIRE(value >= 0 && value <= 6, MFEM_B200_ERR_INVALID, "spmv_kernel must be 0 (auto), 1 (direct loads), 2 (TMA ring), 3 (index-pipelined), 4 (symmetric), 5 (128-bit load timing probe: wrong results) or 6 (mfem_b200_spmv evaluates the mesh-based operator instead of the stored matrix)");
        h->opt_spmv_kernel = (int)value;
    } else if (n == "matrix_free") {
        MFEM_REQUIRE(value >= -1 && value <= 1, MFEM_B200_ERR_INVALID,
                     "matrix_free must be -1 (automatic: 3D quadratic elements), 0 (stored-matrix SpMV) or 1 (mesh-based operator whenever a mesh is set)");
        h->opt_matrix_free = (int)value;
    } else if (n == "mf_slot_pad") {
        h->opt_mf_slot_pad = value != 0;
    } else if (n == "mf_gather_lanes") {
        MFEM_REQUIRE(value == 0 || value == 1 || value == 4 || value == 8, MFEM_B200_ERR_INVALID, "mf_gather_lanes must be 0 (automatic), 1, 4 or 8");
        h->opt_mf_gather_lanes = (int)value;
    } else if (n == "mf_elem_order") {
        h->opt_mf_elem_order = value != 0;
    } else if (n == "mf_gather_policy") {
        MFEM_REQUIRE(value >= 0 && value <= 3, MFEM_B200_ERR_INVALID, "mf_gather_policy must be 0 (evict_first), 1 (evict_last), 2 (evict_normal) or 3 (evict_last + L1 allocation)");
        h->opt_mf_gather_policy = (int)value;
    } else if (n == "mf_chunked") {
        h->opt_mf_chunked = value != 0;
    } else if (n == "mf_chunk_elems") {
        MFEM_REQUIRE(value == 32 || value == 64 || value == 128, MFEM_B200_ERR_INVALID, "mf_chunk_elems must be 32, 64 or 128");
        h->opt_mf_chunk_elems = (int)value;
    } else if (n == "mf_chunk_warps") {
        MFEM_REQUIRE(value == 12 || value == 16 || value == 20, MFEM_B200_ERR_INVALID, "mf_chunk_warps must be 12, 16 or 20");
        h->opt_mf_chunk_warps = (int)value;
    } else if (n == "coarse_aggregates") {
        MFEM_REQUIRE(value >= -1 && value <= 5461, MFEM_B200_ERR_INVALID,
                     "coarse_aggregates must be -1 (automatic), 0 (block-Jacobi only) or 1 .. 5461 large aggregates");
        h->opt_coarse = (int)value;
        h->precondValid = false;
    } else if (n == "coarse_fine_nodes") {
        MFEM_REQUIRE(value >= 0 && value <= 100000, MFEM_B200_ERR_INVALID,
                     "coarse_fine_nodes must be 0 (no level 1) or the number of nodes per small aggregate");
        h->opt_coarse_fine = (int)value;
        h->precondValid = false;
    } else if (n == "coarse_shape") {
        MFEM_REQUIRE(value == 0, MFEM_B200_ERR_INVALID, "coarse_shape: only 0 (nested box grids) is supported");
    } else if (n == "comm_p2p") {
        h->opt_comm_p2p = value != 0;
    } else if (n == "spmv_min_blocks") {
        h->opt_spmv_min_blocks = (int)value;
    } else if (n == "spmv_prefetch") {
        h->opt_spmv_prefetch = value != 0;
    } else if (n == "spmv_lanes") {
        MFEM_REQUIRE(value == 0 || value == 8 || value == 16 || value == 32, MFEM_B200_ERR_INVALID,
                     "spmv_lanes must be 0 (auto), 8, 16 or 32");
        h->opt_spmv_lanes = (int)value;
    } else {
        throw CudaError(MFEM_B200_ERR_INVALID, "unknown option " + n);
    }
    API_END(h)
}

int mfem_b200_set_mesh(mfem_b200_handle h, int dim, int degree, int64_t n_nodes, const double *nodes,
                       int64_t n_elems, const int32_t *elem_nodes, const int64_t *dof_for_node, int64_t n_dofs) {
    API_BEGIN(h)
    setup_mesh(h, dim, degree, n_nodes, nodes, n_elems, elem_nodes, dof_for_node, n_dofs);
    API_END(h)
}

int mfem_b200_set_node_positions(mfem_b200_handle h, const double *nodes) {
    API_BEGIN(h)
    MFEM_REQUIRE(h->nElems > 0 && nodes, MFEM_B200_ERR_INVALID, "set_node_positions: no mesh set");
    MFEM_CUDA(cudaMemcpyAsync(h->nodes, nodes, h->nodes.bytes(), cudaMemcpyHostToDevice, h->stream));
    compute_geometry(h);
    h->precondValid = false;
    h->meshVersion++;
    API_END(h)
}

int mfem_b200_set_material_constant(mfem_b200_handle h, const double *D) {
    API_BEGIN(h)
    MFEM_REQUIRE(h->N > 0 && D, MFEM_B200_ERR_INVALID, "set_material: set the mesh first");
    const int F = flat_len(h->N);
    std::memset(&h->Dconst, 0, sizeof(h->Dconst));
    for (int i = 0; i < F; ++i)
        for (int j = 0; j < F; ++j) {
            // symmetrise from the upper triangle, as every reference setter does
            // (ElasticityTensor.hh:134 selfadjointView<Upper>)
            h->Dconst.d[i * F + j] = (i <= j) ? D[i * F + j] : D[j * F + i];
        }
    h->perElemD = false;
    h->Delem.free();
    h->haveMaterial = true;
    h->valuesValid = false;
    h->precondValid = false;
    API_END(h)
}

int mfem_b200_set_material_per_element(mfem_b200_handle h, const double *D) {
    API_BEGIN(h)
    MFEM_REQUIRE(h->nElems > 0 && D, MFEM_B200_ERR_INVALID, "set_material: set the mesh first");
    const int F = flat_len(h->N);
    std::vector<double> sym((size_t)h->nElems * F * F);
    for (int64_t e = 0; e < h->nElems; ++e)
        for (int i = 0; i < F; ++i)
            for (int j = 0; j < F; ++j)
                sym[(size_t)e * F * F + i * F + j] = (i <= j) ? D[e * F * F + i * F + j] : D[e * F * F + j * F + i];
    h->Delem.alloc(sym.size());
    MFEM_CUDA(cudaMemcpy(h->Delem, sym.data(), h->Delem.bytes(), cudaMemcpyHostToDevice));
    h->perElemD = true;
    h->haveMaterial = true;
    h->valuesValid = false;
    h->precondValid = false;
    API_END(h)
}

int mfem_b200_assemble(mfem_b200_handle h) {
    API_BEGIN(h)
    assemble_values(h);
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_get_bsr_sizes(mfem_b200_handle h, int64_t *n_block_rows, int64_t *nnz_blocks) {
    API_BEGIN(h)
    MFEM_REQUIRE(h->patternValid, MFEM_B200_ERR_INVALID, "get_bsr_sizes: matrix not assembled");
    if (n_block_rows) *n_block_rows = h->nDofs;
    if (nnz_blocks) *nnz_blocks = h->nnzb;
    API_END(h)
}

int mfem_b200_get_bsr(mfem_b200_handle h, int64_t *rowptr, int32_t *colidx, double *vals) {
    API_BEGIN(h)
    MFEM_REQUIRE(rowptr, MFEM_B200_ERR_INVALID, "get_bsr: rowptr required");
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    export_bsr(h, rowptr, colidx, vals);
    API_END(h)
}

int mfem_b200_dump_upper_triplets(mfem_b200_handle h, const char *path) {
    API_BEGIN(h)
    MFEM_REQUIRE(path && h->valuesValid, MFEM_B200_ERR_INVALID, "dump: matrix not assembled");
    const int N = h->N, NN = N * N;
    std::vector<int64_t> rp((size_t)h->nDofs + 1);
    std::vector<int32_t> ci((size_t)h->nnzb);
    std::vector<double> v((size_t)h->nnzb * NN);
    export_bsr(h, rp.data(), ci.data(), v.data());
    // column-major order like the reference's sorted CSC (SparseMatrices.hh:355-361)
    std::vector<uint64_t> rows, cols;
    std::vector<double> vv;
    // K is symmetric: upper triangle of column j == lower triangle entries of row j transposed;
    // emit by (col, row) using the row-major data of row `col`.
    for (int64_t bj = 0; bj < h->nDofs; ++bj)
        for (int cj = 0; cj < N; ++cj)
            for (int64_t k = rp[(size_t)bj]; k < rp[(size_t)bj + 1]; ++k) {
                const int64_t bi = ci[(size_t)k];
                for (int cc = 0; cc < N; ++cc) {
                    const uint64_t row = (uint64_t)(N * bi + cc), col = (uint64_t)(N * bj + cj);
                    if (row > col) continue;
                    const double val = v[(size_t)k * NN + cj * N + cc];   // K[col,row] == K[row,col]
                    if (val * val <= 0.0) continue;                        // pruneTol = 0
                    rows.push_back(row); cols.push_back(col); vv.push_back(val);
                }
            }
    std::ofstream os(path, std::ios::binary);
    MFEM_REQUIRE(os.is_open(), MFEM_B200_ERR_INVALID, std::string("cannot open ") + path);
    const uint64_t nnz = rows.size();
    os.write(reinterpret_cast<const char *>(&nnz), 8);
    os.write(reinterpret_cast<const char *>(rows.data()), 8 * nnz);
    os.write(reinterpret_cast<const char *>(cols.data()), 8 * nnz);
    os.write(reinterpret_cast<const char *>(vv.data()), 8 * nnz);
    API_END(h)
}

__global__ void k_set_fixed(int64_t n, int N, const int64_t *vars, const double *vals, const int32_t *ext2int,
                            uint8_t *mask, double *fixedVals) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = vars[i];
    const int64_t d = v / N;
    const int cc = (int)(v - d * N);
    const int64_t vi = (int64_t)ext2int[d] * N + cc;
    mask[vi] = 1;
    fixedVals[vi] = vals ? vals[i] : 0.0;
}

// SPSDSystem(K) for a matrix assembled elsewhere (SparseMatrices.hh:2332-2348 setConstrained with an
// empty C): COO triplets -> summed block-CSR on the host (the reference's TripletMatrix is a host object
// and its sumRepeated is host work too), then the usual device layout.
int mfem_b200_set_matrix_triplets(mfem_b200_handle h, int block_dim, int64_t n_vars, int64_t nnz, const int64_t *rows,
                                  const int64_t *cols, const double *vals, int upper_triangle_only) {
    API_BEGIN(h)
    MFEM_REQUIRE(block_dim == 2 || block_dim == 3, MFEM_B200_ERR_INVALID, "set_matrix: block_dim must be 2 or 3");
    MFEM_REQUIRE(n_vars > 0 && n_vars % block_dim == 0, MFEM_B200_ERR_INVALID, "set_matrix: n_vars must be a positive multiple of block_dim");
    MFEM_REQUIRE(n_vars / block_dim < INT32_MAX, MFEM_B200_ERR_INVALID, "set_matrix: too many block rows");
    MFEM_REQUIRE(nnz > 0 && rows && cols && vals, MFEM_B200_ERR_INVALID, "set_matrix: empty matrix");
    MFEM_REQUIRE(h->nRanks <= 1, MFEM_B200_ERR_INVALID, "set_matrix: single-GPU handles only");
    const int d = block_dim;
    const int64_t nb = n_vars / d;
    struct Entry { uint64_t key; int rc; double v; };
    std::vector<Entry> e;
    e.reserve((size_t)nnz * (upper_triangle_only ? 2 : 1));
    for (int64_t k = 0; k < nnz; ++k) {
        const int64_t i = rows[k], j = cols[k];
        MFEM_REQUIRE(i >= 0 && i < n_vars && j >= 0 && j < n_vars, MFEM_B200_ERR_INVALID, "set_matrix: index out of range");
        if (upper_triangle_only) MFEM_REQUIRE(i <= j, MFEM_B200_ERR_INVALID, "set_matrix: entry below the diagonal in an upper-triangle matrix");
        e.push_back({((uint64_t)(i / d) << 32) | (uint64_t)(j / d), (int)((i % d) * d + (j % d)), vals[k]});
        if (upper_triangle_only && i != j)
            e.push_back({((uint64_t)(j / d) << 32) | (uint64_t)(i / d), (int)((j % d) * d + (i % d)), vals[k]});
    }
    std::stable_sort(e.begin(), e.end(), [](const Entry &a, const Entry &b) { return a.key < b.key; });
    std::vector<int64_t> rowptr((size_t)nb + 1, 0);
    std::vector<int32_t> colidx;
    std::vector<double> blocks;
    uint64_t prev = ~0ULL;
    for (const Entry &t : e) {
        if (t.key != prev) {
            prev = t.key;
            colidx.push_back((int32_t)(t.key & 0xffffffffULL));
            blocks.insert(blocks.end(), (size_t)d * d, 0.0);
            rowptr[(size_t)(t.key >> 32) + 1]++;
        }
        blocks[blocks.size() - (size_t)d * d + t.rc] += t.v;       // repeated entries are summed (sumRepeated)
    }
    for (int64_t r = 0; r < nb; ++r) rowptr[(size_t)r + 1] += rowptr[(size_t)r];
    upload_external_bsr(h, d, nb, rowptr, colidx, blocks);
    API_END(h)
}

int mfem_b200_fix_variables(mfem_b200_handle h, int64_t n, const int64_t *vars, const double *values) {
    API_BEGIN(h)
    MFEM_REQUIRE(h->nDofs > 0, MFEM_B200_ERR_INVALID, "fix_variables: no mesh or matrix set");
    if (n == 0) return MFEM_B200_OK;
    MFEM_REQUIRE(n > 0 && vars, MFEM_B200_ERR_INVALID, "fix_variables: bad arguments");
    const int64_t nvar = h->nvar();
    for (int64_t i = 0; i < n; ++i) {
        MFEM_REQUIRE(vars[i] >= 0 && vars[i] < nvar, MFEM_B200_ERR_INVALID, "fix_variables: variable out of range");
        MFEM_REQUIRE(!h->fixedHost[(size_t)vars[i]], MFEM_B200_ERR_ALREADY_FIXED, "Variable already fixed.");
        h->fixedHost[(size_t)vars[i]] = 1;
    }
    h->nFixed += n;
    ensure_work(h);
    DevBuf<int64_t> dv((size_t)n);
    DevBuf<double> dx;
    MFEM_CUDA(cudaMemcpyAsync(dv, vars, dv.bytes(), cudaMemcpyHostToDevice, h->stream));
    if (values) {
        dx.alloc((size_t)n);
        MFEM_CUDA(cudaMemcpyAsync(dx, values, dx.bytes(), cudaMemcpyHostToDevice, h->stream));
    }
    k_set_fixed<<<grid_for(n, 256), 256, 0, h->stream>>>(n, h->N, dv, values ? dx.p : nullptr, h->ext2int,
                                                         h->fixedMask, h->fixedVals);
    h->launches++;
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    MFEM_CUDA(cudaGetLastError());
    h->precondValid = false;
    API_END(h)
}

int mfem_b200_clear_fixed_variables(mfem_b200_handle h) {
    API_BEGIN(h)
    std::fill(h->fixedHost.begin(), h->fixedHost.end(), 0);
    h->nFixed = 0;
    if (h->fixedMask.n) {
        MFEM_CUDA(cudaMemsetAsync(h->fixedMask, 0, h->fixedMask.bytes(), h->stream));
        MFEM_CUDA(cudaMemsetAsync(h->fixedVals, 0, h->fixedVals.bytes(), h->stream));
    }
    h->precondValid = false;
    API_END(h)
}

int mfem_b200_solve(mfem_b200_handle h, int nrhs, const double *f, double *u, double rtol, int max_iters,
                    mfem_b200_solve_info *info) {
    API_BEGIN(h)
    MFEM_REQUIRE(nrhs >= 1 && f && u, MFEM_B200_ERR_BAD_RHS, "Bad RHS");
    MFEM_REQUIRE(h->valuesValid, MFEM_B200_ERR_INVALID, "No system to solve");
    MFEM_REQUIRE(rtol > 0 && max_iters > 0, MFEM_B200_ERR_INVALID, "solve: bad tolerance / iteration limit");
    const size_t n = (size_t)h->nvar();
    int firstErr = MFEM_B200_OK;
    std::string firstMsg;
    if (h->opt_batch_rhs && nrhs == flat_len(h->N) && !will_use_coarse(h)) {
        // the cell-problem case without aggregation levels (small problems, or coarse_aggregates = 0): all right-hand
        // sides in one batched block-Jacobi PCG (one matrix stream for all of them).  With the levels on, the systems
        // are solved one after the other -- measured on cfg4 (5.5 M quadratic tets): ~250 multilevel iterations per
        // system against 2574 batched block-Jacobi iterations at half the per-system cost
        DevBuf<double> fext(n), fin(n * nrhs), uin(n * nrhs), uext(n);
        for (int k = 0; k < nrhs; ++k) {
            MFEM_CUDA(cudaMemcpyAsync(fext, f + (size_t)k * n, n * 8, cudaMemcpyHostToDevice, h->stream));
            permute_to_internal(h, fext, fin.p + (size_t)k * n);
        }
        try {
            pcg_solve_multi(h, nrhs, fin, uin, rtol, max_iters, info);
        } catch (const CudaError &e) {
            if (e.status != MFEM_B200_ERR_NO_CONVERGE) throw;
            firstErr = e.status; firstMsg = e.what();
        }
        for (int k = 0; k < nrhs; ++k) {
            permute_to_external(h, uin.p + (size_t)k * n, uext);
            MFEM_CUDA(cudaMemcpyAsync(u + (size_t)k * n, uext, n * 8, cudaMemcpyDeviceToHost, h->stream));
        }
        MFEM_CUDA(cudaStreamSynchronize(h->stream));
        if (firstErr != MFEM_B200_OK) throw CudaError(firstErr, firstMsg);
        return MFEM_B200_OK;
    }
    DevBuf<double> fext(n), fin(n), uin(n), uext(n);
    for (int k = 0; k < nrhs; ++k) {
        MFEM_CUDA(cudaMemcpyAsync(fext, f + (size_t)k * n, n * 8, cudaMemcpyHostToDevice, h->stream));
        permute_to_internal(h, fext, fin);
        try {
            pcg_solve(h, fin, uin, rtol, max_iters, info ? info + k : nullptr);
        } catch (const CudaError &e) {
            if (e.status != MFEM_B200_ERR_NO_CONVERGE) throw;
            if (firstErr == MFEM_B200_OK) { firstErr = e.status; firstMsg = e.what(); }
        }
        permute_to_external(h, uin, uext);
        MFEM_CUDA(cudaMemcpyAsync(u + (size_t)k * n, uext, n * 8, cudaMemcpyDeviceToHost, h->stream));
        MFEM_CUDA(cudaStreamSynchronize(h->stream));
    }
    if (firstErr != MFEM_B200_OK) throw CudaError(firstErr, firstMsg);
    API_END(h)
}

int mfem_b200_spmv(mfem_b200_handle h, const double *x, double *y) {
    API_BEGIN(h)
    MFEM_REQUIRE(x && y, MFEM_B200_ERR_INVALID, "spmv: null argument");
    const size_t n = (size_t)h->nvar();
    DevBuf<double> a(n), b(n);
    MFEM_CUDA(cudaMemcpyAsync(a, x, n * 8, cudaMemcpyHostToDevice, h->stream));
    permute_to_internal(h, a, b);
    spmv_plain(h, b, a);
    permute_to_external(h, a, b);
    MFEM_CUDA(cudaMemcpyAsync(y, b, n * 8, cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_apply_K(mfem_b200_handle h, const double *u_nodes, double *Ku_nodes) {
    API_BEGIN(h)
    MFEM_REQUIRE(u_nodes && Ku_nodes, MFEM_B200_ERR_INVALID, "apply_K: null argument");
    const size_t n = (size_t)h->nNodes * h->N;
    DevBuf<double> a(n), b(n);
    MFEM_CUDA(cudaMemcpyAsync(a, u_nodes, n * 8, cudaMemcpyHostToDevice, h->stream));
    apply_K_nodes(h, a, b);
    MFEM_CUDA(cudaMemcpyAsync(Ku_nodes, b, n * 8, cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_const_strain_load(mfem_b200_handle h, const double *eps_flat, double *f_dofs) {
    API_BEGIN(h)
    MFEM_REQUIRE(eps_flat && f_dofs, MFEM_B200_ERR_INVALID, "const_strain_load: null argument");
    const size_t n = (size_t)h->nvar();
    DevBuf<double> fext(n);
    const_strain_load(h, eps_flat, fext);
    MFEM_CUDA(cudaMemcpyAsync(f_dofs, fext, n * 8, cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_avg_strain_stress(mfem_b200_handle h, const double *u_nodes, double *strain, double *stress) {
    API_BEGIN(h)
    MFEM_REQUIRE(u_nodes && h->nElems > 0, MFEM_B200_ERR_INVALID, "avg_strain_stress: bad arguments");
    const size_t n = (size_t)h->nNodes * h->N, m = (size_t)h->nElems * flat_len(h->N);
    DevBuf<double> u(n), e, s;
    if (strain) e.alloc(m);
    if (stress) s.alloc(m);
    MFEM_CUDA(cudaMemcpyAsync(u, u_nodes, n * 8, cudaMemcpyHostToDevice, h->stream));
    avg_strain_stress(h, u, strain ? e.p : nullptr, stress ? s.p : nullptr);
    if (strain) MFEM_CUDA(cudaMemcpyAsync(strain, e, m * 8, cudaMemcpyDeviceToHost, h->stream));
    if (stress) MFEM_CUDA(cudaMemcpyAsync(stress, s, m * 8, cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_get_volumes(mfem_b200_handle h, double *vol) {
    API_BEGIN(h)
    MFEM_REQUIRE(vol && h->geom.n, MFEM_B200_ERR_INVALID, "get_volumes: no mesh set");
    const int GS = 1 + h->N * (h->N + 1);
    MFEM_CUDA(cudaMemcpy2D(vol, sizeof(double), h->geom, GS * sizeof(double), sizeof(double), (size_t)h->nElems,
                           cudaMemcpyDeviceToHost));
    API_END(h)
}

double mfem_b200_get_timer(mfem_b200_handle h, const char *section) {
    if (!h || !section) return -1.0;
    auto it = h->timers.find(section);
    return it == h->timers.end() ? -1.0 : it->second;
}

int mfem_b200_reset_timers(mfem_b200_handle h) {
    if (!h) return MFEM_B200_ERR_INVALID;
    h->timers.clear();
    h->launches = 0;
    return MFEM_B200_OK;
}

int64_t mfem_b200_launch_count(mfem_b200_handle h) { return h ? h->launches : -1; }

int mfem_b200_time_spmv(mfem_b200_handle h, int iters, double *seconds_per_launch) {
    API_BEGIN(h)
    MFEM_REQUIRE(iters > 0 && seconds_per_launch, MFEM_B200_ERR_INVALID, "time_spmv: bad arguments");
    *seconds_per_launch = time_spmv(h, iters);
    API_END(h)
}

int mfem_b200_time_operator(mfem_b200_handle h, int iters, double *seconds_per_product, double *seconds_parts,
                            int *matrix_free) {
    API_BEGIN(h)
    MFEM_REQUIRE(iters > 0 && seconds_per_product, MFEM_B200_ERR_INVALID, "time_operator: bad arguments");
    *seconds_per_product = time_operator(h, iters, matrix_free, seconds_parts);
    API_END(h)
}

int mfem_b200_apply_preconditioner(mfem_b200_handle h, const double *r, double *z, double *rz) {
    API_BEGIN(h)
    MFEM_REQUIRE(r && z && rz, MFEM_B200_ERR_INVALID, "apply_preconditioner: null argument");
    const size_t n = (size_t)h->nvar();
    DevBuf<double> a(n), b(n);
    MFEM_CUDA(cudaMemcpyAsync(a, r, n * 8, cudaMemcpyHostToDevice, h->stream));
    permute_to_internal(h, a, b);
    apply_preconditioner(h, b, a, rz);
    permute_to_external(h, a, b);
    MFEM_CUDA(cudaMemcpyAsync(z, b, n * 8, cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_get_coarse_array(mfem_b200_handle h, const char *name, double *out, int64_t capacity, int64_t *n) {
    API_BEGIN(h)
    MFEM_REQUIRE(name && n, MFEM_B200_ERR_INVALID, "get_coarse_array: null argument");
    *n = get_coarse_array(h, name, out, capacity);
    API_END(h)
}

// ---- discrete shape derivatives (csrc/shape.cu)
static __global__ void k_max_vertex_node(int64_t nElems, int npe, int nv, const int32_t *__restrict__ elemNodes, int *out) {
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    int m = 0;
    for (int k = 0; k < nv; ++k) m = max(m, elemNodes[e * npe + k]);
    atomicMax(out, m);
}
// delta_p arrives per VERTEX (the first n_vertices nodes); returns a device copy after checking that no element
// reads past it
static void upload_delta_p(mfem_b200_handle h, const double *delta_p, int64_t n_vertices, DevBuf<double> &dp) {
    MFEM_REQUIRE(h->nElems > 0 && !h->externalMatrix, MFEM_B200_ERR_INVALID, "shape derivatives need a mesh");
    MFEM_REQUIRE(delta_p && n_vertices > 0 && n_vertices <= h->nNodes, MFEM_B200_ERR_INVALID, "shape derivatives: bad delta_p / n_vertices");
    if (h->maxVertexNodeVersion != h->meshVersion) {
        DevBuf<int> m(1);
        MFEM_CUDA(cudaMemsetAsync(m, 0, sizeof(int), h->stream));
        k_max_vertex_node<<<grid_for(h->nElems, 256), 256, 0, h->stream>>>(h->nElems, h->npe, h->N + 1, h->elemNodes, m);
        int hm = 0;
        MFEM_CUDA(cudaMemcpyAsync(&hm, m, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        MFEM_CUDA(cudaStreamSynchronize(h->stream));
        h->maxVertexNode = hm;
        h->maxVertexNodeVersion = h->meshVersion;
        h->launches++;
    }
    MFEM_REQUIRE(h->maxVertexNode < n_vertices, MFEM_B200_ERR_INVALID, "shape derivatives: per-vertex perturbation expected (an element vertex lies beyond n_vertices)");
    dp.alloc((size_t)n_vertices * h->N);
    MFEM_CUDA(cudaMemcpyAsync(dp, delta_p, dp.bytes(), cudaMemcpyHostToDevice, h->stream));
}

int mfem_b200_apply_delta_K(mfem_b200_handle h, const double *u_nodes, const double *delta_p, int64_t n_vertices, double *out_dofs) {
    API_BEGIN(h)
    MFEM_REQUIRE(u_nodes && out_dofs, MFEM_B200_ERR_INVALID, "apply_delta_K: null argument");
    DevBuf<double> dp, u((size_t)h->nNodes * h->N), out((size_t)h->nvar());
    upload_delta_p(h, delta_p, n_vertices, dp);
    MFEM_CUDA(cudaMemcpyAsync(u, u_nodes, u.bytes(), cudaMemcpyHostToDevice, h->stream));
    apply_delta_K(h, u, dp, out);
    MFEM_CUDA(cudaMemcpyAsync(out_dofs, out, out.bytes(), cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_delta_const_strain_load(mfem_b200_handle h, const double *eps_flat, const double *delta_p, int64_t n_vertices,
                                      double *out_dofs) {
    API_BEGIN(h)
    MFEM_REQUIRE(eps_flat && out_dofs, MFEM_B200_ERR_INVALID, "delta_const_strain_load: null argument");
    DevBuf<double> dp, out((size_t)h->nvar());
    upload_delta_p(h, delta_p, n_vertices, dp);
    delta_const_strain_load(h, eps_flat, dp, out);
    MFEM_CUDA(cudaMemcpyAsync(out_dofs, out, out.bytes(), cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_delta_avg_strain(mfem_b200_handle h, const double *u_nodes, const double *delta_u_nodes, const double *delta_p,
                               int64_t n_vertices, double *strain) {
    API_BEGIN(h)
    MFEM_REQUIRE(u_nodes && delta_u_nodes && strain, MFEM_B200_ERR_INVALID, "delta_avg_strain: null argument");
    const size_t nn = (size_t)h->nNodes * h->N;
    DevBuf<double> dp, u(nn), du(nn), out((size_t)h->nElems * flat_len(h->N));
    upload_delta_p(h, delta_p, n_vertices, dp);
    MFEM_CUDA(cudaMemcpyAsync(u, u_nodes, u.bytes(), cudaMemcpyHostToDevice, h->stream));
    MFEM_CUDA(cudaMemcpyAsync(du, delta_u_nodes, du.bytes(), cudaMemcpyHostToDevice, h->stream));
    delta_avg_strain(h, u, du, dp, out);
    MFEM_CUDA(cudaMemcpyAsync(strain, out, out.bytes(), cudaMemcpyDeviceToHost, h->stream));
    MFEM_CUDA(cudaStreamSynchronize(h->stream));
    API_END(h)
}

int mfem_b200_release_cached_memory(void) {
    mfem::pool_release_all();
    return MFEM_B200_OK;
}

}  // extern "C"

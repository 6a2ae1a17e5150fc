// Shared state and helpers of libmfem_b200 (internal header; the public surface is
// include/mfem_b200.h).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/mfem_b200.h"
#include "elem_math.cuh"

namespace mfem {

struct CudaError : std::runtime_error {
    int status;
    CudaError(int st, const std::string &m) : std::runtime_error(m), status(st) {}
};

#define MFEM_CUDA(call)                                                                          \
    do {                                                                                         \
        cudaError_t e_ = (call);                                                                 \
        if (e_ != cudaSuccess)                                                                   \
            throw mfem::CudaError(MFEM_B200_ERR_CUDA, std::string(#call) + " failed at " +        \
                                                          __FILE__ + ":" + std::to_string(__LINE__) + \
                                                          ": " + cudaGetErrorString(e_));        \
    } while (0)

#define MFEM_REQUIRE(cond, status, msg)                       \
    do {                                                      \
        if (!(cond)) throw mfem::CudaError((status), (msg));  \
    } while (0)

// Device memory comes from a small process-wide caching pool (capi.cu): cudaMalloc / cudaFree of the multi-GB arrays
// of a handle cost hundreds of milliseconds (page-table work, a device-wide synchronisation per cudaFree), which a
// long-lived process that creates handle after handle (an optimisation loop, bench.py's end-to-end steps) would pay
// every time.  A freed block is kept and handed to the next request of a similar size on the same device; the pool is
// emptied when an allocation fails, on mfem_b200_release_cached_memory(), or never used with MFEM_B200_POOL=0.
void *pool_alloc(size_t bytes);
void pool_free(void *p);

// Owning device buffer (lifetime = handle or a setup scope).
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept {
        if (this != &o) { free(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { free(); }
    void alloc(size_t count) {
        free();
        n = count;
        if (count) p = static_cast<T *>(pool_alloc(count * sizeof(T)));
    }
    void free() {
        if (p) pool_free(p);
        p = nullptr; n = 0;
    }
    size_t bytes() const { return n * sizeof(T); }
    operator T *() const { return p; }
};

struct MatD {            // constant material, lives in the kernel parameter (constant) bank
    double d[36];
};

// One cached PCG workspace (vectors of length nvar).
struct PcgWork {
    DevBuf<double> x, r, z, p, Ap, b, ufix;
    DevBuf<double> partials;     // per-CTA partial sums, 4 slots
    DevBuf<double> scal;         // device scalars (see solver.cu)
    DevBuf<double> dotLoc;       // multi-GPU (batched solver): this rank's partial sums before the all-reduce
    DevBuf<double> red;          // [2 + 32768] r.z, r.r and the coarse residuals c2: ONE all-reduce per PCG iteration
    DevBuf<unsigned> ticket;     // last-block tickets
    DevBuf<int> status;          // [0]=iterations done, [1]=state (0 running, 1 converged, 2 breakdown, 3 nan)
};

struct Halo;   // multi-GPU interface exchange (comm.cu)
struct PcgWorkMulti;   // workspace of the batched (multi right-hand-side) PCG (solver_multi.inl)
struct CoarseSpace;    // aggregation coarse space of the optional two-level preconditioner (coarse.inl)

}  // namespace mfem

struct mfem_b200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    std::string err;

    // options
    int opt_reorder = 1;
    int opt_assembly = 0;
    int opt_graph = 1;
    int opt_spmv_kernel = 0;               // 0 auto, 1 direct-load kernel, 2 TMA-ring kernel, 3 index-pipelined, 4 symmetric (upper tails + atomics)
    int opt_spmm_kernel = 0;               // batched PCG: 0/1 full-warp SpMM, 2 half-warp split SpMM (even batch sizes)
    int opt_batch_rhs = 1;                 // solve flatLen(N) right-hand sides as one batched PCG (SpMM)
    int opt_spmv_min_blocks = 0;           // SpMV occupancy experiment: 3 or 5 CTAs per SM instead of 4 (N = 3, 32 lanes)
    int opt_spmv_prefetch = 0;             // SpMV: L2 prefetch of the row a warp streams next (one prefetch instruction per lane and row)
    int opt_spmv_lanes = 0;                // lanes per block row in the SpMV (0 = choose from the mean row length)
    int opt_matrix_free = -1;              // PCG operator: -1 auto (mesh-based for 3D quadratic elements), 0 assembled SpMV, 1 mesh-based whenever possible
    int opt_mf_slot_pad = 0;               // matrix-free operator, 3D: 32-byte (padded) result slots; 0 = packed 24-byte slots (A/B)
    int opt_mf_gather_lanes = 0;           // matrix-free operator: lanes per DoF row in the gather kernel (4 or 8, the in-loop launch only; 0 = 1 for chunk partials, 8 for slots)
    int opt_mf_elem_order = 0;             // matrix-free operator: 1 = elements processed in the order of their DoFs (build_mf_plan), 0 = caller's order (A/B)
    int opt_mf_gather_policy = 3;          // slot loads of the gather kernel: 0 evict_first, 1 evict_last, 2 evict_normal, 3 evict_last + L1 allocation (A/B)
    int opt_mf_chunked = 1;                // matrix-free operator: 1 = per-chunk partial sums in shared memory (build_mf_chunks), 0 = one slot per (element, node)
    int opt_mf_chunk_elems = 64;           // elements per chunk (= threads per CTA) of the chunked operator: 32, 64 or 128
    int opt_mf_chunk_warps = 16;           // chunked operator: resident warps per SM the element kernel is compiled for (12 or 16)
    int opt_coarse = -1;                   // large aggregates of the multilevel preconditioner: -1 automatic (from the
                                           // problem size; block-Jacobi only below 30k DoFs), 0 = block-Jacobi only
    int opt_coarse_fine = 64;              // DoFs (nodes) per small (level-1) aggregate; 0 = no level 1 (two-level method)
    int64_t maxVertexNode = -1, maxVertexNodeVersion = -1;   // largest node id in a vertex slot of elemNodes (shape derivatives)
    int64_t meshVersion = 0;               // bumped by everything that changes DoFs, positions or the interface

    // mesh
    int N = 0, deg = 0, npe = 0;
    int64_t nNodes = 0, nElems = 0, nDofs = 0;
    bool periodic = false;                 // dof_for_node given
    bool externalMatrix = false;           // K came from mfem_b200_set_matrix_triplets (no mesh, no assembly)
    mfem::DevBuf<double> nodes;            // [nNodes*N]   caller's node order
    mfem::DevBuf<int32_t> elemNodes;       // [nElems*npe] caller's node ids
    mfem::DevBuf<int32_t> elemDof;         // [nElems*npe] INTERNAL dof ids
    mfem::DevBuf<int32_t> nodeDof;         // [nNodes]     internal dof of each node
    mfem::DevBuf<int32_t> ext2int, int2ext;  // [nDofs]
    mfem::DevBuf<double> geom;             // [nElems*(1+N*(N+1))]  vol, G
    bool geomValid = false;
    mfem::DevBuf<double> geomP;            // [nElems*16] packed (G_a, vol) slots for the block-owner assembly
    bool geomPValid = false;
    mfem::DevBuf<double> elemY;            // [nElems*npe*slot] per-element results of the matrix-free operator (matfree.inl)
    // matrix-free operator plan (setup.cu build_mf_plan): the elements in the order of their DoFs (internal DoF ids follow
    // a Morton curve), so that the rows the gather kernel sweeps find their slots in one moving window of elemY
    bool mfPlanValid = false, mfGeomValid = false;
    mfem::DevBuf<int32_t> mfPerm;          // [nElems] element handled by thread k
    mfem::DevBuf<int32_t> mfElemDof;       // [nElems*npe] elemDof in that order
    mfem::DevBuf<double> mfGeomP;          // [nElems*16]  geomP in that order
    mfem::DevBuf<int32_t> mfIncList;       // [totalInc]   k*npe + i sorted by DoF (same extents as incPtr)
    // chunked variant (setup.cu build_mf_chunks): opt_mf_chunk_elems consecutive elements = one CTA, which sums the results of its
    // elements per DISTINCT DoF in shared memory and writes one partial per (chunk, DoF) instead of one slot per (element, node)
    bool mfChunksValid = false;
    int mfChunkElems = 0;                  // elements per chunk the tables were built for
    int64_t mfPartials = 0;                // total (chunk, DoF) partials
    mfem::DevBuf<int32_t> mfChunkBase;     // [nChunks+1] first partial of each chunk
    mfem::DevBuf<int32_t> mfChunkDof;      // [nChunks*S] DoF of each chunk-local index (ascending DoF id), fixed stride, 0 beyond the chunk's count
    mfem::DevBuf<uint16_t> mfLocalIdx;     // [nChunks*S] chunk-local DoF index of slot i*chunk + t   (S = chunk*npe)
    mfem::DevBuf<uint16_t> mfCsrPtr;       // [nChunks*(S+1)] extents of each chunk-local DoF in the sorted slot order
    mfem::DevBuf<uint16_t> mfCsrList;      // [nChunks*S] rank of slot i*chunk + t in the order sorted by chunk-local DoF ((element, local node) inside)
    mfem::DevBuf<int64_t> mfIncPtr2;       // [nDofs+1]   DoF row -> its partials
    mfem::DevBuf<int32_t> mfIncList2;      // [mfPartials]

    // material
    bool haveMaterial = false, perElemD = false;
    mfem::MatD Dconst{};
    mfem::DevBuf<double> Delem;            // [nElems*flat*flat]

    // block-CSR matrix (internal numbering)
    bool patternValid = false, valuesValid = false;
    int64_t nnzb = 0;
    mfem::DevBuf<int64_t> rowptr;          // [nDofs+1]
    mfem::DevBuf<int32_t> colidx;          // [nnzb]
    mfem::DevBuf<double> vals;             // [nnzb*N*N (+2 pad)]  "row-plane" layout, see val_index()
    mfem::DevBuf<int64_t> tileRow;         // [nTiles+1] first row of each kSpmvTileWindow-block window (TMA SpMV)
    int64_t maxRowLen = 0;                 // longest block row
    mfem::DevBuf<int32_t> upperStart;      // [nDofs] first slot with column >= row (symmetric SpMV reads only that tail)
    // DoF -> incident (element, local node) lists
    int64_t totalInc = 0;
    mfem::DevBuf<int64_t> incPtr;          // [nDofs+1]
    mfem::DevBuf<int32_t> incList;         // [totalInc]  e*npe + i
    mfem::DevBuf<int64_t> jobRow;          // [ceil(totalInc/kAsmChunk)+1] first row of each assembly job
    // block-owner assembly plan (setup.cu k_plan_*): per chunk of kBlkChunk blocks
    mfem::DevBuf<uint16_t> planCnt;        // [nnzb] element contributions per block (plan build only)
    mfem::DevBuf<uint16_t> planSegOff;     // [nnzb] first segment of the block inside its chunk
    mfem::DevBuf<uint8_t> planNseg;        // [nnzb] segments of the block
    mfem::DevBuf<uint8_t> planChunkL;      // [nChunks] segment length of the chunk
    mfem::DevBuf<uint16_t> planSegOrder;   // [nChunks*kSegSlots] segment (chunk-local id) of each thread slot, 0xffff = none
    mfem::DevBuf<int64_t> planWarpBase;    // [nChunks*kSegSlots/32+1] first list entry of each warp-round
    mfem::DevBuf<uint32_t> planList;       // [planEntries] interleaved pair ids e*npe^2 + i*npe + j, kPlanSentinel = none
    mfem::DevBuf<int32_t> planChunkRow;    // [nChunks+1] block row containing the chunk's first block
    mfem::DevBuf<int64_t> planElemPtr;     // [nChunks+1] elements staged per chunk (prefix); empty range = unstaged chunk
    mfem::DevBuf<uint32_t> planElems;      // distinct elements of every staged chunk, in chunk-local order
    int64_t planEntries = 0;
    mfem::DevBuf<double> pairW;            // [4][npe*npe] W weights of the (i,j) pair table (assemble.cu)
    mfem::DevBuf<uint32_t> pairIdx;        // [npe*npe]    packed gradient offsets
    int pairTabKey = 0;                    // 10*N + deg the table was built for
    // element colouring (assembly mode 1)
    int nColors = 0;
    std::vector<int64_t> colorPtr;         // host: [nColors+1]
    mfem::DevBuf<int32_t> colorElems;      // elements sorted by colour

    // constraints
    std::vector<uint8_t> fixedHost;        // [nDofs*N] caller's numbering
    int64_t nFixed = 0;
    mfem::DevBuf<uint8_t> fixedMask;       // [nDofs*N] internal numbering, 1 = fixed
    mfem::DevBuf<double> fixedVals;        // [nDofs*N]
    mfem::DevBuf<double> Minv;             // [nDofs*N*N] inverse diagonal blocks after masking
    bool precondValid = false;

    mfem::PcgWork work;
    bool workValid = false;
    mfem::PcgWorkMulti *workMulti = nullptr;
    mfem::CoarseSpace *coarse = nullptr;   // built with the preconditioner when opt_coarse > 0

    // multi-GPU
    int nRanks = 1, rank = 0;
    void *ncclComm = nullptr;
    bool ownsComm = false;                 // false: borrowed from another handle (mfem_b200_comm_share)
    mfem::Halo *halo = nullptr;
    void *peerWin = nullptr;               // PeerWinOwner (comm.cu): window mapped by all ranks for the small collectives
    bool ownsWin = false;
    int opt_comm_p2p = 1;                  // 1: small collectives over the peer window when it is open, 0: always NCCL

    // bookkeeping
    std::map<std::string, double> timers;
    int64_t launches = 0;

    int64_t nvar() const { return nDofs * N; }
};

namespace mfem {

struct ScopedTimer {           // CUDA-event section timer accumulating into ctx->timers
    mfem_b200_ctx *c;
    std::string name;
    cudaEvent_t e0, e1;
    ScopedTimer(mfem_b200_ctx *ctx, const char *n) : c(ctx), name(n) {
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, c->stream);
    }
    double stop() {
        if (!e0) return 0.0;
        cudaEventRecord(e1, c->stream);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        e0 = e1 = nullptr;
        c->timers[name] += ms * 1e-3;
        return ms * 1e-3;
    }
    ~ScopedTimer() { stop(); }
};

// Value layout of the block-CSR matrix ("row-plane"): the N*N*n values of a block row with n
// blocks starting at block b0 are stored as N planes, plane r holding scalar row r of the block
// row as [block j][component c]:   index = N*N*b0 + r*(N*n) + N*j + c.
// A warp streaming one scalar row reads contiguous doubles (coalesced at any block size), and
// the block column index j = f / N is shared by the N lanes that need the same x block.
template <int N>
__host__ __device__ __forceinline__ int64_t val_index(int64_t b0, int64_t n, int64_t j, int r, int c) {
    return (int64_t)N * N * b0 + (int64_t)r * (N * n) + (int64_t)N * j + c;
}

constexpr int kSpmvTileWindow = 512;  // blocks per tile window of the TMA-ring SpMV (solver.cu kTmaWindow)
constexpr int kAsmChunk = 32;       // element incidences per warp job of the owner-gather assembly
constexpr int kBlkChunk = 256;      // BSR blocks per CTA of the block-owner assembly (one thread per block)
constexpr int kSegSlots = 384;      // thread slots (segments) per chunk: one full round of kBlkChunk threads + a partial one
constexpr int kGeomCap = 96;        // element geometry records staged in shared memory per chunk (TMA)
constexpr uint32_t kPlanSentinel = 0xffffffffu;

inline int grid_for(int64_t n, int block) { return static_cast<int>((n + block - 1) / block); }

// setup.cu
void setup_mesh(mfem_b200_ctx *c, int dim, int degree, int64_t nNodes, const double *nodes, int64_t nElems,
                const int32_t *elemNodes, const int64_t *dofForNode, int64_t nDofs);
void compute_geometry(mfem_b200_ctx *c);
void build_pattern(mfem_b200_ctx *c);
void finish_pattern(mfem_b200_ctx *c);
void upload_external_bsr(mfem_b200_ctx *c, int dim, int64_t nb, const std::vector<int64_t> &rowptr,
                         const std::vector<int32_t> &colidx, const std::vector<double> &blocks);
void build_coloring(mfem_b200_ctx *c);
void build_mf_plan(mfem_b200_ctx *c);               // element order + incidence list of the matrix-free operator
void build_mf_chunks(mfem_b200_ctx *c);             // chunk-local DoF tables of the matrix-free operator
// assemble.cu
void assemble_values(mfem_b200_ctx *c);
void ensure_packed_geometry(mfem_b200_ctx *c);     // geomP from geom (128-byte records)
// solver.cu
void spmv_plain(mfem_b200_ctx *c, const double *x_int, double *y_int);
void build_preconditioner(mfem_b200_ctx *c);
void pcg_solve(mfem_b200_ctx *c, const double *f_ext_dev, double *u_ext_dev, double rtol, int maxIters,
               mfem_b200_solve_info *info);
void ensure_work(mfem_b200_ctx *c);
bool pcg_solve_multi(mfem_b200_ctx *c, int nrhs, const double *f_int, double *u_int, double rtol, int maxIters,
                     mfem_b200_solve_info *info);
void free_work_multi(mfem_b200_ctx *c);
void free_coarse_space(mfem_b200_ctx *c);
bool will_use_coarse(mfem_b200_ctx *c);
double time_spmv(mfem_b200_ctx *c, int iters);
double time_operator(mfem_b200_ctx *c, int iters, int *matrixFree, double *secondsParts);   // what the PCG launches per iteration for K*p
int64_t get_coarse_array(mfem_b200_ctx *c, const std::string &name, double *out, int64_t capacity);
void apply_preconditioner(mfem_b200_ctx *c, const double *r_int, double *z_int, double *rz);
// comm.cu
void halo_exchange_add(mfem_b200_ctx *c, double *vec_int, int width);   // no-op on one rank
void allreduce_sum(mfem_b200_ctx *c, const double *in, double *out, int n);
void allgather_inplace(mfem_b200_ctx *c, double *buf, int count);
int comm_peer_error(mfem_b200_ctx *c);                             // non-zero after a peer-window spin timed out
bool comm_uses_peer_window(mfem_b200_ctx *c);
bool halo_exchange_add_allreduce1(mfem_b200_ctx *c, double *vec_int, int width, double *scalar);   // fused; false = not available
struct PeerWin;
const PeerWin *comm_peer_window(mfem_b200_ctx *c);                 // nullptr when the collectives go through NCCL   // slices of `count` doubles in rank order
const uint8_t *halo_owned(mfem_b200_ctx *c);
const uint8_t *halo_shared(mfem_b200_ctx *c);      // [nDofs] 1 = DoF shared with another rank
// aux.cu
void permute_to_internal(mfem_b200_ctx *c, const double *ext, double *in);   // per-DoF vectors
void permute_to_external(mfem_b200_ctx *c, const double *in, double *ext);
void apply_K_nodes(mfem_b200_ctx *c, const double *u_nodes_dev, double *Ku_nodes_dev);
void const_strain_load(mfem_b200_ctx *c, const double *epsFlatHost, double *f_ext_dev);
void avg_strain_stress(mfem_b200_ctx *c, const double *u_nodes_dev, double *strain_dev, double *stress_dev);
void export_bsr(mfem_b200_ctx *c, int64_t *rowptr, int32_t *colidx, double *vals);
// shape.cu (discrete shape derivatives; device pointers, per node / per DoF in the caller's numbering)
void apply_delta_K(mfem_b200_ctx *c, const double *u_nodes, const double *deltaP_nodes, double *out_dofs);
void delta_const_strain_load(mfem_b200_ctx *c, const double *epsFlatHost, const double *deltaP_nodes, double *out_dofs);
void delta_avg_strain(mfem_b200_ctx *c, const double *u_nodes, const double *du_nodes, const double *deltaP_nodes, double *out);

}  // namespace mfem

// K3-K6: block-CSR SpMV, block-Jacobi preconditioner, fused PCG vector kernels.
//
// Replaces SPSDSystem::fixVariables / solve (SparseMatrices.hh:2389-2500, 2516-2606) and the
// CholmodFactorizer behind it (SparseMatrices.hh:1984-2296): the reduced SPD system
//     K_ff u_f = f_f - K_fc u_c ,  u_c = fixed values
// is solved by preconditioned conjugate gradients on FULL-LENGTH vectors whose fixed
// components are held at zero (CG on the free subspace); K itself is never modified, the
// mask is applied where the SpMV writes its result.  The preconditioner is block-Jacobi on the
// dim x dim diagonal blocks taken after masking.
//
// All reductions are two-stage with a fixed order (per-CTA partials, then the last CTA to
// finish sums them in index order): results are bit-reproducible run to run.
#include "core.cuh"
#include "peer.cuh"

#include <cusolverDn.h>   // types only: the library is loaded with dlopen by coarse.inl (optional two-level preconditioner)
#include <dlfcn.h>

#include <cmath>
#include <cstring>

namespace mfem {

constexpr int kSpmvThreads = 256;
constexpr int kVecThreads = 256;
constexpr int kMaxPartials = 4096;      // upper bound on CTAs of any reducing kernel

// device scalar slots (PcgWork::scal)
enum {
    S_RZ = 0,      // r.z of the current iterate
    S_PAP = 1,     // p.Ap
    S_RZ_NEW = 2,  // r.z after the update     } adjacent: one 2-element all-reduce
    S_RR = 3,      // r.r                      }
    S_BB = 4,      // b.b
    S_TOL2 = 5,    // rtol^2
    S_COUNT = 8
};
// Multi-GPU: kernels write their OWNED-DoF partial sums to PcgWork::dotLoc and an ncclAllReduce
// delivers the global value into the scal slot; single-GPU: kernels write the scal slot directly.
enum { DL_PAP = 0, DL_RZ_NEW = 1, DL_RR = 2, DL_COUNT = 4 };
// status slots (PcgWork::status)
enum { ST_ITERS = 0, ST_STATE = 1 };   // state: 0 running, 1 converged, 2 breakdown (p'Ap<=0), 3 nan

// ---------------------------------------------------------------------------
// Deterministic block reduction + "last block finalises" pattern.
// ---------------------------------------------------------------------------
template <int NV>
__device__ __forceinline__ void block_reduce_store(double (&v)[NV], double *partials /* [NV][gridDim.x] */) {
    __shared__ double red[NV][32];                  // up to 1024 threads per block
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double x = v[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
        if (lane == 0) red[k][warp] = x;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[threadIdx.x][w];
        partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

// returns true in exactly one block (the last to arrive); that block then sees all partials
__device__ __forceinline__ bool last_block(unsigned *ticket) {
    __shared__ bool isLast;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(ticket, 1u);
        isLast = (t == gridDim.x - 1);
        if (isLast) *ticket = 0u;        // re-arm for the next launch
    }
    __syncthreads();
    if (isLast) __threadfence();
    return isLast;
}

// sum partials[k][0..n) in fixed order with one warp-shaped tree (whole block participates)
__device__ __forceinline__ double final_sum(const double *partials, int n) {
    __shared__ double red2[32];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += partials[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red2[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red2[w];
    return tot;
}

// ---------------------------------------------------------------------------
// K3  bsr_spmv: y = mask(K x) [, dot = x.y]
// LPR lanes (8/16/32) cooperate on one block row.  With the row-plane value layout
// (core.cuh val_index) lane f of a row reads element f of each of the N scalar rows: N fully
// coalesced streaming loads per step, ONE gathered x load shared by the N lanes of a block
// column, N FMAs -- no per-element index arithmetic beyond f / N.  The N partial sums are
// folded across the lanes so that the reduction costs 2 + log2(LPR/2) shuffles instead of
// N * log2(LPR).
// ---------------------------------------------------------------------------
template <int N, int LPR>
__device__ __forceinline__ double fold_reduce(double (&a)[N], int sl) {
    constexpr unsigned FULL = 0xffffffffu;
    constexpr int half = LPR / 2;
    const bool upper = (sl & half) != 0;
    double v;
    if (N == 3) {
        constexpr int q = LPR / 4;
        // step A: lower half keeps (a0, a1), upper half keeps a2
        const double r1 = __shfl_xor_sync(FULL, upper ? a[0] : a[2], half);
        const double r2 = __shfl_xor_sync(FULL, a[1], half);
        if (upper) a[2] += r1; else { a[0] += r1; a[1] += r2; }
        // step B: lower half splits a0 / a1 between its quarters; upper half keeps folding a2
        const bool bq = (sl & q) != 0;
        const double send = upper ? a[2] : (bq ? a[0] : a[1]);
        const double r3 = __shfl_xor_sync(FULL, send, q);
        v = upper ? a[2] + r3 : (bq ? a[1] + r3 : a[0] + r3);
#pragma unroll
        for (int o = q / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    } else {
        const double r1 = __shfl_xor_sync(FULL, upper ? a[0] : a[1], half);
        v = upper ? a[1] + r1 : a[0] + r1;
#pragma unroll
        for (int o = half / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    }
    return v;   // component k of the row lives in lane sl == owner_lane<N,LPR>(k)
}
template <int N, int LPR>
__device__ __forceinline__ int owner_component(int sl) {   // -1 if this lane owns no result
    if (N == 3) return sl == 0 ? 0 : (sl == LPR / 4 ? 1 : (sl == LPR / 2 ? 2 : -1));
    return sl == 0 ? 0 : (sl == LPR / 2 ? 1 : -1);
}

// L2 residency hints: the matrix (values + column indices) is streamed exactly once per SpMV
// and must not displace the vectors, which are re-read ~27 times through the gather.
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ld_stream_f64(const double *p, uint64_t pol) {
    double v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ int ld_stream_s32(const int32_t *p, uint64_t pol) {
    int v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ double ld_keep_f64(const double *p, uint64_t pol) {
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    return v;
}

// All streaming loads of one chunk (U = 3 sub-chunks: 3 column indices + 3*N matrix scalars per
// lane) are issued from ONE asm block: the hardware issues in order, so nothing that consumes a
// load result may sit between them.  Lanes past the end of the row are predicated off and get 0.
template <int N, int LPR>
struct ChunkLoader;

#define MFEM_LD_S32 "ld.global.nc.L1::no_allocate.L2::cache_hint.s32"
#define MFEM_LD_F64 "ld.global.nc.L1::no_allocate.L2::cache_hint.f64"

template <int LPR>
struct ChunkLoader<3, LPR> {
    static __device__ __forceinline__ void run(const int32_t *c0, const int32_t *c1, const int32_t *c2,
                                               const double *p0, const double *p1, const double *p2, int rem,
                                               uint64_t pol, int (&col)[3], double (&a)[3][3]) {
        asm volatile(
            "{\n\t"
            ".reg .pred q0, q1, q2;\n\t"
            "setp.gt.s32 q0, %18, 0;\n\t"
            "setp.gt.s32 q1, %18, %20;\n\t"
            "setp.gt.s32 q2, %18, %21;\n\t"
            "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0;\n\t"
            "mov.f64 %3, 0d0000000000000000; mov.f64 %4, 0d0000000000000000; mov.f64 %5, 0d0000000000000000;\n\t"
            "mov.f64 %6, 0d0000000000000000; mov.f64 %7, 0d0000000000000000; mov.f64 %8, 0d0000000000000000;\n\t"
            "mov.f64 %9, 0d0000000000000000; mov.f64 %10, 0d0000000000000000; mov.f64 %11, 0d0000000000000000;\n\t"
            "@q0 " MFEM_LD_S32 " %0, [%12], %19;\n\t"
            "@q1 " MFEM_LD_S32 " %1, [%13], %19;\n\t"
            "@q2 " MFEM_LD_S32 " %2, [%14], %19;\n\t"
            "@q0 " MFEM_LD_F64 " %3, [%15], %19;\n\t"
            "@q0 " MFEM_LD_F64 " %4, [%16], %19;\n\t"
            "@q0 " MFEM_LD_F64 " %5, [%17], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %6, [%15+%22], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %7, [%16+%22], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %8, [%17+%22], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %9, [%15+%23], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %10, [%16+%23], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %11, [%17+%23], %19;\n\t"
            "}"
            : "=r"(col[0]), "=r"(col[1]), "=r"(col[2]), "=d"(a[0][0]), "=d"(a[0][1]), "=d"(a[0][2]), "=d"(a[1][0]),
              "=d"(a[1][1]), "=d"(a[1][2]), "=d"(a[2][0]), "=d"(a[2][1]), "=d"(a[2][2])
            : "l"(c0), "l"(c1), "l"(c2), "l"(p0), "l"(p1), "l"(p2), "r"(rem), "l"(pol), "n"(LPR), "n"(2 * LPR),
              "n"(8 * LPR), "n"(16 * LPR)
            : "memory");
    }
};

template <int LPR>
struct ChunkLoader<2, LPR> {
    static __device__ __forceinline__ void run(const int32_t *c0, const int32_t *c1, const int32_t *c2,
                                               const double *p0, const double *p1, const double * /*unused*/, int rem,
                                               uint64_t pol, int (&col)[3], double (&a)[3][2]) {
        asm volatile(
            "{\n\t"
            ".reg .pred q0, q1, q2;\n\t"
            "setp.gt.s32 q0, %14, 0;\n\t"
            "setp.gt.s32 q1, %14, %16;\n\t"
            "setp.gt.s32 q2, %14, %17;\n\t"
            "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0;\n\t"
            "mov.f64 %3, 0d0000000000000000; mov.f64 %4, 0d0000000000000000; mov.f64 %5, 0d0000000000000000;\n\t"
            "mov.f64 %6, 0d0000000000000000; mov.f64 %7, 0d0000000000000000; mov.f64 %8, 0d0000000000000000;\n\t"
            "@q0 " MFEM_LD_S32 " %0, [%9], %15;\n\t"
            "@q1 " MFEM_LD_S32 " %1, [%10], %15;\n\t"
            "@q2 " MFEM_LD_S32 " %2, [%11], %15;\n\t"
            "@q0 " MFEM_LD_F64 " %3, [%12], %15;\n\t"
            "@q0 " MFEM_LD_F64 " %4, [%13], %15;\n\t"
            "@q1 " MFEM_LD_F64 " %5, [%12+%18], %15;\n\t"
            "@q1 " MFEM_LD_F64 " %6, [%13+%18], %15;\n\t"
            "@q2 " MFEM_LD_F64 " %7, [%12+%19], %15;\n\t"
            "@q2 " MFEM_LD_F64 " %8, [%13+%19], %15;\n\t"
            "}"
            : "=r"(col[0]), "=r"(col[1]), "=r"(col[2]), "=d"(a[0][0]), "=d"(a[0][1]), "=d"(a[1][0]), "=d"(a[1][1]),
              "=d"(a[2][0]), "=d"(a[2][1])
            : "l"(c0), "l"(c1), "l"(c2), "l"(p0), "l"(p1), "r"(rem), "l"(pol), "n"(LPR), "n"(2 * LPR), "n"(8 * LPR),
              "n"(16 * LPR)
            : "memory");
    }
};

// the three gathered x loads of a chunk, again back to back
__device__ __forceinline__ void gather3(const double *x0, const double *x1, const double *x2, uint64_t pol,
                                        double (&xv)[3]) {
    asm volatile(
        "ld.global.nc.L2::cache_hint.f64 %0, [%3], %6;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %1, [%4], %6;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %2, [%5], %6;"
        : "=d"(xv[0]), "=d"(xv[1]), "=d"(xv[2])
        : "l"(x0), "l"(x1), "l"(x2), "l"(pol)
        : "memory");
}

// PF = true: the first lane of a row also issues two bulk L2 prefetches (cp.async.bulk.prefetch.L2: values, column
// indices) per row iteration for the row this warp will stream in its NEXT iteration (its extent is fetched two
// iterations ahead).  The prefetch needs no destination registers, so the DRAM round trip of row k+1 overlaps all of
// row k's work at no occupancy cost; the streaming loads then hit L2.
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

template <int N, int LPR, bool MASKED, bool DOT, bool PF, int MINB = 4>
__global__ void __launch_bounds__(MINB == 1 ? 1024 : kSpmvThreads, MINB)   // MINB = 4: <= 64 registers, 32 resident warps per SM;
                                                                          // MINB = 1: ONE 1024-thread CTA per SM (same registers and
                                                                          // occupancy), so the 32 warps of an SM work on 32 consecutive rows
k_bsr_spmv(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
           const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
           const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut,
           const int *status) {
    constexpr int NN = N * N;
    constexpr int RPW = 32 / LPR;                     // rows per warp
    constexpr int U = 3;                              // sub-chunks whose loads are issued together
    constexpr int CH = U * LPR;                       // scalars of one scalar row per chunk
    static_assert(CH % N == 0, "a chunk must cover whole blocks");
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, sl = lane % LPR;
    const int64_t warpGlobal = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int comp = owner_component<N, LPR>(sl);
    const unsigned groupMask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    // lane-constant decomposition of its U scalars of a chunk into (block, component): no
    // division inside the row loop
    int jj[U], cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int f = sl + u * LPR;
        jj[u] = f / N;
        cc[u] = f - jj[u] * N;
    }
    double dot = 0.0;
    // software pipeline over rows: the next row's extent is fetched while this one is computed (PF: two rows ahead)
    int64_t row = warpGlobal * RPW + sub;
    const int64_t rowStride = nWarps * RPW;
    int64_t nb0 = 0, nb1 = 0, pb0 = 0;
    int pn = 0;
    if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
    if (PF && row + rowStride < nb) { pb0 = rowptr[row + rowStride]; pn = (int)(rowptr[row + rowStride + 1] - pb0); }
    for (int64_t rowBase = warpGlobal * RPW; rowBase < nb; rowBase += rowStride) {
        const int64_t b0 = nb0;
        const int L = (int)(nb1 - nb0) * N;
        const int64_t thisRow = row;
        row += rowStride;
        if (PF) {
            // row k+1: extent known since the previous iteration -> prefetch its lines now; row k+2: fetch the extent
            nb0 = pb0; nb1 = pb0 + pn;
            if (pn > 0 && sl == 0) {
                // ONE bulk L2 prefetch per array and row, issued by the row's first lane: handled by the copy engine, no
                // L1 wavefronts (a per-lane prefetch.global.L2 of 32 lines costs 32 L1 transactions per row and made
                // the kernel 10 % slower)
                const uintptr_t v0 = reinterpret_cast<uintptr_t>(vals + pb0 * NN) & ~uintptr_t(15);
                const uintptr_t v1 = (reinterpret_cast<uintptr_t>(vals + (pb0 + pn) * NN) + 15) & ~uintptr_t(15);
                const uintptr_t c0 = reinterpret_cast<uintptr_t>(colidx + pb0) & ~uintptr_t(15);
                const uintptr_t c1 = (reinterpret_cast<uintptr_t>(colidx + pb0 + pn) + 15) & ~uintptr_t(15);
                bulk_prefetch_l2(reinterpret_cast<const void *>(v0), (uint32_t)(v1 - v0));
                bulk_prefetch_l2(reinterpret_cast<const void *>(c0), (uint32_t)(c1 - c0));
            }
            pb0 = 0; pn = 0;
            if (row + rowStride < nb) { pb0 = rowptr[row + rowStride]; pn = (int)(rowptr[row + rowStride + 1] - pb0); }
        } else {
            nb0 = nb1 = 0;
            if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
        }
        const double *v = vals + b0 * NN + sl;        // plane 0, this lane's first scalar
        const int32_t *ci = colidx + b0;
        double acc[N];
#pragma unroll
        for (int r = 0; r < N; ++r) acc[r] = 0.0;
        for (int base = 0; base < L; base += CH, v += CH, ci += CH / N) {
            const int rem = L - base - sl;            // scalar f = base + sl + u*LPR is valid iff u*LPR < rem
            int col[U];
            double a[U][N], xv[U];
            ChunkLoader<N, LPR>::run(ci + jj[0], ci + jj[1], ci + jj[2], v, v + L, v + 2 * L, rem, polStream, col, a);
            // ptxas would otherwise hoist the dependent gather (and the FMAs behind it) in between the
            // streaming loads; with in-order issue that stalls the warp before most loads are out.
            // A warp-level barrier is a scheduling fence for memory instructions.
            __syncwarp(groupMask);
            gather3(x + (col[0] * N + cc[0]), x + (col[1] * N + cc[1]), x + (col[2] * N + cc[2]), polKeep, xv);
            __syncwarp(groupMask);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < N; ++r) acc[r] = fma(a[u][r], xv[u], acc[r]);
        }
        const double out0 = fold_reduce<N, LPR>(acc, sl);
        if (comp >= 0 && thisRow < nb) {
            double out = out0;
            if (MASKED && fixedMask[thisRow * N + comp]) out = 0.0;
            y[thisRow * N + comp] = out;
            if (DOT) dot += out * x[thisRow * N + comp];
        }
    }
    if (DOT) {
        double v1[1] = {dot};
        block_reduce_store<1>(v1, partials);
        if (last_block(ticket)) {
            const double s = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) dotOut[0] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// K3p  bsr_spmv_pipe: the same warp-per-row SpMV with the COLUMN INDICES software-pipelined one
// work item (row chunk) ahead.  In k_bsr_spmv an item costs two dependent memory round trips
// (matrix values + indices, then the gathered x).  Here the indices of item k+1 are fetched
// together with the values of item k, so the x gather of item k (whose indices arrived with
// item k-1) is issued in the same batch as its values: ONE round trip per item, twice the bytes
// in flight per resident warp at the same register budget.  Row extents are prefetched two rows
// ahead for the same reason.  LPR = 32 only (one block row per warp at a time).
// ---------------------------------------------------------------------------
template <int N>
struct PipeLoader;

template <>
struct PipeLoader<3> {
    // values of the current item (predicated by rem) + column indices of the NEXT item (remx)
    static __device__ __forceinline__ void run(const int32_t *c0, const int32_t *c1, const int32_t *c2,
                                               const double *p0, const double *p1, const double *p2, int rem, int remx,
                                               uint64_t pol, int (&col)[3], double (&a)[3][3]) {
        asm volatile(
            "{\n\t"
            ".reg .pred q0, q1, q2, n0, n1, n2;\n\t"
            "setp.gt.s32 q0, %18, 0;\n\t"
            "setp.gt.s32 q1, %18, %20;\n\t"
            "setp.gt.s32 q2, %18, %21;\n\t"
            "setp.gt.s32 n0, %24, 0;\n\t"
            "setp.gt.s32 n1, %24, %20;\n\t"
            "setp.gt.s32 n2, %24, %21;\n\t"
            "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0;\n\t"
            "mov.f64 %3, 0d0000000000000000; mov.f64 %4, 0d0000000000000000; mov.f64 %5, 0d0000000000000000;\n\t"
            "mov.f64 %6, 0d0000000000000000; mov.f64 %7, 0d0000000000000000; mov.f64 %8, 0d0000000000000000;\n\t"
            "mov.f64 %9, 0d0000000000000000; mov.f64 %10, 0d0000000000000000; mov.f64 %11, 0d0000000000000000;\n\t"
            "@q0 " MFEM_LD_F64 " %3, [%15], %19;\n\t"
            "@q0 " MFEM_LD_F64 " %4, [%16], %19;\n\t"
            "@q0 " MFEM_LD_F64 " %5, [%17], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %6, [%15+%22], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %7, [%16+%22], %19;\n\t"
            "@q1 " MFEM_LD_F64 " %8, [%17+%22], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %9, [%15+%23], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %10, [%16+%23], %19;\n\t"
            "@q2 " MFEM_LD_F64 " %11, [%17+%23], %19;\n\t"
            "@n0 " MFEM_LD_S32 " %0, [%12], %19;\n\t"
            "@n1 " MFEM_LD_S32 " %1, [%13], %19;\n\t"
            "@n2 " MFEM_LD_S32 " %2, [%14], %19;\n\t"
            "}"
            : "=r"(col[0]), "=r"(col[1]), "=r"(col[2]), "=d"(a[0][0]), "=d"(a[0][1]), "=d"(a[0][2]), "=d"(a[1][0]),
              "=d"(a[1][1]), "=d"(a[1][2]), "=d"(a[2][0]), "=d"(a[2][1]), "=d"(a[2][2])
            : "l"(c0), "l"(c1), "l"(c2), "l"(p0), "l"(p1), "l"(p2), "r"(rem), "l"(pol), "n"(32), "n"(64), "n"(256),
              "n"(512), "r"(remx)
            : "memory");
    }
};

template <>
struct PipeLoader<2> {
    static __device__ __forceinline__ void run(const int32_t *c0, const int32_t *c1, const int32_t *c2,
                                               const double *p0, const double *p1, const double * /*unused*/, int rem,
                                               int remx, uint64_t pol, int (&col)[3], double (&a)[3][2]) {
        asm volatile(
            "{\n\t"
            ".reg .pred q0, q1, q2, n0, n1, n2;\n\t"
            "setp.gt.s32 q0, %14, 0;\n\t"
            "setp.gt.s32 q1, %14, %16;\n\t"
            "setp.gt.s32 q2, %14, %17;\n\t"
            "setp.gt.s32 n0, %20, 0;\n\t"
            "setp.gt.s32 n1, %20, %16;\n\t"
            "setp.gt.s32 n2, %20, %17;\n\t"
            "mov.b32 %0, 0; mov.b32 %1, 0; mov.b32 %2, 0;\n\t"
            "mov.f64 %3, 0d0000000000000000; mov.f64 %4, 0d0000000000000000; mov.f64 %5, 0d0000000000000000;\n\t"
            "mov.f64 %6, 0d0000000000000000; mov.f64 %7, 0d0000000000000000; mov.f64 %8, 0d0000000000000000;\n\t"
            "@q0 " MFEM_LD_F64 " %3, [%12], %15;\n\t"
            "@q0 " MFEM_LD_F64 " %4, [%13], %15;\n\t"
            "@q1 " MFEM_LD_F64 " %5, [%12+%18], %15;\n\t"
            "@q1 " MFEM_LD_F64 " %6, [%13+%18], %15;\n\t"
            "@q2 " MFEM_LD_F64 " %7, [%12+%19], %15;\n\t"
            "@q2 " MFEM_LD_F64 " %8, [%13+%19], %15;\n\t"
            "@n0 " MFEM_LD_S32 " %0, [%9], %15;\n\t"
            "@n1 " MFEM_LD_S32 " %1, [%10], %15;\n\t"
            "@n2 " MFEM_LD_S32 " %2, [%11], %15;\n\t"
            "}"
            : "=r"(col[0]), "=r"(col[1]), "=r"(col[2]), "=d"(a[0][0]), "=d"(a[0][1]), "=d"(a[1][0]), "=d"(a[1][1]),
              "=d"(a[2][0]), "=d"(a[2][1])
            : "l"(c0), "l"(c1), "l"(c2), "l"(p0), "l"(p1), "r"(rem), "l"(pol), "n"(32), "n"(64), "n"(256), "n"(512),
              "r"(remx)
            : "memory");
    }
};

template <int N, bool MASKED, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, 4)
k_bsr_spmv_pipe(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
                const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut,
                const int *status) {
    constexpr int NN = N * N;
    constexpr int LPR = 32, U = 3, CH = U * LPR;
    static_assert(CH % N == 0, "a chunk must cover whole blocks");
    __shared__ double sZero;
    if (status && status[ST_STATE] != 0) return;
    if (threadIdx.x == 0) sZero = 0.0;
    __syncthreads();
    const int sl = threadIdx.x & 31;
    const int64_t warpGlobal = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t rowStride = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int comp = owner_component<N, LPR>(sl);
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    int jj[U], cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int f = sl + u * LPR;
        jj[u] = f / N;
        cc[u] = f - jj[u] * N;
    }
    double dot = 0.0;
    int64_t row = warpGlobal;
    // extents of the current row and of the next one
    int64_t b0c = 0, b0n = 0;
    int Lc = 0, Ln = 0;
    if (row < nb) { const int64_t a0 = rowptr[row], a1 = rowptr[row + 1]; b0c = a0; Lc = (int)(a1 - a0) * N; }
    if (row + rowStride < nb) {
        const int64_t a0 = rowptr[row + rowStride], a1 = rowptr[row + rowStride + 1];
        b0n = a0; Ln = (int)(a1 - a0) * N;
    }
    // column indices of the first item
    int colc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) colc[u] = (u * LPR < Lc - sl) ? colidx[b0c + jj[u]] : 0;
    int base = 0;
    double acc[N];
#pragma unroll
    for (int r = 0; r < N; ++r) acc[r] = 0.0;
    while (row < nb) {
        const bool lastOfRow = base + CH >= Lc;
        // the item after this one: next chunk of this row, or the first chunk of the next row
        const int64_t b0x = lastOfRow ? b0n : b0c;
        const int Lx = lastOfRow ? Ln : Lc;
        const int basex = lastOfRow ? 0 : base + CH;
        // leaving the row: fetch the extent of the row after next (consumed one item later)
        int64_t b0f = 0;
        int Lf = 0;
        if (lastOfRow) {
            const int64_t r2 = row + 2 * rowStride;
            if (r2 < nb) { const int64_t a0 = rowptr[r2], a1 = rowptr[r2 + 1]; b0f = a0; Lf = (int)(a1 - a0) * N; }
        }
        const double *v = vals + b0c * NN + base + sl;
        const int32_t *cn = colidx + b0x + basex / N;
        int coln[U];
        double a[U][N], xv[U];
        PipeLoader<N>::run(cn + jj[0], cn + jj[1], cn + jj[2], v, v + Lc, v + 2 * Lc, Lc - base - sl, Lx - basex - sl,
                           polStream, coln, a);
        gather3(x + (colc[0] * N + cc[0]), x + (colc[1] * N + cc[1]), x + (colc[2] * N + cc[2]), polKeep, xv);
        // Scheduling fence.  ptxas would otherwise start the FMA chain as soon as the first two loads
        // are back, in between the remaining loads of the batch, and with in-order issue the warp
        // then stalls with most of its loads not yet sent.  Memory instructions do not cross a warp
        // barrier, arithmetic does -- so the FMAs are made to depend on a (zero-valued) shared-memory
        // load placed after the barrier.
        __syncwarp();
        const double z = *reinterpret_cast<volatile double *>(&sZero);
#pragma unroll
        for (int u = 0; u < U; ++u) xv[u] += z;
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
            for (int r = 0; r < N; ++r) acc[r] = fma(a[u][r], xv[u], acc[r]);
#pragma unroll
        for (int u = 0; u < U; ++u) colc[u] = coln[u];
        if (lastOfRow) {
            const double out0 = fold_reduce<N, LPR>(acc, sl);
            if (comp >= 0) {
                double out = out0;
                if (MASKED && fixedMask[row * N + comp]) out = 0.0;
                y[row * N + comp] = out;
                if (DOT) dot += out * x[row * N + comp];
            }
#pragma unroll
            for (int r = 0; r < N; ++r) acc[r] = 0.0;
            row += rowStride;
            b0c = b0n; Lc = Ln;
            b0n = b0f; Ln = Lf;
            base = 0;
        } else {
            base += CH;
        }
    }
    if (DOT) {
        double v1[1] = {dot};
        block_reduce_store<1>(v1, partials);
        if (last_block(ticket)) {
            const double s = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) dotOut[0] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// K3s  bsr_spmv_sym: y += K x reading only the UPPER-triangular tail of every block row.
// K is symmetric (K_ji = K_ij^T), and the row-plane layout stores a row's blocks in column order, so the
// blocks with column >= row are a contiguous tail of each plane: the kernel streams that tail only --
// half the matrix bytes -- and applies every off-diagonal block twice:
//     y_i += K_ij x_j          (row part, reduced across the lanes as in k_bsr_spmv)
//     y_j += K_ij^T x_i        (transposed part: one fire-and-forget red.global.add.f64 per (block, component))
// y must be zero on entry; every update of y is an atomic add (rows receive transposed contributions from
// other warps).  Summation order therefore varies from run to run: results agree to rounding, not
// bit for bit (the assembly stays bit-reproducible).  Masking of fixed rows and the p.Ap dot product
// move to k_mask_dot.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f64(double *p, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

template <int N, int LPR>
__global__ void __launch_bounds__(kSpmvThreads, 4)
k_bsr_spmv_sym(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ upperStart,
               const int32_t *__restrict__ colidx, const double *__restrict__ vals, const double *__restrict__ x,
               double *__restrict__ y, const int *status) {
    constexpr int NN = N * N;
    constexpr int RPW = 32 / LPR;
    constexpr int U = 3;
    constexpr int CH = U * LPR;
    static_assert(CH % N == 0, "a chunk must cover whole blocks");
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, sl = lane % LPR;
    const int64_t warpGlobal = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int comp = owner_component<N, LPR>(sl);
    const unsigned groupMask = (LPR == 32) ? 0xffffffffu : (((1u << LPR) - 1u) << (sub * LPR));
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    int jj[U], cc[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int f = sl + u * LPR;
        jj[u] = f / N;
        cc[u] = f - jj[u] * N;
    }
    int64_t row = warpGlobal * RPW + sub;
    const int64_t rowStride = nWarps * RPW;
    int64_t nb0 = 0, nb1 = 0;
    int us1 = 0;
    if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; us1 = upperStart[row]; }
    for (int64_t rowBase = warpGlobal * RPW; rowBase < nb; rowBase += rowStride) {
        const int64_t b0 = nb0;
        const int n = (int)(nb1 - nb0), us = us1;
        const int L = n * N;                      // plane stride of the full row
        const int Lu = (n - us) * N;              // scalars of the upper tail per plane
        const int64_t thisRow = row;
        row += rowStride;
        nb0 = nb1 = 0; us1 = 0;
        if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; us1 = upperStart[row]; }
        const double *v = vals + b0 * NN + N * us + sl;
        const int32_t *ci = colidx + b0 + us;
        double xi[N];
#pragma unroll
        for (int r = 0; r < N; ++r) xi[r] = (thisRow < nb) ? x[thisRow * N + r] : 0.0;
        double acc[N];
#pragma unroll
        for (int r = 0; r < N; ++r) acc[r] = 0.0;
        for (int base = 0; base < Lu; base += CH, v += CH, ci += CH / N) {
            const int rem = Lu - base - sl;
            int col[U];
            double a[U][N], xv[U];
            ChunkLoader<N, LPR>::run(ci + jj[0], ci + jj[1], ci + jj[2], v, v + L, v + 2 * L, rem, polStream, col, a);
            __syncwarp(groupMask);
            gather3(x + (col[0] * N + cc[0]), x + (col[1] * N + cc[1]), x + (col[2] * N + cc[2]), polKeep, xv);
            __syncwarp(groupMask);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                double t = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    acc[r] = fma(a[u][r], xv[u], acc[r]);
                    t = fma(a[u][r], xi[r], t);
                }
                // transposed part: this lane holds column (col, cc) of the block for all N rows
                if (u * LPR < rem && col[u] != (int)thisRow) red_add_f64(y + ((int64_t)col[u] * N + cc[u]), t);
            }
        }
        const double out0 = fold_reduce<N, LPR>(acc, sl);
        if (comp >= 0 && thisRow < nb) red_add_f64(y + (thisRow * N + comp), out0);
    }
}

// ---------------------------------------------------------------------------
// K3v  bsr_spmv_v2 (N = 3): 128-bit matrix loads.  16 lanes per block row, two rows per warp; lane l of a row owns the
// scalar PAIRS (32u + 2l, 32u + 2l + 1), u = 0..2, of each of the three planes of a 96-scalar chunk: 9 v2.f64 streaming
// loads, 6 column indices, 6 gathered x values, 18 FMAs per lane and chunk -- 10.5 load instructions per row chunk
// instead of 15, twice the bytes in flight per warp, and the fold / row bookkeeping shared by two rows.
// Needs 16-byte aligned planes: ALIGNED = true is for a value layout whose planes start on even indices (not the
// current one); ALIGNED = false rounds every address down to 16 bytes, which reads the wrong doubles whenever a plane
// starts on an odd index -- option spmv_kernel = 5 is a TIMING PROBE of the memory pipeline, not a usable SpMV.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ld_stream_v2(const double *p, uint64_t pol, int pred, double &a, double &b) {
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %3, 0;\n\t"
        "mov.f64 %0, 0d0000000000000000; mov.f64 %1, 0d0000000000000000;\n\t"
        "@q ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0, %1}, [%2], %4;\n\t}"
        : "=d"(a), "=d"(b)
        : "l"(p), "r"(pred), "l"(pol)
        : "memory");
}
__device__ __forceinline__ int ld_stream_s32_pred(const int32_t *p, uint64_t pol, int pred) {
    int v;
    asm volatile(
        "{\n\t.reg .pred q;\n\tsetp.ne.s32 q, %2, 0;\n\tmov.b32 %0, 0;\n\t"
        "@q ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %3;\n\t}"
        : "=r"(v)
        : "l"(p), "r"(pred), "l"(pol)
        : "memory");
    return v;
}
__device__ __forceinline__ double ld_keep_f64_v(const double *p, uint64_t pol) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol) : "memory");
    return v;
}

template <bool MASKED, bool DOT, bool ALIGNED>
__global__ void __launch_bounds__(kSpmvThreads, 3)
k_bsr_spmv_v2(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
              const double *__restrict__ vals, const double *__restrict__ x, double *__restrict__ y,
              const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut, const int *status) {
    constexpr int N = 3, NN = 9, LPR = 16, RPW = 2, U = 3, CH = U * 2 * LPR;     // 96 scalars of a plane per chunk
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane / LPR, sl = lane % LPR;
    const int64_t warpGlobal = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int comp = owner_component<N, LPR>(sl);
    const unsigned groupMask = ((1u << LPR) - 1u) << (sub * LPR);
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    int j0[U], c0[U], j1[U], c1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        const int f = 2 * LPR * u + 2 * sl;
        j0[u] = f / N; c0[u] = f - j0[u] * N;
        j1[u] = (f + 1) / N; c1[u] = f + 1 - j1[u] * N;
    }
    double dot = 0.0;
    int64_t row = warpGlobal * RPW + sub;
    const int64_t rowStride = nWarps * RPW;
    int64_t nb0 = 0, nb1 = 0;
    if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
    for (int64_t rowBase = warpGlobal * RPW; rowBase < nb; rowBase += rowStride) {
        const int64_t b0 = nb0;
        const int L = (int)(nb1 - nb0) * N;
        const int64_t thisRow = row;
        row += rowStride;
        nb0 = nb1 = 0;
        if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
        const double *v = vals + b0 * NN + 2 * sl;
        const int32_t *ci = colidx + b0;
        double acc[N] = {0.0, 0.0, 0.0};
        // the two half-warps may have rows of different lengths: loop to the longer one (shorter one is predicated off)
        const int Lmax = max(L, __shfl_xor_sync(0xffffffffu, L, LPR));
        for (int base = 0; base < Lmax; base += CH, v += CH, ci += CH / N) {
            const int rem = L - base - 2 * sl;        // pair u is valid iff 32u < rem (second scalar iff 32u + 1 < rem)
            int col0[U], col1[U];
            double a0[U][N], a1[U][N];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                col0[u] = ld_stream_s32_pred(ci + j0[u], polStream, 2 * LPR * u < rem);
                col1[u] = ld_stream_s32_pred(ci + j1[u], polStream, 2 * LPR * u + 1 < rem);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < N; ++r) {
                    const double *q = v + r * L + 2 * LPR * u;
                    if (!ALIGNED) q = reinterpret_cast<const double *>(reinterpret_cast<uintptr_t>(q) & ~uintptr_t(15));
                    ld_stream_v2(q, polStream, 2 * LPR * u < rem, a0[u][r], a1[u][r]);
                }
            __syncwarp(groupMask);
            double x0[U], x1[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                x0[u] = ld_keep_f64_v(x + (col0[u] * N + c0[u]), polKeep);
                x1[u] = ld_keep_f64_v(x + (col1[u] * N + c1[u]), polKeep);
            }
            __syncwarp(groupMask);
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int r = 0; r < N; ++r) acc[r] = fma(a1[u][r], x1[u], fma(a0[u][r], x0[u], acc[r]));
        }
        const double out0 = fold_reduce<N, LPR>(acc, sl);
        if (comp >= 0 && thisRow < nb) {
            double out = out0;
            if (MASKED && fixedMask[thisRow * N + comp]) out = 0.0;
            y[thisRow * N + comp] = out;
            if (DOT) dot += out * x[thisRow * N + comp];
        }
    }
    if (DOT) {
        double v1[1] = {dot};
        block_reduce_store<1>(v1, partials);
        if (last_block(ticket)) {
            const double s = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) dotOut[0] = s;
        }
    }
}

// after the symmetric SpMV (and, on several GPUs, the interface exchange): zero Ap on fixed variables and
// reduce p.Ap over the owned DoFs
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_mask_dot(int64_t nb, const uint8_t *__restrict__ fixedMask, const uint8_t *__restrict__ owned,
           const double *__restrict__ p, double *__restrict__ Ap, double *partials, unsigned *ticket, double *dotOut,
           const int *status) {
    if (status && status[ST_STATE] != 0) return;
    double acc[1] = {0.0};
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        const double wgt = (owned && !owned[i]) ? 0.0 : 1.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            if (fixedMask && fixedMask[i * N + k]) Ap[i * N + k] = 0.0;
            else acc[0] += wgt * p[i * N + k] * Ap[i * N + k];
        }
    }
    if (dotOut) {
        block_reduce_store<1>(acc, partials);
        if (last_block(ticket)) {
            const double s = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) dotOut[0] = s;
        }
    }
}

// ---------------------------------------------------------------------------
// K3b  bsr_spmv_tma: the same SpMV with the matrix stream decoupled from the compute warps.
// One persistent CTA per SM.  A producer warp walks the CTA's tiles (a tile = the consecutive
// block rows whose first block falls into a window of kTmaWindow blocks) and moves each tile's
// values and column indices -- two contiguous byte ranges -- into a ring of shared-memory stages
// with 1-D bulk TMA copies (cp.async.bulk, completion on an mbarrier, L2 evict-first policy).
// Consumer warps wait on the stage's "full" barrier, pull rows of the tile off a shared counter
// (dynamic balance), read values/indices with conflict-free LDS, gather x from L2, and signal the
// stage's "empty" barrier when they leave the tile.  HBM requests in flight are then bounded by
// the ring (kTmaStages x ~29 KB per SM), not by the registers of stalled warps.
// ---------------------------------------------------------------------------
constexpr int kTmaWindow = kSpmvTileWindow;     // blocks per tile window (tile table built in setup.cu)
constexpr int kTmaMaxRow = 128;                 // longest block row the ring is sized for
constexpr int kTmaStageBlocks = kTmaWindow + kTmaMaxRow;
constexpr int kTmaStages = 4;
constexpr int kTmaConsumers = 31;               // consumer warps
constexpr int kTmaThreads = 32 * (kTmaConsumers + 1);

template <int N>
struct TmaStage {
    double vals[kTmaStageBlocks * N * N + 2];   // +16 B: the copy starts at the 16 B boundary below the tile
    long long rows[kTmaWindow + 4];             // rowptr[r0 .. r1] of the tile (when it has <= kTmaWindow rows)
    long long hdr[4];                           // r0, r1, first block of the tile (written by the producer)
    int32_t cols[kTmaStageBlocks + 8];
};
template <int N>
struct TmaSmem {
    TmaStage<N> stage[kTmaStages];
    unsigned long long full[kTmaStages], empty[kTmaStages];
    int next[kTmaStages];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void tma_bulk_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar,
                                             uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}

template <int N, bool MASKED, bool DOT>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_bsr_spmv_tma(int64_t nb, int64_t nTiles, const int64_t *__restrict__ tileRow, const int64_t *__restrict__ rowptr,
               const int32_t *__restrict__ colidx, const double *__restrict__ vals, const double *__restrict__ x,
               double *__restrict__ y, const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket,
               double *dotOut, const int *status) {
    constexpr int NN = N * N;
    constexpr int U = 3, LPR = 32, CH = U * LPR;
    extern __shared__ __align__(128) unsigned char tma_smem_raw[];
    TmaSmem<N> &sm = *reinterpret_cast<TmaSmem<N> *>(tma_smem_raw);
    if (status && status[ST_STATE] != 0) return;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kTmaStages; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], kTmaConsumers);
            sm.next[s] = 0;
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // tiles of this CTA: t = blockIdx.x + k * gridDim.x
    const int64_t myTiles = (nTiles > blockIdx.x) ? (nTiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    double dot = 0.0;

    if (warp == kTmaConsumers) {
        // ===== producer warp: lanes fetch tile metadata in batches of 32, lane 0 issues the copies =====
        const uint64_t pol = l2_policy_evict_first();
        for (int64_t kb = 0; kb < myTiles; kb += 32) {
            const int64_t k = kb + lane;
            int64_t b0 = 0, b1 = 0, q0 = 0, q1 = 0;
            if (k < myTiles) {
                const int64_t t = blockIdx.x + k * gridDim.x;
                q0 = tileRow[t];
                q1 = tileRow[t + 1];
                b0 = rowptr[q0];
                b1 = rowptr[q1];
            }
            const int cnt = (int)((myTiles - kb < 32) ? (myTiles - kb) : 32);
            for (int i = 0; i < cnt; ++i) {
                const int64_t tb0 = __shfl_sync(0xffffffffu, b0, i), tb1 = __shfl_sync(0xffffffffu, b1, i);
                const int64_t tr0 = __shfl_sync(0xffffffffu, q0, i), tr1 = __shfl_sync(0xffffffffu, q1, i);
                if (lane == 0) {
                    const int64_t kk = kb + i;
                    const int s = (int)(kk % kTmaStages);
                    if (kk >= kTmaStages) mbar_wait(&sm.empty[s], (uint32_t)(((kk / kTmaStages) - 1) & 1));
                    sm.next[s] = 0;
                    sm.stage[s].hdr[0] = tr0; sm.stage[s].hdr[1] = tr1; sm.stage[s].hdr[2] = tb0;
                    // 16-byte aligned source ranges enclosing the tile's values / column indices
                    const uintptr_t v0 = (uintptr_t)(vals + tb0 * NN), v1 = (uintptr_t)(vals + tb1 * NN);
                    const uintptr_t c0 = (uintptr_t)(colidx + tb0), c1 = (uintptr_t)(colidx + tb1);
                    const uintptr_t v0a = v0 & ~(uintptr_t)15, v1a = (v1 + 15) & ~(uintptr_t)15;
                    const uintptr_t c0a = c0 & ~(uintptr_t)15, c1a = (c1 + 15) & ~(uintptr_t)15;
                    const uint32_t vb = (uint32_t)(v1a - v0a), cb = (uint32_t)(c1a - c0a);
                    // the tile's slice of rowptr travels too (unless it has too many -- empty -- rows)
                    const uintptr_t p0 = (uintptr_t)(rowptr + tr0), p1 = (uintptr_t)(rowptr + tr1 + 1);
                    const uintptr_t p0a = p0 & ~(uintptr_t)15, p1a = (p1 + 15) & ~(uintptr_t)15;
                    const uint32_t pb = (tr1 - tr0 <= kTmaWindow) ? (uint32_t)(p1a - p0a) : 0u;
                    mbar_arrive_expect_tx(&sm.full[s], vb + cb + pb);
                    if (vb) tma_bulk_g2s(sm.stage[s].vals, (const void *)v0a, vb, &sm.full[s], pol);
                    if (cb) tma_bulk_g2s(sm.stage[s].cols, (const void *)c0a, cb, &sm.full[s], pol);
                    if (pb) tma_bulk_g2s(sm.stage[s].rows, (const void *)p0a, pb, &sm.full[s], pol);
                }
            }
        }
    } else {
        // ===== consumer warps =====
        const int comp = owner_component<N, LPR>(lane);
        const uint64_t polKeep = l2_policy_evict_last();
        int jj[U], cc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int f = lane + u * LPR;
            jj[u] = f / N;
            cc[u] = f - jj[u] * N;
        }
        for (int64_t k = 0; k < myTiles; ++k) {
            const int s = (int)(k % kTmaStages);
            mbar_wait(&sm.full[s], (uint32_t)((k / kTmaStages) & 1));
            const int64_t r0 = sm.stage[s].hdr[0], r1 = sm.stage[s].hdr[1], tb0 = sm.stage[s].hdr[2];
            const double *sv = sm.stage[s].vals + (((uintptr_t)(vals + tb0 * NN) & 15) >> 3);
            const int32_t *sc = sm.stage[s].cols + (((uintptr_t)(colidx + tb0) & 15) >> 2);
            const bool rowsStaged = (r1 - r0) <= kTmaWindow;
            const long long *sr = sm.stage[s].rows + (((uintptr_t)(rowptr + r0) & 15) >> 3);
            while (true) {
                int idx = 0;
                if (lane == 0) idx = atomicAdd(&sm.next[s], 1);
                idx = __shfl_sync(0xffffffffu, idx, 0);
                const int64_t row = r0 + idx;
                if (row >= r1) break;
                const int64_t b0 = rowsStaged ? sr[idx] : rowptr[row];
                const int L = (int)((rowsStaged ? sr[idx + 1] : rowptr[row + 1]) - b0) * N;
                const double *v = sv + (b0 - tb0) * NN + lane;
                const int32_t *ci = sc + (b0 - tb0);
                double acc[N];
#pragma unroll
                for (int r = 0; r < N; ++r) acc[r] = 0.0;
                for (int base = 0; base < L; base += CH, v += CH, ci += CH / N) {
                    const int rem = L - base - lane;
                    int col[U];
                    double a[U][N], xv[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const bool ok = u * LPR < rem;
                        col[u] = ok ? ci[jj[u]] : 0;
#pragma unroll
                        for (int r = 0; r < N; ++r) a[u][r] = ok ? v[r * L + u * LPR] : 0.0;
                    }
                    gather3(x + (col[0] * N + cc[0]), x + (col[1] * N + cc[1]), x + (col[2] * N + cc[2]), polKeep, xv);
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int r = 0; r < N; ++r) acc[r] = fma(a[u][r], xv[u], acc[r]);
                }
                const double out0 = fold_reduce<N, LPR>(acc, lane);
                if (comp >= 0) {
                    double out = out0;
                    if (MASKED && fixedMask[row * N + comp]) out = 0.0;
                    y[row * N + comp] = out;
                    if (DOT) dot += out * x[row * N + comp];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[s]);
        }
    }
    if (DOT) {
        __shared__ double redT[kTmaThreads / 32];
        double xsum = dot;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) xsum += __shfl_xor_sync(0xffffffffu, xsum, o);
        if (lane == 0) redT[warp] = xsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            double sum = 0.0;
            for (int w = 0; w < kTmaThreads / 32; ++w) sum += redT[w];
            partials[blockIdx.x] = sum;
        }
        if (last_block(ticket)) {
            const double tot = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) dotOut[0] = tot;
        }
    }
}

// ---------------------------------------------------------------------------
// K5  block-Jacobi: invert the masked diagonal blocks.
// ---------------------------------------------------------------------------
// diagonal block of every row, plain row-major (multi-GPU: summed across sharers before inversion)
template <int N>
__global__ void k_diag_extract(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                               const double *__restrict__ vals, double *__restrict__ out) {
    constexpr int NN = N * N;
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nb) return;
    int64_t lo = rowptr[row], hi = rowptr[row + 1];
    const int64_t b0 = lo, end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (colidx[mid] < row) lo = mid + 1; else hi = mid;
    }
    const bool have = (lo < end) && (colidx[lo] == row);
    for (int r = 0; r < N; ++r)
        for (int c = 0; c < N; ++c) out[row * NN + r * N + c] = have ? vals[val_index<N>(b0, end - b0, lo - b0, r, c)] : 0.0;
}

template <int N>
__global__ void k_jacobi_setup(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                               const double *__restrict__ vals, const uint8_t *__restrict__ fixedMask,
                               double *Minv, int *bad, bool preExtracted) {
    constexpr int NN = N * N;
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nb) return;
    int64_t lo = rowptr[row], hi = rowptr[row + 1];
    const int64_t end = hi;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if (colidx[mid] < row) lo = mid + 1; else hi = mid;
    }
    double a[N][N];
    const bool have = (lo < end) && (colidx[lo] == row);
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c)
            a[r][c] = preExtracted ? Minv[row * NN + r * N + c]
                                   : (have ? vals[val_index<N>(rowptr[row], end - rowptr[row], lo - rowptr[row], r, c)]
                                           : ((r == c) ? 1.0 : 0.0));
    if (preExtracted) {      // DoF without any element on any rank: keep the identity
        bool allZero = true;
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int c = 0; c < N; ++c) allZero = allZero && (a[r][c] == 0.0);
        if (allZero)
#pragma unroll
            for (int r = 0; r < N; ++r) a[r][r] = 1.0;
    }
    bool fx[N];
#pragma unroll
    for (int r = 0; r < N; ++r) fx[r] = fixedMask[row * N + r] != 0;
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c)
            if (fx[r] || fx[c]) a[r][c] = (r == c) ? 1.0 : 0.0;
    double inv[N][N];
    double det;
    if (N == 3) {
        const double c00 = a[1][1] * a[2][2] - a[1][2] * a[2][1];
        const double c01 = a[1][2] * a[2][0] - a[1][0] * a[2][2];
        const double c02 = a[1][0] * a[2][1] - a[1][1] * a[2][0];
        det = a[0][0] * c00 + a[0][1] * c01 + a[0][2] * c02;
        const double id = 1.0 / det;
        inv[0][0] = c00 * id;
        inv[0][1] = (a[0][2] * a[2][1] - a[0][1] * a[2][2]) * id;
        inv[0][2] = (a[0][1] * a[1][2] - a[0][2] * a[1][1]) * id;
        inv[1][0] = c01 * id;
        inv[1][1] = (a[0][0] * a[2][2] - a[0][2] * a[2][0]) * id;
        inv[1][2] = (a[0][2] * a[1][0] - a[0][0] * a[1][2]) * id;
        inv[2][0] = c02 * id;
        inv[2][1] = (a[0][1] * a[2][0] - a[0][0] * a[2][1]) * id;
        inv[2][2] = (a[0][0] * a[1][1] - a[0][1] * a[1][0]) * id;
    } else {
        det = a[0][0] * a[1][1] - a[0][1] * a[1][0];
        const double id = 1.0 / det;
        inv[0][0] = a[1][1] * id; inv[0][1] = -a[0][1] * id;
        inv[1][0] = -a[1][0] * id; inv[1][1] = a[0][0] * id;
    }
    if (!(det > 0.0) || !isfinite(det)) atomicAdd(bad, 1);
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int c = 0; c < N; ++c) Minv[row * NN + r * N + c] = inv[r][c];
}

static int sm_count(mfem_b200_ctx *c);

#include "coarse.inl"

// ---------------------------------------------------------------------------
// K4  fused vector kernels (one thread per DoF block, grid-stride)
// ---------------------------------------------------------------------------
// update (INIT: start-up): alpha = rz/pAp; x += alpha p; r -= alpha Ap  (INIT: x = 0, r = mask(b));  z = Minv r;
// partial sums of r.z and r.r over owned DoFs -> red[0], red[1]; with a coarse space also c1 += P1^T r over owned
// DoFs (coarse.inl).  A warp holds 32 consecutive DoFs, which is what the segmented restriction needs.
template <int N, bool INIT>
__global__ void __launch_bounds__(kVecThreads)
k_pcg_update(int64_t nb, const double *__restrict__ b, const uint8_t *__restrict__ fixedMask, const double *__restrict__ Minv,
             const uint8_t *__restrict__ owned, const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ x,
             double *__restrict__ r, double *__restrict__ z, double *partials, unsigned *ticket, const double *__restrict__ scal,
             double *red /* [0] = r.z, [1] = r.r */, const int *status, const int32_t *__restrict__ agg1,
             const double *__restrict__ Y1, double *c1) {
    constexpr int NN = N * N;
    if (!INIT && status[ST_STATE] != 0) return;
    const double alpha = INIT ? 0.0 : scal[S_RZ] / scal[S_PAP];
    const int lane = threadIdx.x & 31;
    double acc[2] = {0.0, 0.0};
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~31ll; base < nb; base += stride) {
        const int64_t i = base + lane;
        const bool active = i < nb;
        double rv[N], zv[N];
        bool mine = false;
        if (active) {
#pragma unroll
            for (int k = 0; k < N; ++k) {
                if (INIT) {
                    x[i * N + k] = 0.0;
                    rv[k] = fixedMask[i * N + k] ? 0.0 : b[i * N + k];
                } else {
                    x[i * N + k] += alpha * p[i * N + k];
                    rv[k] = r[i * N + k] - alpha * Ap[i * N + k];
                }
            }
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m) s += Minv[i * NN + k * N + m] * rv[m];
                zv[k] = s;
            }
            mine = !(owned && !owned[i]);
            const double wgt = mine ? 1.0 : 0.0;
#pragma unroll
            for (int k = 0; k < N; ++k) {
                r[i * N + k] = rv[k]; z[i * N + k] = zv[k];
                acc[0] += wgt * rv[k] * zv[k];
                acc[1] += wgt * rv[k] * rv[k];
            }
        }
        if (c1) {                       // kernel-uniform
            if (!mine) {
#pragma unroll
                for (int k = 0; k < N; ++k) rv[k] = 0.0;
            }
            coarse_restrict_warp<N>(active, i, rv, agg1, Y1, c1, lane);
        }
    }
    block_reduce_store<2>(acc, partials);
    if (last_block(ticket)) {
        const double rz = final_sum(partials, gridDim.x);
        const double rr = final_sum(partials + gridDim.x, gridDim.x);
        if (threadIdx.x == 0) { red[0] = rz; red[1] = rr; }
    }
}

// after the start-up sums are in rzrr[0..1] (all-reduced, coarse part included)
__global__ void k_pcg_init_finalize(double *scal, const double *rzrr, int *status, double tol2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const double rr = rzrr[1];
    scal[S_RZ] = rzrr[0];
    scal[S_RR] = rr;
    scal[S_BB] = rr;
    scal[S_TOL2] = tol2;
    status[ST_ITERS] = 0;
    status[ST_STATE] = (rr == 0.0) ? 1 : ((rr != rr) ? 3 : 0);
}

// direction: beta = rz_new/rz; p = z + [mask(R1 (y1 + P2 y2))] + beta p  (INIT: p = z + coarse part).  The last CTA then
// closes the iteration on the (global) scalars: rotates rz, counts the iteration, decides convergence / breakdown.
// With a coarse space the kernel also clears c2 (red[2..]) for the next application.
template <int N, bool INIT>
__global__ void __launch_bounds__(kVecThreads)
k_pcg_direction(int64_t nb, const double *__restrict__ z, double *__restrict__ p, double *scal, const double *rzrr, int *status,
                unsigned *ticket, const uint8_t *__restrict__ fixedMask, const int32_t *__restrict__ agg1,
                const double *__restrict__ Y1, const double *__restrict__ y1, const double *__restrict__ y2,
                const double *__restrict__ shift, int64_t S1, int64_t R, int64_t aggBase, bool level1, double *red, int64_t nc2) {
    if (!INIT && status[ST_STATE] != 0) return;
    const double beta = INIT ? 0.0 : rzrr[0] / scal[S_RZ];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < nb; i += stride) {
        double v[N];
#pragma unroll
        for (int k = 0; k < N; ++k) v[k] = 0.0;
        if (agg1) coarse_prolong_dof<N>(i, agg1, Y1, y1, y2, shift, S1, R, aggBase, level1, v);
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const double zk = z[i * N + k] + ((agg1 && !fixedMask[i * N + k]) ? v[k] : 0.0);
            p[i * N + k] = INIT ? zk : zk + beta * p[i * N + k];
        }
    }
    if (INIT) {
        for (int64_t k = t0; k < nc2; k += stride) red[2 + k] = 0.0;
        return;
    }
    // rzrr may alias red[0..1] (no coarse space): read the scalars before anything is cleared, by the last CTA only
    if (last_block(ticket)) {
        if (threadIdx.x == 0) {
            const double pAp = scal[S_PAP], rz = rzrr[0], rr = rzrr[1];
            scal[S_RZ] = rz;
            scal[S_RR] = rr;
            status[ST_ITERS] += 1;
            int st = 0;
            if (!(pAp > 0.0)) st = 2;
            if (rr != rr || rz != rz || pAp != pAp) st = 3;
            if (st == 0 && rr <= scal[S_TOL2] * scal[S_BB]) st = 1;
            status[ST_STATE] = st;     // written last; kernels of later iterations read it first
        }
    }
    for (int64_t k = t0; k < nc2; k += stride) red[2 + k] = 0.0;      // nobody reads c2 in this kernel
}

// b = f - K ufix on free rows (ufix carries the fixed values, zero elsewhere); u = x + ufix
__global__ void k_axpby(int64_t n, double a, const double *__restrict__ x, double bcoef, const double *__restrict__ y,
                        double *__restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = a * x[i] + bcoef * y[i];
}

// ---------------------------------------------------------------------------
static int sm_count(mfem_b200_ctx *c) {
    static int n = 0;
    if (!n) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, c->device);
    return n > 0 ? n : 148;
}

#include "matfree.inl"

static int spmv_lanes(mfem_b200_ctx *c) {
    if (c->opt_spmv_lanes == 8 || c->opt_spmv_lanes == 16 || c->opt_spmv_lanes == 32) return c->opt_spmv_lanes;
    const double meanL = c->nDofs ? double(c->nnzb) * c->N / double(c->nDofs) : 0.0;   // scalars per scalar row
    return meanL > 40.0 ? 32 : (meanL > 18.0 ? 16 : 8);
}

// persistent-style launch: exactly the CTAs that are co-resident (occupancy API), row groups
// grid-strided by warp so that all SMs sweep one moving window of the matrix together
template <class Kern>
static int spmv_grid(mfem_b200_ctx *c, int lpr, Kern kern) {
    int perSM = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kern, kSpmvThreads, 0);
    if (perSM < 1) perSM = 1;
    const int64_t warpsNeeded = (c->nDofs + (32 / lpr) - 1) / (32 / lpr);
    const int64_t ctas = (warpsNeeded + (kSpmvThreads / 32) - 1) / (kSpmvThreads / 32);
    return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, (int64_t)sm_count(c) * perSM));
}

static int vec_grid(mfem_b200_ctx *c, int64_t n) {
    const int64_t ctas = (n + kVecThreads - 1) / kVecThreads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, std::min<int64_t>(kMaxPartials, (int64_t)sm_count(c) * 8)));
}

void ensure_work(mfem_b200_ctx *c) {
    if (c->workValid) return;
    const size_t n = (size_t)c->nvar();
    PcgWork &w = c->work;
    w.x.alloc(n); w.r.alloc(n); w.z.alloc(n); w.p.alloc(n); w.Ap.alloc(n); w.b.alloc(n); w.ufix.alloc(n);
    w.partials.alloc(4 * (size_t)kMaxPartials);
    w.scal.alloc(S_COUNT);
    w.dotLoc.alloc(DL_COUNT);
    w.red.alloc(2 + 32768);          // (r.z, r.r, coarse residuals c2): one all-reduce per iteration
    w.ticket.alloc(8);
    w.status.alloc(4);
    MFEM_CUDA(cudaMemsetAsync(w.ticket, 0, w.ticket.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.status, 0, w.status.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.scal, 0, w.scal.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.dotLoc, 0, w.dotLoc.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.red, 0, w.red.bytes(), c->stream));
    if (c->fixedMask.n != n) {
        c->fixedMask.alloc(n);
        c->fixedVals.alloc(n);
        MFEM_CUDA(cudaMemsetAsync(c->fixedMask, 0, c->fixedMask.bytes(), c->stream));
        MFEM_CUDA(cudaMemsetAsync(c->fixedVals, 0, c->fixedVals.bytes(), c->stream));
    }
    c->workValid = true;
}

template <int N, int LPR, bool PF>
static void launch_spmv_lp(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    PcgWork &w = c->work;
    if (masked && dot)
        k_bsr_spmv<N, LPR, true, true, PF><<<spmv_grid(c, LPR, k_bsr_spmv<N, LPR, true, true, PF>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, w.partials, w.ticket, w.scal.p + S_PAP, w.status);
    else if (masked)
        k_bsr_spmv<N, LPR, true, false, PF><<<spmv_grid(c, LPR, k_bsr_spmv<N, LPR, true, false, PF>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, nullptr, nullptr, nullptr, nullptr);
    else
        k_bsr_spmv<N, LPR, false, false, PF><<<spmv_grid(c, LPR, k_bsr_spmv<N, LPR, false, false, PF>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, nullptr, nullptr, nullptr, nullptr, nullptr);
    c->launches++;
}
// occupancy experiment (option spmv_min_blocks = 3 / 5, N = 3 with 32 lanes only): the same kernel compiled for 3 CTAs
// per SM (80 registers: ptxas keeps the lane constants instead of recomputing them per row) or 5 (48 registers)
template <int MINB>
static void launch_spmv_occ(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    PcgWork &w = c->work;
    const int threads = MINB == 1 ? 1024 : kSpmvThreads;
    int grid;
    if (MINB == 1) {
        const int64_t ctas = (c->nDofs + 31) / 32;
        grid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, sm_count(c)));
    } else {
        grid = spmv_grid(c, 32, k_bsr_spmv<3, 32, true, true, false, MINB>);
    }
    if (masked && dot)
        k_bsr_spmv<3, 32, true, true, false, MINB><<<grid, threads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, w.partials, w.ticket, w.scal.p + S_PAP, w.status);
    else
        k_bsr_spmv<3, 32, false, false, false, MINB><<<grid, threads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, nullptr, nullptr, nullptr, nullptr, nullptr);
    c->launches++;
}
template <int N, int LPR>
static void launch_spmv_l(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    if (N == 3 && LPR == 32 && (masked == dot) && (c->opt_spmv_min_blocks == 3 || c->opt_spmv_min_blocks == 5 || c->opt_spmv_min_blocks == 1)) {
        if (c->opt_spmv_min_blocks == 3) launch_spmv_occ<3>(c, x, y, masked, dot);
        else if (c->opt_spmv_min_blocks == 5) launch_spmv_occ<5>(c, x, y, masked, dot);
        else launch_spmv_occ<1>(c, x, y, masked, dot);
        return;
    }
    if (c->opt_spmv_prefetch) launch_spmv_lp<N, LPR, true>(c, x, y, masked, dot);
    else launch_spmv_lp<N, LPR, false>(c, x, y, masked, dot);
}

template <int N>
static void launch_spmv_pipe(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    PcgWork &w = c->work;
    if (masked && dot)
        k_bsr_spmv_pipe<N, true, true><<<spmv_grid(c, 32, k_bsr_spmv_pipe<N, true, true>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, w.partials, w.ticket, w.scal.p + S_PAP, w.status);
    else if (masked)
        k_bsr_spmv_pipe<N, true, false><<<spmv_grid(c, 32, k_bsr_spmv_pipe<N, true, false>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, nullptr, nullptr, nullptr, nullptr);
    else
        k_bsr_spmv_pipe<N, false, false><<<spmv_grid(c, 32, k_bsr_spmv_pipe<N, false, false>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, nullptr, nullptr, nullptr, nullptr, nullptr);
    c->launches++;
}

static bool spmv_use_sym(mfem_b200_ctx *c) { return c->opt_spmv_kernel == 4; }

// y = K x through the upper tails (y zeroed here); no masking, no dot
template <int N>
static void launch_spmv_sym(mfem_b200_ctx *c, const double *x, double *y, const int *status) {
    MFEM_CUDA(cudaMemsetAsync(y, 0, sizeof(double) * c->nvar(), c->stream));
    const double meanLu = c->nDofs ? double(c->nnzb + c->nDofs) * 0.5 * c->N / double(c->nDofs) : 0.0;
    const int lanes = (c->opt_spmv_lanes == 8 || c->opt_spmv_lanes == 16 || c->opt_spmv_lanes == 32)
                          ? c->opt_spmv_lanes : (meanLu > 40.0 ? 32 : (meanLu > 18.0 ? 16 : 8));
#define MFEM_SYM_LAUNCH(L_)                                                                                         \
    k_bsr_spmv_sym<N, L_><<<spmv_grid(c, L_, k_bsr_spmv_sym<N, L_>), kSpmvThreads, 0, c->stream>>>(                  \
        c->nDofs, c->rowptr, c->upperStart, c->colidx, c->vals, x, y, status)
    if (lanes == 8) MFEM_SYM_LAUNCH(8);
    else if (lanes == 16) MFEM_SYM_LAUNCH(16);
    else MFEM_SYM_LAUNCH(32);
#undef MFEM_SYM_LAUNCH
    c->launches++;
}

static bool spmv_use_tma(mfem_b200_ctx *c) {
    if (c->opt_spmv_kernel == 1 || c->opt_spmv_kernel == 3 || c->opt_spmv_kernel == 4) return false;
    const bool fits = c->maxRowLen <= kTmaMaxRow && c->tileRow.n > 1;
    if (c->opt_spmv_kernel == 2) {
        MFEM_REQUIRE(fits, MFEM_B200_ERR_INVALID, "spmv_kernel=2 (TMA ring) needs block rows of at most 128 blocks");
        return true;
    }
    // auto: the direct-load kernel is faster today (cfg3: 1.30 ms vs 1.63 ms); see DESIGN.md section 4
    return false;
}

template <int N>
static void launch_spmv_tma(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    PcgWork &w = c->work;
    const int64_t nTiles = (int64_t)c->tileRow.n - 1;
    const int grid = (int)std::max<int64_t>(1, std::min<int64_t>(nTiles, sm_count(c)));
    const size_t smem = sizeof(TmaSmem<N>);
#define MFEM_TMA_LAUNCH(M_, D_, MASK_, PART_, TICK_, DOUT_, ST_)                                                        \
    do {                                                                                                               \
        auto kern = k_bsr_spmv_tma<N, M_, D_>;                                                                         \
        static bool attr = false;                                                                                      \
        if (!attr) { MFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); attr = true; } \
        kern<<<grid, kTmaThreads, smem, c->stream>>>(c->nDofs, nTiles, c->tileRow, c->rowptr, c->colidx, c->vals, x, y, MASK_, \
                                                     PART_, TICK_, DOUT_, ST_);                                        \
    } while (0)
    if (masked && dot) MFEM_TMA_LAUNCH(true, true, c->fixedMask, w.partials, w.ticket, w.scal.p + S_PAP, w.status);
    else if (masked) MFEM_TMA_LAUNCH(true, false, c->fixedMask, nullptr, nullptr, nullptr, nullptr);
    else MFEM_TMA_LAUNCH(false, false, nullptr, nullptr, nullptr, nullptr, nullptr);
#undef MFEM_TMA_LAUNCH
    c->launches++;
}

static void launch_spmv_v2probe(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    PcgWork &w = c->work;
    if (masked && dot)
        k_bsr_spmv_v2<true, true, false><<<spmv_grid(c, 16, k_bsr_spmv_v2<true, true, false>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, c->fixedMask, w.partials, w.ticket, w.scal.p + S_PAP, w.status);
    else
        k_bsr_spmv_v2<false, false, false><<<spmv_grid(c, 16, k_bsr_spmv_v2<false, false, false>), kSpmvThreads, 0, c->stream>>>(
            c->nDofs, c->rowptr, c->colidx, c->vals, x, y, nullptr, nullptr, nullptr, nullptr, nullptr);
    c->launches++;
}

template <int N>
static void launch_spmv(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    if (N == 3 && c->opt_spmv_kernel == 5) { launch_spmv_v2probe(c, x, y, masked, dot); return; }
    if (spmv_use_sym(c)) {
        PcgWork &w = c->work;
        launch_spmv_sym<N>(c, x, y, (masked && dot) ? w.status.p : nullptr);
        if (masked || dot) {
            k_mask_dot<N><<<vec_grid(c, c->nDofs), kVecThreads, 0, c->stream>>>(
                c->nDofs, masked ? c->fixedMask.p : nullptr, nullptr, x, y, w.partials, w.ticket, dot ? w.scal.p + S_PAP : nullptr,
                (masked && dot) ? w.status.p : nullptr);
            c->launches++;
        }
        return;
    }
    if (spmv_use_tma(c)) { launch_spmv_tma<N>(c, x, y, masked, dot); return; }
    // index-pipelined kernel (option 3): one block row per warp, i.e. the 32-lane configuration.  Not the
    // default: measured equal on cfg3 (1.33 vs 1.32 ms) and slower on cfg5 (6.75 vs 6.33 ms) -- the direct
    // kernel is not bound by the per-row round trips.
    if (spmv_lanes(c) == 32 && c->opt_spmv_kernel == 3) { launch_spmv_pipe<N>(c, x, y, masked, dot); return; }
    switch (spmv_lanes(c)) {
        case 8: launch_spmv_l<N, 8>(c, x, y, masked, dot); break;
        case 16: launch_spmv_l<N, 16>(c, x, y, masked, dot); break;
        default: launch_spmv_l<N, 32>(c, x, y, masked, dot); break;
    }
}

// what the Krylov loops multiply with: the mesh-based operator (matfree.inl) where it is chosen, else the stored matrix
template <int N>
static void launch_operator(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot) {
    if (use_matrix_free(c)) launch_matrix_free<N>(c, x, y, masked, dot);
    else launch_spmv<N>(c, x, y, masked, dot);
}

void spmv_plain(mfem_b200_ctx *c, const double *x_int, double *y_int) {
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "spmv: matrix not assembled");
    if (c->opt_spmv_kernel == 6) {      // the mesh-based operator on request (parity tests compare it with the stored matrix)
        MFEM_REQUIRE(matrix_free_eligible(c), MFEM_B200_ERR_INVALID, "spmv_kernel 6: the matrix-free operator needs a mesh and a material");
        ensure_work(c);
        if (c->N == 3) launch_matrix_free<3>(c, x_int, y_int, false, false);
        else launch_matrix_free<2>(c, x_int, y_int, false, false);
    } else if (c->N == 3) launch_spmv<3>(c, x_int, y_int, false, false);
    else launch_spmv<2>(c, x_int, y_int, false, false);
    MFEM_CUDA(cudaGetLastError());
}

void free_coarse_space(mfem_b200_ctx *c) { free_coarse(c); }

// will the next solve use the aggregation levels (explicit option, or the automatic rule on a large enough problem)?
bool will_use_coarse(mfem_b200_ctx *c) {
    if (c->opt_coarse == 0) return false;
    if (c->opt_coarse > 0) return true;
    return coarse_budget(c, coarse_global_dofs(c)) > 0;
}

void build_preconditioner(mfem_b200_ctx *c) {
    if (c->precondValid) return;
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "preconditioner: matrix not assembled");
    ensure_work(c);
    ScopedTimer timer(c, "Fix Variables");
    const size_t n = (size_t)c->nDofs * c->N * c->N;
    if (c->Minv.n != n) c->Minv.alloc(n);
    DevBuf<int> bad(1);
    MFEM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), c->stream));
    const bool multi = c->nRanks > 1;
    if (multi) {
        // interface DoFs: the diagonal block is the sum of every sharer's partial block
        if (c->N == 3) k_diag_extract<3><<<grid_for(c->nDofs, 256), 256, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals, c->Minv);
        else k_diag_extract<2><<<grid_for(c->nDofs, 256), 256, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals, c->Minv);
        c->launches++;
        halo_exchange_add(c, c->Minv, c->N * c->N);
    }
    if (c->N == 3)
        k_jacobi_setup<3><<<grid_for(c->nDofs, 256), 256, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals,
                                                                          c->fixedMask, c->Minv, bad, multi);
    else
        k_jacobi_setup<2><<<grid_for(c->nDofs, 256), 256, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals,
                                                                          c->fixedMask, c->Minv, bad, multi);
    c->launches++;
    int nbad = 0;
    MFEM_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MFEM_CUDA(cudaStreamSynchronize(c->stream));
    MFEM_CUDA(cudaGetLastError());
    MFEM_REQUIRE(nbad == 0, MFEM_B200_ERR_NOT_SPD,
                 "block-Jacobi: " + std::to_string(nbad) + " diagonal blocks are not positive definite");
    timer.stop();
    if (c->opt_coarse != 0) build_coarse(c); else free_coarse(c);
    c->precondValid = true;
}

// y = mask(K x) completed across ranks: local SpMV, then the interface sum-exchange
template <int N>
static void spmv_exchanged(mfem_b200_ctx *c, const double *x, double *y, bool masked) {
    launch_operator<N>(c, x, y, masked, false);
    halo_exchange_add(c, y, N);
}

// the preconditioner's tail after k_pcg_update left r.z / r.r in red[0..1] and the level-1 residuals in c1:
// level 1, ONE all-reduce of (r.z, r.r, c2), dense level.  Returns where the final (r.z, r.r) pair lives.
template <int N>
static const double *enqueue_precond_tail(mfem_b200_ctx *c, const int *status) {
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    const bool multi = c->nRanks > 1;
    if (!c->coarse) {
        if (multi) allreduce_sum(c, w.red, w.red, 2);
        return w.red;
    }
    CoarseSpace &cs = *c->coarse;
    const int g1 = (int)std::max<int64_t>(1, std::min<int64_t>((cs.n1 + kVecThreads - 1) / kVecThreads, (int64_t)sm_count(c) * 8));
    k_coarse_level1<N><<<g1, kVecThreads, 0, s>>>(cs.S1, cs.n1, cs.R, cs.aggBase, cs.level1, cs.c1, cs.y1, cs.B1inv, cs.shift, w.red, status);
    if (multi) allreduce_sum(c, w.red, w.red, (int)(2 + cs.nc2));
    const int gg = (int)std::max<int64_t>(1, std::min<int64_t>((cs.nRowsLoc + 7) / 8, std::min<int64_t>(kMaxPartials, (int64_t)sm_count(c) * 8)));
    if (!multi) {
        k_coarse_gemv<true><<<gg, kVecThreads, 0, s>>>(cs.nc2, 0, cs.nRowsLoc, cs.Einv, w.red, cs.y2, w.partials + 2 * (size_t)kMaxPartials,
                                                       w.ticket + 3, w.scal.p + S_RZ_NEW, status);
    } else {
        // row-split dense level: this rank's slice of y2, one in-place all-gather, then c2.y2 in a fixed order
        const PeerWin *pw = comm_peer_window(c);
        if (pw && cs.nc2 <= kAgCap) {
            // the GEMV stores its rows of y2 straight into every rank's window; the 1-CTA kernel that adds c2.y2 waits for
            // everybody's rows and copies the gathered vector out: the all-gather costs no launch of its own
            k_coarse_gemv_ship<<<gg, kVecThreads, 0, s>>>(*pw, cs.nc2, cs.rowBase, cs.nRowsLoc, cs.Einv, w.red, w.ticket + 4, status);
            k_coarse_cy_gather<<<1, 1024, 0, s>>>(*pw, cs.nc2, w.red, cs.y2, w.scal.p + S_RZ_NEW, status);
        } else {
            k_coarse_gemv<false><<<gg, kVecThreads, 0, s>>>(cs.nc2, cs.rowBase, cs.nRowsLoc, cs.Einv, w.red, cs.y2, nullptr, nullptr, nullptr, status);
            allgather_inplace(c, cs.y2, (int)cs.nRowsLoc);
            k_coarse_cy<<<1, 256, 0, s>>>(cs.nc2, w.red, cs.y2, w.scal.p + S_RZ_NEW, status);
        }
        c->launches++;
    }
    c->launches += 2;
    return w.scal.p + S_RZ_NEW;
}

template <int N, bool INIT>
static void launch_direction(mfem_b200_ctx *c, const double *rzrr) {
    PcgWork &w = c->work;
    CoarseSpace *cs = c->coarse;
    k_pcg_direction<N, INIT><<<vec_grid(c, c->nDofs), kVecThreads, 0, c->stream>>>(
        c->nDofs, w.z, w.p, w.scal, rzrr, w.status, w.ticket + 2, c->fixedMask, cs ? cs->agg1.p : nullptr, cs ? cs->Y1.p : nullptr,
        cs ? cs->y1.p : nullptr, cs ? cs->y2.p : nullptr, cs ? cs->shift.p : nullptr, cs ? cs->S1 : 0, cs ? cs->R : 1,
        cs ? cs->aggBase : 0, cs ? cs->level1 : false, w.red, cs ? cs->nc2 : 0);
    c->launches++;
}

// One PCG iteration.  1 rank: SpMV (+ p.Ap) -> update (+ restriction) -> [level 1 -> dense level] -> direction.
// N ranks: p.Ap = sum over ranks of p_loc . (K_loc p_loc) -- the LOCAL products before the exchange, over all local
// rows, because K = sum of the ranks' element matrices -- so the same fused SpMV epilogue serves, and its 1-double
// all-reduce travels next to the interface exchange; the second (and last) all-reduce of the iteration carries
// (r.z, r.r, coarse residuals) together.
template <int N>
static void enqueue_iteration(mfem_b200_ctx *c) {
    PcgWork &w = c->work;
    const int64_t nb = c->nDofs;
    const int vgrid = vec_grid(c, nb);
    const bool multi = c->nRanks > 1;
    const uint8_t *owned = multi ? halo_owned(c) : nullptr;
    CoarseSpace *cs = c->coarse;
    launch_operator<N>(c, w.p, w.Ap, true, true);             // p.Ap fused into the epilogue -> scal[S_PAP]
    if (multi) {
        // masked rows stay zero through the exchange: every sharer masks the same DoFs
        if (!halo_exchange_add_allreduce1(c, w.Ap, N, w.scal.p + S_PAP)) {      // one kernel over the peer window, or NCCL
            halo_exchange_add(c, w.Ap, N);
            allreduce_sum(c, w.scal.p + S_PAP, w.scal.p + S_PAP, 1);
        }
    }
    k_pcg_update<N, false><<<vgrid, kVecThreads, 0, c->stream>>>(nb, nullptr, c->fixedMask, c->Minv, owned, w.p, w.Ap, w.x, w.r, w.z,
                                                                  w.partials, w.ticket + 1, w.scal, w.red, w.status,
                                                                  cs ? cs->agg1.p : nullptr, cs ? cs->Y1.p : nullptr,
                                                                  cs ? cs->c1.p : nullptr);
    c->launches++;
    const double *rzrr = enqueue_precond_tail<N>(c, w.status);
    launch_direction<N, false>(c, rzrr);
}

template <int N>
static void pcg_impl(mfem_b200_ctx *c, const double *f_int, double *u_int, double rtol, int maxIters,
                     mfem_b200_solve_info *info) {
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs, n = c->nvar();
    const bool multi = c->nRanks > 1;
    const uint8_t *owned = multi ? halo_owned(c) : nullptr;
    CoarseSpace *cs = c->coarse;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    // b = f - K ufix  (masked rows are zeroed by the start-up kernel)
    if (multi) spmv_exchanged<N>(c, c->fixedVals, w.Ap, false);
    else launch_operator<N>(c, c->fixedVals, w.Ap, false, false);
    k_axpby<<<vec_grid(c, n), kVecThreads, 0, s>>>(n, 1.0, f_int, -1.0, w.Ap, w.b);
    if (cs) {
        MFEM_CUDA(cudaMemsetAsync(cs->c1, 0, cs->c1.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(w.red, 0, w.red.bytes(), s));
    }
    k_pcg_update<N, true><<<vec_grid(c, nb), kVecThreads, 0, s>>>(nb, w.b, c->fixedMask, c->Minv, owned, nullptr, nullptr, w.x, w.r, w.z,
                                                                   w.partials, w.ticket + 1, w.scal, w.red, nullptr,
                                                                   cs ? cs->agg1.p : nullptr, cs ? cs->Y1.p : nullptr,
                                                                   cs ? cs->c1.p : nullptr);
    const double *rzrr = enqueue_precond_tail<N>(c, nullptr);                                // r.z0 with the coarse part
    k_pcg_init_finalize<<<1, 32, 0, s>>>(w.scal, rzrr, w.status, rtol * rtol);
    launch_direction<N, true>(c, rzrr);                                                      // p0 = z0
    c->launches += 3;
    MFEM_CUDA(cudaGetLastError());

    // The iteration is captured once into a CUDA graph of kBatch iterations; kernels turn into no-ops as soon as the
    // device-side state leaves "running", so the host only polls the 2-int status between graph launches.  The NCCL
    // calls of a multi-rank iteration (grouped send/recv of the interface, two all-reduces) are captured with it.
    const int kBatch = multi ? 10 : 25;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    int64_t launchesPerBatch = 0;
    if (c->opt_graph) {
        const int64_t launchesBefore = c->launches;
        MFEM_CUDA(cudaStreamBeginCapture(s, multi ? cudaStreamCaptureModeRelaxed : cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < kBatch; ++k) enqueue_iteration<N>(c);
        MFEM_CUDA(cudaStreamEndCapture(s, &graph));
        MFEM_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        launchesPerBatch = c->launches - launchesBefore;
        c->launches = launchesBefore;      // capture does not launch
    }
    int hst[2] = {0, 0};
    int done = 0;
    while (done < maxIters) {
        if (exec) {
            MFEM_CUDA(cudaGraphLaunch(exec, s));
            c->launches += launchesPerBatch;
        } else {
            for (int k = 0; k < kBatch; ++k) enqueue_iteration<N>(c);
        }
        MFEM_CUDA(cudaMemcpyAsync(hst, w.status, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
        done = hst[ST_ITERS];
        if (hst[ST_STATE] != 0) break;
    }
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    MFEM_CUDA(cudaGetLastError());
    c->timers["Elasticity Solve"] += ms * 1e-3;

    double hs[S_COUNT];
    MFEM_CUDA(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, s));
    // u = x + ufix
    k_axpby<<<vec_grid(c, n), kVecThreads, 0, s>>>(n, 1.0, w.x, 1.0, c->fixedVals, u_int);
    c->launches++;
    MFEM_CUDA(cudaStreamSynchronize(s));
    if (info) {
        info->iterations = hst[ST_ITERS];
        info->converged = hst[ST_STATE] == 1;
        info->rel_residual = hs[S_BB] > 0 ? std::sqrt(hs[S_RR] / hs[S_BB]) : 0.0;
        info->seconds = ms * 1e-3;
        info->spmv_seconds = 0.0;
    }
    if (multi && comm_peer_error(c))
        throw CudaError(MFEM_B200_ERR_COMM, "PCG: a peer-window collective timed out (a rank stopped responding)");
    if (hst[ST_STATE] == 2)
        throw CudaError(MFEM_B200_ERR_NOT_SPD, "PCG breakdown: p'Ap <= 0 (matrix is not positive definite)");
    if (hst[ST_STATE] == 3) throw CudaError(MFEM_B200_ERR_NAN, "PCG: NaN encountered");
    if (hst[ST_STATE] == 0)
        throw CudaError(MFEM_B200_ERR_NO_CONVERGE, "PCG: no convergence in " + std::to_string(hst[ST_ITERS]) +
                                                       " iterations (rel. residual " +
                                                       std::to_string(info ? info->rel_residual : -1.0) + ")");
}

void pcg_solve(mfem_b200_ctx *c, const double *f_int, double *u_int, double rtol, int maxIters,
               mfem_b200_solve_info *info) {
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "solve: matrix not assembled");
    ensure_work(c);
    build_preconditioner(c);
    if (c->N == 3) pcg_impl<3>(c, f_int, u_int, rtol, maxIters, info);
    else pcg_impl<2>(c, f_int, u_int, rtol, maxIters, info);
}

// z = M^-1 r and r.z for the preconditioner the next solve would use (block-Jacobi [+ level 1 + dense level]); r is
// masked on the fixed variables first.  Runs the PCG's own start-up kernels, so what is returned is exactly the operator
// inside the iteration.  Testing / diagnostics entry (mfem_b200_apply_preconditioner).
void apply_preconditioner(mfem_b200_ctx *c, const double *r_int, double *z_int, double *rz) {
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "apply_preconditioner: matrix not assembled");
    ensure_work(c);
    build_preconditioner(c);
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs;
    const uint8_t *owned = c->nRanks > 1 ? halo_owned(c) : nullptr;
    CoarseSpace *cs = c->coarse;
    if (cs) MFEM_CUDA(cudaMemsetAsync(cs->c1, 0, cs->c1.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(w.red, 0, w.red.bytes(), s));
    const double *rzrr = nullptr;
    if (c->N == 3) {
        k_pcg_update<3, true><<<vec_grid(c, nb), kVecThreads, 0, s>>>(nb, r_int, c->fixedMask, c->Minv, owned, nullptr, nullptr, w.x, w.r, w.z,
                                                                       w.partials, w.ticket + 1, w.scal, w.red, nullptr,
                                                                       cs ? cs->agg1.p : nullptr, cs ? cs->Y1.p : nullptr, cs ? cs->c1.p : nullptr);
        rzrr = enqueue_precond_tail<3>(c, nullptr);
        launch_direction<3, true>(c, rzrr);
    } else {
        k_pcg_update<2, true><<<vec_grid(c, nb), kVecThreads, 0, s>>>(nb, r_int, c->fixedMask, c->Minv, owned, nullptr, nullptr, w.x, w.r, w.z,
                                                                       w.partials, w.ticket + 1, w.scal, w.red, nullptr,
                                                                       cs ? cs->agg1.p : nullptr, cs ? cs->Y1.p : nullptr, cs ? cs->c1.p : nullptr);
        rzrr = enqueue_precond_tail<2>(c, nullptr);
        launch_direction<2, true>(c, rzrr);
    }
    c->launches++;
    MFEM_CUDA(cudaMemcpyAsync(z_int, w.p, sizeof(double) * c->nvar(), cudaMemcpyDeviceToDevice, s));
    MFEM_CUDA(cudaMemcpyAsync(rz, rzrr, sizeof(double), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_CUDA(cudaGetLastError());
}

// diagnostics: copy a named array of the coarse space to the host as doubles (internal DoF order); returns its length
int64_t get_coarse_array(mfem_b200_ctx *c, const std::string &name, double *out, int64_t capacity) {
    MFEM_REQUIRE(c->coarse, MFEM_B200_ERR_INVALID, "no coarse space (solve or apply_preconditioner first)");
    CoarseSpace &cs = *c->coarse;
    std::vector<double> h;
    auto fromD = [&](const DevBuf<double> &b) { h.resize(b.n); if (b.n) MFEM_CUDA(cudaMemcpy(h.data(), b.p, b.bytes(), cudaMemcpyDeviceToHost)); };
    if (name == "agg1") {
        std::vector<int32_t> t(cs.agg1.n);
        if (cs.agg1.n) MFEM_CUDA(cudaMemcpy(t.data(), cs.agg1.p, cs.agg1.bytes(), cudaMemcpyDeviceToHost));
        h.assign(t.begin(), t.end());
    } else if (name == "int2ext") {
        std::vector<int32_t> t(c->int2ext.n);
        if (c->int2ext.n) MFEM_CUDA(cudaMemcpy(t.data(), c->int2ext.p, c->int2ext.bytes(), cudaMemcpyDeviceToHost));
        h.assign(t.begin(), t.end());
    } else if (name == "Y1") fromD(cs.Y1);
    else if (name == "shift") fromD(cs.shift);
    else if (name == "B1inv") fromD(cs.B1inv);
    else if (name == "D1") fromD(cs.D1);
    else if (name == "Einv") fromD(cs.Einv);      // this rank's rows (all of them on one rank)
    else if (name == "y1") fromD(cs.y1);
    else if (name == "y2") fromD(cs.y2);
    else if (name == "sizes") h = {(double)cs.S1, (double)cs.S2, (double)cs.R, (double)cs.n1, (double)cs.aggBase, cs.level1 ? 1.0 : 0.0};
    else throw CudaError(MFEM_B200_ERR_INVALID, "unknown coarse array " + name);
    if (out) std::copy(h.begin(), h.begin() + std::min<int64_t>((int64_t)h.size(), capacity), out);
    return (int64_t)h.size();
}

#include "solver_multi.inl"

// Average duration of the SpMV launch the PCG iteration issues -- masked rows AND the fused p.Ap epilogue
// (k_bsr_spmv<N, lanes, 1, 1>) -- timed with CUDA events on the library's stream after 3 warm-up launches.
double time_spmv(mfem_b200_ctx *c, int iters) {
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "time_spmv: matrix not assembled");
    ensure_work(c);
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    // a non-trivial, reproducible input vector; status "running" (a finished solve leaves it at "converged", which
    // turns the in-loop kernels into no-ops)
    MFEM_CUDA(cudaMemsetAsync(w.p, 0x3f, w.p.bytes(), s));   // 0x3f3f.. ~ 4.8e-4, finite
    MFEM_CUDA(cudaMemsetAsync(w.status, 0, w.status.bytes(), s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&]() {
        if (c->N == 3) launch_spmv<3>(c, w.p, w.Ap, true, true);
        else launch_spmv<2>(c, w.p, w.Ap, true, true);
    };
    for (int k = 0; k < 3; ++k) run();
    cudaEventRecord(e0, s);
    for (int k = 0; k < iters; ++k) run();
    cudaEventRecord(e1, s);
    cudaEventSynchronize(e1);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MFEM_CUDA(cudaGetLastError());
    const double sec = ms * 1e-3 / iters;
    c->timers["SpMV"] += ms * 1e-3;
    return sec;
}

// the product the PCG launches per iteration (mesh-based operator or the stored-matrix SpMV), masked + fused dot;
// secondsParts[0..1]: the operator's two kernels timed alone (0 for the SpMV)
double time_operator(mfem_b200_ctx *c, int iters, int *matrixFree, double *secondsParts) {
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "time_operator: matrix not assembled");
    if (secondsParts) secondsParts[0] = secondsParts[1] = 0.0;
    if (!use_matrix_free(c)) {
        if (matrixFree) *matrixFree = 0;
        return time_spmv(c, iters);
    }
    if (matrixFree) *matrixFree = 1;
    ensure_work(c);
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    MFEM_CUDA(cudaMemsetAsync(w.p, 0x3f, w.p.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(w.status, 0, w.status.bytes(), s));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    double sec[3] = {0, 0, 0};
    const int phases[3] = {3, 1, 2};
    for (int t = 0; t < 3; ++t) {
        auto run = [&]() {
            if (c->N == 3) launch_matrix_free<3>(c, w.p, w.Ap, true, true, phases[t]);
            else launch_matrix_free<2>(c, w.p, w.Ap, true, true, phases[t]);
        };
        for (int k = 0; k < 3; ++k) run();
        cudaEventRecord(e0, s);
        for (int k = 0; k < iters; ++k) run();
        cudaEventRecord(e1, s);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        sec[t] = ms * 1e-3 / iters;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    MFEM_CUDA(cudaGetLastError());
    if (secondsParts) { secondsParts[0] = sec[1]; secondsParts[1] = sec[2]; }
    c->timers["Matrix-free Operator"] += sec[0] * iters;
    // diagnostics for the caller's byte accounting (reset_timers() drops the copy written when the tables were built)
    c->timers["Matrix-free Partials"] = (c->opt_mf_chunked && c->mfChunksValid) ? (double)c->mfPartials : 0.0;
    return sec[0];
}

}  // namespace mfem

// K2: numeric assembly of the block-CSR stiffness matrix.
//
// Replaces Simulator::m_assembleStiffnessMatrix (LinearElasticity.hh:1408-1466: per-element
// Ke then a SERIAL triplet scatter) and the numeric half of TripletMatrix::sumRepeated
// (SparseMatrices.hh:280-374).
//
// Mode 0 "owner-gather" (default).  The write pattern is turned inside out: instead of
// elements scattering 100 blocks each into shared rows (which needs atomics or colouring and
// moves every block ~2.5 times through HBM as read-modify-write), the DoF rows own the work.
// A warp takes a run of consecutive block rows whose element incidences fill ~one warp
// (32 (row, element, local node) incidences), every lane computes the 3 x 3*npe ROW SLICE of
// its element's stiffness with the factorised W (x) S form (elem_math.cuh), and the slices are
// summed into the rows' blocks in shared memory in a fixed order.  Each block of K is then
// written to HBM exactly once, coalesced, with no atomics and bit-reproducible sums.
// HBM traffic = the algorithmic minimum (nnzb * 72 B) + element records (L2-resident reuse).
//
// Mode 1 "coloured scatter" is the classical alternative named in the design brief:
// elements of one colour share no DoF, one launch per colour, plain read-modify-write.
#include "core.cuh"

namespace mfem {

constexpr int kAsmWarps = 4;        // warps per CTA
constexpr int kAsmSlots = 224;      // block accumulators per warp in shared memory

template <int N>
struct AsmSmem {
    double acc[kAsmWarps][kAsmSlots * N * N];
    int32_t cols[kAsmWarps][kAsmSlots];
};

__device__ __forceinline__ int64_t lower_bound_i64(const int64_t *a, int64_t lo, int64_t hi, int64_t v) {
    while (lo < hi) {   // first index in [lo,hi) with a[idx] >= v
        const int64_t mid = (lo + hi) >> 1;
        if (a[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// One run of consecutive rows [r0, r1) whose blocks [s0, s1) are accumulated either in the warp's
// shared-memory buffer (BIG = false) or, for a single row larger than the buffer, directly in HBM.
template <int N, int DEG, bool PER_ELEM_D, bool BIG>
__device__ __forceinline__ void asm_process_run(int lane, int64_t r0, int64_t r1, int64_t s0, int ns, double *acc,
                                                int32_t *cols, const int64_t *__restrict__ incPtr,
                                                const int32_t *__restrict__ incList, const int64_t *__restrict__ rowptr,
                                                const int32_t *__restrict__ colidx, const int32_t *__restrict__ elemDof,
                                                const double *__restrict__ geom, const MatD &Dc,
                                                const double *__restrict__ Delem, double *__restrict__ vals) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int NN = N * N;
    constexpr int GS = 1 + N * (N + 1);
    constexpr int F = flat_len(N);
    double *A = BIG ? vals + s0 * NN : acc;
    for (int64_t k = lane; k < (int64_t)ns * NN; k += 32) A[k] = 0.0;
    if (!BIG)
        for (int k = lane; k < ns; k += 32) cols[k] = colidx[s0 + k];
    __syncwarp();

    const int64_t i0 = incPtr[r0], i1 = incPtr[r1];
    for (int64_t tb = i0; tb < i1; tb += 32) {
        const int64_t t = tb + lane;
        const bool valid = t < i1;
        int64_t e = 0;
        int li = 0, rowBase = 0, rowLen = 0;
        if (valid) {
            const int32_t id = incList[t];
            e = id / NPE;
            li = id - (int)e * NPE;
            int64_t r = r0;                     // row of this incidence: last r with incPtr[r] <= t
            while (r + 1 < r1 && incPtr[r + 1] <= t) ++r;
            rowBase = (int)(rowptr[r] - s0);
            rowLen = (int)(rowptr[r + 1] - rowptr[r]);
        }
        ElemGeom<N> g;
        {
            const double *gp = geom + e * GS;
            g.vol = gp[0];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int a = 0; a <= N; ++a) g.G[r][a] = gp[1 + r * (N + 1) + a];
        }
        int32_t dofs[NPE];
#pragma unroll
        for (int j = 0; j < NPE; ++j) dofs[j] = elemDof[e * NPE + j];
        const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
        double *rowAcc = A + (int64_t)rowBase * NN;
        const int planeStride = N * rowLen;

        // lanes of one row start at different barycentric indices -> they meet a shared column
        // block at different steps (fewer serialised same-block additions)
        ke_row_slice_rot<N, DEG>(g, D, li, lane % (N + 1), [&](int j, const double blk[N][N]) {
            // slot of column DoF dofs[j] inside this incidence's row (sorted colidx)
            int jslot = 0;
            if (valid) {
                int lo = 0, hi = rowLen;
                const int32_t want = dofs[j];
                if (BIG) {
                    const int32_t *rc = colidx + s0 + rowBase;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rc[mid] < want) lo = mid + 1; else hi = mid; }
                } else {
                    const int32_t *rc = cols + rowBase;
                    while (lo < hi) { const int mid = (lo + hi) >> 1; if (rc[mid] < want) lo = mid + 1; else hi = mid; }
                }
                jslot = lo;
            }
            // fixed-order accumulation: lanes hitting the same block add one after the other in
            // lane (= element) order -> no atomics, reproducible sums
            const unsigned key = valid ? (unsigned)(rowBase + jslot) : (0x40000000u | (unsigned)lane);
            const unsigned peers = __match_any_sync(0xffffffffu, key);
            const int rank = __popc(peers & ((1u << lane) - 1u));
            const int maxRank = __reduce_max_sync(0xffffffffu, valid ? rank : 0);
            double *dst = rowAcc + N * jslot;           // row-plane layout (core.cuh val_index)
            for (int rr = 0; rr <= maxRank; ++rr) {
                if (valid && rank == rr) {
#pragma unroll
                    for (int cc = 0; cc < N; ++cc)
#pragma unroll
                        for (int dd = 0; dd < N; ++dd) dst[cc * planeStride + dd] += blk[cc][dd];
                }
                __syncwarp();
            }
        });
    }
    __syncwarp();
    if (!BIG) {
        double *out = vals + s0 * NN;
        for (int k = lane; k < ns * NN; k += 32) out[k] = acc[k];
    }
    __syncwarp();
}

template <int N, int DEG, bool PER_ELEM_D>
__global__ void __launch_bounds__(kAsmWarps * 32)
k_assemble_gather(int64_t nJobs, int64_t nb, const int64_t *__restrict__ jobRow, const int64_t *__restrict__ incPtr,
                  const int32_t *__restrict__ incList, const int64_t *__restrict__ rowptr,
                  const int32_t *__restrict__ colidx, const int32_t *__restrict__ elemDof,
                  const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem,
                  double *__restrict__ vals) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    AsmSmem<N> &sm = *reinterpret_cast<AsmSmem<N> *>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t job = (int64_t)blockIdx.x * kAsmWarps + warp;
    if (job >= nJobs) return;     // warp-uniform; no block-level barrier is used below

    // rows whose first incidence falls into this job's chunk of the incidence list
    // (jobRow[k] = first row r with incPtr[r] >= k * kChunk, precomputed with the pattern)
    const int64_t rBegin = jobRow[job], rEnd = jobRow[job + 1];
    if (rBegin >= rEnd || rBegin >= nb) return;
    const int64_t rStop = rEnd < nb ? rEnd : nb;
    double *acc = sm.acc[warp];
    int32_t *cols = sm.cols[warp];

    int64_t r0 = rBegin;
    while (r0 < rStop) {
        // largest run of rows [r0, r1) whose blocks fit the shared accumulators
        const int64_t s0 = rowptr[r0];
        int64_t r1 = r0 + 1;
        while (r1 < rStop && rowptr[r1 + 1] - s0 <= kAsmSlots) ++r1;
        const int ns = (int)(rowptr[r1] - s0);
        if (ns > kAsmSlots)      // a single row larger than the buffer: accumulate in HBM
            asm_process_run<N, DEG, PER_ELEM_D, true>(lane, r0, r1, s0, ns, acc, cols, incPtr, incList, rowptr, colidx,
                                                      elemDof, geom, Dc, Delem, vals);
        else
            asm_process_run<N, DEG, PER_ELEM_D, false>(lane, r0, r1, s0, ns, acc, cols, incPtr, incList, rowptr, colidx,
                                                       elemDof, geom, Dc, Delem, vals);
        r0 = r1;
    }
}

// ---------------------------------------------------------------------------
// Mode 1: coloured element scatter.  One thread per element of the colour; the element's
// full Ke is produced row slice by row slice and added with plain loads/stores.
template <int N, int DEG, bool PER_ELEM_D>
__global__ void __launch_bounds__(128)
k_assemble_colored(int64_t nInColor, const int32_t *__restrict__ elems, const int64_t *__restrict__ rowptr,
                   const int32_t *__restrict__ colidx, const int32_t *__restrict__ elemDof,
                   const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem,
                   double *__restrict__ vals) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int NN = N * N;
    constexpr int GS = 1 + N * (N + 1);
    constexpr int F = flat_len(N);
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nInColor) return;
    const int64_t e = elems[t];
    ElemGeom<N> g;
    const double *gp = geom + e * GS;
    g.vol = gp[0];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int a = 0; a <= N; ++a) g.G[r][a] = gp[1 + r * (N + 1) + a];
    int32_t dofs[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) dofs[j] = elemDof[e * NPE + j];
    const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
#pragma unroll 1
    for (int i = 0; i < NPE; ++i) {
        const int64_t rb = rowptr[dofs[i]], re = rowptr[dofs[i] + 1];
        ke_row_slice<N, DEG>(g, D, i, [&](int j, const double blk[N][N]) {
            int64_t lo = rb, hi = re;
            const int32_t want = dofs[j];
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (colidx[mid] < want) lo = mid + 1; else hi = mid;
            }
            double *dst = vals + rb * NN + N * (lo - rb);
            const int64_t plane = N * (re - rb);
#pragma unroll
            for (int cc = 0; cc < N; ++cc)
#pragma unroll
                for (int dd = 0; dd < N; ++dd) dst[cc * plane + dd] += blk[cc][dd];
        });
    }
}

template <int N, int DEG>
static void launch_assemble(mfem_b200_ctx *c) {
    cudaStream_t s = c->stream;
    if (c->opt_assembly == 0) {
        const int64_t nJobs = (c->totalInc + kAsmChunk - 1) / kAsmChunk;
        const int grid = (int)((nJobs + kAsmWarps - 1) / kAsmWarps);
        const size_t smem = sizeof(AsmSmem<N>);
        if (c->perElemD) {
            auto kern = k_assemble_gather<N, DEG, true>;
            MFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kAsmWarps * 32, smem, s>>>(nJobs, c->nDofs, c->jobRow, c->incPtr, c->incList, c->rowptr, c->colidx,
                                                    c->elemDof, c->geom, c->Dconst, c->Delem, c->vals);
        } else {
            auto kern = k_assemble_gather<N, DEG, false>;
            MFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kAsmWarps * 32, smem, s>>>(nJobs, c->nDofs, c->jobRow, c->incPtr, c->incList, c->rowptr, c->colidx,
                                                    c->elemDof, c->geom, c->Dconst, nullptr, c->vals);
        }
        c->launches++;
    } else {
        build_coloring(c);
        MFEM_CUDA(cudaMemsetAsync(c->vals, 0, c->vals.bytes(), s));
        for (int col = 0; col < c->nColors; ++col) {
            const int64_t b = c->colorPtr[col], n = c->colorPtr[col + 1] - b;
            if (n == 0) continue;
            if (c->perElemD)
                k_assemble_colored<N, DEG, true><<<grid_for(n, 128), 128, 0, s>>>(
                    n, c->colorElems.p + b, c->rowptr, c->colidx, c->elemDof, c->geom, c->Dconst, c->Delem, c->vals);
            else
                k_assemble_colored<N, DEG, false><<<grid_for(n, 128), 128, 0, s>>>(
                    n, c->colorElems.p + b, c->rowptr, c->colidx, c->elemDof, c->geom, c->Dconst, nullptr, c->vals);
            c->launches++;
        }
    }
    MFEM_CUDA(cudaGetLastError());
}

void assemble_values(mfem_b200_ctx *c) {
    MFEM_REQUIRE(c->geomValid, MFEM_B200_ERR_INVALID, "assemble: no mesh set");
    MFEM_REQUIRE(c->haveMaterial, MFEM_B200_ERR_INVALID, "assemble: no material set");
    build_pattern(c);
    {
        ScopedTimer timer(c, "Assemble System");
        if (c->N == 3 && c->deg == 1) launch_assemble<3, 1>(c);
        else if (c->N == 3 && c->deg == 2) launch_assemble<3, 2>(c);
        else if (c->N == 2 && c->deg == 1) launch_assemble<2, 1>(c);
        else launch_assemble<2, 2>(c);
    }
    c->valuesValid = true;
    c->precondValid = false;
}

}  // namespace mfem

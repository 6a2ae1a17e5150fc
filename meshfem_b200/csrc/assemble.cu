// K2: numeric assembly of the block-CSR stiffness matrix.
//
// Replaces Simulator::m_assembleStiffnessMatrix (LinearElasticity.hh:1408-1466: per-element
// Ke then a SERIAL triplet scatter) and the numeric half of TripletMatrix::sumRepeated
// (SparseMatrices.hh:280-374).
//
// Mode 0 "block-owner" (default).  Every BSR block is owned by one thread.  The symbolic phase
// (setup.cu) sorted all (element, i, j) pairs by their block, so a block's contributions are a
// contiguous list; the thread sums them in list (= element) order and the block is written to
// HBM exactly once -- no atomics, no colouring, bit-reproducible.  The sum is done in "geometry
// space": with a constant material
//     K[r,c] = sum_e Ke[i,j] = C : ( sum_e vol_e sum_{a,b} W[i,a][j,b] G_a (x) G_b ) = C : M[r,c]
// so a contribution costs one 3x3 accumulation of outer products (~36 FMA) and the 81-FMA
// contraction with the elasticity tensor happens once per block (per contribution only with
// per-element materials).  A CTA owns a chunk of 256 consecutive blocks.  The 128-byte geometry records
// of the chunk's distinct elements (a few dozen) are staged in shared memory by TMA bulk copies
// (cp.async.bulk completing on an mbarrier) while the threads do their bookkeeping; the blocks'
// contribution lists are cut into segments of <= 4 contributions handed to the threads in order of
// decreasing length (balanced warps, coalesced interleaved id lists of chunk-local element ids), the
// per-segment partial sums are combined per block in shared memory in list order, and every thread
// writes its block to HBM once (the lanes of a warp hold consecutive blocks of a row, so the stores of a
// plane fill its sectors together).
//
// Mode 2 "owner-gather" (first-generation kernel, kept for A/B).  Instead of
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
// Mode 0: block-owner assembly.
// Pair table (one entry per (i,j) of the element, built on the host from node_terms / w_coeff):
//   idx  = barycentric-gradient indices a of the (up to two) terms of row node i (bytes 0,1) and
//          column node j (bytes 2,3); the kernel reads G_a from the packed geometry records
//   w[q] = W[i,a_ta][j,b_tb] for q = 2*ta + tb (0 where a term does not exist)
struct PairTable {
    std::vector<double> w;        // [4][npe*npe]
    std::vector<uint32_t> idx;    // [npe*npe]
};

template <int N, int DEG>
static PairTable make_pair_table() {
    constexpr int NPE = nodes_per_elem(N, DEG), PP = NPE * NPE;
    PairTable t;
    t.w.assign(4 * PP, 0.0);
    t.idx.assign(PP, 0);
    for (int i = 0; i < NPE; ++i)
        for (int j = 0; j < NPE; ++j) {
            const NodeTerms<N, DEG> ti = node_terms<N, DEG>(i), tj = node_terms<N, DEG>(j);
            const int ij = i * NPE + j;
            const int a0 = ti.a[0], a1 = ti.n == 2 ? ti.a[1] : ti.a[0];
            const int b0 = tj.a[0], b1 = tj.n == 2 ? tj.a[1] : tj.a[0];
            t.idx[ij] = (uint32_t)a0 | ((uint32_t)a1 << 8) | ((uint32_t)b0 << 16) | ((uint32_t)b1 << 24);
            for (int ta = 0; ta < ti.n; ++ta)
                for (int tb = 0; tb < tj.n; ++tb)
                    t.w[(2 * ta + tb) * PP + ij] = (DEG == 1) ? 1.0 : w_coeff<N>(ti.edge, ti.p[ta], tj.edge, tj.p[tb]);
        }
    return t;
}

// B[c][d] (+)= sum_{r,t} D[flat(r,c)][flat(d,t)] M[r][t]
// ORTHO: D has the orthotropic sparsity pattern (normal-normal block + diagonal shear entries; covers
// isotropic materials) -- the other 2/3 of the 81 products are structurally zero and are skipped.
template <int N, bool ACC, bool ORTHO = false>
__device__ __forceinline__ void contract_CM(const double *D, const double (&M)[N][N], double (&B)[N][N]) {
    constexpr int F = flat_len(N);
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int d = 0; d < N; ++d) {
            double s = ACC ? B[c][d] : 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int t = 0; t < N; ++t) {
                    const int a = flat_idx<N>(r, c), b = flat_idx<N>(d, t);
                    if (ORTHO && !((a < N && b < N) || a == b)) continue;
                    s = fma(D[a * F + b], M[r][t], s);
                }
            B[c][d] = s;
        }
}

// Packed element geometry for the block-owner kernel: 4 slots of 32 bytes per element, slot a =
// (G[0][a], G[1][a], G[2][a] (0 in 2D), vol) -- one 128-byte record per element, the unit of the TMA staging.
template <int N>
__global__ void k_pack_geom(int64_t nElems, const double *__restrict__ geom, double *__restrict__ geomP) {
    const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (t >= nElems * 4) return;
    const int64_t e = t >> 2;
    const int a = (int)(t & 3);
    constexpr int GS = 1 + N * (N + 1);
    const double *g = geom + e * GS;
    double o[4] = {0.0, 0.0, 0.0, g[0]};
    if (a <= N)
        for (int r = 0; r < N; ++r) o[r] = g[1 + r * (N + 1) + a];
    for (int q = 0; q < 4; ++q) geomP[t * 4 + q] = o[q];
}

__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// Accumulate the contributions of one segment (pair ids read interleaved from `lp`, nIt trips of
// the warp) into acc: M in geometry space, or -- with per-element materials -- the block itself.
struct GeomSlot { double g[3], vol; };
__device__ __forceinline__ GeomSlot ld_geom_slot(const double *p) {
    GeomSlot s;
    asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(s.g[0]), "=d"(s.g[1]), "=d"(s.g[2]), "=d"(s.vol) : "l"(p));
    return s;
}
__device__ __forceinline__ GeomSlot lds_geom_slot(const double *p) {      // 32-byte aligned shared-memory slot
    GeomSlot s;
    const double2 a = *reinterpret_cast<const double2 *>(p), b = *reinterpret_cast<const double2 *>(p + 2);
    s.g[0] = a.x; s.g[1] = a.y; s.g[2] = b.x; s.vol = b.y;
    return s;
}

// STAGED: the entries carry chunk-local element ids and the records are read from the shared-memory copy
// the TMA made (sGeom); otherwise global element ids and 256-bit gathers from HBM/L2.
template <int N, int DEG, bool PER_ELEM_D, int PP, bool STAGED>
__device__ __forceinline__ void accumulate_segment(const uint32_t *lp, int nIt, const double *__restrict__ geomP,
                                                   const double *sGeom, const uint32_t *sElems,
                                                   const double *__restrict__ Delem, const double (*sW)[PP],
                                                   const uint32_t *sIdx, double (&acc)[N][N]) {
    constexpr int F = flat_len(N);
    uint32_t vnext = nIt > 0 ? ld_stream_u32(lp) : kPlanSentinel;
    for (int it = 0; it < nIt; ++it) {
        const uint32_t v = vnext;
        if (it + 1 < nIt) vnext = ld_stream_u32(lp + 32 * (it + 1));
        if (v == kPlanSentinel) continue;       // slots are sorted by length: only trailing lanes idle
        const uint32_t e = v / (uint32_t)PP;
        const int ij = (int)(v - e * (uint32_t)PP);
        const uint32_t idx = sIdx[ij];
        const double *g = STAGED ? sGeom + e * 16 : geomP + (int64_t)e * 16;
        const GeomSlot A0 = STAGED ? lds_geom_slot(g + 4 * (idx & 0xffu)) : ld_geom_slot(g + 4 * (idx & 0xffu));
        const GeomSlot B0 = STAGED ? lds_geom_slot(g + 4 * ((idx >> 16) & 0xffu)) : ld_geom_slot(g + 4 * ((idx >> 16) & 0xffu));
        const double vol = A0.vol;
        double u0[N];
        double Mc[N][N];
        if (DEG == 1) {
#pragma unroll
            for (int r = 0; r < N; ++r) u0[r] = vol * A0.g[r];
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    if (PER_ELEM_D) Mc[r][q] = u0[r] * B0.g[q]; else acc[r][q] = fma(u0[r], B0.g[q], acc[r][q]);
                }
        } else {
            const GeomSlot A1 = STAGED ? lds_geom_slot(g + 4 * ((idx >> 8) & 0xffu)) : ld_geom_slot(g + 4 * ((idx >> 8) & 0xffu));
            const GeomSlot B1 = STAGED ? lds_geom_slot(g + 4 * (idx >> 24)) : ld_geom_slot(g + 4 * (idx >> 24));
            double u1[N];
            const double w00 = vol * sW[0][ij], w01 = vol * sW[1][ij], w10 = vol * sW[2][ij], w11 = vol * sW[3][ij];
#pragma unroll
            for (int r = 0; r < N; ++r) {
                u0[r] = fma(w10, A1.g[r], w00 * A0.g[r]);     // pairs with column term 0
                u1[r] = fma(w11, A1.g[r], w01 * A0.g[r]);     // pairs with column term 1
            }
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int q = 0; q < N; ++q) {
                    if (PER_ELEM_D) Mc[r][q] = fma(u1[r], B1.g[q], u0[r] * B0.g[q]);
                    else acc[r][q] = fma(u1[r], B1.g[q], fma(u0[r], B0.g[q], acc[r][q]));
                }
        }
        if (PER_ELEM_D) {
            double De[F * F];
            const double *dp = Delem + (int64_t)(STAGED ? sElems[e] : e) * (F * F);     // per-element D by GLOBAL element id
#pragma unroll
            for (int q = 0; q < F * F; ++q) De[q] = __ldg(dp + q);
            contract_CM<N, true>(De, Mc, acc);
        }
    }
}

// ---- TMA staging helpers (1-D bulk copies completing on an mbarrier) ----
__device__ __forceinline__ uint32_t asm_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void asm_mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(asm_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void asm_mbar_expect_tx(unsigned long long *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(asm_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void asm_mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(asm_smem_u32(bar)), "r"(parity)
            : "memory");
    }
}
__device__ __forceinline__ void asm_tma_g2s(void *dst, const void *src, uint32_t bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(asm_smem_u32(dst)), "l"(src), "r"(bytes), "r"(asm_smem_u32(bar))
                 : "memory");
}

template <int N, int DEG, bool PER_ELEM_D, bool ORTHO>
__global__ void __launch_bounds__(kBlkChunk, PER_ELEM_D ? 2 : 4)
k_assemble_blocks(int64_t nnzb, const int32_t *__restrict__ chunkRow, const int64_t *__restrict__ rowptr,
                  const uint16_t *__restrict__ planSegOff, const uint8_t *__restrict__ planNseg,
                  const uint16_t *__restrict__ segOrder, const int64_t *__restrict__ warpBase,
                  const uint32_t *__restrict__ list, const double *__restrict__ geomP, const MatD Dc,
                  const double *__restrict__ Delem, const double *__restrict__ pairW,
                  const uint32_t *__restrict__ pairIdx, const int64_t *__restrict__ elemPtr,
                  const uint32_t *__restrict__ elems, double *__restrict__ vals) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int PP = NPE * NPE;
    constexpr int NN = N * N;
    __shared__ __align__(128) double sGeom[kGeomCap * 16];   // TMA destination: the chunk's element records (G_a, vol) x 4
    __shared__ uint32_t sElems[kGeomCap];         // their global element ids
    __shared__ unsigned long long sBar;           // mbarrier the bulk copies complete on
    __shared__ double sW[4][PP];
    __shared__ uint32_t sIdx[PP];
    __shared__ double sPart[kSegSlots * NN];      // per-segment partial sums
    __shared__ long long sRowPtr[kBlkChunk + 2];  // the chunk's slice of rowptr
    const int t = threadIdx.x, lane = t & 31;
    const int64_t chunk = blockIdx.x;
    const int64_t k = chunk * kBlkChunk + t;
    // ---- set-up, one barrier: pair table, mbarrier, rowptr slice (one global round trip for all of it)
    for (int q = t; q < PP; q += kBlkChunk) {
        sIdx[q] = pairIdx[q];
#pragma unroll
        for (int w = 0; w < 4; ++w) sW[w][q] = pairW[w * PP + q];
    }
    const int64_t ePtr = elemPtr[chunk];
    const int nLocal = (int)(elemPtr[chunk + 1] - ePtr);          // 0: unstaged chunk (global-load path)
    if (t == 0) asm_mbar_init(&sBar, 1);
    int mySegs = 0, mySegOff = 0;
    if (k < nnzb) { mySegs = planNseg[k]; mySegOff = planSegOff[k]; }
    const int64_t rLo = chunkRow[chunk], rHi = chunkRow[chunk + 1];
    const bool rowsStaged = rHi - rLo + 1 <= kBlkChunk;
    const int nRowPtr = (int)min((int64_t)kBlkChunk, rHi - rLo + 1) + 1;      // rowptr[rLo .. rLo+nRowPtr-1]
    if (t < nRowPtr) sRowPtr[t] = rowptr[rLo + t];
    uint32_t myElem = 0;
    if (t < nLocal) myElem = elems[ePtr + t];
    __syncthreads();
    // TMA staging of the element geometry: one 128-byte bulk copy per distinct element of the chunk, issued
    // by the first nLocal threads, all completing on one mbarrier (waited on right before phase 1)
    if (nLocal > 0) {
        if (t == 0) asm_mbar_expect_tx(&sBar, (uint32_t)nLocal * 128u);
        if (t < nLocal) {
            sElems[t] = myElem;
            asm_tma_g2s(sGeom + t * 16, geomP + (int64_t)myElem * 16, 128u, &sBar);
        }
    }
    // where block (chunk, t) lives in the row-plane layout (kept in registers for the write-out)
    double *myDst = nullptr;
    int64_t myStride = 0;
    if (k < nnzb) {
        int64_t b0, n;
        if (rowsStaged) {
            int l = 0, h = (int)(rHi - rLo);
            while (l < h) {                      // last row with rowptr[row] <= k
                const int mid = (l + h + 1) >> 1;
                if (sRowPtr[mid] <= k) l = mid; else h = mid - 1;
            }
            b0 = sRowPtr[l]; n = sRowPtr[l + 1] - b0;
        } else {                                 // more than 256 (empty) rows inside one chunk: search in HBM
            int64_t lo = rLo, hi = rHi;
            while (lo < hi) {
                const int64_t mid = (lo + hi + 1) >> 1;
                if (rowptr[mid] <= k) lo = mid; else hi = mid - 1;
            }
            b0 = rowptr[lo]; n = rowptr[lo + 1] - b0;
        }
        myDst = vals + NN * b0 + N * (k - b0);
        myStride = N * n;
    }

    // phase 1: one segment per thread slot.  kSegSlots/32 warp-rounds sorted by decreasing trip count:
    // round 0 gives warp w the w-th longest; the remaining (shorter) ones go, longest first, to the warps
    // that had the least to do in round 0, so the warps of the CTA reach the barrier together.
    constexpr int nW = kBlkChunk / 32, nWR = kSegSlots / 32;
    static_assert(nWR >= nW && nWR <= 2 * nW, "one full and at most one more round");
    if (nLocal > 0) asm_mbar_wait(&sBar, 0);      // geometry records have landed
#pragma unroll 1
    for (int round = 0; round < 2; ++round) {
        const int w = t >> 5;
        if (round == 1 && w < 2 * nW - nWR) break;              // warp-uniform: no second item for this warp
        const int wrLocal = round == 0 ? w : nW + (nW - 1 - w);
        const int slot = wrLocal * 32 + lane;
        const int64_t wr = chunk * nWR + wrLocal;
        const int64_t base = warpBase[wr];
        const int nIt = (int)((warpBase[wr + 1] - base) >> 5);
        if (nIt == 0) continue;                  // warp-uniform
        const int sg = segOrder[chunk * kSegSlots + slot];
        double acc[N][N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int q = 0; q < N; ++q) acc[r][q] = 0.0;
        if (nLocal > 0)
            accumulate_segment<N, DEG, PER_ELEM_D, PP, true>(list + base + lane, nIt, geomP, sGeom, sElems, Delem, sW, sIdx, acc);
        else
            accumulate_segment<N, DEG, PER_ELEM_D, PP, false>(list + base + lane, nIt, geomP, sGeom, sElems, Delem, sW, sIdx, acc);
        if (sg != 0xffff) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int q = 0; q < N; ++q) sPart[(r * N + q) * kSegSlots + sg] = acc[r][q];
        }
    }
    __syncthreads();

    // phase 2: block (chunk, t) sums its segments in list order, then the material contraction
    double out[N][N];
    {
        double M[N][N];
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int q = 0; q < N; ++q) M[r][q] = 0.0;
        for (int p = 0; p < mySegs; ++p)
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int q = 0; q < N; ++q) M[r][q] += sPart[(r * N + q) * kSegSlots + mySegOff + p];
        if (PER_ELEM_D) {
#pragma unroll
            for (int r = 0; r < N; ++r)
#pragma unroll
                for (int q = 0; q < N; ++q) out[r][q] = M[r][q];
        } else {
            contract_CM<N, false, ORTHO>(Dc.d, M, out);
        }
    }
    // write-out straight from registers: a block is N runs of N doubles (one per plane); the lanes of a warp
    // hold consecutive blocks of a row, so the N stores of a plane fill its sectors together and every
    // block of K is written exactly once
    if (myDst) {
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int q = 0; q < N; ++q) myDst[r * myStride + q] = out[r][q];
    }
}

void ensure_packed_geometry(mfem_b200_ctx *c) {
    if (c->geomPValid) return;
    MFEM_REQUIRE(c->geomValid, MFEM_B200_ERR_INVALID, "packed geometry: no mesh geometry");
    if (c->geomP.n != (size_t)c->nElems * 16) c->geomP.alloc((size_t)c->nElems * 16);
    if (c->N == 3) k_pack_geom<3><<<grid_for(c->nElems * 4, 256), 256, 0, c->stream>>>(c->nElems, c->geom, c->geomP);
    else k_pack_geom<2><<<grid_for(c->nElems * 4, 256), 256, 0, c->stream>>>(c->nElems, c->geom, c->geomP);
    c->launches++;
    c->geomPValid = true;
}

template <int N, int DEG>
static void launch_assemble_blocks(mfem_b200_ctx *c) {
    cudaStream_t s = c->stream;
    constexpr int PP = nodes_per_elem(N, DEG) * nodes_per_elem(N, DEG);
    if (c->pairTabKey != N * 10 + DEG) {
        const PairTable tab = make_pair_table<N, DEG>();
        c->pairW.alloc(tab.w.size());
        c->pairIdx.alloc(tab.idx.size());
        MFEM_CUDA(cudaMemcpyAsync(c->pairW, tab.w.data(), tab.w.size() * 8, cudaMemcpyHostToDevice, s));
        MFEM_CUDA(cudaMemcpyAsync(c->pairIdx, tab.idx.data(), tab.idx.size() * 4, cudaMemcpyHostToDevice, s));
        MFEM_CUDA(cudaStreamSynchronize(s));     // `tab` is pageable host memory going out of scope
        c->pairTabKey = N * 10 + DEG;
    }
    static_assert(PP <= 100, "pair table sized for at most 10 nodes per element");
    ensure_packed_geometry(c);
    const int64_t nChunks = (c->nnzb + kBlkChunk - 1) / kBlkChunk;
    // orthotropic sparsity pattern of the constant tensor (isotropic included)?
    constexpr int F = flat_len(N);
    bool ortho = !c->perElemD;
    for (int a = 0; a < F && ortho; ++a)
        for (int b = 0; b < F; ++b)
            if (!((a < N && b < N) || a == b) && c->Dconst.d[a * F + b] != 0.0) { ortho = false; break; }
#define MFEM_BLK_LAUNCH(PE_, OR_, DELEM_)                                                                              \
    do {                                                                                                              \
        auto kern = k_assemble_blocks<N, DEG, PE_, OR_>;                                                              \
        /* static shared memory ~45 KB per CTA: ask for the largest carve-out so 4 CTAs fit an SM */                  \
        MFEM_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));                   \
        kern<<<(unsigned)nChunks, kBlkChunk, 0, s>>>(c->nnzb, c->planChunkRow, c->rowptr, c->planSegOff, c->planNseg,  \
                                                     c->planSegOrder, c->planWarpBase, c->planList, c->geomP, c->Dconst, \
                                                     DELEM_, c->pairW, c->pairIdx, c->planElemPtr, c->planElems, c->vals); \
    } while (0)
    if (c->perElemD) MFEM_BLK_LAUNCH(true, false, c->Delem.p);
    else if (ortho) MFEM_BLK_LAUNCH(false, true, nullptr);
    else MFEM_BLK_LAUNCH(false, false, nullptr);
#undef MFEM_BLK_LAUNCH
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
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
        launch_assemble_blocks<N, DEG>(c);
    } else if (c->opt_assembly == 2) {
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
    MFEM_REQUIRE(!c->externalMatrix, MFEM_B200_ERR_INVALID, "assemble: the matrix of this handle was set with set_matrix_triplets");
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

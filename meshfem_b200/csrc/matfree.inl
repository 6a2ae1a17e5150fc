// K10: the PCG's operator WITHOUT the assembled matrix (quadratic elements).
//
// y = mask(K x) [, x.y] evaluated from the mesh: the quantity Simulator::applyStiffnessMatrix
// (LinearElasticity.hh:801-823) sums element by element, with the per-element product
// perElementStiffness * x (LinearElasticity.hh:165-232) evaluated by elem_apply (elem_math.cuh)
// on the degree-2 rule instead of through a stored 30x30 matrix.  Why: the block-CSR SpMV streams
// 76 bytes per 3x3 block -- 2060 bytes per node of a quadratic-tet mesh, 30.9 GB per product on the
// 10.2 M-element workload -- and is HBM-bound at 6.8 ms.  The mesh-based operator reads 128 bytes of
// packed geometry + 40 bytes of 16-bit tables per element and moves one partial sum per (chunk, DoF)
// once through HBM: 4.4 GB per product in all.
//
// Two kernels, no atomics, bit-reproducible.  Default (option mf_chunked = 1):
//   k_mf_chunk     one CTA per chunk of 64 consecutive elements (tables: setup.cu build_mf_chunks): stage the chunk's
//                  DISTINCT x blocks in shared memory, one thread per element evaluates elem_apply (~770 FP64
//                  instructions) and puts its results back into shared memory, the CTA sums them per distinct DoF
//                  in a fixed order and writes ONE partial per (chunk, DoF) -- 25.6 M partials instead of 102 M
//                  (element, node) slots on the 10.2 M-element workload;
//   k_mf_gather    one thread per DoF row: sum the row's partials in the fixed order of a per-row list, apply the
//                  Dirichlet mask, write y and accumulate x.y (same two-stage deterministic reduction as the SpMV
//                  epilogue, same scal slot).
// Measured on that workload: 1.15 + 0.33 ms per product against 6.8 ms for the SpMV.
// A/B variant (mf_chunked = 0): k_mf_elements, one thread per element writing its 30 results into its own slots
// elemY[e*npe + i][c] (packed, or padded to one 32-byte sector), gathered by 4 or 8 lanes per row through the
// incidence list of the symbolic phase -- 1.11 + 1.47 ms.
// The assembled matrix stays what the preconditioner set-up, export and the C-ABI spmv read; the operator is used where
// the Krylov loop multiplies (option `matrix_free`: -1 auto = quadratic tetrahedra of a mesh, 0 never, 1 whenever a
// mesh + material are present).
//
// Included by solver.cu (needs its reduction helpers and load wrappers).

constexpr int kMfThreads = 128;

__device__ __forceinline__ void ld_slot4(const double *p, uint64_t pol, double &a, double &b, double &c, double &d) {
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
        : "=d"(a), "=d"(b), "=d"(c), "=d"(d) : "l"(p), "l"(pol));
}

__device__ __forceinline__ void st_slot4(double *p, double a, double b, double c, double d) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" :: "l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
// one result slot of elemY: N doubles, or -- PAD, 3D -- a 32-byte aligned record of 4 (one sector, one 256-bit access)
template <int N, bool PAD>
struct MfSlot { static constexpr int stride = (N == 3 && PAD) ? 4 : N; };

template <int N, int DEG, bool PER_ELEM_D, bool PAD>
__global__ void __launch_bounds__(kMfThreads, 3)
k_mf_elements(int64_t nElems, const int32_t *__restrict__ elemDof, const double *__restrict__ geomP, const MatD Dc,
              const double *__restrict__ Delem, const int32_t *__restrict__ perm, const double *__restrict__ x,
              double *__restrict__ elemY, const int *status) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int F = flat_len(N);
    if (status && status[ST_STATE] != 0) return;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    int32_t nd[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) nd[j] = ld_stream_s32(elemDof + e * NPE + j, polStream);
    double Ga[N + 1][N], vol = 0.0;
#pragma unroll
    for (int a = 0; a <= N; ++a) {
        double g0, g1, g2, v;
        ld_slot4(geomP + e * 16 + a * 4, polStream, g0, g1, g2, v);
        Ga[a][0] = g0; Ga[a][1] = g1;
        if (N == 3) Ga[a][2] = g2;
        vol = v;
    }
    double xe[NPE][N];
#pragma unroll
    for (int j = 0; j < NPE; ++j)
#pragma unroll
        for (int d = 0; d < N; ++d) xe[j][d] = ld_keep_f64(x + (int64_t)nd[j] * N + d, polKeep);
    const double *D = PER_ELEM_D ? Delem + (perm ? (int64_t)perm[e] : e) * (F * F) : Dc.d;   // materials keep the caller's element order
    double ye[NPE * N];
    elem_apply<N, DEG>(Ga, vol, D, [&](int j, int d) { return xe[j][d]; },
                       [&](int i, int c, double v) { ye[i * N + c] = v; });
    if (N == 3 && PAD) {
        double *out = elemY + e * (NPE * 4);
#pragma unroll
        for (int i = 0; i < NPE; ++i) st_slot4(out + 4 * i, ye[i * N], ye[i * N + 1], ye[i * N + N - 1], 0.0);
    } else {
        // the element's record is 16-byte aligned (NPE * N is even): 128-bit stores
        static_assert((NPE * N) % 2 == 0, "element record must hold an even number of doubles");
        double2 *out = reinterpret_cast<double2 *>(elemY + e * (NPE * N));
#pragma unroll
        for (int k = 0; k < NPE * N / 2; ++k) out[k] = make_double2(ye[2 * k], ye[2 * k + 1]);
    }
}

// Chunked element kernel: one CTA = CH = 32 / 64 / 128 consecutive elements (tables: setup.cu build_mf_chunks).
//   1. the chunk's DISTINCT x blocks are staged in shared memory once (buf[c][u], u = chunk-local DoF index);
//   2. every thread (element) picks its npe blocks through the 16-bit local indices, evaluates elem_apply and puts its
//      npe results back into the same buffer, at their rank in the order sorted by chunk-local DoF;
//   3. the results are summed per chunk-local DoF -- a contiguous range, (element, local node) order inside -- and written as
//      ONE partial per (chunk, DoF), contiguous per chunk -- 2.2-3x fewer bytes than one slot per (element, node), for
//      this kernel's stores and for the gather kernel's loads alike.
template <int N, int DEG, bool PER_ELEM_D, int kMfChunk, int WARPS = 12>      // WARPS resident per SM: 12 (<= 168 registers), 16 (128, a few spills) or 20 (96, ~600 bytes of spills)
__global__ void __launch_bounds__(kMfChunk, WARPS * 32 / kMfChunk)
k_mf_chunk(int64_t nElems, const int32_t *__restrict__ chunkBase, const int32_t *__restrict__ chunkDof,
           const uint16_t *__restrict__ localIdx, const uint16_t *__restrict__ csrPtr, const uint16_t *__restrict__ rankOfSlot,
           const double *__restrict__ geomP, const MatD Dc, const double *__restrict__ Delem,
           const double *__restrict__ x, double *__restrict__ partialY, const int *status) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int F = flat_len(N);
    constexpr int S = kMfChunk * NPE;
    __shared__ double buf[N * S];
    if (status && status[ST_STATE] != 0) return;
    const int64_t b = blockIdx.x;
    const int t = threadIdx.x;
    const int64_t e = b * kMfChunk + t;
    const bool valid = e < nElems;
    const uint64_t polStream = l2_policy_evict_first(), polKeep = l2_policy_evict_last();
    // Loads are issued in program order and a warp stalls at the first USE of a pending result, so everything that does
    // not depend on the chunk's extent goes out first (geometry, the 16-bit tables), then the extent, then -- in one
    // batch per dependency level -- the DoF ids, the x blocks and the extents of the final sums.
    double Ga[N + 1][N], vol = 0.0;
    uint32_t lr[NPE];                           // chunk-local DoF index | rank of the result << 16
    if (valid) {
#pragma unroll
        for (int a = 0; a <= N; ++a) {
            double g0, g1, g2, v;
            ld_slot4(geomP + e * 16 + a * 4, polStream, g0, g1, g2, v);
            Ga[a][0] = g0; Ga[a][1] = g1;
            if (N == 3) Ga[a][2] = g2;
            vol = v;
        }
#pragma unroll
        for (int i = 0; i < NPE; ++i)
            lr[i] = (uint32_t)localIdx[b * S + i * kMfChunk + t] | ((uint32_t)rankOfSlot[b * S + i * kMfChunk + t] << 16);
    }
    // the chunk's fixed-stride tables are read without knowing its DoF count (entries past it are 0 = valid)
    const uint16_t *ptr = csrPtr + b * (S + 1);
    const int32_t *dofs = chunkDof + b * S;
    constexpr int UB = 3;                       // chunk-local DoFs per thread handled in one batch (typical count / chunk = 2.5)
    int pf0[UB], pf1[UB];
    int64_t d[UB];
#pragma unroll
    for (int j = 0; j < UB; ++j) {
        const int u = t + j * kMfChunk;
        d[j] = u < S ? dofs[u] : 0;
        pf0[j] = u < S ? ptr[u] : 0;
        pf1[j] = u < S ? ptr[u + 1] : 0;
    }
    const int base = chunkBase[b], nu = chunkBase[b + 1] - base;
    {
        double xv[UB][N];
#pragma unroll
        for (int j = 0; j < UB; ++j)
#pragma unroll
            for (int c = 0; c < N; ++c) xv[j][c] = ld_keep_f64(x + d[j] * N + c, polKeep);
#pragma unroll
        for (int j = 0; j < UB; ++j)
            if (t + j * kMfChunk < nu) {
#pragma unroll
                for (int c = 0; c < N; ++c) buf[c * S + t + j * kMfChunk] = xv[j][c];
            }
    }
    for (int u = t + UB * kMfChunk; u < nu; u += kMfChunk) {      // rare: more than UB distinct DoFs per element of the chunk
        const int64_t du = dofs[u];
#pragma unroll
        for (int c = 0; c < N; ++c) buf[c * S + u] = ld_keep_f64(x + du * N + c, polKeep);
    }
    __syncthreads();
    double xe[NPE][N];
    if (valid) {
#pragma unroll
        for (int i = 0; i < NPE; ++i)
#pragma unroll
            for (int c = 0; c < N; ++c) xe[i][c] = buf[c * S + (lr[i] & 0xffffu)];
    }
    __syncthreads();                            // everyone holds its x blocks: the buffer now takes the results
    if (valid) {
        const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
        elem_apply<N, DEG>(Ga, vol, D, [&](int j, int d) { return xe[j][d]; },
                           [&](int i, int c, double v) { buf[c * S + (lr[i] >> 16)] = v; });
    }
    __syncthreads();
    // results sit in the order sorted by chunk-local DoF: DoF u owns the contiguous range ptr[u] .. ptr[u+1]
    auto reduce_one = [&](int u, int k0, int k1) {
        double acc[N];
#pragma unroll
        for (int c = 0; c < N; ++c) acc[c] = 0.0;
        for (int k = k0; k < k1; ++k)
#pragma unroll
            for (int c = 0; c < N; ++c) acc[c] += buf[c * S + k];
#pragma unroll
        for (int c = 0; c < N; ++c) partialY[(int64_t)(base + u) * N + c] = acc[c];
    };
#pragma unroll
    for (int j = 0; j < UB; ++j)
        if (t + j * kMfChunk < nu) reduce_one(t + j * kMfChunk, pf0[j], pf1[j]);
    for (int u = t + UB * kMfChunk; u < nu; u += kMfChunk) reduce_one(u, ptr[u], ptr[u + 1]);
}

// one slot of elemY -> a[0..N)
template <int N, bool PAD>
__device__ __forceinline__ void ld_result_slot(const double *elemY, int slot, uint64_t pol, bool l1, double (&a)[N]) {
    if (N == 3 && PAD) {
        double pad;
        ld_slot4(elemY + (int64_t)slot * 4, pol, a[0], a[1], a[N - 1], pad);
    } else if (N == 2) {
        asm("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;"
            : "=d"(a[0]), "=d"(a[1]) : "l"(elemY + (int64_t)slot * 2), "l"(pol));
    } else {
#pragma unroll
        for (int c = 0; c < N; ++c)
            a[c] = l1 ? ld_keep_f64(elemY + (int64_t)slot * N + c, pol) : ld_stream_f64(elemY + (int64_t)slot * N + c, pol);
    }
}

template <int N, bool MASKED, bool DOT, bool PAD, int LPR = 4>
__global__ void __launch_bounds__(kVecThreads)
k_mf_gather(int64_t nb, const int64_t *__restrict__ incPtr, const int32_t *__restrict__ incList,
            const double *__restrict__ elemY, const double *__restrict__ x, double *__restrict__ y,
            const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut,
            const int *status, int slotPolicy) {
    static_assert((LPR >= N || LPR == 1) && (LPR & (LPR - 1)) == 0, "lanes per DoF row: 1, or a power of two >= N");
    constexpr unsigned FULL = 0xffffffffu;
    if (status && status[ST_STATE] != 0) return;
    // The slots are the only data of this kernel with reuse in L2 (the other slots of a fetched 64/128-byte line belong
    // to rows swept a little later): they must outlive the pure streams (lists, extents, x, y).
    const uint64_t polStream = l2_policy_evict_first();
    const uint64_t polSlot = (slotPolicy == 1 || slotPolicy == 3) ? l2_policy_evict_last() : (slotPolicy == 2 ? l2_policy_evict_normal() : polStream);
    const bool l1 = slotPolicy == 3;            // packed slots: let the three 8-byte loads of a slot share its L1 line
    const int sl = threadIdx.x & (LPR - 1);
    const int64_t stride = (int64_t)gridDim.x * blockDim.x / LPR;
    double dot = 0.0;
    for (int64_t rowBase = ((int64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31)) / LPR; rowBase < nb; rowBase += stride) {
        const int64_t row = rowBase + ((threadIdx.x & 31) / LPR);
        int64_t k0 = 0, k1 = 0;
        if (row < nb) { k0 = incPtr[row]; k1 = incPtr[row + 1]; }
        double acc[N];
#pragma unroll
        for (int c = 0; c < N; ++c) acc[c] = 0.0;
        int64_t k = k0 + sl;
        for (; k + LPR < k1; k += 2 * LPR) {     // two incidences per lane in flight
            const int s0 = ld_stream_s32(incList + k, polStream), s1 = ld_stream_s32(incList + k + LPR, polStream);
            double a0[N], a1[N];
            ld_result_slot<N, PAD>(elemY, s0, polSlot, l1, a0);
            ld_result_slot<N, PAD>(elemY, s1, polSlot, l1, a1);
#pragma unroll
            for (int c = 0; c < N; ++c) acc[c] = (acc[c] + a0[c]) + a1[c];
        }
        if (k < k1) {
            double a0[N];
            ld_result_slot<N, PAD>(elemY, ld_stream_s32(incList + k, polStream), polSlot, l1, a0);
#pragma unroll
            for (int c = 0; c < N; ++c) acc[c] += a0[c];
        }
        // fixed-shape tree over the LPR lanes (the same value in every lane of the group)
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int o = 1; o < LPR; o <<= 1) acc[c] += __shfl_xor_sync(FULL, acc[c], o);
        if (LPR == 1) {
            if (row < nb) {
#pragma unroll
                for (int c = 0; c < N; ++c) {
                    double out = acc[c];
                    if (MASKED && fixedMask[row * N + c]) out = 0.0;
                    y[row * N + c] = out;
                    if (DOT) dot = fma(out, x[row * N + c], dot);
                }
            }
        } else if (row < nb && sl < N) {
            double out = sl == 0 ? acc[0] : (sl == 1 ? acc[1] : acc[N - 1]);
            if (MASKED && fixedMask[row * N + sl]) out = 0.0;
            y[row * N + sl] = out;
            if (DOT) dot = fma(out, x[row * N + sl], dot);
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

// is the mesh-based operator usable / chosen for this handle?
static bool matrix_free_eligible(mfem_b200_ctx *c) {
    return !c->externalMatrix && c->nElems > 0 && c->patternValid && c->geomValid && c->haveMaterial && c->elemDof.p && c->incPtr.p &&
           c->incList.p && c->totalInc == c->nElems * (int64_t)c->npe;
}
static bool use_matrix_free(mfem_b200_ctx *c) {
    if (c->opt_matrix_free == 0 || !matrix_free_eligible(c)) return false;
    if (c->opt_matrix_free > 0) return true;
    return c->deg == 2 && c->N == 3;            // auto: where the stored matrix is 5x the mesh (27 blocks per row)
}

template <int N, int DEG, bool PAD>
static void launch_matrix_free_nd(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot, int phases) {
    PcgWork &w = c->work;
    cudaStream_t s = c->stream;
    const bool chunked = c->opt_mf_chunked != 0;
    const bool ordered = !chunked && c->opt_mf_elem_order != 0;
    if (chunked) { build_mf_chunks(c); ensure_packed_geometry(c); }
    else if (ordered) build_mf_plan(c);
    else ensure_packed_geometry(c);
    const int32_t *elemDof = ordered ? c->mfElemDof.p : c->elemDof.p;
    const double *geomP = ordered ? c->mfGeomP.p : c->geomP.p;
    const int32_t *incList = chunked ? c->mfIncList2.p : (ordered ? c->mfIncList.p : c->incList.p);
    const int64_t *incPtr = chunked ? c->mfIncPtr2.p : c->incPtr.p;
    const int32_t *perm = ordered ? c->mfPerm.p : nullptr;
    const size_t need = chunked ? (size_t)c->mfPartials * N : (size_t)c->nElems * c->npe * MfSlot<N, PAD>::stride;
    if (c->elemY.n != need) c->elemY.alloc(need);
    const int *status = (masked && dot) ? w.status.p : nullptr;      // in-loop launches turn into no-ops once the solve left "running"
    const int grid = grid_for(c->nElems, kMfThreads);
    if (!(phases & 1)) {}
    else if (chunked) {
        const int ch = c->mfChunkElems;
        const unsigned nChunks = (unsigned)((c->nElems + ch - 1) / ch);
#define MFEM_CHUNK(PE_, CH_, DELEM_)                                                                                        \
    do {                                                                                                                    \
        if (c->opt_mf_chunk_warps == 20)                                                                                    \
            k_mf_chunk<N, DEG, PE_, CH_, 20><<<nChunks, CH_, 0, s>>>(c->nElems, c->mfChunkBase, c->mfChunkDof, c->mfLocalIdx, c->mfCsrPtr, \
                                                                     c->mfCsrList, c->geomP, c->Dconst, DELEM_, x, c->elemY, status); \
        else if (c->opt_mf_chunk_warps == 16)                                                                               \
            k_mf_chunk<N, DEG, PE_, CH_, 16><<<nChunks, CH_, 0, s>>>(c->nElems, c->mfChunkBase, c->mfChunkDof, c->mfLocalIdx, c->mfCsrPtr, \
                                                                     c->mfCsrList, c->geomP, c->Dconst, DELEM_, x, c->elemY, status); \
        else                                                                                                                \
            k_mf_chunk<N, DEG, PE_, CH_, 12><<<nChunks, CH_, 0, s>>>(c->nElems, c->mfChunkBase, c->mfChunkDof, c->mfLocalIdx, c->mfCsrPtr, \
                                                                     c->mfCsrList, c->geomP, c->Dconst, DELEM_, x, c->elemY, status); \
    } while (0)
        if (c->perElemD) {
            if (ch == 32) MFEM_CHUNK(true, 32, c->Delem.p);
            else if (ch == 64) MFEM_CHUNK(true, 64, c->Delem.p);
            else MFEM_CHUNK(true, 128, c->Delem.p);
        } else {
            if (ch == 32) MFEM_CHUNK(false, 32, nullptr);
            else if (ch == 64) MFEM_CHUNK(false, 64, nullptr);
            else MFEM_CHUNK(false, 128, nullptr);
        }
#undef MFEM_CHUNK
    }
    else if (c->perElemD)
        k_mf_elements<N, DEG, true, PAD><<<grid, kMfThreads, 0, s>>>(c->nElems, elemDof, geomP, c->Dconst, c->Delem, perm, x, c->elemY, status);
    else
        k_mf_elements<N, DEG, false, PAD><<<grid, kMfThreads, 0, s>>>(c->nElems, elemDof, geomP, c->Dconst, nullptr, nullptr, x, c->elemY, status);
    const int lpr = c->opt_mf_gather_lanes == 8 ? 8 : (c->opt_mf_gather_lanes == 4 ? 4 : (c->opt_mf_gather_lanes == 1 ? 1 : (chunked ? 1 : 8)));   // 0 = auto
    const int pol = c->opt_mf_gather_policy;
    const int64_t ctas = (c->nDofs * lpr + kVecThreads - 1) / kVecThreads;
    const int ggrid = (int)std::max<int64_t>(1, std::min<int64_t>(ctas, std::min<int64_t>(kMaxPartials, (int64_t)sm_count(c) * 8)));
    if (!(phases & 2)) {}
    else if (masked && dot && lpr == 1)
        k_mf_gather<N, true, true, PAD, 1><<<ggrid, kVecThreads, 0, s>>>(c->nDofs, incPtr, incList, c->elemY, x, y, c->fixedMask,
                                                                         w.partials, w.ticket, w.scal.p + S_PAP, status, pol);
    else if (masked && dot && lpr == 8)
        k_mf_gather<N, true, true, PAD, 8><<<ggrid, kVecThreads, 0, s>>>(c->nDofs, incPtr, incList, c->elemY, x, y, c->fixedMask,
                                                                         w.partials, w.ticket, w.scal.p + S_PAP, status, pol);
    else if (masked && dot)
        k_mf_gather<N, true, true, PAD><<<ggrid, kVecThreads, 0, s>>>(c->nDofs, incPtr, incList, c->elemY, x, y, c->fixedMask,
                                                                      w.partials, w.ticket, w.scal.p + S_PAP, status, pol);
    else if (masked)
        k_mf_gather<N, true, false, PAD><<<ggrid, kVecThreads, 0, s>>>(c->nDofs, incPtr, incList, c->elemY, x, y, c->fixedMask,
                                                                       nullptr, nullptr, nullptr, nullptr, pol);
    else
        k_mf_gather<N, false, false, PAD><<<ggrid, kVecThreads, 0, s>>>(c->nDofs, incPtr, incList, c->elemY, x, y, nullptr,
                                                                        nullptr, nullptr, nullptr, nullptr, pol);
    c->launches += (phases & 1) + ((phases >> 1) & 1);
}

template <int N>
static void launch_matrix_free(mfem_b200_ctx *c, const double *x, double *y, bool masked, bool dot, int phases = 3) {
    const bool pad = N == 3 && c->opt_mf_slot_pad != 0 && !c->opt_mf_chunked;      // chunk partials are always packed
    if (c->deg == 2) {
        if (pad) launch_matrix_free_nd<N, 2, N == 3>(c, x, y, masked, dot, phases);
        else launch_matrix_free_nd<N, 2, false>(c, x, y, masked, dot, phases);
    } else {
        if (pad) launch_matrix_free_nd<N, 1, N == 3>(c, x, y, masked, dot, phases);
        else launch_matrix_free_nd<N, 1, false>(c, x, y, masked, dot, phases);
    }
}

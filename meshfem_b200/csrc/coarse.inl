// Multilevel aggregation preconditioner of the PCG (option coarse_aggregates != 0):
//
//     M^-1 = B0^-1  +  P1 B1^-1 P1^T  +  Z2 E2^-1 Z2^T ,          Z2 = P1 P2 ,   E2 = Z2^T K_ff Z2
//
// B0 = the block-Jacobi of solver.cu; level 1 = rigid-body modes (6 per aggregate in 3D, 3 in 2D: the tentative
// prolongator of aggregation AMG) of SMALL aggregates (option coarse_fine_nodes, ~tens of nodes) with B1 = the
// diagonal 6x6 blocks of P1^T K_ff P1; level 2 = rigid-body modes of LARGE aggregates (at most coarse_aggregates of
// them), E2 dense and inverted explicitly.  Aggregates are the cells of two NESTED regular box grids over the bounding
// box of the DoFs (every small box lies in exactly one large box), so the rigid modes of a large box are exact
// combinations of the modes of its small boxes (t = T + W x d, w = W with d = centre(small) - centre(large)):
// the fine-level transfers touch level 1 only, and levels 1 <-> 2 talk through a table of a few hundred thousand
// 6-vectors that lives in L2.  Everything is additive: an iteration still costs ONE fine SpMV.
//
// Why: the cantilever workloads are bending-dominated; block-Jacobi PCG needs 6074 iterations on cfg5, the large
// boxes alone (round 1's two-level method) 324 with 4096 boxes, and tools/proto_three_level.py measures another
// 1.6x fewer with the level-1 term (40x8x8 quadratic: 1106 -> 271 -> 168).
//
// Per application (all inside the captured CUDA graph, fused with the PCG's own vector passes):
//   k_pcg_update      x += a p, r -= a Ap, z = B0^-1 r, partial r.z / r.r, AND c1 += P1^T r   (segmented warp
//                     reduction over runs of equal slots, head lanes add with FP64 atomics)
//   k_coarse_level1   per level-1 slot: y1 = B1^-1 c1, r.z += c1.y1, c2 += P2^T c1
//   [N ranks: ONE all-reduce of (r.z, r.r, c2) -- 2 + 6 S2 doubles]
//   k_coarse_gemv     y2 = E2^-1 c2 (warp per row), r.z += c2.y2 in a fixed order (same bits on every rank)
//   k_pcg_direction   p = z + mask(R1 (y1 + P2 y2)) + beta p, closes the iteration
// Set-up per assembly: E2 (one pass over K, segmented shuffle reductions + atomics), cuSOLVER potrf/potri through
// dlopen (no link-time dependency; one handle per device for the life of the process), B1 (a second pass over K)
// and its 6x6 inverses (Cholesky per aggregate that drops the modes with a vanishing pivot: tiny aggregates whose nodes
// do not span six modes simply lose the missing directions).  The box grids, slots and centred positions depend on the mesh
// only and are kept across assemblies.
//
// Several GPUs: every rank lays its box grids over the DoFs it OWNS (large ids = rank * stride + box).  DoFs shared
// between ranks take no part in level 1; they are attached directly to their OWNER's large box ("pass-through" slots
// S1 + A, centred at the large box) on every sharer -- one owner-contributes interface sum-exchange at set-up makes
// the rows of Z2 agree on all sharers -- so E2 = sum over ranks of Z_loc^T K_loc Z_loc is one all-reduce at set-up,
// the restriction runs over owned DoFs, the prolongation over all local DoFs, and level 1 needs no communication.
// The restriction's FP64 atomics make the solve not bit-reproducible run to run with this option on.
// Included by solver.cu inside namespace mfem (which includes <cusolverDn.h> and <dlfcn.h> for it).

struct CoarseBoxes {                 // nested grids: small box q1 = min(b*r - 1, floor((x - lo) * scale1)), large = q1 / r
    double lo[3], scale1[3];
    int b[3], r[3];
};

struct CoarseSpace {
    int M = 0;                        // modes per aggregate
    int64_t S2 = 0, nc2 = 0;          // large aggregates over all ranks, M * S2
    int64_t S1 = 0, R = 1;            // this rank's small boxes (slots [0, S1), large-box-major: slot / R = local large box)
    int64_t n1 = 0;                   // S1 + S2: level-1 slots (the last S2 are the pass-through slots of the large boxes)
    int64_t aggBase = 0;              // first large id of this rank
    bool level1 = false;              // B1 term active
    int64_t meshVersion = -1;         // structure below was built for this mesh / interface / option state
    int optCoarse = 0, optFine = 0;
    DevBuf<int32_t> agg1;             // [nDofs] level-1 slot of every local DoF
    DevBuf<double> Y1;                // [nDofs*N] position relative to the centre of its slot
    DevBuf<double> shift;             // [S1*N] centre(small) - centre(its large box)
    DevBuf<double> B1inv;             // [S1*M*M]
    DevBuf<double> D1;                // [S1*M(M+1)/2] the blocks B1inv was computed from (upper triangles, column by column)
    DevBuf<double> c1, y1;            // [n1*M]
    DevBuf<double> Einv;              // [nRowsLoc*nc2] rows rowBase.. of E2^-1 (row-major).  One rank: all of it; N ranks:
                                      // the rows of the large boxes this rank owns (the dense level is ROW-SPLIT: every
                                      // rank computes its slice of y2 = E2^-1 c2, one all-gather completes it)
    DevBuf<double> Efac;              // [nc2*nc2] N ranks only: the assembled / all-reduced / factorised E2
    int64_t rowBase = 0, nRowsLoc = 0;
    DevBuf<double> y2;                // [nc2]
};

// q = R_i^T v (M values) for the rigid modes at relative position y: translations, then rotations
// 3D: (0,-z,y), (z,0,-x), (-y,x,0); 2D: (-y,x)
template <int N>
__device__ __forceinline__ void coarse_Rt(const double *y, const double *v, double *q) {
    if (N == 3) {
        q[0] = v[0]; q[1] = v[1]; q[2] = v[2];
        q[3] = y[1] * v[2] - y[2] * v[1];
        q[4] = y[2] * v[0] - y[0] * v[2];
        q[5] = y[0] * v[1] - y[1] * v[0];
    } else {
        q[0] = v[0]; q[1] = v[1];
        q[2] = y[0] * v[1] - y[1] * v[0];
    }
}
// v = R_i c
template <int N>
__device__ __forceinline__ void coarse_R(const double *y, const double *c, double *v) {
    if (N == 3) {
        v[0] = c[0] + y[2] * c[4] - y[1] * c[5];
        v[1] = c[1] - y[2] * c[3] + y[0] * c[5];
        v[2] = c[2] + y[1] * c[3] - y[0] * c[4];
    } else {
        v[0] = c[0] - y[1] * c[2];
        v[1] = c[1] + y[0] * c[2];
    }
}
// level 1 -> 2 restriction of the coefficients of a small box whose centre sits at d from its large box's: the
// moment about the large centre is cw + d x ct
template <int N>
__device__ __forceinline__ void coarse_shift_restrict(const double *d, double *c) {
    if (N == 3) {
        c[3] += d[1] * c[2] - d[2] * c[1];
        c[4] += d[2] * c[0] - d[0] * c[2];
        c[5] += d[0] * c[1] - d[1] * c[0];
    } else {
        c[2] += d[0] * c[1] - d[1] * c[0];
    }
}
// level 2 -> 1 prolongation (the transpose): q += (T + W x d, W)
template <int N>
__device__ __forceinline__ void coarse_shift_prolong(const double *d, const double *Y, double *q) {
    if (N == 3) {
        q[0] += Y[0] + Y[4] * d[2] - Y[5] * d[1];
        q[1] += Y[1] + Y[5] * d[0] - Y[3] * d[2];
        q[2] += Y[2] + Y[3] * d[1] - Y[4] * d[0];
        q[3] += Y[3]; q[4] += Y[4]; q[5] += Y[5];
    } else {
        q[0] += Y[0] - Y[2] * d[1];
        q[1] += Y[1] + Y[2] * d[0];
        q[2] += Y[2];
    }
}

// lanes [lane+1, lane+o] hold no run head  <=>  lane+o belongs to lane's run
__device__ __forceinline__ bool coarse_same_run(unsigned heads, int lane, int o) {
    return (lane + o < 32) && (((heads >> (lane + 1)) & ((1u << o) - 1u)) == 0u);
}
// segmented sum of q[0..M) over runs of equal keys inside the warp (all 32 lanes call); returns true on run heads,
// which then hold the run's sums
template <int M>
__device__ __forceinline__ bool coarse_segmented_sum(long long key, double (&q)[M], int lane) {
    const long long prev = __shfl_up_sync(0xffffffffu, key, 1);
    const bool head = (lane == 0) || (prev != key);
    const unsigned heads = __ballot_sync(0xffffffffu, head);
    if (heads != 0xffffffffu) {          // warp-uniform: every lane its own run -> nothing to add
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const bool take = coarse_same_run(heads, lane, o);
#pragma unroll
            for (int m = 0; m < M; ++m) {
                const double other = __shfl_down_sync(0xffffffffu, q[m], o);
                if (take) q[m] += other;
            }
        }
    }
    return head;
}

// ---------------------------------------------------------------------------------------------------------------
// structure (per mesh): box grids, slots, centred positions

// lowest node of every DoF (periodic DoFs have several nodes; any would do, the lowest is deterministic)
__global__ void k_coarse_first_node(int64_t nNodes, const int32_t *__restrict__ nodeDof, int32_t *firstNode) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < nNodes) atomicMin(&firstNode[nodeDof[n]], (int32_t)n);
}

// order-preserving map double -> uint64 for atomicMin / atomicMax
__device__ __forceinline__ unsigned long long coarse_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
static double coarse_unkey(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double x;
    std::memcpy(&x, &b, sizeof(double));
    return x;
}
// keys[0..N) = min, keys[N..2N) = max over the owned DoFs that have a node; keys[2N] = their number
template <int N>
__global__ void k_coarse_bbox(int64_t nb, const uint8_t *__restrict__ owned, int64_t nNodes, const int32_t *__restrict__ firstNode,
                              const double *__restrict__ nodes, unsigned long long *keys) {
    double lo[N], hi[N];
    unsigned long long cnt = 0;
    for (int k = 0; k < N; ++k) { lo[k] = 1e300; hi[k] = -1e300; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        if (owned && !owned[i]) continue;
        const int64_t node = firstNode[i];
        if (node < 0 || node >= nNodes) continue;
        ++cnt;
        for (int k = 0; k < N; ++k) { const double x = nodes[node * N + k]; lo[k] = fmin(lo[k], x); hi[k] = fmax(hi[k], x); }
    }
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[k] = fmin(lo[k], __shfl_xor_sync(0xffffffffu, lo[k], o));
            hi[k] = fmax(hi[k], __shfl_xor_sync(0xffffffffu, hi[k], o));
        }
        if ((threadIdx.x & 31) == 0) {
            atomicMin(&keys[k], coarse_key(lo[k]));
            atomicMax(&keys[N + k], coarse_key(hi[k]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&keys[2 * N], cnt);
}

// per owned DoF: its large box (global id; -1 for DoFs owned elsewhere) and, when it takes part in level 1 (owned, not
// shared with another rank, has a node), its small slot (large-box-major); sums for the box centres
template <int N>
__global__ void k_coarse_box_agg(int64_t nb, const uint8_t *__restrict__ owned, const uint8_t *__restrict__ shared, int64_t nNodes,
                                 const int32_t *__restrict__ firstNode, const double *__restrict__ nodes, const CoarseBoxes bx,
                                 int64_t aggBase, int64_t R, bool level1, int32_t *__restrict__ aggBig, int32_t *__restrict__ slot,
                                 double *cen2 /* [Sr*(N+1)] */, double *cen1 /* [S1*(N+1)] */) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    if (owned && !owned[i]) { aggBig[i] = -1; slot[i] = -1; return; }
    const int64_t node = firstNode[i];
    const bool positioned = node >= 0 && node < nNodes;
    int64_t big = 0, loc = 0;
    if (positioned)
        for (int k = 0; k < N; ++k) {
            int q = (int)floor((nodes[node * N + k] - bx.lo[k]) * bx.scale1[k]);
            q = max(0, min(bx.b[k] * bx.r[k] - 1, q));
            big = big * bx.b[k] + q / bx.r[k];
            loc = loc * bx.r[k] + q % bx.r[k];
        }
    aggBig[i] = (int32_t)(aggBase + big);
    const bool l1 = level1 && positioned && !(shared && shared[i]);
    const int64_t s = big * R + loc;
    slot[i] = l1 ? (int32_t)s : -1;
    if (!positioned) return;
    for (int k = 0; k < N; ++k) {
        const double x = nodes[node * N + k];
        atomicAdd(&cen2[big * (N + 1) + k], x);
        if (l1) atomicAdd(&cen1[s * (N + 1) + k], x);
    }
    atomicAdd(&cen2[big * (N + 1) + N], 1.0);
    if (l1) atomicAdd(&cen1[s * (N + 1) + N], 1.0);
}
// owned DoFs: T[i] = (large id + 1, position - centre of the large box); rows of DoFs owned elsewhere stay zero
template <int N>
__global__ void k_coarse_pack_T(int64_t nb, int64_t aggBase, const int32_t *__restrict__ aggBig, int64_t nNodes,
                                const int32_t *__restrict__ firstNode, const double *__restrict__ nodes,
                                const double *__restrict__ cen2, double *__restrict__ T) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t a = aggBig[i];
    if (a < 0) return;
    T[i * (N + 1)] = (double)(a + 1);
    const int64_t node = firstNode[i];
    if (node < 0 || node >= nNodes) return;
    const double cnt = cen2[(a - aggBase) * (N + 1) + N];
    for (int k = 0; k < N; ++k) T[i * (N + 1) + 1 + k] = nodes[node * N + k] - cen2[(a - aggBase) * (N + 1) + k] / cnt;
}
// shift[s] = centre(small box s) - centre(its large box); empty boxes: 0
template <int N>
__global__ void k_coarse_shift(int64_t S1, int64_t R, const double *__restrict__ cen1, const double *__restrict__ cen2,
                               double *__restrict__ shift) {
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S1) return;
    const double c1 = cen1[s * (N + 1) + N], c2 = cen2[(s / R) * (N + 1) + N];
    for (int k = 0; k < N; ++k)
        shift[s * N + k] = (c1 > 0.0 && c2 > 0.0) ? cen1[s * (N + 1) + k] / c1 - cen2[(s / R) * (N + 1) + k] / c2 : 0.0;
}
// after the owner-contributes sum-exchange of T: level-1 slot and centred position of every local DoF
template <int N>
__global__ void k_coarse_finish_dofs(int64_t nb, int64_t S2, int64_t S1, const double *__restrict__ T, const int32_t *__restrict__ slot,
                                     const double *__restrict__ shift, int32_t *__restrict__ agg1, double *__restrict__ Y1, int *bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    int64_t a = llrint(T[i * (N + 1)]) - 1;
    if (a < 0 || a >= S2) { atomicAdd(bad, 1); a = 0; }
    const int64_t s = slot[i];
    if (s >= 0) {
        agg1[i] = (int32_t)s;
        for (int k = 0; k < N; ++k) Y1[i * N + k] = T[i * (N + 1) + 1 + k] - shift[s * N + k];
    } else {
        agg1[i] = (int32_t)(S1 + a);
        for (int k = 0; k < N; ++k) Y1[i * N + k] = T[i * (N + 1) + 1 + k];
    }
}

// large box and position relative to its centre, from the level-1 description
template <int N>
__device__ __forceinline__ int64_t coarse_big_of(int64_t s, int64_t S1, int64_t R, int64_t aggBase, const double *__restrict__ Y1i,
                                                 const double *__restrict__ shift, double *y2) {
    if (s < S1) {
        for (int k = 0; k < N; ++k) y2[k] = Y1i[k] + shift[s * N + k];
        return aggBase + s / R;
    }
    for (int k = 0; k < N; ++k) y2[k] = Y1i[k];
    return s - S1;
}

// ---------------------------------------------------------------------------------------------------------------
// values (per assembly)

// C = R_i^T K_ij R_j (M x M, row-major) for block j of row `row`, fixed rows / columns masked out
template <int N>
__device__ __forceinline__ void coarse_block_product(const double *__restrict__ vals, int64_t b0, int64_t n, int64_t j, const bool *fi,
                                                     const uint8_t *__restrict__ fixedMaskCol, const double *yi, const double *yj,
                                                     double *C) {
    constexpr int M = N == 3 ? 6 : 3;
    double T[N][M];                                      // T[r][b] = sum_c K[r][c] R_j[c][b]
    for (int r = 0; r < N; ++r) {
        double krow[N];
        for (int cc = 0; cc < N; ++cc) krow[cc] = (fi[r] || fixedMaskCol[cc]) ? 0.0 : vals[val_index<N>(b0, n, j, r, cc)];
        coarse_Rt<N>(yj, krow, T[r]);                    // (K_row R_j) = R_j^T K_row^T
    }
    for (int b = 0; b < M; ++b) {
        double tcol[N], q[M];
        for (int r = 0; r < N; ++r) tcol[r] = T[r][b];
        coarse_Rt<N>(yi, tcol, q);
        for (int a = 0; a < M; ++a) C[a * M + b] = q[a];
    }
}

// E_loc += Z2_loc^T K_loc Z2_loc: one warp per block row, one lane per block; lanes whose columns fall into the same
// large box are combined by a segmented shuffle reduction, run heads add the M x M result to E.  The factorisation
// reads only the row-major UPPER triangle of E (= column-major lower), so block columns left of the diagonal block
// are skipped.
template <int N>
__global__ void __launch_bounds__(256)
k_coarse_matrix(int64_t nb, int64_t S2, int64_t S1, int64_t R, int64_t aggBase, const int32_t *__restrict__ agg1,
                const double *__restrict__ Y1, const double *__restrict__ shift, const int64_t *__restrict__ rowptr,
                const int32_t *__restrict__ colidx, const double *__restrict__ vals, const uint8_t *__restrict__ fixedMask, double *E) {
    constexpr int M = N == 3 ? 6 : 3;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nc = (int64_t)M * S2;
    for (int64_t row = warp; row < nb; row += nWarps) {
        const int64_t b0 = rowptr[row], n = rowptr[row + 1] - b0;
        double yi[N];
        bool fi[N];
        const int64_t ai = coarse_big_of<N>(agg1[row], S1, R, aggBase, Y1 + row * N, shift, yi);
        for (int k = 0; k < N; ++k) fi[k] = fixedMask[row * N + k] != 0;
        for (int64_t j0 = 0; j0 < n; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool active = j < n;
            double C[M * M];
            long long aj = -1 - lane;                    // inactive lanes: unique keys, never merged
            for (int a = 0; a < M * M; ++a) C[a] = 0.0;
            if (active) {
                const int64_t col = colidx[b0 + j];
                double yj[N];
                aj = coarse_big_of<N>(agg1[col], S1, R, aggBase, Y1 + col * N, shift, yj);
                coarse_block_product<N>(vals, b0, n, j, fi, fixedMask + col * N, yi, yj, C);
            }
            const bool head = coarse_segmented_sum<M * M>(aj, C, lane);
            if (active && head && aj >= ai)
                for (int a = 0; a < M; ++a)
                    for (int b = 0; b < M; ++b)
                        if (C[a * M + b] != 0.0) atomicAdd(&E[(ai * M + a) * nc + aj * M + b], C[a * M + b]);
        }
    }
}

// dead modes (all-zero rows: boxes without free variables, rotations without coordinates) get a unit diagonal;
// a relative diagonal shift keeps E positive definite when a box has too few free nodes for six independent
// modes (1e-8: far above rounding, far below anything that matters for a preconditioner)
__global__ void k_coarse_regularize(int64_t nc, double *E, double shift) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc) return;
    const double d = E[i * nc + i];
    E[i * nc + i] = (d == 0.0) ? 1.0 : d * (1.0 + shift);
}
// potri leaves the inverse in one triangle (column-major lower = row-major upper): mirror it
__global__ void k_coarse_symmetrize(int64_t nc, double *E) {
    const int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c < r && c < nc) E[r * nc + c] = E[c * nc + r];
}

// D1[s] += sym(R1_i^T K_ij R1_j) over the block pairs inside small box s (upper triangle of the symmetric M x M
// block, MS = M (M + 1) / 2 values, column by column): one warp per block row, lanes over its blocks, one warp sum,
// MS atomics per row.  The sum over all pairs of a box is symmetric, a single pair is not, hence the symmetrisation.
template <int N>
__global__ void __launch_bounds__(256)
k_coarse_diag1(int64_t nb, int64_t S1, const int32_t *__restrict__ agg1, const double *__restrict__ Y1,
               const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx, const double *__restrict__ vals,
               const uint8_t *__restrict__ fixedMask, double *D1 /* [S1*MS] */) {
    constexpr int M = N == 3 ? 6 : 3, MS = M * (M + 1) / 2;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < nb; row += nWarps) {
        const int64_t s = agg1[row];
        if (s >= S1) continue;                           // warp-uniform
        const int64_t b0 = rowptr[row], n = rowptr[row + 1] - b0;
        double yi[N];
        bool fi[N];
        for (int k = 0; k < N; ++k) { yi[k] = Y1[row * N + k]; fi[k] = fixedMask[row * N + k] != 0; }
        double acc[MS];
        for (int a = 0; a < MS; ++a) acc[a] = 0.0;
        for (int64_t j = lane; j < n; j += 32) {
            const int64_t col = colidx[b0 + j];
            if (agg1[col] != s) continue;
            double yj[N], C[M * M];
            for (int k = 0; k < N; ++k) yj[k] = Y1[col * N + k];
            coarse_block_product<N>(vals, b0, n, j, fi, fixedMask + col * N, yi, yj, C);
            int t = 0;
            for (int b = 0; b < M; ++b)
                for (int a = 0; a <= b; ++a, ++t) acc[t] += 0.5 * (C[a * M + b] + C[b * M + a]);
        }
#pragma unroll
        for (int a = 0; a < MS; ++a) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_xor_sync(0xffffffffu, acc[a], o);
        }
        if (lane < MS) {
            double v = 0.0;
#pragma unroll
            for (int a = 0; a < MS; ++a) if (a == lane) v = acc[a];
            if (v != 0.0) atomicAdd(&D1[s * MS + lane], v);
        }
    }
}

// B1inv[s] = inverse of the symmetric M x M block D1[s] by a Cholesky factorisation that DROPS the modes whose pivot is
// below 1e-10 of the largest diagonal entry (dead modes of fixed boxes, rotations of boxes whose nodes are collinear):
// a dropped mode is removed from the level-1 space of its box, i.e. B1inv is the inverse of the remaining principal
// submatrix embedded in zeros -- still symmetric positive semi-definite.  All loops unroll: registers only.
// (A cyclic-Jacobi pseudo-inverse with dynamically indexed local arrays gave wrong 6x6 results on the device while the
// same source was right on the host; the factorisation is cheaper anyway.)
template <int M>
__global__ void k_coarse_invert1(int64_t S1, const double *__restrict__ D1, double *__restrict__ B1inv) {
    constexpr int MS = M * (M + 1) / 2;
    const int64_t s = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (s >= S1) return;
    double L[M][M], W[M][M];
    {
        int t = 0;
#pragma unroll
        for (int b = 0; b < M; ++b)
#pragma unroll
            for (int a = 0; a <= b; ++a, ++t) L[a][b] = L[b][a] = D1[s * MS + t];
    }
    double dmax = 0.0;
#pragma unroll
    for (int a = 0; a < M; ++a) dmax = fmax(dmax, L[a][a]);
    const double tol = 1e-10 * dmax;
    // in-place Cholesky on the lower triangle; a dropped mode gets a zero column and L[k][k] = 0
#pragma unroll
    for (int k = 0; k < M; ++k) {
        double d = L[k][k];
#pragma unroll
        for (int j = 0; j < k; ++j) d -= L[k][j] * L[k][j];
        const bool live = d > tol && dmax > 0.0;
        const double piv = live ? sqrt(d) : 0.0, ipiv = live ? 1.0 / piv : 0.0;
        L[k][k] = piv;
#pragma unroll
        for (int i = k + 1; i < M; ++i) {
            double v = L[i][k];
#pragma unroll
            for (int j = 0; j < k; ++j) v -= L[i][j] * L[k][j];
            L[i][k] = v * ipiv;
        }
    }
    // W = L^-1 on the live modes (lower triangular; rows / columns of dropped modes stay zero)
#pragma unroll
    for (int c = 0; c < M; ++c) {
#pragma unroll
        for (int i = 0; i < M; ++i) W[i][c] = 0.0;
        const double icc = L[c][c] > 0.0 ? 1.0 / L[c][c] : 0.0;
        W[c][c] = icc;
#pragma unroll
        for (int i = c + 1; i < M; ++i) {
            double v = 0.0;
#pragma unroll
            for (int j = c; j < i; ++j) v -= L[i][j] * W[j][c];
            W[i][c] = (L[i][i] > 0.0 && icc > 0.0) ? v / L[i][i] : 0.0;
        }
    }
#pragma unroll
    for (int a = 0; a < M; ++a)
#pragma unroll
        for (int b = 0; b < M; ++b) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < M; ++k) v += W[k][a] * W[k][b];
            B1inv[s * M * M + a * M + b] = v;
        }
}

// ---------------------------------------------------------------------------------------------------------------
// per application

// c1 += P1^T v for the 32 consecutive DoFs of a warp (called by k_pcg_update with v = the new residual, zero on
// fixed components and on DoFs owned by another rank)
template <int N>
__device__ __forceinline__ void coarse_restrict_warp(bool active, int64_t i, const double *v, const int32_t *__restrict__ agg1,
                                                     const double *__restrict__ Y1, double *c1, int lane) {
    constexpr int M = N == 3 ? 6 : 3;
    double q[M];
    long long key = -1 - lane;
#pragma unroll
    for (int m = 0; m < M; ++m) q[m] = 0.0;
    if (active) {
        key = agg1[i];
        double y[N];
#pragma unroll
        for (int k = 0; k < N; ++k) y[k] = Y1[i * N + k];
        coarse_Rt<N>(y, v, q);
    }
    const bool head = coarse_segmented_sum<M>(key, q, lane);
    if (active && head)
#pragma unroll
        for (int m = 0; m < M; ++m)
            if (q[m] != 0.0) atomicAdd(&c1[key * M + m], q[m]);
}

// level 1, one thread per slot: y1 = B1^-1 c1, red[0] += c1.y1, c2 += P2^T c1 (segmented over the slots of one large
// box, which are consecutive); clears c1 for the next application
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_level1(int64_t S1, int64_t n1, int64_t R, int64_t aggBase, bool level1, double *c1, double *__restrict__ y1,
                const double *__restrict__ B1inv, const double *__restrict__ shift, double *red /* [2 + nc2] */, const int *status) {
    constexpr int M = N == 3 ? 6 : 3;
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    double rz = 0.0;
    for (int64_t base = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) & ~31ll; base < n1; base += (int64_t)gridDim.x * blockDim.x) {
        const int64_t s = base + lane;
        double c[M];
        long long key = -1 - lane;
#pragma unroll
        for (int m = 0; m < M; ++m) c[m] = 0.0;
        if (s < n1) {
#pragma unroll
            for (int m = 0; m < M; ++m) { c[m] = c1[s * M + m]; c1[s * M + m] = 0.0; }
            if (s < S1) {
                key = aggBase + s / R;
                if (level1) {
                    const double *B = B1inv + s * M * M;
#pragma unroll
                    for (int a = 0; a < M; ++a) {
                        double y = 0.0;
#pragma unroll
                        for (int b = 0; b < M; ++b) y = fma(B[a * M + b], c[b], y);
                        y1[s * M + a] = y;
                        rz = fma(c[a], y, rz);
                    }
                }
                double d[N];
#pragma unroll
                for (int k = 0; k < N; ++k) d[k] = shift[s * N + k];
                coarse_shift_restrict<N>(d, c);
            } else {
                key = s - S1;
            }
        }
        const bool head = coarse_segmented_sum<M>(key, c, lane);
        if (s < n1 && head)
#pragma unroll
            for (int m = 0; m < M; ++m)
                if (c[m] != 0.0) atomicAdd(&red[2 + key * M + m], c[m]);
    }
    if (level1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) rz += __shfl_xor_sync(0xffffffffu, rz, o);
        __shared__ double sh[kVecThreads / 32];
        if (lane == 0) sh[threadIdx.x >> 5] = rz;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w];
            if (t != 0.0) atomicAdd(&red[0], t);
        }
    }
}

// y2[rowBase + k] = Einv[k, :] . c2 for the nRows local rows (one warp per row).  One rank (FINAL): the last CTA adds
// c2.y2 in a fixed order to the block-Jacobi + level-1 part of r.z: out[0] = red[0] + c2.y2, out[1] = red[1].
template <bool FINAL>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_gemv(int64_t nc, int64_t rowBase, int64_t nRows, const double *__restrict__ Einv, const double *__restrict__ red,
              double *__restrict__ y2, double *partials, unsigned *ticket, double *out, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    const double *c2 = red + 2;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (int64_t row = warp; row < nRows; row += nWarps) {
        const double *e = Einv + row * nc;
        double s0 = 0.0, s1 = 0.0;
        int64_t k = lane;
        for (; k + 32 < nc; k += 64) { s0 = fma(e[k], c2[k], s0); s1 = fma(e[k + 32], c2[k + 32], s1); }
        if (k < nc) s0 = fma(e[k], c2[k], s0);
        double s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { y2[rowBase + row] = s; dot = fma(s, c2[rowBase + row], dot); }
    }
    if (FINAL) {
        double v1[1] = {dot};
        block_reduce_store<1>(v1, partials);
        if (last_block(ticket)) {
            const double s = final_sum(partials, gridDim.x);
            if (threadIdx.x == 0) { out[0] = red[0] + s; out[1] = red[1]; }
        }
    }
}
// N ranks, after the all-gather of y2: out[0] = red[0] + c2.y2 in a fixed order (one CTA) -- every rank holds the same
// c2 and y2, hence the same bits; out[1] = red[1]
__global__ void __launch_bounds__(256) k_coarse_cy(int64_t nc, const double *__restrict__ red, const double *__restrict__ y2,
                                                   double *out, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    __shared__ double sh[256];
    const double *c2 = red + 2;
    double s = 0.0;
    for (int64_t k = threadIdx.x; k < nc; k += 256) s = fma(c2[k], y2[k], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = red[0] + sh[0]; out[1] = red[1]; }
}

// N ranks with the peer window: this rank's rows of y2 = E2^-1 c2, every row stored straight into ALL ranks' windows
// (lane q of the row's warp stores to rank q); the last CTA publishes the sequence number to the peers.
__global__ void __launch_bounds__(kVecThreads)
k_coarse_gemv_ship(PeerWin w, int64_t nc, int64_t rowBase, int64_t nRows, const double *__restrict__ Einv, const double *__restrict__ red,
                   unsigned *ticket, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    const double *c2 = red + 2;
    const unsigned long long sq = *w.seq(SET_AG) + 1;
    const int phase = (int)(sq & 1);
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = warp; row < nRows; row += nWarps) {
        const double *e = Einv + row * nc;
        double s0 = 0.0, s1 = 0.0;
        int64_t k = lane;
        for (; k + 32 < nc; k += 64) { s0 = fma(e[k], c2[k], s0); s1 = fma(e[k + 32], c2[k + 32], s1); }
        if (k < nc) s0 = fma(e[k], c2[k], s0);
        double s = s0 + s1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane < w.R) w.ag(lane, phase)[rowBase + row] = s;
    }
    __threadfence_system();
    if (last_block(ticket)) {
        if ((int)threadIdx.x < w.R) peer_publish(w.flags(threadIdx.x, SET_AG) + w.rank, sq);
    }
}
// one CTA: wait for every rank's rows, copy the gathered y2 out of the window, out[0] = red[0] + c2.y2 in a fixed order
// (every rank holds the same c2 and y2, hence the same bits), out[1] = red[1]; advances the all-gather sequence number
__global__ void __launch_bounds__(1024)
k_coarse_cy_gather(PeerWin w, int64_t nc, const double *__restrict__ red, double *__restrict__ y2, double *out, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    __shared__ double sh[1024];
    const unsigned long long sq = *w.seq(SET_AG) + 1;
    const int phase = (int)(sq & 1);
    if ((int)threadIdx.x < w.R && !peer_wait(w.flags(w.rank, SET_AG) + threadIdx.x, sq)) *w.err() = 1;
    __syncthreads();
    const double *c2 = red + 2, *g = w.ag(w.rank, phase);
    double s = 0.0;
    for (int64_t k = threadIdx.x; k < nc; k += 1024) {
        const double y = g[k];
        y2[k] = y;
        s = fma(c2[k], y, s);
    }
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 512; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { out[0] = red[0] + sh[0]; out[1] = red[1]; *w.seq(SET_AG) = sq; }
}

// coarse part of z for DoF i: R1_i (y1[slot] + P2 y2[large box])
template <int N>
__device__ __forceinline__ void coarse_prolong_dof(int64_t i, const int32_t *__restrict__ agg1, const double *__restrict__ Y1,
                                                   const double *__restrict__ y1, const double *__restrict__ y2,
                                                   const double *__restrict__ shift, int64_t S1, int64_t R, int64_t aggBase,
                                                   bool level1, double *v) {
    constexpr int M = N == 3 ? 6 : 3;
    const int64_t s = agg1[i];
    double q[M], y[N];
#pragma unroll
    for (int k = 0; k < N; ++k) y[k] = Y1[i * N + k];
    if (s < S1) {
        double d[N];
#pragma unroll
        for (int k = 0; k < N; ++k) d[k] = shift[s * N + k];
#pragma unroll
        for (int m = 0; m < M; ++m) q[m] = level1 ? y1[s * M + m] : 0.0;
        coarse_shift_prolong<N>(d, y2 + (aggBase + s / R) * M, q);
    } else {
#pragma unroll
        for (int m = 0; m < M; ++m) q[m] = y2[(s - S1) * M + m];
    }
    coarse_R<N>(y, q, v);
}

// ---- cuSOLVER through dlopen (setup only)
struct CusolverApi {
    void *lib = nullptr;
    cusolverStatus_t (*create)(cusolverDnHandle_t *) = nullptr;
    cusolverStatus_t (*destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*setStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*potrfBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potrf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potriBuf)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potri)(cusolverDnHandle_t, cublasFillMode_t, int, double *, int, double *, int, int *) = nullptr;
    cusolverStatus_t (*potrs)(cusolverDnHandle_t, cublasFillMode_t, int, int, const double *, int, double *, int, int *) = nullptr;
};
static CusolverApi &cusolver_api() {
    static CusolverApi api;
    if (api.lib) return api;
    for (const char *name : {"libcusolver.so.11", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    MFEM_REQUIRE(api.lib, MFEM_B200_ERR_INVALID, std::string("coarse_aggregates: cannot load libcusolver (") + dlerror() + ")");
    auto sym = [&](const char *n) {
        void *p = dlsym(api.lib, n);
        MFEM_REQUIRE(p, MFEM_B200_ERR_INVALID, std::string("coarse_aggregates: libcusolver lacks ") + n);
        return p;
    };
    api.create = reinterpret_cast<decltype(api.create)>(sym("cusolverDnCreate"));
    api.destroy = reinterpret_cast<decltype(api.destroy)>(sym("cusolverDnDestroy"));
    api.setStream = reinterpret_cast<decltype(api.setStream)>(sym("cusolverDnSetStream"));
    api.potrfBuf = reinterpret_cast<decltype(api.potrfBuf)>(sym("cusolverDnDpotrf_bufferSize"));
    api.potrf = reinterpret_cast<decltype(api.potrf)>(sym("cusolverDnDpotrf"));
    api.potriBuf = reinterpret_cast<decltype(api.potriBuf)>(sym("cusolverDnDpotri_bufferSize"));
    api.potri = reinterpret_cast<decltype(api.potri)>(sym("cusolverDnDpotri"));
    api.potrs = reinterpret_cast<decltype(api.potrs)>(sym("cusolverDnDpotrs"));
    return api;
}

static void free_coarse(mfem_b200_ctx *c) {
    delete c->coarse;
    c->coarse = nullptr;
}

__global__ void k_coarse_identity_cols(int64_t nc, int64_t rowBase, int64_t nCols, double *B /* [nc x nCols] column-major */) {
    const int64_t j = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (j < nCols) B[j * nc + rowBase + j] = 1.0;
}

// E (assembled in `Ebuf`: the row-major upper block triangle, which is all potrf reads) -> regularised -> Cholesky factor.
// One rank: explicit inverse in place (potri + mirror), Ebuf is cs.Einv.  N ranks: every rank factorises the same
// all-reduced E and then solves for ITS rows of the (symmetric) inverse only: E X = I[:, own rows] (potrs), so the
// O(n^3) part of the inversion and the per-iteration GEMV are split N ways.
// Returns false when E is not positive definite (the caller falls back to block-Jacobi in automatic mode, throws otherwise).
static bool invert_coarse_matrix(mfem_b200_ctx *c, CoarseSpace &cs, double *Ebuf, std::string &why) {
    cudaStream_t s = c->stream;
    const int64_t nc = cs.nc2;
    const bool multi = c->nRanks > 1;
    k_coarse_regularize<<<grid_for(nc, 256), 256, 0, s>>>(nc, Ebuf, 1e-8);
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
    CusolverApi &api = cusolver_api();
    // one cuSOLVER handle per device for the life of the process: creating one costs 0.2-1 s (module loading), which
    // every re-assembly would pay again (measured on cfg5: 1.17 s per step with a handle per call)
    static std::map<int, cusolverDnHandle_t> handles;
    cusolverDnHandle_t &h = handles[c->device];
    if (!h) MFEM_REQUIRE(api.create(&h) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "cusolverDnCreate failed");
    MFEM_REQUIRE(api.setStream(h, s) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "cusolverDnSetStream failed");
    int lw1 = 0, lw2 = 0;
    MFEM_REQUIRE(api.potrfBuf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, Ebuf, (int)nc, &lw1) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potrf_bufferSize failed");
    if (!multi)
        MFEM_REQUIRE(api.potriBuf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, Ebuf, (int)nc, &lw2) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potri_bufferSize failed");
    DevBuf<double> wbuf((size_t)std::max(lw1, lw2) + 1);
    DevBuf<int> info(1);
    int hinfo = 0;
    MFEM_REQUIRE(api.potrf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, Ebuf, (int)nc, wbuf, lw1, info) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potrf failed");
    MFEM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    if (hinfo != 0) {
        why = "coarse matrix is not positive definite (leading minor " + std::to_string(hinfo) + "): fewer aggregates, or a singular system";
        return false;
    }
    if (multi) {
        // rows rowBase .. rowBase + nRowsLoc of E^-1 = columns of the solution of E X = I[:, those columns]
        // (column-major nc x nRowsLoc == row-major nRowsLoc x nc)
        MFEM_CUDA(cudaMemsetAsync(cs.Einv, 0, cs.Einv.bytes(), s));
        k_coarse_identity_cols<<<grid_for(cs.nRowsLoc, 256), 256, 0, s>>>(nc, cs.rowBase, cs.nRowsLoc, cs.Einv);
        c->launches++;
        MFEM_REQUIRE(api.potrs(h, CUBLAS_FILL_MODE_LOWER, (int)nc, (int)cs.nRowsLoc, Ebuf, (int)nc, cs.Einv, (int)nc, info) == CUSOLVER_STATUS_SUCCESS,
                     MFEM_B200_ERR_CUDA, "potrs failed");
        MFEM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
        if (hinfo != 0) { why = "coarse matrix solve failed (" + std::to_string(hinfo) + ")"; return false; }
        MFEM_CUDA(cudaGetLastError());
        return true;
    }
    MFEM_REQUIRE(api.potri(h, CUBLAS_FILL_MODE_LOWER, (int)nc, Ebuf, (int)nc, wbuf, lw2, info) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potri failed");
    MFEM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    if (hinfo != 0) { why = "coarse matrix inversion failed (" + std::to_string(hinfo) + ")"; return false; }
    k_coarse_symmetrize<<<dim3((unsigned)grid_for(nc, 256), (unsigned)nc), 256, 0, s>>>(nc, Ebuf);
    c->launches++;
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_CUDA(cudaGetLastError());
    return true;
}

// near-cubic boxes, at most `budget` of them, over the extents L (flat directions get one layer)
static void coarse_choose_boxes(int N, const double *L, int64_t budget, int *b) {
    double vol = 1.0;
    int nd = 0;
    for (int k = 0; k < N; ++k) if (L[k] > 0.0) { vol *= L[k]; ++nd; }
    const double h = nd ? std::pow(vol / (double)std::max<int64_t>(budget, 1), 1.0 / nd) : 1.0;
    for (int k = 0; k < N; ++k) b[k] = L[k] > 0.0 ? (int)std::max<int64_t>(1, std::llround(L[k] / h)) : 1;
    auto count = [&]() { int64_t p = 1; for (int k = 0; k < N; ++k) p *= b[k]; return p; };
    while (count() > budget) {                       // rounding up overshot: shrink the direction with the most layers
        int kmax = 0;
        for (int k = 1; k < N; ++k) if (b[k] > b[kmax]) kmax = k;
        if (b[kmax] == 1) break;
        --b[kmax];
    }
}
// refinement of the large grid b into small boxes of about `target` DoFs each (density assumed uniform over the
// boxes that hold the nDofs DoFs); r[k] >= 1
static void coarse_choose_refinement(int N, const double *L, const int *b, int64_t nDofs, int target, int *r) {
    for (int k = 0; k < 3; ++k) r[k] = 1;
    if (target <= 0) return;
    double vol = 1.0;
    int nd = 0;
    for (int k = 0; k < N; ++k) if (L[k] > 0.0) { vol *= L[k]; ++nd; }
    if (!nd || nDofs <= 0) return;
    const double h1 = std::pow(vol * (double)target / (double)nDofs, 1.0 / nd);      // edge of a small box
    for (int k = 0; k < N; ++k)
        if (L[k] > 0.0) r[k] = (int)std::max<int64_t>(1, std::llround(L[k] / b[k] / h1));
}

// requested number of large boxes: explicit (> 0), or automatic (-1): one per ~400 DoFs (nodes), 16..2048, none below
// 30k DoFs.  Measured: 92,785 nodes: 26 boxes 414 iterations, 256 boxes 170 (block-Jacobi 1240); cfg5 (14.65 M nodes):
// 1024 / 2048 / 4096 boxes 307 / 254 / 213 iterations, but the dense inverse grows like boxes^3 (37 / 131 / 738 ms).
static int64_t coarse_budget(mfem_b200_ctx *c, int64_t nDofsGlobal) {
    if (c->opt_coarse > 0) return c->opt_coarse;
    if (c->opt_coarse == 0 || nDofsGlobal < 30000) return 0;
    return std::max<int64_t>(16, std::min<int64_t>(2048, nDofsGlobal / 400));
}

// Box grids, slots, centred positions (1..N ranks; on one rank the exchange is a no-op).
template <int N>
static void build_coarse_structure(mfem_b200_ctx *c, int64_t budget) {
    constexpr int M = N == 3 ? 6 : 3;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs;
    const bool multi = c->nRanks > 1;
    const uint8_t *ownedDev = multi ? halo_owned(c) : nullptr;
    const uint8_t *sharedDev = multi ? halo_shared(c) : nullptr;
    // large ids per rank: a fixed stride (the option and nRanks are the same everywhere), so global ids need no
    // negotiation; a rank that uses fewer boxes than its stride leaves dead modes behind
    int64_t Sr = std::max<int64_t>(1, std::min<int64_t>(budget, 32768 / M) / c->nRanks);
    if (!multi) Sr = std::min<int64_t>(Sr, std::max<int64_t>(1, nb / 8));
    const int64_t aggBase = Sr * c->rank;
    const bool havePositions = !c->externalMatrix && c->nNodes > 0;
    const int64_t nNodesEff = havePositions ? c->nNodes : 0;
    DevBuf<int32_t> firstNode((size_t)nb);
    MFEM_CUDA(cudaMemsetAsync(firstNode, 0x7f, firstNode.bytes(), s));
    CoarseBoxes bx;
    for (int k = 0; k < 3; ++k) { bx.lo[k] = 0.0; bx.scale1[k] = 0.0; bx.b[k] = 1; bx.r[k] = 1; }
    if (havePositions) {
        k_coarse_first_node<<<grid_for(c->nNodes, 256), 256, 0, s>>>(c->nNodes, c->nodeDof, firstNode);
        DevBuf<unsigned long long> keys(2 * N + 1);
        MFEM_CUDA(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * N, s));                 // min slots: +inf keys
        MFEM_CUDA(cudaMemsetAsync(keys.p + N, 0x00, sizeof(unsigned long long) * (N + 1), s));     // max slots: -inf keys; counter
        k_coarse_bbox<N><<<(int)std::min<int64_t>(grid_for(nb, 256), (int64_t)sm_count(c) * 8), 256, 0, s>>>(nb, ownedDev, c->nNodes, firstNode,
                                                                                                  c->nodes, keys);
        c->launches += 2;
        unsigned long long hk[2 * N + 1];
        MFEM_CUDA(cudaMemcpyAsync(hk, keys, sizeof(hk), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));
        double L[3] = {0.0, 0.0, 0.0};
        bool any = true;
        for (int k = 0; k < N; ++k) {
            const double lo = coarse_unkey(hk[k]), hi = coarse_unkey(hk[N + k]);
            if (!(hi >= lo)) { any = false; break; }                                          // this rank owns no positioned DoF
            bx.lo[k] = lo; L[k] = hi - lo;
        }
        if (any) {
            coarse_choose_boxes(N, L, Sr, bx.b);
            coarse_choose_refinement(N, L, bx.b, (int64_t)hk[2 * N], c->opt_coarse_fine, bx.r);
            for (int k = 0; k < N; ++k) bx.scale1[k] = L[k] > 0.0 ? (bx.b[k] * bx.r[k]) / L[k] : 0.0;
            if (!multi) { Sr = 1; for (int k = 0; k < N; ++k) Sr *= bx.b[k]; }     // one rank: no stride to keep, no dead ids
        }
    }
    int64_t nBoxes = 1, R = 1;
    for (int k = 0; k < N; ++k) { nBoxes *= bx.b[k]; R *= bx.r[k]; }
    const bool level1 = havePositions && c->opt_coarse_fine > 0 && R > 1;
    const int64_t S2 = Sr * c->nRanks;
    const int64_t S1 = level1 ? nBoxes * R : 0;
    free_coarse(c);
    c->coarse = new CoarseSpace();
    CoarseSpace &cs = *c->coarse;
    cs.M = M; cs.S2 = S2; cs.nc2 = M * S2; cs.S1 = S1; cs.R = level1 ? R : 1; cs.n1 = S1 + S2; cs.aggBase = aggBase; cs.level1 = level1;
    cs.agg1.alloc((size_t)nb);
    cs.Y1.alloc((size_t)nb * N);
    cs.shift.alloc((size_t)std::max<int64_t>(S1, 1) * N);
    cs.c1.alloc((size_t)cs.n1 * M); cs.y1.alloc((size_t)cs.n1 * M);
    cs.y2.alloc((size_t)cs.nc2);
    MFEM_CUDA(cudaMemsetAsync(cs.c1, 0, cs.c1.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(cs.y1, 0, cs.y1.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(cs.shift, 0, cs.shift.bytes(), s));
    {
        DevBuf<int32_t> aggBig((size_t)nb), slot((size_t)nb);
        DevBuf<double> T((size_t)nb * (N + 1));
        DevBuf<double> cen2((size_t)Sr * (N + 1)), cen1((size_t)std::max<int64_t>(S1, 1) * (N + 1));
        DevBuf<int> bad(1);
        MFEM_CUDA(cudaMemsetAsync(T, 0, T.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(cen2, 0, cen2.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(cen1, 0, cen1.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
        k_coarse_box_agg<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, ownedDev, sharedDev, nNodesEff, firstNode, c->nodes, bx, aggBase, cs.R,
                                                              level1, aggBig, slot, cen2, cen1);
        k_coarse_pack_T<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, aggBase, aggBig, nNodesEff, firstNode, c->nodes, cen2, T);
        if (S1) k_coarse_shift<N><<<grid_for(S1, 256), 256, 0, s>>>(S1, cs.R, cen1, cen2, cs.shift);
        c->launches += 3;
        halo_exchange_add(c, T, N + 1);                       // only the owner's rows are non-zero (no-op on one rank)
        k_coarse_finish_dofs<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, S2, S1, T, slot, cs.shift, cs.agg1, cs.Y1, bad);
        c->launches++;
        int nbad = 0;
        MFEM_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));                  // also: the temporaries go out of scope
        MFEM_CUDA(cudaGetLastError());
        // every rank reaches the collectives of the value phase even if this one is inconsistent: it poisons E instead
        cs.meshVersion = nbad ? -2 : c->meshVersion;
    }
    cs.optCoarse = c->opt_coarse; cs.optFine = c->opt_coarse_fine;
}

// E2^-1 and B1^-1 for the current matrix values and constraints.  Returns false (with the reason) if E2 is not SPD.
template <int N>
static bool build_coarse_values(mfem_b200_ctx *c, std::string &why) {
    constexpr int M = N == 3 ? 6 : 3, MS = M * (M + 1) / 2;
    cudaStream_t s = c->stream;
    CoarseSpace &cs = *c->coarse;
    const int64_t nb = c->nDofs, nc = cs.nc2;
    const bool multi = c->nRanks > 1;
    // one rank: E is assembled, factorised and inverted in cs.Einv; N ranks: in cs.Efac, and cs.Einv gets this rank's rows
    cs.rowBase = multi ? cs.aggBase * M : 0;
    cs.nRowsLoc = multi ? nc / c->nRanks : nc;
    if (cs.Einv.n != (size_t)cs.nRowsLoc * nc) cs.Einv.alloc((size_t)cs.nRowsLoc * nc);
    if (multi && cs.Efac.n != (size_t)nc * nc) cs.Efac.alloc((size_t)nc * nc);
    double *Ebuf = multi ? cs.Efac.p : cs.Einv.p;
    const int grid = (int)std::min<int64_t>((nb + 7) / 8, (int64_t)sm_count(c) * 8);
    {
        ScopedTimer t(c, "Coarse Matrix");
        MFEM_CUDA(cudaMemsetAsync(Ebuf, 0, sizeof(double) * nc * nc, s));
        if (cs.meshVersion == -2) MFEM_CUDA(cudaMemsetAsync(Ebuf, 0xff, sizeof(double), s));   // NaN in E[0]: the inversion fails on all ranks
        k_coarse_matrix<N><<<grid, 256, 0, s>>>(nb, cs.S2, cs.S1, cs.R, cs.aggBase, cs.agg1, cs.Y1, cs.shift, c->rowptr, c->colidx, c->vals,
                                                c->fixedMask, Ebuf);
        c->launches++;
        MFEM_CUDA(cudaGetLastError());
        if (multi) {
            // one all-reduce of the partial coarse matrices (chunked: keep single calls moderate)
            const size_t total = (size_t)nc * nc, chunk = (size_t)1 << 27;
            for (size_t off = 0; off < total; off += chunk)
                allreduce_sum(c, Ebuf + off, Ebuf + off, (int)std::min(chunk, total - off));
        }
    }
    {
        ScopedTimer t(c, "Coarse Inverse");
        if (!invert_coarse_matrix(c, cs, Ebuf, why)) return false;
    }
    if (cs.level1) {
        ScopedTimer t(c, "Coarse Level 1");
        if (cs.B1inv.n != (size_t)cs.S1 * M * M) cs.B1inv.alloc((size_t)cs.S1 * M * M);
        if (cs.D1.n != (size_t)cs.S1 * MS) cs.D1.alloc((size_t)cs.S1 * MS);
        MFEM_CUDA(cudaMemsetAsync(cs.D1, 0, cs.D1.bytes(), s));
        k_coarse_diag1<N><<<grid, 256, 0, s>>>(nb, cs.S1, cs.agg1, cs.Y1, c->rowptr, c->colidx, c->vals, c->fixedMask, cs.D1);
        k_coarse_invert1<M><<<grid_for(cs.S1, 128), 128, 0, s>>>(cs.S1, cs.D1, cs.B1inv);
        c->launches += 2;
        MFEM_CUDA(cudaStreamSynchronize(s));
        MFEM_CUDA(cudaGetLastError());
    }
    return true;
}

__global__ void k_count_owned(int64_t nb, const uint8_t *__restrict__ owned, double *out) {
    double cnt = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x)
        if (owned[i]) cnt += 1.0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    if ((threadIdx.x & 31) == 0 && cnt != 0.0) atomicAdd(out, cnt);
}
// number of DoFs over all ranks (owned ones counted once; integers below 2^53: exact in any order)
static int64_t coarse_global_dofs(mfem_b200_ctx *c) {
    if (c->nRanks <= 1) return c->nDofs;
    DevBuf<double> cnt(1);
    MFEM_CUDA(cudaMemsetAsync(cnt, 0, sizeof(double), c->stream));
    k_count_owned<<<(int)std::min<int64_t>(grid_for(c->nDofs, 256), 1024), 256, 0, c->stream>>>(c->nDofs, halo_owned(c), cnt);
    c->launches++;
    allreduce_sum(c, cnt, cnt, 1);
    double h = 0.0;
    MFEM_CUDA(cudaMemcpyAsync(&h, cnt, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    MFEM_CUDA(cudaStreamSynchronize(c->stream));
    return (int64_t)llround(h);
}

static void build_coarse(mfem_b200_ctx *c) {
    ScopedTimer timer(c, "Coarse Space");
    const int64_t budget = coarse_budget(c, c->opt_coarse < 0 ? coarse_global_dofs(c) : c->nDofs);
    if (budget <= 0) { free_coarse(c); return; }
    if (!c->coarse || c->coarse->meshVersion != c->meshVersion || c->coarse->optCoarse != c->opt_coarse ||
        c->coarse->optFine != c->opt_coarse_fine) {
        ScopedTimer t(c, "Coarse Structure");
        if (c->N == 3) build_coarse_structure<3>(c, budget); else build_coarse_structure<2>(c, budget);
    }
    std::string why;
    const bool ok = c->N == 3 ? build_coarse_values<3>(c, why) : build_coarse_values<2>(c, why);
    if (!ok) {
        free_coarse(c);
        // automatic mode: a coarse matrix that is not SPD (a singular system, e.g. a floating body handled by the
        // Lagrange-row algebra on the host) falls back to block-Jacobi alone -- still the device PCG, never a CPU path
        MFEM_REQUIRE(c->opt_coarse < 0, MFEM_B200_ERR_NOT_SPD, why);
    }
}

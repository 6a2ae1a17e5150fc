// Batched PCG: R right-hand sides advanced in lockstep against ONE stream of the matrix.
// (#included at the end of solver.cu: it reuses the SpMV building blocks and the reduction helpers.)
//
// The reference solves the flatLen(N) cell problems of periodic homogenization with one factorisation
// and flatLen(N) back-substitutions (PeriodicHomogenization.hh:34-54).  The iterative analogue: the
// R systems K u_r = f_r share K, and the PCG is bandwidth-bound on streaming K -- so the R Krylov
// iterations run together and every SpMV becomes an SpMM that reads the 72 B/block matrix once for all R
// vectors.  Each system keeps its own alpha/beta/residual and its own convergence state (a finished
// system is frozen: alpha = 0, direction not updated); the batch ends when all are finished.
//
// Vector layout: interleaved, X[(dof*N + c)*R + r] -- the R values a matrix scalar multiplies are
// contiguous (one or three 16-byte gathers), and the interface exchange moves them together
// (halo width N*R).

// scalar slots (doubles): RZ[R], PAP[R], RZ_NEW[R], RR[R], BB[R], TOL2
template <int R> struct MS { enum { RZ = 0, PAP = R, RZ_NEW = 2 * R, RR = 3 * R, BB = 4 * R, TOL2 = 5 * R, COUNT = 5 * R + 1 }; };
// status ints: [0] iterations, [1] global state, [2+r] state of system r, [2+R+r] iterations of system r

struct PcgWorkMulti {
    int R = 0;
    size_t n = 0;
    DevBuf<double> x, r, z, p, Ap, b;
    DevBuf<double> partials, scal, dotLoc;
    DevBuf<unsigned> ticket;
    DevBuf<int> status;
};

void free_work_multi(mfem_b200_ctx *c) {
    delete c->workMulti;
    c->workMulti = nullptr;
}

static PcgWorkMulti &ensure_work_multi(mfem_b200_ctx *c, int R) {
    const size_t n = (size_t)c->nvar() * R;
    if (c->workMulti && c->workMulti->R == R && c->workMulti->n == n) return *c->workMulti;
    free_work_multi(c);
    c->workMulti = new PcgWorkMulti();
    PcgWorkMulti &w = *c->workMulti;
    w.R = R; w.n = n;
    w.x.alloc(n); w.r.alloc(n); w.z.alloc(n); w.p.alloc(n); w.Ap.alloc(n); w.b.alloc(n);
    w.partials.alloc((size_t)2 * R * kMaxPartials);
    w.scal.alloc(8 * R + 8);
    w.dotLoc.alloc(4 * R);
    w.ticket.alloc(4);
    w.status.alloc(2 + 2 * R + 2);
    MFEM_CUDA(cudaMemsetAsync(w.ticket, 0, w.ticket.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.status, 0, w.status.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.scal, 0, w.scal.bytes(), c->stream));
    MFEM_CUDA(cudaMemsetAsync(w.dotLoc, 0, w.dotLoc.bytes(), c->stream));
    return w;
}

// R gathered values (contiguous) with the keep-in-L2 policy
template <int R>
__device__ __forceinline__ void gather_R(const double *p, uint64_t pol, double (&v)[R]) {
    if (R % 2 == 0) {
#pragma unroll
        for (int q = 0; q < R / 2; ++q)
            asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0,%1}, [%2], %3;" : "=d"(v[2 * q]), "=d"(v[2 * q + 1]) : "l"(p + 2 * q), "l"(pol) : "memory");
    } else {
#pragma unroll
        for (int q = 0; q < R; ++q)
            asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v[q]) : "l"(p + q), "l"(pol) : "memory");
    }
}

// Y = mask(K X) for R interleaved vectors [, dot[r] = X_r . Y_r].  Same row decomposition as k_bsr_spmv
// (one warp per block row, 32 lanes, chunks of 96 scalars per plane).
template <int N, int R, bool MASKED, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, 2)
k_bsr_spmm(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
           const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ Y,
           const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut, const int *status) {
    constexpr int NN = N * N;
    constexpr int LPR = 32, U = 3, CH = U * LPR;
    if (status && status[ST_STATE] != 0) return;
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
    double dot[R];
#pragma unroll
    for (int r = 0; r < R; ++r) dot[r] = 0.0;
    int64_t row = warpGlobal;
    int64_t nb0 = 0, nb1 = 0;
    if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
    for (; row < nb; ) {
        const int64_t b0 = nb0;
        const int L = (int)(nb1 - nb0) * N;
        const int64_t thisRow = row;
        row += rowStride;
        nb0 = nb1 = 0;
        if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
        const double *v = vals + b0 * NN + sl;
        const int32_t *ci = colidx + b0;
        double acc[N][R];
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int r = 0; r < R; ++r) acc[k][r] = 0.0;
        for (int base = 0; base < L; base += CH, v += CH, ci += CH / N) {
            const int rem = L - base - sl;
            int col[U];
            double a[U][N];
            ChunkLoader<N, LPR>::run(ci + jj[0], ci + jj[1], ci + jj[2], v, v + L, v + 2 * L, rem, polStream, col, a);
            __syncwarp();     // scheduling fence: all streaming loads leave before the dependent gathers
            double xr[U][R];
#pragma unroll
            for (int u = 0; u < U; ++u) gather_R<R>(X + ((int64_t)col[u] * N + cc[u]) * R, polKeep, xr[u]);
            __syncwarp();     // ... and all gathers before the first FMA
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < N; ++k)
#pragma unroll
                    for (int r = 0; r < R; ++r) acc[k][r] = fma(a[u][k], xr[u][r], acc[k][r]);
        }
        const bool fixedHere = MASKED && comp >= 0 && fixedMask[thisRow * N + (comp >= 0 ? comp : 0)];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            double t[N];
#pragma unroll
            for (int k = 0; k < N; ++k) t[k] = acc[k][r];
            const double out0 = fold_reduce<N, LPR>(t, sl);
            if (comp >= 0) {
                const double out = fixedHere ? 0.0 : out0;
                const int64_t idx = (thisRow * N + comp) * R + r;
                Y[idx] = out;
                if (DOT) dot[r] += out * X[idx];
            }
        }
    }
    if (DOT) {
        block_reduce_store<R>(dot, partials);
        if (last_block(ticket)) {
            for (int r = 0; r < R; ++r) {
                const double s = final_sum(partials + (size_t)r * gridDim.x, gridDim.x);
                if (threadIdx.x == 0) dotOut[r] = s;
            }
        }
    }
}

// Split variant for even R: the two half-warps of a warp work on the SAME block row, each for R/2 of the
// systems.  Both halves issue identical matrix loads (merged by the coalescer: no extra traffic), a lane
// carries N*R/2 accumulators instead of N*R (80 instead of 128 registers: three CTAs per SM), and the
// cross-lane reduction runs over 16 lanes.  ~2x fewer instructions per row than k_bsr_spmm.
template <int N, int R, bool MASKED, bool DOT>
__global__ void __launch_bounds__(kSpmvThreads, 3)
k_bsr_spmm_split(int64_t nb, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                 const double *__restrict__ vals, const double *__restrict__ X, double *__restrict__ Y,
                 const uint8_t *__restrict__ fixedMask, double *partials, unsigned *ticket, double *dotOut,
                 const int *status) {
    static_assert(R % 2 == 0, "split SpMM needs an even number of systems");
    constexpr int NN = N * N;
    constexpr int RH = R / 2;
    constexpr int LPR = 16, U = 3, CH = U * LPR;
    static_assert(CH % N == 0, "a chunk must cover whole blocks");
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int half = lane >> 4, sl = lane & 15;
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
    const int r0 = half * RH;                         // first system of this half-warp
    double dot[RH];
#pragma unroll
    for (int r = 0; r < RH; ++r) dot[r] = 0.0;
    int64_t row = warpGlobal;
    int64_t nb0 = 0, nb1 = 0;
    if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
    for (; row < nb; ) {
        const int64_t b0 = nb0;
        const int L = (int)(nb1 - nb0) * N;
        const int64_t thisRow = row;
        row += rowStride;
        nb0 = nb1 = 0;
        if (row < nb) { nb0 = rowptr[row]; nb1 = rowptr[row + 1]; }
        const double *v = vals + b0 * NN + sl;
        const int32_t *ci = colidx + b0;
        double acc[N][RH];
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int r = 0; r < RH; ++r) acc[k][r] = 0.0;
        for (int base = 0; base < L; base += CH, v += CH, ci += CH / N) {
            const int rem = L - base - sl;
            int col[U];
            double a[U][N];
            ChunkLoader<N, LPR>::run(ci + jj[0], ci + jj[1], ci + jj[2], v, v + L, v + 2 * L, rem, polStream, col, a);
            __syncwarp();     // scheduling fence: all streaming loads leave before the dependent gathers
            double xr[U][RH];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double *xp = X + ((int64_t)col[u] * N + cc[u]) * R + r0;
#pragma unroll
                for (int q = 0; q < RH; ++q)
                    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(xr[u][q]) : "l"(xp + q), "l"(polKeep) : "memory");
            }
            __syncwarp();     // ... and all gathers before the first FMA
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < N; ++k)
#pragma unroll
                    for (int r = 0; r < RH; ++r) acc[k][r] = fma(a[u][k], xr[u][r], acc[k][r]);
        }
        const bool fixedHere = MASKED && comp >= 0 && fixedMask[thisRow * N + (comp >= 0 ? comp : 0)];
#pragma unroll
        for (int r = 0; r < RH; ++r) {
            double t[N];
#pragma unroll
            for (int k = 0; k < N; ++k) t[k] = acc[k][r];
            const double out0 = fold_reduce<N, LPR>(t, sl);
            if (comp >= 0) {
                const double out = fixedHere ? 0.0 : out0;
                const int64_t idx = (thisRow * N + comp) * R + r0 + r;
                Y[idx] = out;
                if (DOT) dot[r] += out * X[idx];
            }
        }
    }
    if (DOT) {
        // the two halves hold different systems: lay the R sums out as [system] before the block reduction
        double d2[R];
#pragma unroll
        for (int r = 0; r < R; ++r) d2[r] = (r / RH == half) ? dot[r % RH] : 0.0;
        block_reduce_store<R>(d2, partials);
        if (last_block(ticket)) {
            for (int r = 0; r < R; ++r) {
                const double s = final_sum(partials + (size_t)r * gridDim.x, gridDim.x);
                if (threadIdx.x == 0) dotOut[r] = s;
            }
        }
    }
}

// init: r = mask(b); z = Minv r; p = z; x = 0; sums of r.z and r.r over owned DoFs, per system
template <int N, int R>
__global__ void __launch_bounds__(kVecThreads)
k_pcg_init_multi(int64_t nb, const double *__restrict__ b, const uint8_t *__restrict__ fixedMask,
                 const uint8_t *__restrict__ owned, const double *__restrict__ Minv, double *__restrict__ x,
                 double *__restrict__ r, double *__restrict__ z, double *__restrict__ p, double *partials,
                 unsigned *ticket, double *dotOut /* [2R]: rz[R], rr[R] */) {
    constexpr int NN = N * N;
    double acc[2 * R];
#pragma unroll
    for (int q = 0; q < 2 * R; ++q) acc[q] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        double M[NN];
#pragma unroll
        for (int q = 0; q < NN; ++q) M[q] = Minv[i * NN + q];
        const double wgt = (owned && !owned[i]) ? 0.0 : 1.0;
        bool fx[N];
#pragma unroll
        for (int k = 0; k < N; ++k) fx[k] = fixedMask[i * N + k] != 0;
#pragma unroll
        for (int q = 0; q < R; ++q) {
            double rv[N], zv[N];
#pragma unroll
            for (int k = 0; k < N; ++k) rv[k] = fx[k] ? 0.0 : b[(i * N + k) * R + q];
#pragma unroll
            for (int k = 0; k < N; ++k) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m) s += M[k * N + m] * rv[m];
                zv[k] = s;
            }
#pragma unroll
            for (int k = 0; k < N; ++k) {
                const int64_t idx = (i * N + k) * R + q;
                x[idx] = 0.0; r[idx] = rv[k]; z[idx] = zv[k]; p[idx] = zv[k];
                acc[q] += wgt * rv[k] * zv[k];
                acc[R + q] += wgt * rv[k] * rv[k];
            }
        }
    }
    block_reduce_store<2 * R>(acc, partials);
    if (last_block(ticket)) {
        for (int q = 0; q < 2 * R; ++q) {
            const double s = final_sum(partials + (size_t)q * gridDim.x, gridDim.x);
            if (threadIdx.x == 0) dotOut[q] = s;
        }
    }
}

template <int R>
__global__ void k_pcg_init_finalize_multi(double *scal, int *status, double tol2) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int running = 0, nan = 0;
    for (int r = 0; r < R; ++r) {
        const double rr = scal[MS<R>::RR + r];
        scal[MS<R>::RZ + r] = scal[MS<R>::RZ_NEW + r];
        scal[MS<R>::BB + r] = rr;
        const int st = (rr == 0.0) ? 1 : ((rr != rr) ? 3 : 0);
        status[2 + r] = st;
        status[2 + R + r] = 0;
        running += st == 0;
        nan += st == 3;
    }
    scal[MS<R>::TOL2] = tol2;
    status[ST_ITERS] = 0;
    status[ST_STATE] = nan ? 3 : (running ? 0 : 1);
}

template <int N, int R>
__global__ void __launch_bounds__(kVecThreads)
k_dot_owned_multi(int64_t nb, const double *__restrict__ a, const double *__restrict__ b2,
                  const uint8_t *__restrict__ owned, double *partials, unsigned *ticket, double *dotOut, const int *status) {
    if (status[ST_STATE] != 0) return;
    double acc[R];
#pragma unroll
    for (int q = 0; q < R; ++q) acc[q] = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        if (!owned[i]) continue;
#pragma unroll
        for (int k = 0; k < N; ++k)
#pragma unroll
            for (int q = 0; q < R; ++q) acc[q] += a[(i * N + k) * R + q] * b2[(i * N + k) * R + q];
    }
    block_reduce_store<R>(acc, partials);
    if (last_block(ticket)) {
        for (int q = 0; q < R; ++q) {
            const double s = final_sum(partials + (size_t)q * gridDim.x, gridDim.x);
            if (threadIdx.x == 0) dotOut[q] = s;
        }
    }
}

// Vector kernels of the batch: one thread per (DoF, system).  With the interleaved layout consecutive
// threads then touch consecutive addresses (a thread per DoF would stride 144 B).  The block size is a
// multiple of R and of 32, and the grid stride a multiple of R, so a thread keeps its system r for the
// whole loop; the per-system sums are combined through shared memory in a fixed order.
constexpr int kMultiThreads = 192;      // multiple of 32 and of R = 3, 6

template <int R>
__device__ __forceinline__ void multi_reduce_store(double a0, double a1, double *partials /* [2R][grid] */) {
    __shared__ double sh[2][kMultiThreads];
    sh[0][threadIdx.x] = a0;
    sh[1][threadIdx.x] = a1;
    __syncthreads();
    if (threadIdx.x < 2 * R) {
        const int which = threadIdx.x / R, r = threadIdx.x % R;
        double s = 0.0;
        for (int t = r; t < kMultiThreads; t += R) s += sh[which][t];
        partials[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
}

// update: alpha_r = rz_r / pAp_r (0 for a finished system); x += alpha p; r -= alpha Ap; z = Minv r; sums
template <int N, int R>
__global__ void __launch_bounds__(kMultiThreads)
k_pcg_update_multi(int64_t nb, const double *__restrict__ Minv, const uint8_t *__restrict__ owned,
                   const double *__restrict__ p, const double *__restrict__ Ap, double *__restrict__ x,
                   double *__restrict__ r, double *__restrict__ z, double *partials, unsigned *ticket,
                   const double *__restrict__ scal, double *dotOut /* [2R] */, const int *status) {
    constexpr int NN = N * N;
    static_assert(kMultiThreads % R == 0, "block size must be a multiple of the batch size");
    if (status[ST_STATE] != 0) return;
    const int q = threadIdx.x % R;                    // this thread's system, constant over the loop
    const double alpha = status[2 + q] == 0 ? scal[MS<R>::RZ + q] / scal[MS<R>::PAP + q] : 0.0;
    double accRZ = 0.0, accRR = 0.0;
    const int64_t total = nb * R;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / R;                      // t % R == q
        double rv[N], zv[N];
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int64_t idx = (i * N + k) * R + q;
            x[idx] += alpha * p[idx];
            rv[k] = r[idx] - alpha * Ap[idx];
        }
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double s = 0.0;
#pragma unroll
            for (int m = 0; m < N; ++m) s += __ldg(Minv + i * NN + k * N + m) * rv[m];
            zv[k] = s;
        }
        const double wgt = (owned && !owned[i]) ? 0.0 : 1.0;
#pragma unroll
        for (int k = 0; k < N; ++k) {
            const int64_t idx = (i * N + k) * R + q;
            r[idx] = rv[k]; z[idx] = zv[k];
            accRZ += wgt * rv[k] * zv[k];
            accRR += wgt * rv[k] * rv[k];
        }
    }
    multi_reduce_store<R>(accRZ, accRR, partials);
    if (last_block(ticket)) {
        for (int k = 0; k < 2 * R; ++k) {
            const double s = final_sum(partials + (size_t)k * gridDim.x, gridDim.x);
            if (threadIdx.x == 0) dotOut[k] = s;
        }
    }
}

// direction: p_r = z_r + beta_r p_r for the running systems; the last CTA closes the iteration
template <int R>
__global__ void __launch_bounds__(kMultiThreads)
k_pcg_direction_multi(int64_t n /* nvar*R */, const double *__restrict__ z, double *__restrict__ p, double *scal, int *status,
                      unsigned *ticket) {
    if (status[ST_STATE] != 0) return;
    const int q = threadIdx.x % R;                    // index % R of every element this thread visits
    const bool run = status[2 + q] == 0;
    const double beta = run ? scal[MS<R>::RZ_NEW + q] / scal[MS<R>::RZ + q] : 0.0;
    if (run)
        for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
            p[i] = z[i] + beta * p[i];
    if (last_block(ticket)) {
        if (threadIdx.x == 0) {
            const int it = status[ST_ITERS] + 1;
            status[ST_ITERS] = it;
            int running = 0, bad = 0;
            for (int q = 0; q < R; ++q) {
                if (status[2 + q] != 0) continue;
                const double pAp = scal[MS<R>::PAP + q], rz = scal[MS<R>::RZ_NEW + q], rr = scal[MS<R>::RR + q];
                scal[MS<R>::RZ + q] = rz;
                int st = 0;
                if (!(pAp > 0.0)) st = 2;
                if (rr != rr || rz != rz || pAp != pAp) st = 3;
                if (st == 0 && rr <= scal[MS<R>::TOL2] * scal[MS<R>::BB + q]) st = 1;
                status[2 + q] = st;
                if (st != 0) status[2 + R + q] = it;
                running += st == 0;
                if (st >= 2 && bad < st) bad = st;
            }
            status[ST_STATE] = bad ? bad : (running ? 0 : 1);     // written last
        }
    }
}

// [R separate vectors] <-> interleaved
__global__ void k_interleave(int64_t n, int R, int r, const double *__restrict__ src, double *__restrict__ dst) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i * R + r] = src[i];
}
__global__ void k_deinterleave_add(int64_t n, int R, int r, const double *__restrict__ src, const double *__restrict__ add,
                                   double *__restrict__ dst) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i * R + r] + add[i];
}
__global__ void k_sub_broadcast(int64_t n, int R, const double *__restrict__ v, double *__restrict__ B) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;     // B[i*R + r] -= v[i]
    if (i >= n) return;
    const double s = v[i];
    for (int r = 0; r < R; ++r) B[i * R + r] -= s;
}

template <int N, int R, bool SPLIT>
struct SpmmKernel {
    template <bool DOT> static auto get() { return k_bsr_spmm<N, R, true, DOT>; }
};
template <int N, int R>
struct SpmmKernel<N, R, true> {
    template <bool DOT> static auto get() { return k_bsr_spmm_split<N, R, true, DOT>; }
};

template <int N, int R>
static void launch_spmm(mfem_b200_ctx *c, PcgWorkMulti &w, const double *X, double *Y, bool dot) {
    // option "spmm_kernel" = 2 selects the half-warp split kernel (even R); measured on cfg3, R = 6:
    // 4.3 ms per SpMM for both -- each is bound by the L1 tag stage of the x gathers (ncu: l1tex 71 %)
    constexpr bool canSplit = R % 2 == 0;
    typedef SpmmKernel<N, R, canSplit> Split;
    typedef SpmmKernel<N, R, false> Full;
    const bool split = canSplit && c->opt_spmm_kernel == 2;
    if (dot) {
        auto kern = split ? Split::template get<true>() : Full::template get<true>();
        kern<<<spmv_grid(c, 32, kern), kSpmvThreads, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals, X, Y, c->fixedMask,
                                                                    w.partials, w.ticket, w.scal.p + MS<R>::PAP, w.status);
    } else {
        auto kern = split ? Split::template get<false>() : Full::template get<false>();
        kern<<<spmv_grid(c, 32, kern), kSpmvThreads, 0, c->stream>>>(c->nDofs, c->rowptr, c->colidx, c->vals, X, Y, c->fixedMask,
                                                                    nullptr, nullptr, nullptr, w.status);
    }
    c->launches++;
}

static int multi_grid(mfem_b200_ctx *c, int64_t n) {
    const int64_t ctas = (n + kMultiThreads - 1) / kMultiThreads;
    return (int)std::max<int64_t>(1, std::min<int64_t>(ctas, std::min<int64_t>(kMaxPartials, (int64_t)sm_count(c) * 10)));
}

template <int N, int R>
static void enqueue_iteration_multi(mfem_b200_ctx *c, PcgWorkMulti &w) {
    const int64_t nb = c->nDofs;
    const int vgrid = vec_grid(c, nb);
    const bool multi = c->nRanks > 1;
    const uint8_t *owned = multi ? halo_owned(c) : nullptr;
    if (!multi) {
        launch_spmm<N, R>(c, w, w.p, w.Ap, true);
    } else {
        launch_spmm<N, R>(c, w, w.p, w.Ap, false);
        halo_exchange_add(c, w.Ap, N * R);
        k_dot_owned_multi<N, R><<<vgrid, kVecThreads, 0, c->stream>>>(nb, w.p, w.Ap, owned, w.partials, w.ticket, w.dotLoc.p,
                                                                    w.status);
        c->launches++;
        allreduce_sum(c, w.dotLoc.p, w.scal.p + MS<R>::PAP, R);
    }
    const int mgrid = multi_grid(c, nb * R);
    k_pcg_update_multi<N, R><<<mgrid, kMultiThreads, 0, c->stream>>>(nb, c->Minv, owned, w.p, w.Ap, w.x, w.r, w.z, w.partials,
                                                                 w.ticket + 1, w.scal,
                                                                 multi ? w.dotLoc.p + R : w.scal.p + MS<R>::RZ_NEW, w.status);
    if (multi) allreduce_sum(c, w.dotLoc.p + R, w.scal.p + MS<R>::RZ_NEW, 2 * R);
    k_pcg_direction_multi<R><<<multi_grid(c, (int64_t)w.n), kMultiThreads, 0, c->stream>>>((int64_t)w.n, w.z, w.p, w.scal,
                                                                                           w.status, w.ticket + 2);
    c->launches += 2;
}

// f_int / u_int: R separate vectors of length nvar, back to back, internal numbering
template <int N, int R>
static void pcg_impl_multi(mfem_b200_ctx *c, const double *f_int, double *u_int, double rtol, int maxIters,
                           mfem_b200_solve_info *info) {
    PcgWorkMulti &w = ensure_work_multi(c, R);
    PcgWork &w1 = c->work;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs, n = c->nvar();
    const bool multi = c->nRanks > 1;
    const uint8_t *owned = multi ? halo_owned(c) : nullptr;
    // b_r = f_r - K ufix (the same correction for every system)
    if (multi) spmv_exchanged<N>(c, c->fixedVals, w1.Ap, false);
    else launch_spmv<N>(c, c->fixedVals, w1.Ap, false, false);
    for (int r = 0; r < R; ++r) k_interleave<<<grid_for(n, 256), 256, 0, s>>>(n, R, r, f_int + (size_t)r * n, w.b);
    k_sub_broadcast<<<grid_for(n, 256), 256, 0, s>>>(n, R, w1.Ap, w.b);
    k_pcg_init_multi<N, R><<<vec_grid(c, nb), kVecThreads, 0, s>>>(nb, w.b, c->fixedMask, owned, c->Minv, w.x, w.r, w.z, w.p,
                                                                 w.partials, w.ticket + 1,
                                                                 multi ? w.dotLoc.p + R : w.scal.p + MS<R>::RZ_NEW);
    if (multi) allreduce_sum(c, w.dotLoc.p + R, w.scal.p + MS<R>::RZ_NEW, 2 * R);
    k_pcg_init_finalize_multi<R><<<1, 32, 0, s>>>(w.scal, w.status, rtol * rtol);
    c->launches += R + 3;
    MFEM_CUDA(cudaGetLastError());

    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0, s);
    const int kBatch = 25;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (c->opt_graph && !multi) {
        const int64_t launchesBefore = c->launches;
        MFEM_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal));
        for (int k = 0; k < kBatch; ++k) enqueue_iteration_multi<N, R>(c, w);
        MFEM_CUDA(cudaStreamEndCapture(s, &graph));
        MFEM_CUDA(cudaGraphInstantiate(&exec, graph, 0));
        c->launches = launchesBefore;
    }
    int hst[2 + 2 * R];
    for (int q = 0; q < 2 + 2 * R; ++q) hst[q] = 0;
    int done = 0;
    while (done < maxIters) {
        if (exec) {
            MFEM_CUDA(cudaGraphLaunch(exec, s));
            c->launches += 3 * kBatch;
        } else {
            for (int k = 0; k < kBatch; ++k) enqueue_iteration_multi<N, R>(c, w);
        }
        MFEM_CUDA(cudaMemcpyAsync(hst, w.status, sizeof(hst), cudaMemcpyDeviceToHost, s));
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

    double hs[MS<R>::COUNT];
    MFEM_CUDA(cudaMemcpyAsync(hs, w.scal, sizeof(hs), cudaMemcpyDeviceToHost, s));
    for (int r = 0; r < R; ++r)       // u_r = x_r + ufix
        k_deinterleave_add<<<grid_for(n, 256), 256, 0, s>>>(n, R, r, w.x, c->fixedVals, u_int + (size_t)r * n);
    c->launches += R;
    MFEM_CUDA(cudaStreamSynchronize(s));
    bool allConverged = true;
    for (int r = 0; r < R; ++r) {
        const int st = hst[2 + r];
        allConverged = allConverged && st == 1;
        if (info) {
            info[r].iterations = st != 0 ? hst[2 + R + r] : hst[ST_ITERS];
            info[r].converged = st == 1;
            info[r].rel_residual = hs[MS<R>::BB + r] > 0 ? std::sqrt(hs[MS<R>::RR + r] / hs[MS<R>::BB + r]) : 0.0;
            info[r].seconds = ms * 1e-3 / R;         // the batch time, split evenly
            info[r].spmv_seconds = 0.0;
        }
    }
    if (hst[ST_STATE] == 2)
        throw CudaError(MFEM_B200_ERR_NOT_SPD, "PCG breakdown: p'Ap <= 0 (matrix is not positive definite)");
    if (hst[ST_STATE] == 3) throw CudaError(MFEM_B200_ERR_NAN, "PCG: NaN encountered");
    if (!allConverged)
        throw CudaError(MFEM_B200_ERR_NO_CONVERGE, "PCG: no convergence of the batched solve in " +
                                                       std::to_string(hst[ST_ITERS]) + " iterations");
}

// Batched solve of nrhs systems; true if handled (nrhs == flatLen(N), the cell-problem case), false if
// the caller should fall back to one solve per right-hand side.
bool pcg_solve_multi(mfem_b200_ctx *c, int nrhs, const double *f_int, double *u_int, double rtol, int maxIters,
                     mfem_b200_solve_info *info) {
    if (!c->opt_batch_rhs || nrhs != flat_len(c->N)) return false;
    MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "solve: matrix not assembled");
    ensure_work(c);
    build_preconditioner(c);
    if (c->N == 3) pcg_impl_multi<3, 6>(c, f_int, u_int, rtol, maxIters, info);
    else pcg_impl_multi<2, 3>(c, f_int, u_int, rtol, maxIters, info);
    return true;
}

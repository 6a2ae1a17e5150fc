// Two-level preconditioner (option coarse_aggregates = S > 0, default off):  M^-1 = B^-1 + Z E^-1 Z^T,
// B = the block-Jacobi of solver.cu, Z = rigid-body modes of S aggregates (6 per aggregate in 3D, 3 in 2D: the
// tentative prolongator of smoothed-aggregation AMG), E = Z^T K_ff Z.  Aggregates are contiguous runs of the internal
// (Morton-ordered) DoF numbering, agg(i) = floor(i S / nDofs), so restriction and prolongation are segmented
// reductions / broadcasts over index ranges with no indirection.
//
// Why: the cantilever workloads are bending-dominated and block-Jacobi PCG needs thousands of iterations (cfg5: 6074);
// tools/proto_two_level.py (CPU, numpy) measures 4.8x fewer iterations with 180 nodes per aggregate and 7x with 45,
// independent of the mesh size at fixed aggregate size.  Cost per iteration: one pass over r and z (120 B per DoF)
// plus a dense (6S)^2 GEMV, a few percent of the SpMV.
//
// STATUS: written after the round-1 GPU budget was spent -- compiled, never run on a GPU; opt-in only.
//   * single right-hand side (the batched PCG ignores it); one GPU here, the multi-GPU variant is further down;
//   * E is dense and inverted explicitly with cuSOLVER (potrf + potri), loaded with dlopen so that the library has no
//     link-time dependency on it: a setup step of O((6S)^3), not on the per-iteration path;
//   * restriction uses FP64 atomics: the solve is no longer bit-reproducible run to run with this option on.
// Included by solver.cu inside namespace mfem (which includes <cusolverDn.h> and <dlfcn.h> for it).

struct CoarseSpace {
    int S = 0, M = 0;                 // aggregates, modes per aggregate
    int64_t nc = 0;                   // M * S
    DevBuf<double> Y;                 // [nDofs*N] DoF position relative to its aggregate's centroid (internal order)
    DevBuf<double> Einv;              // [nc*nc] row-major, symmetric
    DevBuf<double> cvec, yvec;        // [nc]
    bool indexed = false;             // multi-GPU: aggregate ids from the array below (owner-based aggregates)
    DevBuf<int32_t> agg;              // [nDofs] global aggregate of every local DoF (indexed only)
};

__host__ __device__ __forceinline__ int64_t coarse_agg(int64_t i, int64_t S, int64_t nb) { return (i * S) / nb; }

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

// lowest node of every DoF (periodic DoFs have several nodes; any would do, the lowest is deterministic)
__global__ void k_coarse_first_node(int64_t nNodes, const int32_t *__restrict__ nodeDof, int32_t *firstNode) {
    const int64_t n = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (n < nNodes) atomicMin(&firstNode[nodeDof[n]], (int32_t)n);
}
template <int N>
__global__ void k_coarse_positions(int64_t nb, int64_t S, int64_t nNodes, const int32_t *__restrict__ firstNode,
                                   const double *__restrict__ nodes, double *__restrict__ Y, double *cen /* [S*(N+1)] sums and counts */) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t a = coarse_agg(i, S, nb);
    const int64_t node = firstNode[i];
    if (node < 0 || node >= nNodes) return;              // a DoF without a node: stays at the origin
    for (int k = 0; k < N; ++k) {
        const double x = nodes[node * N + k];
        Y[i * N + k] = x;
        atomicAdd(&cen[a * (N + 1) + k], x);
    }
    atomicAdd(&cen[a * (N + 1) + N], 1.0);
}
template <int N>
__global__ void k_coarse_center(int64_t nb, int64_t S, const double *__restrict__ cen, double *__restrict__ Y) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t a = coarse_agg(i, S, nb);
    const double cnt = cen[a * (N + 1) + N];
    if (cnt > 0.0)
        for (int k = 0; k < N; ++k) Y[i * N + k] -= cen[a * (N + 1) + k] / cnt;
}

// E += Z^T K_ff Z: one warp per block row, one lane per block; lanes whose columns fall into the same aggregate
// (contiguous: columns are sorted and agg is monotone) are combined by a segmented shuffle reduction, the head lane
// of each segment adds the M x M result to E.
template <int N>
__global__ void __launch_bounds__(256)
k_coarse_matrix(int64_t nb, int64_t S, const int64_t *__restrict__ rowptr, const int32_t *__restrict__ colidx,
                const double *__restrict__ vals, const uint8_t *__restrict__ fixedMask, const double *__restrict__ Y,
                double *E) {
    constexpr int M = N == 3 ? 6 : 3;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nc = (int64_t)M * S;
    for (int64_t row = warp; row < nb; row += nWarps) {
        const int64_t b0 = rowptr[row], n = rowptr[row + 1] - b0;
        const int64_t ai = coarse_agg(row, S, nb);
        double yi[N];
        bool fi[N];
        for (int k = 0; k < N; ++k) { yi[k] = Y[row * N + k]; fi[k] = fixedMask[row * N + k] != 0; }
        for (int64_t j0 = 0; j0 < n; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool active = j < n;
            double C[M][M];
            int64_t aj = -1 - lane;                      // inactive lanes: unique keys, never merged
            for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) C[a][b] = 0.0;
            if (active) {
                const int64_t col = colidx[b0 + j];
                aj = coarse_agg(col, S, nb);
                double yj[N];
                for (int k = 0; k < N; ++k) yj[k] = Y[col * N + k];
                // T[r][b] = sum_c K[r][c] R_j[c][b] with fixed rows / columns masked out
                double T[N][M];
                for (int r = 0; r < N; ++r) {
                    double krow[N];
                    for (int cc = 0; cc < N; ++cc)
                        krow[cc] = (fi[r] || fixedMask[col * N + cc]) ? 0.0 : vals[val_index<N>(b0, n, j, r, cc)];
                    coarse_Rt<N>(yj, krow, T[r]);      // (K_row R_j) = R_j^T K_row^T
                }
                for (int b = 0; b < M; ++b) {
                    double tcol[N], q[M];
                    for (int r = 0; r < N; ++r) tcol[r] = T[r][b];
                    coarse_Rt<N>(yi, tcol, q);
                    for (int a = 0; a < M; ++a) C[a][b] = q[a];
                }
            }
            // segmented reduction over contiguous lanes with equal aj
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int64_t ajo = __shfl_down_sync(0xffffffffu, aj, o);
                const bool take = (lane + o < 32) && (ajo == aj);
                for (int a = 0; a < M; ++a)
                    for (int b = 0; b < M; ++b) {
                        const double other = __shfl_down_sync(0xffffffffu, C[a][b], o);
                        if (take) C[a][b] += other;
                    }
            }
            const int64_t ajPrev = __shfl_up_sync(0xffffffffu, aj, 1);
            const bool head = active && (lane == 0 || ajPrev != aj);
            // the factorisation reads only the row-major UPPER triangle of E (= column-major lower): block columns
            // left of the diagonal block are never looked at, so half of the reductions can be skipped
            if (head && aj >= ai)
                for (int a = 0; a < M; ++a)
                    for (int b = 0; b < M; ++b)
                        if (C[a][b] != 0.0) atomicAdd(&E[(ai * M + a) * nc + aj * M + b], C[a][b]);
        }
    }
}

// dead modes (all-zero rows: aggregates without free variables, rotations without coordinates) get a unit diagonal;
// a relative diagonal shift keeps E positive definite when an aggregate has too few free nodes for six independent
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

// c += Z^T r (c zeroed by k_coarse_prolong of the previous application)
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_restrict(int64_t nb, int64_t S, const double *__restrict__ r, const uint8_t *__restrict__ fixedMask,
                  const double *__restrict__ Y, double *cvec, const int *status) {
    constexpr int M = N == 3 ? 6 : 3;
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    // contiguous slabs per warp iteration: the 32 lanes hold consecutive DoFs
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 32; base < nb; base += nWarps * 32) {
        const int64_t i = base + lane;
        double q[M];
        int64_t a = -1 - lane;
        for (int m = 0; m < M; ++m) q[m] = 0.0;
        if (i < nb) {
            a = coarse_agg(i, S, nb);
            double v[N], y[N];
            for (int k = 0; k < N; ++k) { v[k] = fixedMask[i * N + k] ? 0.0 : r[i * N + k]; y[k] = Y[i * N + k]; }
            coarse_Rt<N>(y, v, q);
        }
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int64_t ao = __shfl_down_sync(0xffffffffu, a, o);
            const bool take = (lane + o < 32) && (ao == a);
            for (int m = 0; m < M; ++m) {
                const double other = __shfl_down_sync(0xffffffffu, q[m], o);
                if (take) q[m] += other;
            }
        }
        const int64_t aPrev = __shfl_up_sync(0xffffffffu, a, 1);
        if (i < nb && (lane == 0 || aPrev != a))
            for (int m = 0; m < M; ++m) atomicAdd(&cvec[a * M + m], q[m]);
    }
}

// y = Einv c (one warp per row) and rz += c.y
__global__ void __launch_bounds__(256)
k_coarse_gemv(int64_t nc, const double *__restrict__ Einv, const double *__restrict__ cvec, double *__restrict__ yvec,
              double *rzOut, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    double dot = 0.0;
    for (int64_t row = warp; row < nc; row += nWarps) {
        const double *e = Einv + row * nc;
        double s = 0.0;
        for (int64_t k = lane; k < nc; k += 32) s = fma(e[k], cvec[k], s);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) { yvec[row] = s; dot += s * cvec[row]; }
    }
    if (rzOut && lane == 0 && dot != 0.0) atomicAdd(rzOut, dot);
}

// z += mask(Z y) [and p = z for the initial direction]; clears c for the next application
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_prolong(int64_t nb, int64_t S, const double *__restrict__ yvec, const uint8_t *__restrict__ fixedMask,
                 const double *__restrict__ Y, double *__restrict__ z, double *p /* may be null */, double *cvec, int64_t nc,
                 const int *status) {
    constexpr int M = N == 3 ? 6 : 3;
    if (status && status[ST_STATE] != 0) return;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t a = coarse_agg(i, S, nb);
        double c[M], y[N], v[N];
        for (int m = 0; m < M; ++m) c[m] = yvec[a * M + m];
        for (int k = 0; k < N; ++k) y[k] = Y[i * N + k];
        coarse_R<N>(y, c, v);
        for (int k = 0; k < N; ++k) {
            const double zk = z[i * N + k] + (fixedMask[i * N + k] ? 0.0 : v[k]);
            z[i * N + k] = zk;
            if (p) p[i * N + k] = zk;
        }
        if (i < nc) cvec[i] = 0.0;
    }
    // nc > nb cannot happen (S <= nb / 8), so the loop above clears all of c
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
    return api;
}

static void free_coarse(mfem_b200_ctx *c) {
    delete c->coarse;
    c->coarse = nullptr;
}

// E (assembled: the row-major upper block triangle, which is all potrf reads) -> regularised -> explicit inverse in place
// (potrf + potri + mirror)
static void invert_coarse_matrix(mfem_b200_ctx *c, CoarseSpace &cs) {
    cudaStream_t s = c->stream;
    const int64_t nc = cs.nc;
    k_coarse_regularize<<<grid_for(nc, 256), 256, 0, s>>>(nc, cs.Einv, 1e-8);
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
    // explicit inverse: E = L L^T, E^-1 from the factor
    CusolverApi &api = cusolver_api();
    cusolverDnHandle_t h = nullptr;
    MFEM_REQUIRE(api.create(&h) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "cusolverDnCreate failed");
    struct Guard { CusolverApi &a; cusolverDnHandle_t h; ~Guard() { if (h) a.destroy(h); } } guard{api, h};
    MFEM_REQUIRE(api.setStream(h, s) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "cusolverDnSetStream failed");
    int lw1 = 0, lw2 = 0;
    MFEM_REQUIRE(api.potrfBuf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, cs.Einv, (int)nc, &lw1) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potrf_bufferSize failed");
    MFEM_REQUIRE(api.potriBuf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, cs.Einv, (int)nc, &lw2) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potri_bufferSize failed");
    DevBuf<double> wbuf((size_t)std::max(lw1, lw2) + 1);
    DevBuf<int> info(1);
    int hinfo = 0;
    MFEM_REQUIRE(api.potrf(h, CUBLAS_FILL_MODE_LOWER, (int)nc, cs.Einv, (int)nc, wbuf, lw1, info) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potrf failed");
    MFEM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_REQUIRE(hinfo == 0, MFEM_B200_ERR_NOT_SPD, "coarse matrix is not positive definite (leading minor " + std::to_string(hinfo) +
                                                         "): fewer aggregates, or a singular system");
    MFEM_REQUIRE(api.potri(h, CUBLAS_FILL_MODE_LOWER, (int)nc, cs.Einv, (int)nc, wbuf, lw2, info) == CUSOLVER_STATUS_SUCCESS, MFEM_B200_ERR_CUDA, "potri failed");
    MFEM_CUDA(cudaMemcpyAsync(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost, s));
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_REQUIRE(hinfo == 0, MFEM_B200_ERR_NOT_SPD, "coarse matrix inversion failed (" + std::to_string(hinfo) + ")");
    k_coarse_symmetrize<<<dim3((unsigned)grid_for(nc, 256), (unsigned)nc), 256, 0, s>>>(nc, cs.Einv);
    c->launches++;
    MFEM_CUDA(cudaStreamSynchronize(s));
    MFEM_CUDA(cudaGetLastError());
}

template <int N>
static void build_coarse_impl(mfem_b200_ctx *c) {
    constexpr int M = N == 3 ? 6 : 3;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs;
    int64_t S = std::min<int64_t>(c->opt_coarse, std::max<int64_t>(1, nb / 8));
    S = std::min<int64_t>(S, 32768 / M);
    free_coarse(c);
    c->coarse = new CoarseSpace();
    CoarseSpace &cs = *c->coarse;
    cs.S = (int)S; cs.M = M; cs.nc = M * S;
    const int64_t nc = cs.nc;
    cs.Y.alloc((size_t)nb * N);
    cs.Einv.alloc((size_t)nc * nc);
    cs.cvec.alloc((size_t)nc); cs.yvec.alloc((size_t)nc);
    MFEM_CUDA(cudaMemsetAsync(cs.Y, 0, cs.Y.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(cs.Einv, 0, cs.Einv.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(cs.cvec, 0, cs.cvec.bytes(), s));
    if (!c->externalMatrix && c->nNodes > 0) {
        DevBuf<int32_t> firstNode((size_t)nb);
        DevBuf<double> cen((size_t)S * (N + 1));
        MFEM_CUDA(cudaMemsetAsync(firstNode, 0x7f, firstNode.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(cen, 0, cen.bytes(), s));
        k_coarse_first_node<<<grid_for(c->nNodes, 256), 256, 0, s>>>(c->nNodes, c->nodeDof, firstNode);
        k_coarse_positions<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, S, c->nNodes, firstNode, c->nodes, cs.Y, cen);
        k_coarse_center<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, S, cen, cs.Y);
        c->launches += 3;
        MFEM_CUDA(cudaStreamSynchronize(s));          // firstNode / cen go out of scope
    }
    const int grid = (int)std::min<int64_t>((nb + 7) / 8, (int64_t)sm_count(c) * 8);
    k_coarse_matrix<N><<<grid, 256, 0, s>>>(nb, S, c->rowptr, c->colidx, c->vals, c->fixedMask, cs.Y, cs.Einv);
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
    invert_coarse_matrix(c, cs);
}


// ---------------------------------------------------------------------------------------------------------------
// Multi-GPU variant (nRanks > 1).  Every rank aggregates the DoFs it OWNS (S / nRanks aggregates per rank, contiguous
// runs of its owned DoFs in internal order; global aggregate id = rank * S_r + local id).  A DoF shared with other
// ranks carries its OWNER's aggregate id and its owner's centred position on every sharer (one interface
// sum-exchange at setup in which only the owner contributes), so the rows of Z agree on all sharers and
//     E = Z^T K Z = sum over ranks of Z_loc^T K_loc Z_loc      (K_loc holds partial sums on interface rows)
// is one all-reduce of the (M S)^2 partial matrices; every rank then inverts E redundantly.  Per application:
// restriction over owned DoFs, all-reduce of the M S coarse residuals (98 kB at S = 2048), the dense GEMV replicated
// on every rank, prolongation on all local DoFs (consistent on shared DoFs, like the block-Jacobi part).  The scalar
// c.y is added to the already all-reduced r.z by a single-CTA fixed-order reduction so that every rank holds the same
// bits.  Aggregate ids come from an array here (the single-GPU kernels above use the closed form), and runs of equal
// ids inside a warp are found with head flags because foreign aggregates may interleave.
// STATUS: written without GPU access -- compiled, never run; opt-in like the single-GPU variant.

// lanes [lane+1, lane+o] hold no run head  <=>  lane+o belongs to lane's run
__device__ __forceinline__ bool coarse_same_run(unsigned heads, int lane, int o) {
    return (lane + o < 32) && (((heads >> (lane + 1)) & ((1u << o) - 1u)) == 0u);
}

// owned DoFs: T[i] = (aggregate id + 1, position); centroid sums per LOCAL aggregate.  Non-owned rows stay zero.
template <int N>
__global__ void k_coarse_positions_idx(int64_t nb, int64_t aggBase, const int32_t *__restrict__ agg, int64_t nNodes,
                                       const int32_t *__restrict__ firstNode, const double *__restrict__ nodes,
                                       double *__restrict__ T, double *cen) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t a = agg[i];
    if (a < 0) return;
    T[i * (N + 1)] = (double)(a + 1);
    const int64_t node = firstNode[i];
    if (node < 0 || node >= nNodes) return;
    for (int k = 0; k < N; ++k) {
        const double x = nodes[node * N + k];
        T[i * (N + 1) + 1 + k] = x;
        atomicAdd(&cen[(a - aggBase) * (N + 1) + k], x);
    }
    atomicAdd(&cen[(a - aggBase) * (N + 1) + N], 1.0);
}
template <int N>
__global__ void k_coarse_center_idx(int64_t nb, int64_t aggBase, const int32_t *__restrict__ agg, int64_t nNodes,
                                    const int32_t *__restrict__ firstNode, const double *__restrict__ cen, double *__restrict__ T) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    const int64_t a = agg[i];
    if (a < 0) return;
    const int64_t node = firstNode[i];
    if (node < 0 || node >= nNodes) return;
    const double cnt = cen[(a - aggBase) * (N + 1) + N];
    if (cnt > 0.0)
        for (int k = 0; k < N; ++k) T[i * (N + 1) + 1 + k] -= cen[(a - aggBase) * (N + 1) + k] / cnt;
}
// after the owner-contributes sum-exchange: split T into agg / Y on every local DoF
template <int N>
__global__ void k_coarse_unpack(int64_t nb, int64_t S, const double *__restrict__ T, int32_t *__restrict__ agg,
                                double *__restrict__ Y, int *bad) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    int64_t a = llrint(T[i * (N + 1)]) - 1;
    if (a < 0 || a >= S) { atomicAdd(bad, 1); a = 0; }
    agg[i] = (int32_t)a;
    for (int k = 0; k < N; ++k) Y[i * N + k] = T[i * (N + 1) + 1 + k];
}

// E_loc += Z_loc^T K_loc Z_loc with aggregate ids from the array
template <int N>
__global__ void __launch_bounds__(256)
k_coarse_matrix_idx(int64_t nb, int64_t S, const int32_t *__restrict__ agg, const int64_t *__restrict__ rowptr,
                    const int32_t *__restrict__ colidx, const double *__restrict__ vals, const uint8_t *__restrict__ fixedMask,
                    const double *__restrict__ Y, double *E) {
    constexpr int M = N == 3 ? 6 : 3;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int64_t nc = (int64_t)M * S;
    for (int64_t row = warp; row < nb; row += nWarps) {
        const int64_t b0 = rowptr[row], n = rowptr[row + 1] - b0;
        const int64_t ai = agg[row];
        double yi[N];
        bool fi[N];
        for (int k = 0; k < N; ++k) { yi[k] = Y[row * N + k]; fi[k] = fixedMask[row * N + k] != 0; }
        for (int64_t j0 = 0; j0 < n; j0 += 32) {
            const int64_t j = j0 + lane;
            const bool active = j < n;
            double C[M][M];
            int64_t aj = -1 - lane;
            for (int a = 0; a < M; ++a) for (int b = 0; b < M; ++b) C[a][b] = 0.0;
            if (active) {
                const int64_t col = colidx[b0 + j];
                aj = agg[col];
                double yj[N];
                for (int k = 0; k < N; ++k) yj[k] = Y[col * N + k];
                double T[N][M];
                for (int r = 0; r < N; ++r) {
                    double krow[N];
                    for (int cc = 0; cc < N; ++cc)
                        krow[cc] = (fi[r] || fixedMask[col * N + cc]) ? 0.0 : vals[val_index<N>(b0, n, j, r, cc)];
                    coarse_Rt<N>(yj, krow, T[r]);
                }
                for (int b = 0; b < M; ++b) {
                    double tcol[N], q[M];
                    for (int r = 0; r < N; ++r) tcol[r] = T[r][b];
                    coarse_Rt<N>(yi, tcol, q);
                    for (int a = 0; a < M; ++a) C[a][b] = q[a];
                }
            }
            const int64_t ajPrev = __shfl_up_sync(0xffffffffu, aj, 1);
            const bool head = (lane == 0) || (ajPrev != aj);
            const unsigned heads = __ballot_sync(0xffffffffu, head);
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const bool take = coarse_same_run(heads, lane, o);
                for (int a = 0; a < M; ++a)
                    for (int b = 0; b < M; ++b) {
                        const double other = __shfl_down_sync(0xffffffffu, C[a][b], o);
                        if (take) C[a][b] += other;
                    }
            }
            if (active && head && aj >= ai)          // upper block triangle only (see k_coarse_matrix)
                for (int a = 0; a < M; ++a)
                    for (int b = 0; b < M; ++b)
                        if (C[a][b] != 0.0) atomicAdd(&E[(ai * M + a) * nc + aj * M + b], C[a][b]);
        }
    }
}

// c += Z^T r over the OWNED free variables
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_restrict_idx(int64_t nb, const int32_t *__restrict__ agg, const uint8_t *__restrict__ owned,
                      const double *__restrict__ r, const uint8_t *__restrict__ fixedMask, const double *__restrict__ Y,
                      double *cvec, const int *status) {
    constexpr int M = N == 3 ? 6 : 3;
    if (status && status[ST_STATE] != 0) return;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nWarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t base = warp * 32; base < nb; base += nWarps * 32) {
        const int64_t i = base + lane;
        double q[M];
        int64_t a = -1 - lane;
        for (int m = 0; m < M; ++m) q[m] = 0.0;
        if (i < nb) {
            a = agg[i];
            const bool mine = !owned || owned[i] != 0;
            double v[N], y[N];
            for (int k = 0; k < N; ++k) { v[k] = (!mine || fixedMask[i * N + k]) ? 0.0 : r[i * N + k]; y[k] = Y[i * N + k]; }
            coarse_Rt<N>(y, v, q);
        }
        const int64_t aPrev = __shfl_up_sync(0xffffffffu, a, 1);
        const bool head = (lane == 0) || (aPrev != a);
        const unsigned heads = __ballot_sync(0xffffffffu, head);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const bool take = coarse_same_run(heads, lane, o);
            for (int m = 0; m < M; ++m) {
                const double other = __shfl_down_sync(0xffffffffu, q[m], o);
                if (take) q[m] += other;
            }
        }
        if (i < nb && head)
            for (int m = 0; m < M; ++m)
                if (q[m] != 0.0) atomicAdd(&cvec[a * M + m], q[m]);
    }
}

// rz += c.y in a fixed order (one CTA): bit-identical on every rank, which holds the same c and y
__global__ void __launch_bounds__(256) k_coarse_cy(int64_t nc, const double *__restrict__ cvec, const double *__restrict__ yvec,
                                                   double *rzOut, const int *status) {
    if (status && status[ST_STATE] != 0) return;
    __shared__ double sh[256];
    double s = 0.0;
    for (int64_t k = threadIdx.x; k < nc; k += 256) s = fma(cvec[k], yvec[k], s);
    sh[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *rzOut += sh[0];
}

// z += mask(Z y) on every local DoF [and p = z]; clears c for the next application
template <int N>
__global__ void __launch_bounds__(kVecThreads)
k_coarse_prolong_idx(int64_t nb, const int32_t *__restrict__ agg, const double *__restrict__ yvec,
                     const uint8_t *__restrict__ fixedMask, const double *__restrict__ Y, double *__restrict__ z, double *p,
                     double *cvec, int64_t nc, const int *status) {
    constexpr int M = N == 3 ? 6 : 3;
    if (status && status[ST_STATE] != 0) return;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x, t0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    for (int64_t i = t0; i < nb; i += stride) {
        const int64_t a = agg[i];
        double c[M], y[N], v[N];
        for (int m = 0; m < M; ++m) c[m] = yvec[a * M + m];
        for (int k = 0; k < N; ++k) y[k] = Y[i * N + k];
        coarse_R<N>(y, c, v);
        for (int k = 0; k < N; ++k) {
            const double zk = z[i * N + k] + (fixedMask[i * N + k] ? 0.0 : v[k]);
            z[i * N + k] = zk;
            if (p) p[i * N + k] = zk;
        }
    }
    for (int64_t k = t0; k < nc; k += stride) cvec[k] = 0.0;      // nobody reads c in this kernel
}

// ---- box aggregates.  tools/proto_two_level.py-style experiments on the CPU show that the SHAPE of the aggregates
// matters as much as their number: contiguous runs of the Morton order are ragged (a run cuts across the cells of
// the curve), and compact boxes of the same count need 1.6-1.7x fewer iterations (quadratic cantilever 20x4x4:
// 32 aggregates 293 -> 169, 64: 234 -> 140, 256: 145 -> 90).  So the indexed path bins the (owned) DoFs into a
// regular grid of near-cubic boxes over their bounding box; empty boxes are dead modes (unit diagonal in E).

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
// keys[0..N) = min, keys[N..2N) = max over the owned DoFs that have a node
template <int N>
__global__ void k_coarse_bbox(int64_t nb, const uint8_t *__restrict__ owned, int64_t nNodes, const int32_t *__restrict__ firstNode,
                              const double *__restrict__ nodes, unsigned long long *keys) {
    double lo[N], hi[N];
    for (int k = 0; k < N; ++k) { lo[k] = 1e300; hi[k] = -1e300; }
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < nb; i += (int64_t)gridDim.x * blockDim.x) {
        if (owned && !owned[i]) continue;
        const int64_t node = firstNode[i];
        if (node < 0 || node >= nNodes) continue;
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
}
struct CoarseBoxes { double lo[3], scale[3]; int b[3]; };     // box of x along k: min(b-1, floor((x - lo) * scale))
// agg[i] = aggBase + box of the DoF's first node (owned DoFs; -1 for the others)
template <int N>
__global__ void k_coarse_box_agg(int64_t nb, const uint8_t *__restrict__ owned, int64_t nNodes, const int32_t *__restrict__ firstNode,
                                 const double *__restrict__ nodes, const CoarseBoxes bx, int64_t aggBase, int32_t *__restrict__ agg) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb) return;
    if (owned && !owned[i]) { agg[i] = -1; return; }
    const int64_t node = firstNode[i];
    int64_t id = 0;
    if (node >= 0 && node < nNodes)
        for (int k = 0; k < N; ++k) {
            int q = (int)floor((nodes[node * N + k] - bx.lo[k]) * bx.scale[k]);
            q = max(0, min(bx.b[k] - 1, q));
            id = id * bx.b[k] + q;
        }
    agg[i] = (int32_t)(aggBase + id);
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

// Indexed coarse space (1..N ranks): box aggregates over the DoFs this rank owns, owner's aggregate id / centred
// position on every sharer, all-reduced partial coarse matrix.  On one rank the exchange and the all-reduces are no-ops.
template <int N>
static void build_coarse_indexed_impl(mfem_b200_ctx *c) {
    constexpr int M = N == 3 ? 6 : 3;
    cudaStream_t s = c->stream;
    const int64_t nb = c->nDofs;
    const bool multi = c->nRanks > 1;
    const uint8_t *ownedDev = multi ? halo_owned(c) : nullptr;
    // aggregate ids per rank: a fixed stride (the option and nRanks are the same everywhere), so global ids need no
    // negotiation; a rank that uses fewer boxes than its stride leaves dead modes behind
    int64_t Sr = std::max<int64_t>(1, std::min<int64_t>(c->opt_coarse, 32768 / M) / c->nRanks);
    if (!multi) Sr = std::min<int64_t>(Sr, std::max<int64_t>(1, nb / 8));
    const int64_t aggBase = Sr * c->rank;
    const bool havePositions = !c->externalMatrix && c->nNodes > 0;
    const int64_t nNodesEff = havePositions ? c->nNodes : 0;
    DevBuf<int32_t> firstNode((size_t)nb);
    MFEM_CUDA(cudaMemsetAsync(firstNode, 0x7f, firstNode.bytes(), s));
    CoarseBoxes bx;
    for (int k = 0; k < 3; ++k) { bx.lo[k] = 0.0; bx.scale[k] = 0.0; bx.b[k] = 1; }
    if (havePositions) {
        k_coarse_first_node<<<grid_for(c->nNodes, 256), 256, 0, s>>>(c->nNodes, c->nodeDof, firstNode);
        DevBuf<unsigned long long> keys(2 * N);
        MFEM_CUDA(cudaMemsetAsync(keys, 0xff, sizeof(unsigned long long) * N, s));           // min slots: +inf keys
        MFEM_CUDA(cudaMemsetAsync(keys.p + N, 0x00, sizeof(unsigned long long) * N, s));     // max slots: -inf keys
        k_coarse_bbox<N><<<(int)std::min<int64_t>(grid_for(nb, 256), (int64_t)sm_count(c) * 8), 256, 0, s>>>(nb, ownedDev, c->nNodes, firstNode,
                                                                                                  c->nodes, keys);
        c->launches += 2;
        unsigned long long hk[2 * N];
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
            for (int k = 0; k < N; ++k) bx.scale[k] = L[k] > 0.0 ? bx.b[k] / L[k] : 0.0;
            if (!multi) { Sr = 1; for (int k = 0; k < N; ++k) Sr *= bx.b[k]; }     // one rank: no stride to keep, no dead ids
        }
    }
    const int64_t S = Sr * c->nRanks;
    free_coarse(c);
    c->coarse = new CoarseSpace();
    CoarseSpace &cs = *c->coarse;
    cs.S = (int)S; cs.M = M; cs.nc = M * S; cs.indexed = true;
    const int64_t nc = cs.nc;
    cs.agg.alloc((size_t)nb);
    cs.Y.alloc((size_t)nb * N);
    cs.Einv.alloc((size_t)nc * nc);
    cs.cvec.alloc((size_t)nc); cs.yvec.alloc((size_t)nc);
    MFEM_CUDA(cudaMemsetAsync(cs.Einv, 0, cs.Einv.bytes(), s));
    MFEM_CUDA(cudaMemsetAsync(cs.cvec, 0, cs.cvec.bytes(), s));
    k_coarse_box_agg<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, ownedDev, nNodesEff, firstNode, c->nodes, bx, aggBase, cs.agg);
    c->launches++;
    {
        DevBuf<double> T((size_t)nb * (N + 1));
        DevBuf<double> cen((size_t)Sr * (N + 1));
        DevBuf<int> bad(1);
        MFEM_CUDA(cudaMemsetAsync(T, 0, T.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(cen, 0, cen.bytes(), s));
        MFEM_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
        k_coarse_positions_idx<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, aggBase, cs.agg, nNodesEff, firstNode, c->nodes, T, cen);
        k_coarse_center_idx<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, aggBase, cs.agg, nNodesEff, firstNode, cen, T);
        c->launches += 2;
        halo_exchange_add(c, T, N + 1);                       // only the owner's rows are non-zero (no-op on one rank)
        k_coarse_unpack<N><<<grid_for(nb, 256), 256, 0, s>>>(nb, S, T, cs.agg, cs.Y, bad);
        c->launches++;
        int nbad = 0;
        MFEM_CUDA(cudaMemcpyAsync(&nbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
        MFEM_CUDA(cudaStreamSynchronize(s));                  // also: T / cen go out of scope
        MFEM_CUDA(cudaGetLastError());
        // no throw before the collectives below: every rank must reach them; a bad rank poisons E instead
        if (nbad) MFEM_CUDA(cudaMemsetAsync(cs.Einv, 0xff, sizeof(double), s));   // NaN in E[0] -> the inversion fails on all ranks
    }
    const int grid = (int)std::min<int64_t>((nb + 7) / 8, (int64_t)sm_count(c) * 8);
    k_coarse_matrix_idx<N><<<grid, 256, 0, s>>>(nb, S, cs.agg, c->rowptr, c->colidx, c->vals, c->fixedMask, cs.Y, cs.Einv);
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
    if (multi) {
        // one all-reduce of the partial coarse matrices (chunked: keep single calls moderate)
        const size_t total = (size_t)nc * nc, chunk = (size_t)1 << 27;
        for (size_t off = 0; off < total; off += chunk)
            allreduce_sum(c, cs.Einv.p + off, cs.Einv.p + off, (int)std::min(chunk, total - off));
    }
    invert_coarse_matrix(c, cs);
}

static void build_coarse(mfem_b200_ctx *c) {
    ScopedTimer timer(c, "Coarse Space");
    // default: box aggregates through the indexed kernels; coarse_shape = 1 keeps the first version (contiguous runs of
    // the internal numbering, closed-form aggregate ids, one rank only) for A/B runs
    if (c->nRanks > 1 || c->opt_coarse_shape == 0) {
        if (c->N == 3) build_coarse_indexed_impl<3>(c); else build_coarse_indexed_impl<2>(c);
        return;
    }
    if (c->N == 3) build_coarse_impl<3>(c); else build_coarse_impl<2>(c);
}

// z += Z Einv Z^T r, rz += (Z^T r).(Einv Z^T r); p = z when p != null (initial direction)
template <int N>
static void apply_coarse(mfem_b200_ctx *c, const double *r, double *z, double *p, double *rzSlot, const int *status) {
    CoarseSpace &cs = *c->coarse;
    cudaStream_t s = c->stream;
    const int vgrid = vec_grid(c, c->nDofs);
    if (cs.indexed) {
        const int ggridM = (int)std::min<int64_t>((cs.nc + 7) / 8, (int64_t)sm_count(c) * 8);
        k_coarse_restrict_idx<N><<<vgrid, kVecThreads, 0, s>>>(c->nDofs, cs.agg, c->nRanks > 1 ? halo_owned(c) : nullptr, r, c->fixedMask, cs.Y,
                                                               cs.cvec, status);
        if (c->nRanks > 1) allreduce_sum(c, cs.cvec, cs.cvec, (int)cs.nc);
        k_coarse_gemv<<<ggridM, 256, 0, s>>>(cs.nc, cs.Einv, cs.cvec, cs.yvec, nullptr, status);
        k_coarse_cy<<<1, 256, 0, s>>>(cs.nc, cs.cvec, cs.yvec, rzSlot, status);
        k_coarse_prolong_idx<N><<<vgrid, kVecThreads, 0, s>>>(c->nDofs, cs.agg, cs.yvec, c->fixedMask, cs.Y, z, p, cs.cvec, cs.nc, status);
        c->launches += 4;
        return;
    }
    k_coarse_restrict<N><<<vgrid, kVecThreads, 0, s>>>(c->nDofs, cs.S, r, c->fixedMask, cs.Y, cs.cvec, status);
    const int ggrid = (int)std::min<int64_t>((cs.nc + 7) / 8, (int64_t)sm_count(c) * 8);
    k_coarse_gemv<<<ggrid, 256, 0, s>>>(cs.nc, cs.Einv, cs.cvec, cs.yvec, rzSlot, status);
    k_coarse_prolong<N><<<vgrid, kVecThreads, 0, s>>>(c->nDofs, cs.S, cs.yvec, c->fixedMask, cs.Y, z, p, cs.cvec, cs.nc, status);
    c->launches += 3;
}


// K9: discrete shape derivatives of the elasticity operators under a per-vertex perturbation delta_p -- what an
// optimisation loop evaluates once per iteration next to its solves.
//
// Reference being replaced (all host loops there; the last one is the one TBB reduction the reference parallelises):
//   Simulator::applyDeltaStiffnessMatrix     LinearElasticity.hh:1301-1330
//   Simulator::deltaConstantStrainLoad       LinearElasticity.hh:1333-1348
//   Simulator::deltaAverageStrainField       LinearElasticity.hh:1365-1375
//   (callers: PeriodicHomogenization.hh:383-563)
//
// The straight-sided elements follow their vertices and nodal values are transported (Lagrangian derivative).  With
// the piecewise-linear velocity dp_h = sum_k delta_p_k lambda_k the geometric rules are
//     delta grad phi_i = -(grad dp_h)^T grad phi_i ,      delta vol = vol div dp_h
// (EmbeddedElement.hh:269-278, 338-372), so instead of the reference's per-element 30x30 delta-stiffness matrices each
// element contributes through the displacement gradient at the quadrature points:
//   (delta K u)_i = int [ div dp sigma(u) + C : delta eps(u) ] grad phi_i + sigma(u) delta grad phi_i ,
//   delta eps(u) = -sym(grad u grad dp_h).
// All integrands are polynomials of degree 2 (Deg - 1), integrated exactly by the reference's rules
// (GaussQuadrature.hh:115-127, 283-295).  One thread per element; per-DoF results are accumulated with FP64 atomics
// (element loops of O(elements) work; the host mirror of the same formulas is include/MeshFEM/LinearElasticity.hh
// applyDeltaStiffnessMatrix & co., used in host-only mode and as the CPU parity reference of these kernels).
#include "core.cuh"

namespace mfem {

template <int N>
__device__ __forceinline__ void sd_load_geom(const double *geom, int64_t e, double &vol, double g[N + 1][N]) {
    constexpr int GS = 1 + N * (N + 1);
    const double *gp = geom + e * GS;
    vol = gp[0];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int a = 0; a <= N; ++a) g[a][r] = gp[1 + r * (N + 1) + a];      // g[k] = grad lambda_k
}

// grad phi_i at the barycentric point lam (EmbeddedElement.hh:315-332)
template <int N, int DEG>
__device__ __forceinline__ void sd_grad_phi(const double g[N + 1][N], const double *lam, int i, double *out) {
    if (DEG == 1) {
#pragma unroll
        for (int r = 0; r < N; ++r) out[r] = g[i][r];
    } else if (i <= N) {
#pragma unroll
        for (int r = 0; r < N; ++r) out[r] = (4.0 * lam[i] - 1.0) * g[i][r];
    } else {
        const int s = edge_start(i - (N + 1)), e = edge_end(i - (N + 1));
#pragma unroll
        for (int r = 0; r < N; ++r) out[r] = 4.0 * (lam[e] * g[s][r] + lam[s] * g[e][r]);
    }
}

// G[a][r] = sum_k delta_p_k[a] grad lambda_k[r]
template <int N, int NPE>
__device__ __forceinline__ void sd_velocity_gradient(const int32_t *nd, const double g[N + 1][N], const double *__restrict__ deltaP,
                                                     double G[N][N], double &div) {
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int r = 0; r < N; ++r) G[a][r] = 0.0;
#pragma unroll
    for (int k = 0; k <= N; ++k)
#pragma unroll
        for (int a = 0; a < N; ++a) {
            const double dp = deltaP[(int64_t)nd[k] * N + a];
#pragma unroll
            for (int r = 0; r < N; ++r) G[a][r] += dp * g[k][r];
        }
    div = 0.0;
#pragma unroll
    for (int a = 0; a < N; ++a) div += G[a][a];
}

// sigma = D : sym(A) as a full symmetric matrix (ElasticityTensor.hh:435-447: D * shear-doubled strain)
template <int N>
__device__ __forceinline__ void sd_stress_of_gradient(const double *D, const double A[N][N], double sig[N][N]) {
    constexpr int F = flat_len(N);
    double eps[F];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = a; b < N; ++b) eps[flat_idx<N>(a, b)] = 0.5 * (A[a][b] + A[b][a]);
    double sf[F];
#pragma unroll
    for (int i = 0; i < F; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) s += D[i * F + j] * (j >= N ? 2.0 : 1.0) * eps[j];
        sf[i] = s;
    }
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) sig[a][b] = sf[flat_idx<N>(a, b)];
}

template <int N, int DEG>
__device__ __forceinline__ void sd_quadrature_point(int q, double *lam) {
    if (DEG == 1) {
#pragma unroll
        for (int v = 0; v <= N; ++v) lam[v] = 1.0 / (N + 1);
    } else {
        const double c0 = (N == 3) ? 0.58541019662496845446 : 2.0 / 3.0, c1 = (N == 3) ? 0.13819660112501051518 : 1.0 / 6.0;
#pragma unroll
        for (int v = 0; v <= N; ++v) lam[v] = (v == q) ? c0 : c1;
    }
}

// (delta K) u accumulated per (internal) DoF
template <int N, int DEG, bool PER_ELEM_D>
__global__ void __launch_bounds__(128)
k_apply_delta_K(int64_t nElems, const int32_t *__restrict__ elemNodes, const int32_t *__restrict__ nodeDof,
                const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem, const double *__restrict__ u,
                const double *__restrict__ deltaP, double *__restrict__ out) {
    constexpr int NPE = nodes_per_elem(N, DEG), F = flat_len(N), NQ = (DEG == 1) ? 1 : N + 1;
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    double vol, g[N + 1][N], G[N][N], div;
    sd_load_geom<N>(geom, e, vol, g);
    int32_t nd[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) nd[j] = elemNodes[e * NPE + j];
    sd_velocity_gradient<N, NPE>(nd, g, deltaP, G, div);
    const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
    double f[NPE][N];
#pragma unroll
    for (int i = 0; i < NPE; ++i)
#pragma unroll
        for (int c = 0; c < N; ++c) f[i][c] = 0.0;
#pragma unroll 1
    for (int q = 0; q < NQ; ++q) {
        double lam[N + 1];
        sd_quadrature_point<N, DEG>(q, lam);
        double gu[N][N], dgu[N][N];
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) gu[c][r] = 0.0;
#pragma unroll 1
        for (int i = 0; i < NPE; ++i) {
            double gp[N];
            sd_grad_phi<N, DEG>(g, lam, i, gp);
#pragma unroll
            for (int c = 0; c < N; ++c) {
                const double uic = u[(int64_t)nd[i] * N + c];
#pragma unroll
                for (int r = 0; r < N; ++r) gu[c][r] += uic * gp[r];
            }
        }
#pragma unroll
        for (int c = 0; c < N; ++c)
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m) s -= gu[c][m] * G[m][r];
                dgu[c][r] = s;
            }
        double sig[N][N], dsig[N][N];
        sd_stress_of_gradient<N>(D, gu, sig);
        sd_stress_of_gradient<N>(D, dgu, dsig);
        const double wq = vol / NQ;
#pragma unroll 1
        for (int i = 0; i < NPE; ++i) {
            double gp[N], dgp[N];
            sd_grad_phi<N, DEG>(g, lam, i, gp);
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double s = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m) s -= G[m][r] * gp[m];
                dgp[r] = s;
            }
#pragma unroll
            for (int c = 0; c < N; ++c) {
                double acc = 0.0;
#pragma unroll
                for (int r = 0; r < N; ++r) acc += (div * sig[c][r] + dsig[c][r]) * gp[r] + sig[c][r] * dgp[r];
                f[i][c] += wq * acc;
            }
        }
    }
#pragma unroll 1
    for (int i = 0; i < NPE; ++i)
#pragma unroll
        for (int c = 0; c < N; ++c) atomicAdd(out + (int64_t)nodeDof[nd[i]] * N + c, f[i][c]);
}

// change in constantStrainLoad(eps) under delta_p, accumulated per (internal) DoF
template <int N, int DEG, bool PER_ELEM_D>
__global__ void __launch_bounds__(128)
k_delta_const_strain_load(int64_t nElems, const int32_t *__restrict__ elemNodes, const int32_t *__restrict__ nodeDof,
                          const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem, const MatD eps,
                          const double *__restrict__ deltaP, double *__restrict__ out) {
    constexpr int NPE = nodes_per_elem(N, DEG), F = flat_len(N);
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    double vol, g[N + 1][N], G[N][N], div;
    sd_load_geom<N>(geom, e, vol, g);
    int32_t nd[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) nd[j] = elemNodes[e * NPE + j];
    sd_velocity_gradient<N, NPE>(nd, g, deltaP, G, div);
    const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
    // s = D : eps (eps given flattened: shear entries are the plain off-diagonal components)
    double sf[F], s[N][N];
#pragma unroll
    for (int i = 0; i < F; ++i) {
        double t = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) t += D[i * F + j] * (j >= N ? 2.0 : 1.0) * eps.d[j];
        sf[i] = t;
    }
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) s[a][b] = sf[flat_idx<N>(a, b)];
    double lam[N + 1];
#pragma unroll
    for (int v = 0; v <= N; ++v) lam[v] = 1.0 / (N + 1);          // grad phi_i is (at most) linear: element average = centroid value
#pragma unroll 1
    for (int i = 0; i < NPE; ++i) {
        double gp[N];
        sd_grad_phi<N, DEG>(g, lam, i, gp);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            double l = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) {
                double dg = 0.0;
#pragma unroll
                for (int m = 0; m < N; ++m) dg -= G[m][r] * gp[m];
                l += vol * s[c][r] * (div * gp[r] + dg);
            }
            atomicAdd(out + (int64_t)nodeDof[nd[i]] * N + c, l);
        }
    }
}

// change in the element-averaged strain: avg (delta strain)(u) + avg strain(delta_u), flattened per element
template <int N, int DEG>
__global__ void __launch_bounds__(128)
k_delta_avg_strain(int64_t nElems, const int32_t *__restrict__ elemNodes, const double *__restrict__ geom,
                   const double *__restrict__ u, const double *__restrict__ du, const double *__restrict__ deltaP,
                   double *__restrict__ out /* [nElems*F] */) {
    constexpr int NPE = nodes_per_elem(N, DEG), F = flat_len(N);
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    double vol, g[N + 1][N], G[N][N], div;
    sd_load_geom<N>(geom, e, vol, g);
    int32_t nd[NPE];
#pragma unroll
    for (int j = 0; j < NPE; ++j) nd[j] = elemNodes[e * NPE + j];
    sd_velocity_gradient<N, NPE>(nd, g, deltaP, G, div);
    double lam[N + 1];
#pragma unroll
    for (int v = 0; v <= N; ++v) lam[v] = 1.0 / (N + 1);
    double gu[N][N], total[N][N];
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r) { gu[c][r] = 0.0; total[c][r] = 0.0; }
#pragma unroll 1
    for (int i = 0; i < NPE; ++i) {
        double gp[N];
        sd_grad_phi<N, DEG>(g, lam, i, gp);
#pragma unroll
        for (int c = 0; c < N; ++c) {
            const double uic = u[(int64_t)nd[i] * N + c], duic = du[(int64_t)nd[i] * N + c];
#pragma unroll
            for (int r = 0; r < N; ++r) { gu[c][r] += uic * gp[r]; total[c][r] += duic * gp[r]; }
        }
    }
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int r = 0; r < N; ++r)
#pragma unroll
            for (int m = 0; m < N; ++m) total[c][r] -= gu[c][m] * G[m][r];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = a; b < N; ++b) out[e * F + flat_idx<N>(a, b)] = 0.5 * (total[a][b] + total[b][a]);
}

#define MFEM_SD_DISPATCH(LAUNCH)                      \
    do {                                              \
        if (c->N == 3 && c->deg == 1) LAUNCH(3, 1);   \
        else if (c->N == 3 && c->deg == 2) LAUNCH(3, 2); \
        else if (c->N == 2 && c->deg == 1) LAUNCH(2, 1); \
        else LAUNCH(2, 2);                            \
    } while (0)

// u_nodes, delta_p_nodes: device, per node in the caller's numbering (delta_p is read at vertex nodes only);
// out: device, per DoF in the caller's numbering
void apply_delta_K(mfem_b200_ctx *c, const double *u_nodes, const double *deltaP_nodes, double *out_ext) {
    MFEM_REQUIRE(c->geomValid && c->haveMaterial, MFEM_B200_ERR_INVALID, "apply_delta_K: mesh and material required");
    ensure_work(c);
    cudaStream_t s = c->stream;
    double *f_int = c->work.b;
    MFEM_CUDA(cudaMemsetAsync(f_int, 0, sizeof(double) * c->nvar(), s));
    const int grid = grid_for(c->nElems, 128);
#define LAUNCH(NN_, DD_)                                                                                                       \
    do {                                                                                                                       \
        if (c->perElemD)                                                                                                       \
            k_apply_delta_K<NN_, DD_, true><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->nodeDof, c->geom, c->Dconst, c->Delem, \
                                                                 u_nodes, deltaP_nodes, f_int);                                \
        else                                                                                                                   \
            k_apply_delta_K<NN_, DD_, false><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->nodeDof, c->geom, c->Dconst, nullptr, \
                                                                  u_nodes, deltaP_nodes, f_int);                               \
    } while (0)
    MFEM_SD_DISPATCH(LAUNCH);
#undef LAUNCH
    c->launches++;
    if (c->nRanks > 1) halo_exchange_add(c, f_int, c->N);
    permute_to_external(c, f_int, out_ext);
    MFEM_CUDA(cudaGetLastError());
}

void delta_const_strain_load(mfem_b200_ctx *c, const double *epsFlatHost, const double *deltaP_nodes, double *out_ext) {
    MFEM_REQUIRE(c->geomValid && c->haveMaterial, MFEM_B200_ERR_INVALID, "delta_const_strain_load: mesh and material required");
    ensure_work(c);
    MatD eps{};
    for (int i = 0; i < flat_len(c->N); ++i) eps.d[i] = epsFlatHost[i];
    cudaStream_t s = c->stream;
    double *f_int = c->work.b;
    MFEM_CUDA(cudaMemsetAsync(f_int, 0, sizeof(double) * c->nvar(), s));
    const int grid = grid_for(c->nElems, 128);
#define LAUNCH(NN_, DD_)                                                                                                                 \
    do {                                                                                                                                 \
        if (c->perElemD)                                                                                                                 \
            k_delta_const_strain_load<NN_, DD_, true><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->nodeDof, c->geom, c->Dconst, c->Delem, \
                                                                           eps, deltaP_nodes, f_int);                                    \
        else                                                                                                                             \
            k_delta_const_strain_load<NN_, DD_, false><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->nodeDof, c->geom, c->Dconst, nullptr, \
                                                                            eps, deltaP_nodes, f_int);                                   \
    } while (0)
    MFEM_SD_DISPATCH(LAUNCH);
#undef LAUNCH
    c->launches++;
    if (c->nRanks > 1) halo_exchange_add(c, f_int, c->N);
    permute_to_external(c, f_int, out_ext);
    MFEM_CUDA(cudaGetLastError());
}

void delta_avg_strain(mfem_b200_ctx *c, const double *u_nodes, const double *du_nodes, const double *deltaP_nodes, double *out) {
    MFEM_REQUIRE(c->geomValid, MFEM_B200_ERR_INVALID, "delta_avg_strain: mesh required");
    cudaStream_t s = c->stream;
    const int grid = grid_for(c->nElems, 128);
#define LAUNCH(NN_, DD_) k_delta_avg_strain<NN_, DD_><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->geom, u_nodes, du_nodes, deltaP_nodes, out)
    MFEM_SD_DISPATCH(LAUNCH);
#undef LAUNCH
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
}

}  // namespace mfem

// K7: operators around the solve -- K*u on nodes, constant-strain loads, strain/stress
// averages -- and the numbering conversions at the ABI boundary.
//
// Reference being replaced:
//   Simulator::applyStiffnessMatrix          LinearElasticity.hh:801-823
//   Simulator::constantStrainLoad            LinearElasticity.hh:551-562, 135-162
//   Simulator::averageStrainField/Stress     LinearElasticity.hh:528-549, 99-123
//   Simulator::dofToNodeField                LinearElasticity.hh:665-677
#include <algorithm>
#include <numeric>

#include "core.cuh"

namespace mfem {

__global__ void k_perm_in(int64_t nb, int N, const int32_t *__restrict__ int2ext, const double *__restrict__ ext,
                          double *__restrict__ in) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb * N) return;
    const int64_t d = i / N;
    const int c = (int)(i - d * N);
    in[i] = ext[(int64_t)int2ext[d] * N + c];
}
__global__ void k_perm_out(int64_t nb, int N, const int32_t *__restrict__ int2ext, const double *__restrict__ in,
                           double *__restrict__ ext) {
    const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nb * N) return;
    const int64_t d = i / N;
    const int c = (int)(i - d * N);
    ext[(int64_t)int2ext[d] * N + c] = in[i];
}

void permute_to_internal(mfem_b200_ctx *c, const double *ext, double *in) {
    k_perm_in<<<grid_for(c->nvar(), 256), 256, 0, c->stream>>>(c->nDofs, c->N, c->int2ext, ext, in);
    c->launches++;
}
void permute_to_external(mfem_b200_ctx *c, const double *in, double *ext) {
    k_perm_out<<<grid_for(c->nvar(), 256), 256, 0, c->stream>>>(c->nDofs, c->N, c->int2ext, in, ext);
    c->launches++;
}

template <int N>
__device__ __forceinline__ void load_geom(const double *geom, int64_t e, ElemGeom<N> &g) {
    constexpr int GS = 1 + N * (N + 1);
    const double *gp = geom + e * GS;
    g.vol = gp[0];
#pragma unroll
    for (int r = 0; r < N; ++r)
#pragma unroll
        for (int a = 0; a <= N; ++a) g.G[r][a] = gp[1 + r * (N + 1) + a];
}

// Matrix-free K*u on NODES: one thread per element, fp64 atomics into the nodal result.
// (Post-processing operator; independent of the assembled matrix, so it doubles as the
//  cross-check of the BSR SpMV at sizes no CPU oracle reaches.)
template <int N, int DEG, bool PER_ELEM_D>
__global__ void __launch_bounds__(128)
k_apply_K_elementwise(int64_t nElems, const int32_t *__restrict__ elemNodes, const double *__restrict__ geom,
                      const MatD Dc, const double *__restrict__ Delem, const double *__restrict__ u,
                      double *__restrict__ out) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int F = flat_len(N);
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    ElemGeom<N> g;
    load_geom<N>(geom, e, g);
    int32_t nd[NPE];
    double ue[NPE][N];
#pragma unroll
    for (int j = 0; j < NPE; ++j) {
        nd[j] = elemNodes[e * NPE + j];
#pragma unroll
        for (int k = 0; k < N; ++k) ue[j][k] = u[(int64_t)nd[j] * N + k];
    }
    const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
#pragma unroll 1
    for (int i = 0; i < NPE; ++i) {
        double fi[N];
#pragma unroll
        for (int k = 0; k < N; ++k) fi[k] = 0.0;
        ke_row_slice<N, DEG>(g, D, i, [&](int j, const double blk[N][N]) {
#pragma unroll
            for (int cc = 0; cc < N; ++cc)
#pragma unroll
                for (int dd = 0; dd < N; ++dd) fi[cc] += blk[cc][dd] * ue[j][dd];
        });
#pragma unroll
        for (int k = 0; k < N; ++k) atomicAdd(out + (int64_t)nd[i] * N + k, fi[k]);
    }
}

void apply_K_nodes(mfem_b200_ctx *c, const double *u_nodes_dev, double *Ku_nodes_dev) {
    MFEM_REQUIRE(c->geomValid && c->haveMaterial, MFEM_B200_ERR_INVALID, "apply_K: mesh and material required");
    cudaStream_t s = c->stream;
    MFEM_CUDA(cudaMemsetAsync(Ku_nodes_dev, 0, sizeof(double) * c->nNodes * c->N, s));
    const int grid = grid_for(c->nElems, 128);
#define LAUNCH(NN_, DD_)                                                                                           \
    do {                                                                                                           \
        if (c->perElemD)                                                                                           \
            k_apply_K_elementwise<NN_, DD_, true><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->geom, c->Dconst, \
                                                                       c->Delem, u_nodes_dev, Ku_nodes_dev);       \
        else                                                                                                       \
            k_apply_K_elementwise<NN_, DD_, false><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->geom, c->Dconst, \
                                                                        nullptr, u_nodes_dev, Ku_nodes_dev);       \
    } while (0)
    if (c->N == 3 && c->deg == 1) LAUNCH(3, 1);
    else if (c->N == 3 && c->deg == 2) LAUNCH(3, 2);
    else if (c->N == 2 && c->deg == 1) LAUNCH(2, 1);
    else LAUNCH(2, 2);
#undef LAUNCH
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
}

// sigma = D * shearDoubled(eps)   (ElasticityTensor.hh:435-447), as a full symmetric matrix
template <int N>
__device__ __forceinline__ void stress_of(const double *D, const double *epsFlat, double sig[N][N]) {
    constexpr int F = flat_len(N);
    double sf[F];
#pragma unroll
    for (int i = 0; i < F; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) s += D[i * F + j] * (j >= N ? 2.0 : 1.0) * epsFlat[j];
        sf[i] = s;
    }
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = 0; b < N; ++b) sig[a][b] = sf[flat_idx<N>(a, b)];
}

// Constant-strain load, gathered per DoF row over its element incidences (fixed order).
template <int N, int DEG, bool PER_ELEM_D>
__global__ void k_const_strain_load(int64_t nb, const int64_t *__restrict__ incPtr, const int32_t *__restrict__ incList,
                                    const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem,
                                    const MatD eps /* first flat entries */, double *__restrict__ f) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int F = flat_len(N);
    const int64_t row = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (row >= nb) return;
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = 0.0;
    double sig[N][N];
    if (!PER_ELEM_D) stress_of<N>(Dc.d, eps.d, sig);
    for (int64_t t = incPtr[row]; t < incPtr[row + 1]; ++t) {
        const int32_t id = incList[t];
        const int64_t e = id / NPE;
        const int i = id - (int)e * NPE;
        ElemGeom<N> g;
        load_geom<N>(geom, e, g);
        if (PER_ELEM_D) stress_of<N>(Delem + e * (F * F), eps.d, sig);
        double gi[N];
        int_grad_phi<N, DEG>(g, i, gi);
#pragma unroll
        for (int cc = 0; cc < N; ++cc)
#pragma unroll
            for (int r = 0; r < N; ++r) acc[cc] += sig[cc][r] * gi[r];
    }
#pragma unroll
    for (int k = 0; k < N; ++k) f[row * N + k] = acc[k];
}

void const_strain_load(mfem_b200_ctx *c, const double *epsFlatHost, double *f_ext_dev) {
    MFEM_REQUIRE(c->geomValid && c->haveMaterial, MFEM_B200_ERR_INVALID, "const_strain_load: mesh and material required");
    build_pattern(c);
    ensure_work(c);
    MatD eps{};
    for (int i = 0; i < flat_len(c->N); ++i) eps.d[i] = epsFlatHost[i];
    cudaStream_t s = c->stream;
    double *f_int = c->work.b;
    const int grid = grid_for(c->nDofs, 128);
#define LAUNCH(NN_, DD_)                                                                                        \
    do {                                                                                                        \
        if (c->perElemD)                                                                                        \
            k_const_strain_load<NN_, DD_, true><<<grid, 128, 0, s>>>(c->nDofs, c->incPtr, c->incList, c->geom,  \
                                                                     c->Dconst, c->Delem, eps, f_int);          \
        else                                                                                                    \
            k_const_strain_load<NN_, DD_, false><<<grid, 128, 0, s>>>(c->nDofs, c->incPtr, c->incList, c->geom, \
                                                                      c->Dconst, nullptr, eps, f_int);          \
    } while (0)
    if (c->N == 3 && c->deg == 1) LAUNCH(3, 1);
    else if (c->N == 3 && c->deg == 2) LAUNCH(3, 2);
    else if (c->N == 2 && c->deg == 1) LAUNCH(2, 1);
    else LAUNCH(2, 2);
#undef LAUNCH
    c->launches++;
    // element-partitioned run: every rank integrated its own elements only, the shared DoFs hold
    // partial sums -> one interface sum-exchange makes the load consistent on all sharers
    if (c->nRanks > 1) halo_exchange_add(c, f_int, c->N);
    permute_to_external(c, f_int, f_ext_dev);
    MFEM_CUDA(cudaGetLastError());
}

// Average strain / stress per element: mean of the strain interpolant's nodal values
// = sym( sum_i u_i (x) int grad phi_i ) / vol.
template <int N, int DEG, bool PER_ELEM_D>
__global__ void k_avg_strain_stress(int64_t nElems, const int32_t *__restrict__ elemNodes,
                                    const double *__restrict__ geom, const MatD Dc, const double *__restrict__ Delem,
                                    const double *__restrict__ u, double *__restrict__ strain,
                                    double *__restrict__ stress) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    constexpr int F = flat_len(N);
    const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= nElems) return;
    ElemGeom<N> g;
    load_geom<N>(geom, e, g);
    double grad[N][N];     // du_c / dx_r
#pragma unroll
    for (int cc = 0; cc < N; ++cc)
#pragma unroll
        for (int r = 0; r < N; ++r) grad[cc][r] = 0.0;
#pragma unroll 1
    for (int i = 0; i < NPE; ++i) {
        double gi[N];
        int_grad_phi<N, DEG>(g, i, gi);
        const int64_t nd = elemNodes[e * NPE + i];
#pragma unroll
        for (int cc = 0; cc < N; ++cc) {
            const double uc = u[nd * N + cc];
#pragma unroll
            for (int r = 0; r < N; ++r) grad[cc][r] += uc * gi[r];
        }
    }
    const double iv = 1.0 / g.vol;
    double ef[F];
#pragma unroll
    for (int a = 0; a < N; ++a)
#pragma unroll
        for (int b = a; b < N; ++b) ef[flat_idx<N>(a, b)] = 0.5 * (grad[a][b] + grad[b][a]) * iv;
    if (strain)
#pragma unroll
        for (int k = 0; k < F; ++k) strain[e * F + k] = ef[k];
    if (stress) {
        const double *D = PER_ELEM_D ? Delem + e * (F * F) : Dc.d;
#pragma unroll
        for (int i = 0; i < F; ++i) {
            double s = 0.0;
#pragma unroll
            for (int j = 0; j < F; ++j) s += D[i * F + j] * (j >= N ? 2.0 : 1.0) * ef[j];
            stress[e * F + i] = s;
        }
    }
}

void avg_strain_stress(mfem_b200_ctx *c, const double *u_nodes_dev, double *strain_dev, double *stress_dev) {
    MFEM_REQUIRE(c->geomValid, MFEM_B200_ERR_INVALID, "avg_strain_stress: no mesh set");
    MFEM_REQUIRE(!stress_dev || c->haveMaterial, MFEM_B200_ERR_INVALID, "avg_strain_stress: no material set");
    cudaStream_t s = c->stream;
    const int grid = grid_for(c->nElems, 128);
#define LAUNCH(NN_, DD_)                                                                                             \
    do {                                                                                                             \
        if (c->perElemD)                                                                                             \
            k_avg_strain_stress<NN_, DD_, true><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->geom, c->Dconst,    \
                                                                     c->Delem, u_nodes_dev, strain_dev, stress_dev); \
        else                                                                                                         \
            k_avg_strain_stress<NN_, DD_, false><<<grid, 128, 0, s>>>(c->nElems, c->elemNodes, c->geom, c->Dconst,   \
                                                                      nullptr, u_nodes_dev, strain_dev, stress_dev); \
    } while (0)
    if (c->N == 3 && c->deg == 1) LAUNCH(3, 1);
    else if (c->N == 3 && c->deg == 2) LAUNCH(3, 2);
    else if (c->N == 2 && c->deg == 1) LAUNCH(2, 1);
    else LAUNCH(2, 2);
#undef LAUNCH
    c->launches++;
    MFEM_CUDA(cudaGetLastError());
}

// Export the block-CSR in the caller's numbering (host-side re-sort; parity tests and
// --dumpMatrix only).
void export_bsr(mfem_b200_ctx *c, int64_t *rowptrOut, int32_t *colidxOut, double *valsOut) {
    MFEM_REQUIRE(c->patternValid, MFEM_B200_ERR_INVALID, "get_bsr: matrix not assembled");
    const int64_t nb = c->nDofs, nnzb = c->nnzb;
    const int NN = c->N * c->N;
    std::vector<int64_t> rp((size_t)nb + 1);
    std::vector<int32_t> ci((size_t)nnzb), i2e((size_t)nb), e2i((size_t)nb);
    std::vector<double> v;
    MFEM_CUDA(cudaMemcpy(rp.data(), c->rowptr, rp.size() * 8, cudaMemcpyDeviceToHost));
    MFEM_CUDA(cudaMemcpy(ci.data(), c->colidx, ci.size() * 4, cudaMemcpyDeviceToHost));
    MFEM_CUDA(cudaMemcpy(i2e.data(), c->int2ext, i2e.size() * 4, cudaMemcpyDeviceToHost));
    MFEM_CUDA(cudaMemcpy(e2i.data(), c->ext2int, e2i.size() * 4, cudaMemcpyDeviceToHost));
    if (valsOut) {
        MFEM_REQUIRE(c->valuesValid, MFEM_B200_ERR_INVALID, "get_bsr: values not assembled");
        v.resize((size_t)nnzb * NN);
        MFEM_CUDA(cudaMemcpy(v.data(), c->vals, v.size() * 8, cudaMemcpyDeviceToHost));
    }
    rowptrOut[0] = 0;
    for (int64_t re = 0; re < nb; ++re) {
        const int64_t ri = e2i[(size_t)re];
        rowptrOut[re + 1] = rowptrOut[re] + (rp[(size_t)ri + 1] - rp[(size_t)ri]);
    }
    std::vector<std::pair<int32_t, int64_t>> tmp;
    for (int64_t re = 0; re < nb; ++re) {
        const int64_t ri = e2i[(size_t)re];
        const int64_t b = rp[(size_t)ri], e = rp[(size_t)ri + 1];
        tmp.clear();
        for (int64_t k = b; k < e; ++k) tmp.emplace_back(i2e[(size_t)ci[(size_t)k]], k);
        std::sort(tmp.begin(), tmp.end());
        int64_t o = rowptrOut[re];
        const int N = c->N;
        for (auto &pr : tmp) {
            if (colidxOut) colidxOut[o] = pr.first;
            if (valsOut)      // device "row-plane" layout -> plain row-major blocks
                for (int r = 0; r < N; ++r)
                    for (int cc = 0; cc < N; ++cc)
                        valsOut[o * NN + r * N + cc] =
                            v[(size_t)(NN * b + (int64_t)r * (N * (e - b)) + (int64_t)N * (pr.second - b) + cc)];
            ++o;
        }
    }
}

}  // namespace mfem

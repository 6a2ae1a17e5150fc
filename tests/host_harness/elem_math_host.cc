// Host-side harness: compiles meshfem_b200/csrc/elem_math.cuh with g++ so the
// row-slice formulation of Ke can be checked against the oracle without a GPU.
#include "../../meshfem_b200/csrc/elem_math.cuh"
#include <cstring>
using namespace mfem;

template <int N, int DEG>
static void ke_full(const double *pts, const double *D, double *Ke, double *geom_out) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    double p[N + 1][N];
    for (int v = 0; v <= N; ++v) for (int r = 0; r < N; ++r) p[v][r] = pts[v * N + r];
    ElemGeom<N> g;
    embed(p, g);
    geom_out[0] = g.vol;
    for (int r = 0; r < N; ++r) for (int a = 0; a <= N; ++a) geom_out[1 + r * (N + 1) + a] = g.G[r][a];
    const int n = N * NPE;
    std::memset(Ke, 0, sizeof(double) * n * n);
    for (int i = 0; i < NPE; ++i) {
        ke_row_slice<N, DEG>(g, D, i, [&](int j, const double blk[N][N]) {
            for (int c = 0; c < N; ++c) for (int d = 0; d < N; ++d) Ke[(N * i + c) * n + N * j + d] += blk[c][d];
        });
    }
}

template <int N, int DEG>
static void int_grads(const double *pts, double *out) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    double p[N + 1][N];
    for (int v = 0; v <= N; ++v) for (int r = 0; r < N; ++r) p[v][r] = pts[v * N + r];
    ElemGeom<N> g;
    embed(p, g);
    for (int i = 0; i < NPE; ++i) int_grad_phi<N, DEG>(g, i, out + i * N);
}

template <int N, int DEG>
static void ke_full_rot(const double *pts, const double *D, double *Ke, int rot0) {
    constexpr int NPE = nodes_per_elem(N, DEG);
    double p[N + 1][N];
    for (int v = 0; v <= N; ++v) for (int r = 0; r < N; ++r) p[v][r] = pts[v * N + r];
    ElemGeom<N> g;
    embed(p, g);
    const int n = N * NPE;
    std::memset(Ke, 0, sizeof(double) * n * n);
    for (int i = 0; i < NPE; ++i) {
        ke_row_slice_rot<N, DEG>(g, D, i, (rot0 + i) % (N + 1), [&](int j, const double blk[N][N]) {
            for (int c = 0; c < N; ++c) for (int d = 0; d < N; ++d) Ke[(N * i + c) * n + N * j + d] += blk[c][d];
        });
    }
}
extern "C" int harness_ke_rot(int N, int deg, const double *pts, const double *D, double *Ke, int rot0) {
    if (N == 2 && deg == 1) ke_full_rot<2, 1>(pts, D, Ke, rot0);
    else if (N == 2 && deg == 2) ke_full_rot<2, 2>(pts, D, Ke, rot0);
    else if (N == 3 && deg == 1) ke_full_rot<3, 1>(pts, D, Ke, rot0);
    else if (N == 3 && deg == 2) ke_full_rot<3, 2>(pts, D, Ke, rot0);
    else return 1;
    return 0;
}

extern "C" int harness_ke(int N, int deg, const double *pts, const double *D, double *Ke, double *geom) {
    if (N == 2 && deg == 1) ke_full<2, 1>(pts, D, Ke, geom);
    else if (N == 2 && deg == 2) ke_full<2, 2>(pts, D, Ke, geom);
    else if (N == 3 && deg == 1) ke_full<3, 1>(pts, D, Ke, geom);
    else if (N == 3 && deg == 2) ke_full<3, 2>(pts, D, Ke, geom);
    else return 1;
    return 0;
}
extern "C" int harness_int_grads(int N, int deg, const double *pts, double *out) {
    if (N == 2 && deg == 1) int_grads<2, 1>(pts, out);
    else if (N == 2 && deg == 2) int_grads<2, 2>(pts, out);
    else if (N == 3 && deg == 1) int_grads<3, 1>(pts, out);
    else if (N == 3 && deg == 2) int_grads<3, 2>(pts, out);
    else return 1;
    return 0;
}

// ye = Ke * xe through the matrix-free element operator (elem_apply)
template <int N, int DEG>
static void apply_elem(const double *pts, const double *D, const double *xe, double *ye) {
    double p[N + 1][N];
    for (int v = 0; v <= N; ++v) for (int r = 0; r < N; ++r) p[v][r] = pts[v * N + r];
    ElemGeom<N> g;
    embed(p, g);
    double Ga[N + 1][N];
    for (int a = 0; a <= N; ++a) for (int r = 0; r < N; ++r) Ga[a][r] = g.G[r][a];
    elem_apply<N, DEG>(Ga, g.vol, D, [&](int j, int d) { return xe[j * N + d]; },
                       [&](int i, int c, double v) { ye[i * N + c] = v; });
}
extern "C" int harness_elem_apply(int N, int deg, const double *pts, const double *D, const double *xe, double *ye) {
    if (N == 2 && deg == 1) apply_elem<2, 1>(pts, D, xe, ye);
    else if (N == 2 && deg == 2) apply_elem<2, 2>(pts, D, xe, ye);
    else if (N == 3 && deg == 1) apply_elem<3, 1>(pts, D, xe, ye);
    else if (N == 3 && deg == 2) apply_elem<3, 2>(pts, D, xe, ye);
    else return 1;
    return 0;
}

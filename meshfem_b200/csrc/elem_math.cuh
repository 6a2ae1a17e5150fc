// Element math shared by every kernel on the assemble-and-solve path.
//
// Restates, for the GPU, the arithmetic of
//   EmbeddedElement.hh:170-190, 211-231   (volume + barycentric gradients)
//   EmbeddedElement.hh:288-332            (deg 1/2 shape-function gradients)
//   GaussQuadrature.hh:115-127, 283-295   (deg-2 tri / tet rules)
//   LinearElasticity.hh:165-232           (perElementStiffness)
// in the factorised form  Ke[i,c][j,d] = vol * sum_{a,b} W[i,a][j,b] S[a,c][b,d],
//   S[a,c][b,d] = sum_{r,t} G[r,a] C(r,c,d,t) G[t,b]   (geometry + material)
//   W[i,a][j,b] = sum_q w_q alpha_q[i][a] alpha_q[j][b] (constant, closed form below)
// which is exact (the integrand is quadratic and the rules are degree-2 exact).
//
// Everything is FP64.  Functions are __host__ __device__ so the CPU unit test
// (tests/test_elem_math_host.py) can check them against the oracle without a GPU.
#pragma once
#include <cmath>
#include <cstdint>

#ifdef __CUDACC__
#define MFEM_HD __host__ __device__ __forceinline__
#else
#define MFEM_HD inline
#endif

namespace mfem {

// Simplex.hh:39-41
MFEM_HD constexpr int edge_start(int k) { return (k < 3) ? k : (6 - k) % 3; }
MFEM_HD constexpr int edge_end(int k) { return (k < 3) ? (k + 1) % 3 : 3; }

MFEM_HD constexpr int nodes_per_elem(int N, int deg) {
    return deg == 1 ? N + 1 : (N == 2 ? 6 : 10);
}
MFEM_HD constexpr int flat_len(int N) { return N * (N + 1) / 2; }

// Flattening.hh:47-60
template <int N>
MFEM_HD constexpr int flat_idx(int i, int j) {
    if (i == j) return i;
    if (N == 2) return 2;
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    return hi == 2 ? 4 - lo : 5;
}

// Per-element geometry record kept in HBM: vol, then G[r][a] row-major (r < N, a <= N).
template <int N>
struct ElemGeom {
    double vol;
    double G[N][N + 1];
};

// EmbeddedElement.hh:211-231 (tet) / :170-190 (tri). p[v][r].
MFEM_HD void embed(const double p[4][3], ElemGeom<3> &g) {
    double a[3], b[3], n[4][3];
    // n0 = (p3-p1) x (p2-p1)
    for (int r = 0; r < 3; ++r) { a[r] = p[3][r] - p[1][r]; b[r] = p[2][r] - p[1][r]; }
    n[0][0] = a[1] * b[2] - a[2] * b[1]; n[0][1] = a[2] * b[0] - a[0] * b[2]; n[0][2] = a[0] * b[1] - a[1] * b[0];
    const double V6 = (p[0][0] - p[1][0]) * n[0][0] + (p[0][1] - p[1][1]) * n[0][1] + (p[0][2] - p[1][2]) * n[0][2];
    // n1 = (p2-p0) x (p3-p0)
    for (int r = 0; r < 3; ++r) { a[r] = p[2][r] - p[0][r]; b[r] = p[3][r] - p[0][r]; }
    n[1][0] = a[1] * b[2] - a[2] * b[1]; n[1][1] = a[2] * b[0] - a[0] * b[2]; n[1][2] = a[0] * b[1] - a[1] * b[0];
    // n2 = (p3-p0) x (p1-p0)
    for (int r = 0; r < 3; ++r) { a[r] = p[3][r] - p[0][r]; b[r] = p[1][r] - p[0][r]; }
    n[2][0] = a[1] * b[2] - a[2] * b[1]; n[2][1] = a[2] * b[0] - a[0] * b[2]; n[2][2] = a[0] * b[1] - a[1] * b[0];
    // n3 = (p1-p0) x (p2-p0)
    for (int r = 0; r < 3; ++r) { a[r] = p[1][r] - p[0][r]; b[r] = p[2][r] - p[0][r]; }
    n[3][0] = a[1] * b[2] - a[2] * b[1]; n[3][1] = a[2] * b[0] - a[0] * b[2]; n[3][2] = a[0] * b[1] - a[1] * b[0];
    g.vol = V6 / 6.0;
    for (int k = 0; k < 4; ++k)
        for (int r = 0; r < 3; ++r) g.G[r][k] = n[k][r] / V6;
}

MFEM_HD void embed(const double p[3][2], ElemGeom<2> &g) {
    const double e[3][2] = {{p[2][0] - p[1][0], p[2][1] - p[1][1]},
                            {p[0][0] - p[2][0], p[0][1] - p[2][1]},
                            {p[1][0] - p[0][0], p[1][1] - p[0][1]}};
    const double dblA = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    g.vol = dblA / 2.0;
    for (int k = 0; k < 3; ++k) {
        g.G[0][k] = -e[k][1] / dblA;
        g.G[1][k] = e[k][0] / dblA;
    }
}

// ---------------------------------------------------------------------------
// Shape-gradient terms.  grad phi_i(x) = sum over terms (a, kind, p) of
//   alpha(kind, x_p) * G[:,a]   with  alpha(V, x) = 4x - 1,  alpha(E, x) = 4x.
// Vertex function i:  one term (a = i, V, p = i).
// Edge function k = (s,e): two terms (a = s, E, p = e), (a = e, E, p = s).
// Degree 1: one term (a = i, constant 1).
// ---------------------------------------------------------------------------
template <int N, int DEG>
struct NodeTerms {
    int n;        // number of terms (1 or 2)
    int a[2];     // barycentric-gradient index
    int p[2];     // barycentric coordinate the coefficient depends on
    bool edge;    // kind E (true) or V (false); ignored for DEG 1
};

template <int N, int DEG>
MFEM_HD NodeTerms<N, DEG> node_terms(int i) {
    NodeTerms<N, DEG> t;
    if (DEG == 1 || i <= N) {
        t.n = 1; t.a[0] = i; t.p[0] = i; t.a[1] = 0; t.p[1] = 0; t.edge = false;
    } else {
        const int k = i - (N + 1);
        const int s = edge_start(k), e = edge_end(k);
        t.n = 2; t.a[0] = s; t.p[0] = e; t.a[1] = e; t.p[1] = s; t.edge = true;
    }
    return t;
}

// W = sum_q w_q alpha1(x_{p1}(q)) alpha2(x_{p2}(q)) for the (K+1)-point degree-2 rule
// whose q-th point has barycentric coordinate c0 at vertex q and c1 elsewhere
// (GaussQuadrature.hh:115-127: c0 = 2/3, c1 = 1/6; :283-295: c0, c1 below), w_q = 1/(K+1).
template <int N>
MFEM_HD double w_coeff(bool edge1, int p1, bool edge2, int p2) {
    constexpr double c0 = (N == 3) ? 0.58541019662496845446 : 2.0 / 3.0;
    constexpr double c1 = (N == 3) ? 0.13819660112501051518 : 1.0 / 6.0;
    const double A1 = edge1 ? 4.0 * c0 : 4.0 * c0 - 1.0, A0 = edge1 ? 4.0 * c1 : 4.0 * c1 - 1.0;
    const double B1 = edge2 ? 4.0 * c0 : 4.0 * c0 - 1.0, B0 = edge2 ? 4.0 * c1 : 4.0 * c1 - 1.0;
    constexpr double wq = 1.0 / (N + 1);
    return (p1 == p2) ? wq * (A1 * B1 + N * A0 * B0)
                      : wq * (A1 * B0 + A0 * B1 + (N - 1) * A0 * B0);
}

// Q[fl][d] = sum_t D[fl][flat(d,t)] G[t][b]          (contract the tensor with grad lambda_b)
template <int N>
MFEM_HD void contract_Q(const double *D /* flat x flat row-major, symmetric */,
                        const ElemGeom<N> &g, int b, double Q[flat_len(N)][N]) {
    constexpr int F = flat_len(N);
#pragma unroll
    for (int fl = 0; fl < F; ++fl)
#pragma unroll
        for (int d = 0; d < N; ++d) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < N; ++t) s += D[fl * F + flat_idx<N>(d, t)] * g.G[t][b];
            Q[fl][d] = s;
        }
}

// S[c][d] = sum_r G[r][a] Q[flat(r,c)][d]   ( = S[a,c][b,d] )
template <int N>
MFEM_HD void contract_S(const ElemGeom<N> &g, int a, const double Q[flat_len(N)][N], double S[N][N]) {
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int d = 0; d < N; ++d) {
            double s = 0.0;
#pragma unroll
            for (int r = 0; r < N; ++r) s += g.G[r][a] * Q[flat_idx<N>(r, c)][d];
            S[c][d] = s;
        }
}

// Row slice of the element stiffness: for local node i, calls
//   emit(j, blk)  with blk[c][d] a partial contribution to Ke[i,c][j,d]
// (vertex columns receive one call, edge columns two; partials must be summed).
template <int N, int DEG, class Emit>
MFEM_HD void ke_row_slice(const ElemGeom<N> &g, const double *D, int i, Emit &&emit) {
    const NodeTerms<N, DEG> ti = node_terms<N, DEG>(i);
    double Q[flat_len(N)][N];
    double S0[N][N], S1[N][N];
    double blk[N][N];
#pragma unroll 1
    for (int b = 0; b <= N; ++b) {
        contract_Q<N>(D, g, b, Q);
        contract_S<N>(g, ti.a[0], Q, S0);
        if (ti.n == 2) contract_S<N>(g, ti.a[1], Q, S1);
        if (DEG == 1) {
#pragma unroll
            for (int c = 0; c < N; ++c)
#pragma unroll
                for (int d = 0; d < N; ++d) blk[c][d] = g.vol * S0[c][d];
            emit(b, blk);
        } else {
            // columns that carry a term on grad lambda_b: vertex b, and every edge touching b
            {
                const double w0 = g.vol * w_coeff<N>(ti.edge, ti.p[0], false, b);
                const double w1 = (ti.n == 2) ? g.vol * w_coeff<N>(ti.edge, ti.p[1], false, b) : 0.0;
#pragma unroll
                for (int c = 0; c < N; ++c)
#pragma unroll
                    for (int d = 0; d < N; ++d)
                        blk[c][d] = w0 * S0[c][d] + ((ti.n == 2) ? w1 * S1[c][d] : 0.0);
                emit(b, blk);
            }
            constexpr int NE = N * (N + 1) / 2;
#pragma unroll 1
            for (int k = 0; k < NE; ++k) {
                const int s = edge_start(k), e = edge_end(k);
                int pB;
                if (s == b) pB = e; else if (e == b) pB = s; else continue;
                const double w0 = g.vol * w_coeff<N>(ti.edge, ti.p[0], true, pB);
                const double w1 = (ti.n == 2) ? g.vol * w_coeff<N>(ti.edge, ti.p[1], true, pB) : 0.0;
#pragma unroll
                for (int c = 0; c < N; ++c)
#pragma unroll
                    for (int d = 0; d < N; ++d)
                        blk[c][d] = w0 * S0[c][d] + ((ti.n == 2) ? w1 * S1[c][d] : 0.0);
                emit(N + 1 + k, blk);
            }
        }
    }
}

// Edges incident to each vertex (tet: 3 per vertex, triangle: 2), as local edge indices k with
// edge_start(k) == v or edge_end(k) == v.
template <int N>
MFEM_HD int vertex_edge(int v, int t) {
    // tet edges: 0:(0,1) 1:(1,2) 2:(2,0) 3:(0,3) 4:(2,3) 5:(1,3); triangle: first three
    if (N == 3) {
        const int tab[4][3] = {{0, 2, 3}, {0, 1, 5}, {1, 2, 4}, {3, 4, 5}};
        return tab[v][t];
    }
    const int tab[3][2] = {{0, 2}, {0, 1}, {1, 2}};
    return tab[v][t];
}

// Same row slice as ke_row_slice, but with a control flow that is UNIFORM across the lanes of a
// warp whatever b each lane works on: (N+1) outer steps x (1 vertex column + N edge columns)
// emits, and the barycentric index visited at outer step s is (s + rot) mod (N+1).  Lanes of a
// warp that process incidences of the same DoF row use different `rot`, so they reach a shared
// column block at different steps (fewer same-address accumulations per step).
template <int N, int DEG, class Emit>
MFEM_HD void ke_row_slice_rot(const ElemGeom<N> &g, const double *D, int i, int rot, Emit &&emit) {
    const NodeTerms<N, DEG> ti = node_terms<N, DEG>(i);
    double Q[flat_len(N)][N];
    double S0[N][N], S1[N][N];
    double blk[N][N];
#pragma unroll
    for (int c = 0; c < N; ++c)
#pragma unroll
        for (int d = 0; d < N; ++d) S1[c][d] = 0.0;
#pragma unroll 1
    for (int s = 0; s <= N; ++s) {
        int b = s + rot;
        if (b > N) b -= (N + 1);
        contract_Q<N>(D, g, b, Q);
        contract_S<N>(g, ti.a[0], Q, S0);
        if (ti.n == 2) contract_S<N>(g, ti.a[1], Q, S1);
        if (DEG == 1) {
#pragma unroll
            for (int c = 0; c < N; ++c)
#pragma unroll
                for (int d = 0; d < N; ++d) blk[c][d] = g.vol * S0[c][d];
            emit(b, blk);
        } else {
#pragma unroll 1
            for (int t = 0; t <= N; ++t) {
                // t == 0: vertex column b; t >= 1: the t-th edge column touching vertex b
                int col, pB;
                bool colEdge;
                if (t == 0) { col = b; pB = b; colEdge = false; }
                else {
                    const int k = vertex_edge<N>(b, t - 1);
                    const int es = edge_start(k), ee = edge_end(k);
                    col = N + 1 + k; pB = (es == b) ? ee : es; colEdge = true;
                }
                const double w0 = g.vol * w_coeff<N>(ti.edge, ti.p[0], colEdge, pB);
                const double w1 = (ti.n == 2) ? g.vol * w_coeff<N>(ti.edge, ti.p[1], colEdge, pB) : 0.0;
#pragma unroll
                for (int c = 0; c < N; ++c)
#pragma unroll
                    for (int d = 0; d < N; ++d) blk[c][d] = w0 * S0[c][d] + w1 * S1[c][d];
                emit(col, blk);
            }
        }
    }
}

// ---------------------------------------------------------------------------
// Matrix-free element operator  ye = Ke * xe  without forming Ke (the PCG's operator for quadratic
// elements, csrc/matfree.inl).  Same quantity as perElementStiffness (LinearElasticity.hh:165-232)
// applied to a vector, i.e. what Simulator::applyStiffnessMatrix (LinearElasticity.hh:801-823) sums,
// evaluated as  ye[i] = vol sum_q w_q sigma(q) grad phi_i(q),  sigma(q) = C : grad u(q),
// on the (N+1)-point degree-2 rule of w_coeff() -- exact for the quadratic integrand, so it equals
// Ke*xe up to rounding.  The rule's q-th point has lambda_b(q) = c1 + (c0 - c1) [b == q]; with
//   k0 = 4 c1 - 1, k1 = 4 c1, kd = 4 (c0 - c1)
// the coefficient of grad lambda_a in grad phi at point q is k0 + kd [a == q] (vertex function a) and
// k1 + kd [e == q] (edge function (a, e)), which leaves per element: one "base" displacement
// gradient, N+1 rank-one corrections, N+1 stress evaluations and 2 small contractions per output node
// (~700 FMA for a quadratic tet instead of the 900 multiply-adds of a dense 30x30 product, and
// 40 + 128 bytes of input instead of 7200).
// Ga[a][r] = d lambda_a / d x_r (ElemGeom::G transposed: the packed 32-byte slots of geomP).
// ---------------------------------------------------------------------------
template <int N>
MFEM_HD constexpr int edge_between(int a, int b) {       // local edge index of the edge {a, b}, a != b
    const int lo = a < b ? a : b, hi = a < b ? b : a;
    if (N == 2) return hi - lo == 1 ? lo : 2;            // (0,1)->0 (1,2)->1 (0,2)->2
    return hi == 3 ? (lo == 0 ? 3 : (lo == 1 ? 5 : 4))   // (0,3)->3 (1,3)->5 (2,3)->4
                   : (hi - lo == 1 ? lo : 2);            // (0,1)->0 (1,2)->1 (0,2)->2
}

// sigma_flat = scale * D * (shear-doubled flat strain of the displacement gradient H[d][t] = d u_d / d x_t)
template <int N>
MFEM_HD void stress_flat(const double *D, const double H[N][N], double scale, double sig[flat_len(N)]) {
    constexpr int F = flat_len(N);
    double e[F];
#pragma unroll
    for (int d = 0; d < N; ++d) e[d] = H[d][d];
    if (N == 2) e[2] = H[0][1] + H[1][0];
    else { e[3] = H[1][2] + H[2][1]; e[4] = H[0][2] + H[2][0]; e[5] = H[0][1] + H[1][0]; }
#pragma unroll
    for (int i = 0; i < F; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < F; ++j) s = fma(D[i * F + j], e[j], s);
        sig[i] = scale * s;
    }
}

// out[c] = sum_r sigma_{rc} g[r]
template <int N>
MFEM_HD void sig_dot(const double sig[flat_len(N)], const double g[N], double out[N]) {
#pragma unroll
    for (int c = 0; c < N; ++c) {
        double s = 0.0;
#pragma unroll
        for (int r = 0; r < N; ++r) s = fma(sig[flat_idx<N>(r, c)], g[r], s);
        out[c] = s;
    }
}

template <int N, int DEG, class GetX, class PutY>
MFEM_HD void elem_apply(const double Ga[N + 1][N], double vol, const double *D, GetX &&getx, PutY &&puty) {
    constexpr int F = flat_len(N);
    if (DEG == 1) {
        double H[N][N];
#pragma unroll
        for (int d = 0; d < N; ++d)
#pragma unroll
            for (int t = 0; t < N; ++t) H[d][t] = 0.0;
#pragma unroll
        for (int a = 0; a <= N; ++a)
#pragma unroll
            for (int d = 0; d < N; ++d) {
                const double xv = getx(a, d);
#pragma unroll
                for (int t = 0; t < N; ++t) H[d][t] = fma(xv, Ga[a][t], H[d][t]);
            }
        double sig[F];
        stress_flat<N>(D, H, vol, sig);
#pragma unroll
        for (int a = 0; a <= N; ++a) {
            double o[N];
            sig_dot<N>(sig, Ga[a], o);
#pragma unroll
            for (int c = 0; c < N; ++c) puty(a, c, o[c]);
        }
        return;
    }
    constexpr double c0 = (N == 3) ? 0.58541019662496845446 : 2.0 / 3.0;
    constexpr double c1 = (N == 3) ? 0.13819660112501051518 : 1.0 / 6.0;
    constexpr double k0 = 4.0 * c1 - 1.0, k1 = 4.0 * c1, kd = 4.0 * (c0 - c1);
    const double wv = vol / (N + 1);
    // base gradient: coefficient vectors with lambda == c1 everywhere
    double Hb[N][N];
#pragma unroll
    for (int d = 0; d < N; ++d)
#pragma unroll
        for (int t = 0; t < N; ++t) Hb[d][t] = 0.0;
#pragma unroll
    for (int a = 0; a <= N; ++a)
#pragma unroll
        for (int d = 0; d < N; ++d) {
            double es = 0.0;
#pragma unroll
            for (int b = 0; b <= N; ++b)
                if (b != a) es += getx(N + 1 + edge_between<N>(a, b), d);
            const double bv = fma(k1, es, k0 * getx(a, d));
#pragma unroll
            for (int t = 0; t < N; ++t) Hb[d][t] = fma(bv, Ga[a][t], Hb[d][t]);
        }
    double sig[N + 1][F], ssum[F];
#pragma unroll
    for (int i = 0; i < F; ++i) ssum[i] = 0.0;
#pragma unroll
    for (int q = 0; q <= N; ++q) {
        double H[N][N];
#pragma unroll
        for (int d = 0; d < N; ++d)
#pragma unroll
            for (int t = 0; t < N; ++t) H[d][t] = Hb[d][t];
#pragma unroll
        for (int a = 0; a <= N; ++a)
#pragma unroll
            for (int d = 0; d < N; ++d) {
                const double xv = kd * getx(a == q ? a : N + 1 + edge_between<N>(a, q), d);
#pragma unroll
                for (int t = 0; t < N; ++t) H[d][t] = fma(xv, Ga[a][t], H[d][t]);
            }
        stress_flat<N>(D, H, wv, sig[q]);
#pragma unroll
        for (int i = 0; i < F; ++i) ssum[i] += sig[q][i];
    }
    double T[N + 1][N];
#pragma unroll
    for (int a = 0; a <= N; ++a) sig_dot<N>(ssum, Ga[a], T[a]);
#pragma unroll
    for (int a = 0; a <= N; ++a) {
        double o[N];
        sig_dot<N>(sig[a], Ga[a], o);
#pragma unroll
        for (int c = 0; c < N; ++c) puty(a, c, fma(kd, o[c], k0 * T[a][c]));
    }
    constexpr int NE = N * (N + 1) / 2;
#pragma unroll
    for (int k = 0; k < NE; ++k) {
        const int s = edge_start(k), e = edge_end(k);
        double o1[N], o2[N];
        sig_dot<N>(sig[e], Ga[s], o1);
        sig_dot<N>(sig[s], Ga[e], o2);
#pragma unroll
        for (int c = 0; c < N; ++c) puty(N + 1 + k, c, fma(kd, o1[c] + o2[c], k1 * (T[s][c] + T[e][c])));
    }
}

// Integrated shape-function gradient  int grad phi_i dV = vol * sum_a t[a] G[:,a]
// (EmbeddedElement.hh:288-313 interpolant, integrated: Functions.hh:247-253).
// Degree 2: vertex i -> (3 - N)/(N+1) * ... evaluates to mean of nodal values:
//   vertex fn: (3*1 + (-1)*N)/(N+1) on a = i ;  edge fn (s,e): 4/(N+1) on a = e and a = s.
template <int N, int DEG>
MFEM_HD void int_grad_phi(const ElemGeom<N> &g, int i, double out[N]) {
    if (DEG == 1) {
        for (int r = 0; r < N; ++r) out[r] = g.vol * g.G[r][i];
        return;
    }
    if (i <= N) {
        const double w = g.vol * (3.0 - N) / (N + 1.0);
        for (int r = 0; r < N; ++r) out[r] = w * g.G[r][i];
    } else {
        const int k = i - (N + 1);
        const int s = edge_start(k), e = edge_end(k);
        const double w = g.vol * 4.0 / (N + 1.0);
        for (int r = 0; r < N; ++r) out[r] = w * (g.G[r][s] + g.G[r][e]);
    }
}

}  // namespace mfem

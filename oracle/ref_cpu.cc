// CPU restatement of the reference's assembly path WITH ITS DATA STRUCTURES AND LOOP
// STRUCTURE, used only as (a) a second oracle for the assembled matrix and (b) the timed CPU
// baseline of bench.py (cpu_baseline / --impl reference).  TEST INFRASTRUCTURE: nothing under
// meshfem_b200/, include/ or src/ may link or call this file.
//
// The reference itself cannot be built in this image (needs Eigen, SuiteSparse, TBB, Boost,
// nlohmann/json, tinyexpr; none present, no network), so this file follows it line by line:
//   Element::perElementStiffness            LinearElasticity.hh:165-232   (loop nest, upper triangle,
//                                                                          one heap allocation per call)
//   Simulator::m_assembleStiffnessMatrix    LinearElasticity.hh:1408-1466 (parallel Ke into a vector,
//                                                                          then SERIAL triplet scatter)
//   TripletMatrix::sumRepeated              SparseMatrices.hh:280-374     (serial column binning,
//                                                                          parallel per-column sort+sum)
//   EmbeddedElement gradPhi interpolants    EmbeddedElement.hh:288-313
//   Quadrature<K, 2(Deg-1)>                 GaussQuadrature.hh:115-127, 283-295
// std::thread over blocked element ranges stands in for tbb::parallel_for.
//
// Build: g++ -O3 -std=c++17 -shared -fPIC -pthread oracle/ref_cpu.cc -o oracle/_build/libref_cpu.so
#include <algorithm>
#include <array>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

using Clock = std::chrono::steady_clock;
double secs(Clock::time_point a, Clock::time_point b) { return std::chrono::duration<double>(b - a).count(); }

constexpr int edgeStart(int k) { return k < 3 ? k : (6 - k) % 3; }
constexpr int edgeEnd(int k) { return k < 3 ? (k + 1) % 3 : 3; }

template <int N> int flatIdx(int i, int j) {
    if (i == j) return i;
    if (N == 2) return 2;
    const int lo = std::min(i, j), hi = std::max(i, j);
    return hi == 2 ? 4 - lo : 5;
}

template <class F> void parallel_for(size_t n, int nThreads, F &&f) {
    if (nThreads <= 1 || n < 2) { f(0, n); return; }
    std::vector<std::thread> th;
    const size_t chunk = (n + nThreads - 1) / nThreads;
    for (int t = 0; t < nThreads; ++t) {
        const size_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
        if (b < e) th.emplace_back([=, &f] { f(b, e); });
    }
    for (auto &t : th) t.join();
}

template <int N> struct Vec { double v[N]; };

// An element as the reference stores it: volume + barycentric gradients + its own tensor copy
// (ETensorStoreGetter, LinearElasticity.hh:20-29).
template <int N> struct Element {
    double vol;
    double G[N][N + 1];
    double D[N * (N + 1) / 2][N * (N + 1) / 2];
    double C(int i, int j, int k, int l) const { return D[flatIdx<N>(i, j)][flatIdx<N>(k, l)]; }
};

template <int N> void embed(const double *p /* (N+1) x N */, Element<N> &el);
template <> void embed<3>(const double *p, Element<3> &el) {
    auto P = [&](int v, int r) { return p[v * 3 + r]; };
    auto crossInto = [&](int a0, int a1, int b0, int b1, double out[3]) {   // (p[a0]-p[a1]) x (p[b0]-p[b1])
        double a[3], b[3];
        for (int r = 0; r < 3; ++r) { a[r] = P(a0, r) - P(a1, r); b[r] = P(b0, r) - P(b1, r); }
        out[0] = a[1] * b[2] - a[2] * b[1]; out[1] = a[2] * b[0] - a[0] * b[2]; out[2] = a[0] * b[1] - a[1] * b[0];
    };
    double n[4][3];
    crossInto(3, 1, 2, 1, n[0]);
    const double V6 = (P(0, 0) - P(1, 0)) * n[0][0] + (P(0, 1) - P(1, 1)) * n[0][1] + (P(0, 2) - P(1, 2)) * n[0][2];
    crossInto(2, 0, 3, 0, n[1]);
    crossInto(3, 0, 1, 0, n[2]);
    crossInto(1, 0, 2, 0, n[3]);
    el.vol = V6 / 6.0;
    for (int k = 0; k < 4; ++k) for (int r = 0; r < 3; ++r) el.G[r][k] = n[k][r] / V6;
}
template <> void embed<2>(const double *p, Element<2> &el) {
    const double e[3][2] = {{p[4] - p[2], p[5] - p[3]}, {p[0] - p[4], p[1] - p[5]}, {p[2] - p[0], p[3] - p[1]}};
    const double dblA = e[1][0] * e[2][1] - e[1][1] * e[2][0];
    el.vol = dblA / 2.0;
    for (int k = 0; k < 3; ++k) { el.G[0][k] = -e[k][1] / dblA; el.G[1][k] = e[k][0] / dblA; }
}

// Interpolant<VecN, K, Deg-1>: nodal values at the K+1 vertices (Deg 2) or one constant (Deg 1).
template <int N, int DEG> struct SFGradient {
    static constexpr int NV = DEG == 1 ? 1 : N + 1;
    Vec<N> val[NV];
    Vec<N> operator()(const double *bary) const {
        Vec<N> r;
        for (int a = 0; a < N; ++a) r.v[a] = 0.0;
        if (DEG == 1) return val[0];
        for (int v = 0; v < NV; ++v) for (int a = 0; a < N; ++a) r.v[a] += bary[v] * val[v].v[a];
        return r;
    }
};

// EmbeddedElement::gradPhi(i) (EmbeddedElement.hh:288-313)
template <int N, int DEG> SFGradient<N, DEG> gradPhi(const Element<N> &el, int i) {
    SFGradient<N, DEG> g;
    if (DEG == 1) { for (int a = 0; a < N; ++a) g.val[0].v[a] = el.G[a][i]; return g; }
    for (int v = 0; v <= N; ++v) for (int a = 0; a < N; ++a) g.val[v % SFGradient<N, DEG>::NV].v[a] = 0.0;
    if (i <= N) {
        for (int v = 0; v <= N; ++v)
            for (int a = 0; a < N; ++a) g.val[v % SFGradient<N, DEG>::NV].v[a] = (v == i ? 3.0 : -1.0) * el.G[a][i];
    } else {
        const int k = i - (N + 1), s = edgeStart(k), e = edgeEnd(k);
        for (int a = 0; a < N; ++a) {
            g.val[s % SFGradient<N, DEG>::NV].v[a] = 4.0 * el.G[a][e];
            g.val[e % SFGradient<N, DEG>::NV].v[a] = 4.0 * el.G[a][s];
        }
    }
    return g;
}

template <int N, int DEG> struct Quad {
    static constexpr int NQ = DEG == 1 ? 1 : N + 1;
    double pts[NQ][N + 1];
    double w;
    Quad() {
        if (DEG == 1) { for (int v = 0; v <= N; ++v) pts[0][v] = 1.0 / (N + 1); w = 1.0; return; }
        const double c0 = N == 3 ? 0.58541019662496845446 : 2.0 / 3.0, c1 = N == 3 ? 0.13819660112501051518 : 1.0 / 6.0;
        for (int q = 0; q < NQ; ++q) for (int v = 0; v <= N; ++v) pts[q % NQ][v] = (v == q) ? c0 : c1;
        w = 1.0 / (N + 1);
    }
};

// Element::perElementStiffness (LinearElasticity.hh:165-232): upper triangle only.
template <int N, int DEG>
void perElementStiffness(const Element<N> &el, double *Ke /* n x n row-major */) {
    constexpr int nNodes = DEG == 1 ? N + 1 : (N == 2 ? 6 : 10);
    constexpr int n = N * nNodes;
    static const Quad<N, DEG> quad;
    std::vector<SFGradient<N, DEG>> grad_phis(nNodes);           // the reference heap-allocates here (:195)
    for (int k = 0; k < nNodes; ++k) grad_phis[k] = gradPhi<N, DEG>(el, k);
    double M[N][N];
    for (int c = 0; c < N; ++c) {
        for (int d = c; d < N; ++d) {
            for (int a = 0; a < N; ++a) for (int b = 0; b < N; ++b) M[a][b] = el.C(a, c, d, b);
            for (int j = 0; j < nNodes; ++j) {
                const int vj = j * N + d;
                SFGradient<N, DEG> Mgpj;
                for (int inode = 0; inode < SFGradient<N, DEG>::NV; ++inode)
                    for (int a = 0; a < N; ++a) {
                        double s = 0.0;
                        for (int b = 0; b < N; ++b) s += M[a][b] * grad_phis[j].val[inode].v[b];
                        Mgpj.val[inode].v[a] = s;
                    }
                for (int i = 0; i < nNodes; ++i) {
                    const int vi = i * N + c;
                    if (c == d && vi > vj) continue;
                    double val = 0.0;
                    for (int q = 0; q < Quad<N, DEG>::NQ; ++q) {
                        const Vec<N> gi = grad_phis[i](quad.pts[q]), mg = Mgpj(quad.pts[q]);
                        double dot = 0.0;
                        for (int a = 0; a < N; ++a) dot += gi.v[a] * mg.v[a];
                        val += quad.w * dot;
                    }
                    val *= el.vol;
                    if (vi <= vj) Ke[vi * n + vj] = val; else Ke[vj * n + vi] = val;
                }
            }
        }
    }
}

struct Triplet { size_t i, j; double v; };

struct Result {
    std::vector<int64_t> colptr, rowidx;
    std::vector<double> vals;
    double t_ke = 0, t_scatter = 0, t_compress = 0;
};

template <int N, int DEG>
void assemble(int64_t nNodes, const double *nodes, int64_t nElems, const int32_t *elemNodes, const double *D,
              int perElemD, const int64_t *dofForNode, int64_t nDofs, int nThreads, Result &res) {
    constexpr int nNodesE = DEG == 1 ? N + 1 : (N == 2 ? 6 : 10);
    constexpr int KeSize = N * nNodesE;
    constexpr int F = N * (N + 1) / 2;
    (void)nNodes;
    // mesh elements with embedded geometry and their own tensor copy
    std::vector<Element<N>> elems(nElems);
    for (int64_t e = 0; e < nElems; ++e) {
        double p[(N + 1) * N];
        for (int v = 0; v <= N; ++v) for (int r = 0; r < N; ++r) p[v * N + r] = nodes[(int64_t)elemNodes[e * nNodesE + v] * N + r];
        embed<N>(p, elems[e]);
        const double *De = perElemD ? D + e * F * F : D;
        for (int a = 0; a < F; ++a) for (int b = 0; b < F; ++b) elems[e].D[a][b] = (a <= b) ? De[a * F + b] : De[b * F + a];
    }
    auto DoF = [&](int64_t node) { return dofForNode ? dofForNode[node] : node; };
    const size_t nvars = (size_t)N * (size_t)nDofs;

    // "Assemble System": all Ke in parallel, then the serial scatter (LinearElasticity.hh:1444-1455)
    auto t0 = Clock::now();
    std::vector<double> elemMatrices((size_t)nElems * KeSize * KeSize);
    parallel_for((size_t)nElems, nThreads, [&](size_t b, size_t e) {
        for (size_t ei = b; ei < e; ++ei) perElementStiffness<N, DEG>(elems[ei], &elemMatrices[ei * KeSize * KeSize]);
    });
    auto t1 = Clock::now();
    std::vector<Triplet> nz;
    nz.reserve((size_t)KeSize * KeSize * nElems);                       // :1441-1443
    for (int64_t ei = 0; ei < nElems; ++ei) {
        const double *Ke = &elemMatrices[(size_t)ei * KeSize * KeSize];
        for (int i = 0; i < nNodesE; ++i) {
            const int64_t di = DoF(elemNodes[ei * nNodesE + i]);
            for (int j = 0; j < nNodesE; ++j) {
                const int64_t dj = DoF(elemNodes[ei * nNodesE + j]);
                if (di > dj) continue;
                for (int ci = 0; ci < N; ++ci)
                    for (int cj = 0; cj < N; ++cj) {
                        if (N * di + ci > N * dj + cj) continue;
                        const int row = N * i + ci, col = N * j + cj;
                        const double val = (row <= col) ? Ke[row * KeSize + col] : Ke[col * KeSize + row];
                        nz.push_back(Triplet{(size_t)(N * di + ci), (size_t)(N * dj + cj), val});
                    }
            }
        }
    }
    auto t2 = Clock::now();
    std::vector<double>().swap(elemMatrices);

    // "Compress Matrix": sumRepeated (SparseMatrices.hh:280-374) -- serial binning by column,
    // parallel per-column sort by row + sum of duplicates, prune v*v <= 0.
    struct RV { size_t i; double v; };
    std::vector<size_t> colCount(nvars + 1, 0);
    for (const auto &t : nz) colCount[t.j + 1]++;
    for (size_t c = 0; c < nvars; ++c) colCount[c + 1] += colCount[c];
    std::vector<RV> buckets(nz.size());
    {
        std::vector<size_t> pos(colCount.begin(), colCount.end() - 1);
        for (const auto &t : nz) buckets[pos[t.j]++] = RV{t.i, t.v};     // serial (:288)
    }
    std::vector<Triplet>().swap(nz);
    std::vector<size_t> colNnz(nvars, 0);
    parallel_for(nvars, nThreads, [&](size_t b, size_t e) {
        for (size_t c = b; c < e; ++c) {
            RV *beg = &buckets[colCount[c]], *end = &buckets[colCount[c + 1]];
            std::stable_sort(beg, end, [](const RV &a, const RV &b2) { return a.i < b2.i; });
            RV *out = beg;
            for (RV *it = beg; it != end;) {
                size_t row = it->i; double s = 0.0;
                while (it != end && it->i == row) { s += it->v; ++it; }
                if (s * s > 0.0) { out->i = row; out->v = s; ++out; }
            }
            colNnz[c] = (size_t)(out - beg);
        }
    });
    res.colptr.assign(nvars + 1, 0);
    for (size_t c = 0; c < nvars; ++c) res.colptr[c + 1] = res.colptr[c] + (int64_t)colNnz[c];
    res.rowidx.resize((size_t)res.colptr[nvars]);
    res.vals.resize((size_t)res.colptr[nvars]);
    parallel_for(nvars, nThreads, [&](size_t b, size_t e) {
        for (size_t c = b; c < e; ++c)
            for (size_t k = 0; k < colNnz[c]; ++k) {
                res.rowidx[(size_t)res.colptr[c] + k] = (int64_t)buckets[colCount[c] + k].i;
                res.vals[(size_t)res.colptr[c] + k] = buckets[colCount[c] + k].v;
            }
    });
    auto t3 = Clock::now();
    res.t_ke = secs(t0, t1); res.t_scatter = secs(t1, t2); res.t_compress = secs(t2, t3);
}

}  // namespace

extern "C" {

// Returns an opaque result (upper-triangle CSC) or NULL.  timings3 = {Ke, scatter, compress} seconds.
void *refcpu_assemble(int N, int deg, int64_t nNodes, const double *nodes, int64_t nElems, const int32_t *elemNodes,
                      const double *D, int perElemD, const int64_t *dofForNode, int64_t nDofs, int nThreads,
                      double *timings3, int64_t *nnzOut) {
    auto *r = new Result();
    if (!dofForNode) nDofs = nNodes;
    if (N == 3 && deg == 1) assemble<3, 1>(nNodes, nodes, nElems, elemNodes, D, perElemD, dofForNode, nDofs, nThreads, *r);
    else if (N == 3 && deg == 2) assemble<3, 2>(nNodes, nodes, nElems, elemNodes, D, perElemD, dofForNode, nDofs, nThreads, *r);
    else if (N == 2 && deg == 1) assemble<2, 1>(nNodes, nodes, nElems, elemNodes, D, perElemD, dofForNode, nDofs, nThreads, *r);
    else if (N == 2 && deg == 2) assemble<2, 2>(nNodes, nodes, nElems, elemNodes, D, perElemD, dofForNode, nDofs, nThreads, *r);
    else { delete r; return nullptr; }
    if (timings3) { timings3[0] = r->t_ke; timings3[1] = r->t_scatter; timings3[2] = r->t_compress; }
    if (nnzOut) *nnzOut = (int64_t)r->vals.size();
    return r;
}

void refcpu_copy(void *res, int64_t *colptr, int64_t *rowidx, double *vals) {
    auto *r = static_cast<Result *>(res);
    std::memcpy(colptr, r->colptr.data(), r->colptr.size() * 8);
    std::memcpy(rowidx, r->rowidx.data(), r->rowidx.size() * 8);
    std::memcpy(vals, r->vals.data(), r->vals.size() * 8);
}

void refcpu_free(void *res) { delete static_cast<Result *>(res); }

// One element's Ke by the reference loop nest (upper triangle written, rest untouched).
int refcpu_element_stiffness(int N, int deg, const double *pts, const double *D, double *Ke) {
    const int F = N * (N + 1) / 2;
    if (N == 3) {
        Element<3> el; embed<3>(pts, el);
        for (int a = 0; a < F; ++a) for (int b = 0; b < F; ++b) el.D[a][b] = (a <= b) ? D[a * F + b] : D[b * F + a];
        if (deg == 1) perElementStiffness<3, 1>(el, Ke); else perElementStiffness<3, 2>(el, Ke);
    } else if (N == 2) {
        Element<2> el; embed<2>(pts, el);
        for (int a = 0; a < F; ++a) for (int b = 0; b < F; ++b) el.D[a][b] = (a <= b) ? D[a * F + b] : D[b * F + a];
        if (deg == 1) perElementStiffness<2, 1>(el, Ke); else perElementStiffness<2, 2>(el, Ke);
    } else return 1;
    return 0;
}

// Block-Jacobi PCG on the host cores (same algorithm as the GPU solver; stands in for the
// reference's CHOLMOD solve, which cannot be built here).  A: full symmetric scalar CSR of the
// REDUCED system (fixed rows/cols already removed or replaced by identity); Minv: n/bs dense
// bs x bs inverse diagonal blocks.  Persistent worker threads, one barrier per phase.
int refcpu_pcg(int64_t n, const int64_t *rowptr, const int32_t *colidx, const double *vals, int bs,
               const double *Minv, const double *b, double *x, double rtol, int maxIters, int nThreads,
               int *itersOut, double *secondsOut, double *relresOut) {
    std::vector<double> r(b, b + n), z(n), p(n), Ap(n);
    std::fill(x, x + n, 0.0);
    const int64_t nb = n / bs;
    nThreads = std::max(1, nThreads);
    std::vector<double> part((size_t)nThreads * 8, 0.0);
    struct Barrier {
        std::atomic<int> count{0}, gen{0};
        int n;
        void wait() {
            const int g = gen.load();
            if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
            else while (gen.load() == g) std::this_thread::yield();
        }
    } bar;
    bar.n = nThreads;
    double bb = 0, rz = 0, rr = 0, pAp = 0;
    int iters = 0, state = 0;
    auto applyM = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i)
            for (int k = 0; k < bs; ++k) {
                double s = 0;
                for (int m = 0; m < bs; ++m) s += Minv[(i * bs + k) * bs + m] * r[i * bs + m];
                z[i * bs + k] = s;
            }
    };
    auto worker = [&](int t) {
        const int64_t c = (nb + nThreads - 1) / nThreads, i0 = std::min(nb, t * c), i1 = std::min(nb, i0 + c);
        const int64_t v0 = i0 * bs, v1 = i1 * bs;
        applyM(i0, i1);
        double a = 0, d = 0;
        for (int64_t i = v0; i < v1; ++i) { p[i] = z[i]; a += r[i] * z[i]; d += r[i] * r[i]; }
        part[t * 8] = a; part[t * 8 + 1] = d;
        bar.wait();
        if (t == 0) { rz = rr = 0; for (int k = 0; k < nThreads; ++k) { rz += part[k * 8]; rr += part[k * 8 + 1]; } bb = rr; if (bb == 0) state = 1; }
        bar.wait();
        while (state == 0 && iters < maxIters) {
            double s = 0;
            for (int64_t i = v0; i < v1; ++i) {
                double y = 0;
                for (int64_t k = rowptr[i]; k < rowptr[i + 1]; ++k) y += vals[k] * p[colidx[k]];
                Ap[i] = y; s += y * p[i];
            }
            part[t * 8] = s;
            bar.wait();
            if (t == 0) { pAp = 0; for (int k = 0; k < nThreads; ++k) pAp += part[k * 8]; }
            bar.wait();
            const double alpha = rz / pAp;
            for (int64_t i = v0; i < v1; ++i) { x[i] += alpha * p[i]; r[i] -= alpha * Ap[i]; }
            applyM(i0, i1);
            double a2 = 0, d2 = 0;
            for (int64_t i = v0; i < v1; ++i) { a2 += r[i] * z[i]; d2 += r[i] * r[i]; }
            part[t * 8] = a2; part[t * 8 + 1] = d2;
            bar.wait();
            if (t == 0) {
                double rzn = 0; rr = 0;
                for (int k = 0; k < nThreads; ++k) { rzn += part[k * 8]; rr += part[k * 8 + 1]; }
                part[nThreads * 8 - 1] = rzn / rz;   // beta
                rz = rzn; ++iters;
                if (!(pAp > 0)) state = 2;
                else if (rr <= rtol * rtol * bb) state = 1;
            }
            bar.wait();
            const double beta = part[nThreads * 8 - 1];
            for (int64_t i = v0; i < v1; ++i) p[i] = z[i] + beta * p[i];
            bar.wait();
        }
    };
    auto t0 = Clock::now();
    std::vector<std::thread> th;
    for (int t = 1; t < nThreads; ++t) th.emplace_back(worker, t);
    worker(0);
    for (auto &t : th) t.join();
    auto t1 = Clock::now();
    if (itersOut) *itersOut = iters;
    if (secondsOut) *secondsOut = secs(t0, t1);
    if (relresOut) *relresOut = bb > 0 ? std::sqrt(rr / bb) : 0.0;
    return state == 1 ? 0 : (state == 2 ? -6 : -7);
}

}  // extern "C"

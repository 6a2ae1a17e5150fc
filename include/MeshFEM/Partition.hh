// Element partitioning for multi-GPU execution (new design -- the reference is single-process;
// SURVEY 5.8 / 8e): non-overlapping element sets, nodes on partition interfaces are SHARED by
// every rank that touches them (owner = lowest rank).  Each rank assembles only its own elements,
// so interface rows of its local K hold partial sums; one sum-exchange per SpMV completes them.
//
// Partitioners (METIS is not available offline):
//   slabPartition  equal-count slabs of element centroids along the longest bounding-box axis -- the default; for the
//                  bar-shaped benchmark meshes slabs are also the minimum-interface cut;
//   rcbPartition   recursive coordinate bisection of the centroids for compact domains (a cube cut into 8 slabs has
//                  7 interfaces of a full cross-section each, into 2x2x2 boxes 12 quarter-sections: less interface
//                  per rank, but up to 7 neighbours instead of 2).  Opt-in (MESHFEM_PARTITIONER=rcb / partition(...,
//                  method="rcb")): the device path with more than two neighbours per rank is covered on CPU by the
//                  gloo emulation only (tests/test_multirank_cpu.py).
#ifndef MESHFEM_B200_PARTITION_HH
#define MESHFEM_B200_PARTITION_HH
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>
#include <vector>

namespace Partition {

// elemPart[e] in [0, nParts)
inline std::vector<int32_t> slabPartition(int dim, int64_t nNodes, const double *nodes, int64_t nElems, int npe,
                                          const int32_t *elemNodes, int nParts) {
    if (nParts < 1 || nParts > 64) throw std::runtime_error("slabPartition: 1..64 parts supported");
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < nNodes; ++i)
        for (int r = 0; r < dim; ++r) { mn[r] = std::min(mn[r], nodes[i * dim + r]); mx[r] = std::max(mx[r], nodes[i * dim + r]); }
    int axis = 0;
    for (int r = 1; r < dim; ++r) if (mx[r] - mn[r] > mx[axis] - mn[axis]) axis = r;
    const int nv = dim + 1;
    std::vector<double> key((size_t)nElems);
    for (int64_t e = 0; e < nElems; ++e) {
        double c = 0;
        for (int v = 0; v < nv; ++v) c += nodes[(int64_t)elemNodes[e * npe + v] * dim + axis];
        key[(size_t)e] = c / nv;
    }
    std::vector<int64_t> order((size_t)nElems);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return key[(size_t)a] < key[(size_t)b]; });
    std::vector<int32_t> part((size_t)nElems);
    for (int64_t k = 0; k < nElems; ++k) part[(size_t)order[(size_t)k]] = (int32_t)std::min<int64_t>(nParts - 1, k * nParts / nElems);
    return part;
}

// Recursive coordinate bisection: the part range [p0, p1) is split into floor / ceil halves, the elements -- ordered by
// their centroid coordinate along the longest extent of THIS subset, ties by element id -- in the same proportion.
inline std::vector<int32_t> rcbPartition(int dim, int64_t nNodes, const double *nodes, int64_t nElems, int npe,
                                         const int32_t *elemNodes, int nParts) {
    if (nParts < 1 || nParts > 64) throw std::runtime_error("rcbPartition: 1..64 parts supported");
    (void)nNodes;
    const int nv = dim + 1;
    std::vector<double> cen((size_t)nElems * 3, 0.0);
    for (int64_t e = 0; e < nElems; ++e)
        for (int r = 0; r < dim; ++r) {
            double c = 0;
            for (int v = 0; v < nv; ++v) c += nodes[(int64_t)elemNodes[e * npe + v] * dim + r];
            cen[(size_t)e * 3 + r] = c / nv;
        }
    std::vector<int32_t> part((size_t)nElems, 0);
    std::vector<int64_t> ids((size_t)nElems);
    std::iota(ids.begin(), ids.end(), 0);
    struct Job { int64_t b, e; int p0, p1; };
    std::vector<Job> stack{{0, nElems, 0, nParts}};
    while (!stack.empty()) {
        const Job j = stack.back();
        stack.pop_back();
        const int np = j.p1 - j.p0;
        if (np <= 1 || j.e - j.b == 0) {
            for (int64_t k = j.b; k < j.e; ++k) part[(size_t)ids[(size_t)k]] = (int32_t)j.p0;
            continue;
        }
        double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
        for (int64_t k = j.b; k < j.e; ++k)
            for (int r = 0; r < dim; ++r) {
                const double c = cen[(size_t)ids[(size_t)k] * 3 + r];
                mn[r] = std::min(mn[r], c); mx[r] = std::max(mx[r], c);
            }
        int axis = 0;
        for (int r = 1; r < dim; ++r) if (mx[r] - mn[r] > mx[axis] - mn[axis]) axis = r;
        const int npL = np / 2;
        const int64_t nL = (j.e - j.b) * npL / np;
        auto less = [&](int64_t a, int64_t b) {
            const double ca = cen[(size_t)a * 3 + axis], cb = cen[(size_t)b * 3 + axis];
            return ca < cb || (ca == cb && a < b);
        };
        std::nth_element(ids.begin() + j.b, ids.begin() + j.b + nL, ids.begin() + j.e, less);
        stack.push_back({j.b, j.b + nL, j.p0, j.p0 + npL});
        stack.push_back({j.b + nL, j.e, j.p0 + npL, j.p1});
    }
    return part;
}

struct LocalPart {
    std::vector<int64_t> elems;            // global element ids, ascending
    std::vector<int64_t> nodes;            // global node ids of the local nodes, ascending
    std::vector<int32_t> elemNodes;        // [elems.size() * npe] LOCAL node ids
    // The entities shared between ranks are DoFs: identical to the nodes unless periodic conditions
    // identify nodes (then nodes of one DoF may live on different ranks).
    std::vector<int64_t> dofs;             // global DoF ids of the local DoFs, ascending
    std::vector<int64_t> dofForNode;       // [nodes.size()] LOCAL DoF id of each local node (empty: identity)
    std::vector<uint8_t> owned;            // [dofs.size()] 1 if this rank is the lowest rank sharing the DoF
    std::vector<int32_t> neighborRanks;    // ascending
    std::vector<int64_t> neighborOffsets;  // [nNeighbors + 1] into sharedLocal
    std::vector<int32_t> sharedLocal;      // per neighbour: LOCAL DoF ids shared with it, ascending GLOBAL id
};

inline LocalPart extractPart(int rank, int nParts, int64_t nNodes, int64_t nElems, int npe, const int32_t *elemNodes,
                             const std::vector<int32_t> &elemPart, const int64_t *dofForNode = nullptr,
                             int64_t nDofs = 0) {
    const bool periodic = dofForNode != nullptr;
    if (!periodic) nDofs = nNodes;
    auto dofOf = [&](int64_t n) { return periodic ? dofForNode[n] : n; };
    std::vector<uint64_t> mask((size_t)nDofs, 0);      // ranks touching each DoF
    std::vector<uint8_t> mine((size_t)nNodes, 0);      // nodes of this rank's elements
    for (int64_t e = 0; e < nElems; ++e) {
        const uint64_t bit = 1ULL << elemPart[(size_t)e];
        for (int j = 0; j < npe; ++j) {
            const int64_t n = elemNodes[e * npe + j];
            mask[(size_t)dofOf(n)] |= bit;
            if (elemPart[(size_t)e] == rank) mine[(size_t)n] = 1;
        }
    }
    LocalPart lp;
    const uint64_t me = 1ULL << rank;
    std::vector<int32_t> localNode((size_t)nNodes, -1), localDof((size_t)nDofs, -1);
    for (int64_t n = 0; n < nNodes; ++n)
        if (mine[(size_t)n]) { localNode[(size_t)n] = (int32_t)lp.nodes.size(); lp.nodes.push_back(n); }
    for (int64_t d = 0; d < nDofs; ++d)
        if (mask[(size_t)d] & me) { localDof[(size_t)d] = (int32_t)lp.dofs.size(); lp.dofs.push_back(d); }
    if (periodic) {
        lp.dofForNode.resize(lp.nodes.size());
        for (size_t l = 0; l < lp.nodes.size(); ++l) lp.dofForNode[l] = localDof[(size_t)dofForNode[lp.nodes[l]]];
    }
    lp.owned.resize(lp.dofs.size());
    for (size_t l = 0; l < lp.dofs.size(); ++l) lp.owned[l] = (mask[(size_t)lp.dofs[l]] & (me - 1)) == 0;
    for (int64_t e = 0; e < nElems; ++e)
        if (elemPart[(size_t)e] == rank) {
            lp.elems.push_back(e);
            for (int j = 0; j < npe; ++j) lp.elemNodes.push_back(localNode[(size_t)elemNodes[e * npe + j]]);
        }
    lp.neighborOffsets.push_back(0);
    for (int q = 0; q < nParts; ++q) {
        if (q == rank) continue;
        const uint64_t qb = 1ULL << q;
        size_t before = lp.sharedLocal.size();
        for (size_t l = 0; l < lp.dofs.size(); ++l)
            if (mask[(size_t)lp.dofs[l]] & qb) lp.sharedLocal.push_back((int32_t)l);
        if (lp.sharedLocal.size() > before) { lp.neighborRanks.push_back(q); lp.neighborOffsets.push_back((int64_t)lp.sharedLocal.size()); }
    }
    return lp;
}

}  // namespace Partition
#endif

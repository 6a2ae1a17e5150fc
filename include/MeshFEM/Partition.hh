// Element partitioning for multi-GPU execution (new design -- the reference is single-process;
// SURVEY 5.8 / 8e): non-overlapping element sets, nodes on partition interfaces are SHARED by
// every rank that touches them (owner = lowest rank).  Each rank assembles only its own elements,
// so interface rows of its local K hold partial sums; one sum-exchange per SpMV completes them.
//
// Partitioner: equal-count slabs of element centroids along the longest bounding-box axis
// (METIS is not available offline; for the bar-shaped benchmark meshes slabs are also the
// minimum-interface cut).
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

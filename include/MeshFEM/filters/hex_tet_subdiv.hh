// Hex -> 24 tets (filters/hex_tet_subdiv.hh:32-104 of the reference): one centre vertex per
// hex, one per face (shared between neighbours), tets (e[f[v+1]], e[f[v]], faceCentre, hexCentre).
// The reference stitches face centres through a std::map<UnorderedQuadruplet>; a hash map with
// the same first-encounter numbering is used here so 10M-tet benchmark meshes build in seconds.
#ifndef MESHFEM_B200_HEX_TET_SUBDIV_HH
#define MESHFEM_B200_HEX_TET_SUBDIV_HH
#include <MeshFEM/Types.hh>

#include <algorithm>
#include <array>
#include <stdexcept>
#include <unordered_map>
#include <vector>

namespace hex_tet_detail {
struct QuadKey {
    std::array<size_t, 4> v;
    bool operator==(const QuadKey &o) const { return v == o.v; }
};
struct QuadHash {
    size_t operator()(const QuadKey &k) const {
        uint64_t h = 0x9e3779b97f4a7c15ULL;
        for (size_t x : k.v) { h ^= x + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); h *= 0xff51afd7ed558ccdULL; }
        return (size_t)h;
    }
};
}  // namespace hex_tet_detail

template <class Vertex, class Element>
void hex_tet_subdiv(const std::vector<Vertex> &inVertices, const std::vector<Element> &inElements,
                    std::vector<Vertex> &outVertices, std::vector<Element> &outElements, std::vector<size_t> &hexIdx) {
    using namespace hex_tet_detail;
    outVertices.clear(), outElements.clear();
    outVertices.reserve(inVertices.size() + 4 * inElements.size());
    outVertices = inVertices;
    outElements.reserve(24 * inElements.size());
    std::unordered_map<QuadKey, size_t, QuadHash> faceCenter;
    faceCenter.reserve(3 * inElements.size() + 16);

    std::vector<size_t> oldHexIdx(hexIdx);
    if (oldHexIdx.empty()) for (size_t i = 0; i < inElements.size(); ++i) oldHexIdx.push_back(i);
    if (oldHexIdx.size() != inElements.size()) throw std::runtime_error("Invalid hexIdx");
    hexIdx.clear(), hexIdx.reserve(24 * inElements.size());

    static const size_t faces[6][4] = {{0, 3, 2, 1}, {0, 4, 7, 3}, {4, 5, 6, 7}, {1, 2, 6, 5}, {0, 1, 5, 4}, {2, 3, 7, 6}};
    for (size_t i = 0; i < inElements.size(); ++i) {
        const auto &e = inElements[i];
        if (e.size() != 8) throw std::runtime_error("Non-hex encountered.");
        Point3D hexCenter = Point3D::Zero();
        for (size_t vi = 0; vi < 8; ++vi) hexCenter += Point3D(inVertices[e[vi]]);
        hexCenter /= 8;
        const size_t hCenterIdx = outVertices.size();
        outVertices.emplace_back(hexCenter);
        for (const auto &f : faces) {
            QuadKey q{{e[f[0]], e[f[1]], e[f[2]], e[f[3]]}};
            std::sort(q.v.begin(), q.v.end());
            size_t fCenterIdx;
            auto it = faceCenter.find(q);
            if (it == faceCenter.end()) {
                fCenterIdx = outVertices.size();
                Point3D mid = 0.25 * (Point3D(inVertices[e[f[0]]]) + Point3D(inVertices[e[f[1]]]) +
                                      Point3D(inVertices[e[f[2]]]) + Point3D(inVertices[e[f[3]]]));
                outVertices.emplace_back(mid);
                faceCenter.emplace(q, fCenterIdx);
            } else fCenterIdx = it->second;
            for (size_t v = 0; v < 4; ++v) {
                outElements.emplace_back(e[f[(v + 1) % 4]], e[f[v]], fCenterIdx, hCenterIdx);
                hexIdx.push_back(oldHexIdx[i]);
            }
        }
    }
}
#endif

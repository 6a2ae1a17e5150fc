// FEMMesh<K, Deg>: the mesh queries the assemble-and-solve path needs, with node, edge-node,
// boundary-element and boundary-node numbering IDENTICAL to the reference
// (FEMMesh.inl:11-82, TetMesh.inl:16-120, TriMesh.inl:16-140, FEMMesh.hh:221-237, 366-451):
//   * vertex nodes 0..nV-1 coincide with vertices; edge node k = nV + k with k the order of
//     first encounter sweeping elements in input order and local edges in Simplex.hh order;
//   * boundary faces in the iteration order of the reference's std::map<UnorderedTriplet,int>
//     leftovers (sorted vertex triples), boundary-face corner c = volume half-face corner 2-c;
//   * boundary vertices / boundary edge nodes numbered by first appearance in that sweep.
// Storage is flat SoA (what the C ABI takes), not the reference's half-face handle graph: the
// traversal/circulator API is out of scope for this path.  Hash tables replace std::map so that
// 10M-element meshes build in seconds; numbering does not depend on the container.
#ifndef MESHFEM_B200_FEMMESH_HH
#define MESHFEM_B200_FEMMESH_HH
#include <MeshFEM/Geometry.hh>
#include <MeshFEM/MeshIO.hh>
#include <MeshFEM/Simplex.hh>
#include <MeshFEM/Types.hh>

#include <algorithm>
#include <cstdlib>
#include <new>
#if defined(__linux__)
#include <sys/mman.h>
#endif
#include <memory>
#include <exception>
#include <stdexcept>
#include <thread>
#include <vector>

namespace femmesh_detail {

// open-addressing hash: 64-bit key (+32-bit tag for 96-bit keys) -> int32 value
struct Hash96 {
    struct Slot { uint64_t k; uint32_t t; int32_t val; int32_t aux; };
    // the slot array: 2 MB-aligned and advised into transparent huge pages where the system allows it (the
    // tables of a 10M-element mesh are GBs of randomly probed memory: with 4 kB pages every probe is also a
    // TLB miss)
    struct SlotArray {
        Slot *p = nullptr;
        size_t n = 0;
        ~SlotArray() { std::free(p); }
        SlotArray() = default;
        SlotArray(const SlotArray &) = delete;
        SlotArray &operator=(const SlotArray &) = delete;
        void assign(size_t count, const Slot &v) {
            std::free(p);
            const size_t huge = size_t(2) << 20, bytes = (count * sizeof(Slot) + huge - 1) / huge * huge;
            p = static_cast<Slot *>(std::aligned_alloc(huge, bytes));
            if (!p) throw std::bad_alloc();
#if defined(__linux__) && defined(MADV_HUGEPAGE)
            madvise(p, bytes, MADV_HUGEPAGE);
#endif
            n = count;
            for (size_t i = 0; i < count; ++i) p[i] = v;
        }
        Slot &operator[](size_t i) { return p[i]; }
        const Slot &operator[](size_t i) const { return p[i]; }
        const Slot *begin() const { return p; }
        const Slot *end() const { return p + n; }
    };
    SlotArray slots;
    uint64_t mask;
    size_t count = 0;
    // `expected` is an estimate, not a bound: the table doubles when it passes 70 % load (values are
    // kept, so nothing the callers number depends on the growth)
    explicit Hash96(size_t expected) {
        size_t cap = 16;
        while (2 * cap < 3 * expected + 32) cap <<= 1;
        slots.assign(cap, emptySlot());
        mask = cap - 1;
    }
    static Slot emptySlot() { return Slot{~0ULL, ~0u, -1, 0}; }
    static bool isEmpty(const Slot &s) { return s.val == -1 && s.k == ~0ULL && s.t == ~0u; }
    static uint64_t mix(uint64_t k, uint32_t t) {
        uint64_t h = k * 0x9e3779b97f4a7c15ULL + t;
        h ^= h >> 32; h *= 0xd6e8feb86659fd93ULL; h ^= h >> 32;
        return h;
    }
    // the build loops issue the cache-line fetch of a probe a few elements ahead of its use (same probes,
    // same order -- numbering is unchanged)
    void prefetch(uint64_t k, uint32_t t) const { __builtin_prefetch(&slots[mix(k, t) & mask], 1, 1); }
    void grow() {
        SlotArray old;
        std::swap(old.p, slots.p); std::swap(old.n, slots.n);
        slots.assign(2 * old.n, emptySlot());
        mask = slots.n - 1;
        for (const Slot &o : old) {
            if (isEmpty(o)) continue;
            uint64_t i = mix(o.k, o.t) & mask;
            while (!isEmpty(slots[i])) i = (i + 1) & mask;
            slots[i] = o;
        }
    }
    // returns slot reference (valid until the next insertion); `inserted` tells whether the key was new.
    // A new slot must be given a value != -1 or a key != ~0 by the caller (all callers set val >= 0).
    Slot &findOrInsert(uint64_t k, uint32_t t, bool &inserted) {
        if (10 * (count + 1) > 7 * slots.n) grow();
        uint64_t i = mix(k, t) & mask;
        while (true) {
            Slot &s = slots[i];
            if (isEmpty(s)) { s.k = k; s.t = t; inserted = true; ++count; return s; }
            if (s.k == k && s.t == t) { inserted = false; return s; }
            i = (i + 1) & mask;
        }
    }
    const Slot *find(uint64_t k, uint32_t t) const {
        uint64_t i = mix(k, t) & mask;
        while (true) {
            const Slot &s = slots[i];
            if (isEmpty(s)) return nullptr;
            if (s.k == k && s.t == t) return &s;
            i = (i + 1) & mask;
        }
    }
};

// threads for the table-building sweeps of large meshes: MESHFEM_NUM_THREADS=n (default 1 = the plain sweep: the
// sweeps are bound by random memory probes, and whether n threads help depends on the host -- on the 8-vCPU
// development VM eight threads probing private 24 MB tables run 10x slower EACH than one thread alone, so the
// parallel build is opt-in); MESHFEM_PARALLEL_MIN_ELEMENTS moves the size threshold (the tests set it to 1)
inline unsigned buildThreads(size_t numElements) {
    unsigned n = 1;
    if (const char *e = std::getenv("MESHFEM_NUM_THREADS")) n = (unsigned)std::max(1, std::atoi(e));
    size_t minElements = 200000;
    if (const char *e = std::getenv("MESHFEM_PARALLEL_MIN_ELEMENTS")) minElements = (size_t)std::max(1L, std::atol(e));
    if (numElements < minElements) return 1;
    return std::max(1u, std::min(n, 64u));
}
// f(chunk, begin, end) over T contiguous chunks of [0, n) (the same chunks on every call with the same n, T);
// the first exception thrown by a chunk is rethrown here
template <class F>
inline void parallelChunks(size_t n, unsigned T, F &&f) {
    std::vector<std::thread> threads;
    std::unique_ptr<std::exception_ptr[]> errors(new std::exception_ptr[T]);
    for (unsigned c = 0; c < T; ++c)
        threads.emplace_back([&, c]() {
            try { f(c, n * c / T, n * (c + 1) / T); } catch (...) { errors[c] = std::current_exception(); }
        });
    for (auto &t : threads) t.join();
    for (unsigned c = 0; c < T; ++c) if (errors[c]) std::rethrow_exception(errors[c]);
}

inline uint64_t pairKey(int a, int b) {
    const uint32_t lo = (uint32_t)std::min(a, b), hi = (uint32_t)std::max(a, b);
    return ((uint64_t)lo << 32) | hi;
}

}  // namespace femmesh_detail

template <size_t _K, size_t _Deg, class EmbeddingSpace = VectorND<_K>>
class FEMMesh {
public:
    static constexpr size_t K = _K;
    static constexpr size_t Deg = _Deg;
    static constexpr size_t N = _K;
    static constexpr size_t nodesPerElement = Simplex::numNodes(_K, _Deg);
    static constexpr size_t nodesPerBoundaryElement = Simplex::numNodes(_K - 1, _Deg);
    static constexpr size_t verticesPerElement = _K + 1;
    typedef EmbeddingSpace Point;
    static_assert(_K == 2 || _K == 3, "triangle and tet meshes only");
    static_assert(_Deg == 1 || _Deg == 2, "degree 1 or 2");

    template <typename Elements, typename Vertices>
    FEMMesh(const Elements &elems, const Vertices &vertices, bool suppressNonmanifoldWarning = false) {
        m_build(elems, vertices, suppressNonmanifoldWarning);
    }

    static std::unique_ptr<FEMMesh> load(const std::string &path) {
        std::vector<MeshIO::IOVertex> vertices;
        std::vector<MeshIO::IOElement> elements;
        MeshIO::load(path, vertices, elements);
        return std::unique_ptr<FEMMesh>(new FEMMesh(elements, vertices));
    }

    // ---- entity counts (FEMMesh.hh:150-180)
    size_t numVertices() const { return m_nV; }
    size_t numVertexNodes() const { return m_nV; }
    size_t numEdgeNodes() const { return m_nEdgeNodes; }
    size_t numNodes() const { return m_nV + m_nEdgeNodes; }
    size_t numElements() const { return m_ne; }
    size_t numBoundaryVertices() const { return m_bV.size(); }
    size_t numBoundaryEdgeNodes() const { return m_volEdgeForBdryEdge.size(); }
    size_t numBoundaryNodes() const { return m_bV.size() + m_volEdgeForBdryEdge.size(); }
    size_t numBoundaryElements() const { return m_nbe; }

    // ---- flat views (what mfem_b200_set_mesh takes)
    const std::vector<Real> &nodePositions() const { return m_nodes; }                 // [numNodes*K]
    const std::vector<int32_t> &elementNodes() const { return m_elemNodes; }           // [numElements*nodesPerElement]
    const std::vector<int32_t> &boundaryElementNodes() const { return m_bdryElemNodes; }     // volume node ids
    const std::vector<int32_t> &boundaryElementVertices() const { return m_bdryElemVerts; }  // volume vertex ids

    // ---- per-entity queries
    Point nodePosition(size_t n) const {
        Point p;
        for (size_t c = 0; c < _K; ++c) p[c] = m_nodes[n * _K + c];
        return p;
    }
    int elementNode(size_t e, size_t i) const { return m_elemNodes[e * nodesPerElement + i]; }
    int elementVertex(size_t e, size_t c) const { return m_elemNodes[e * nodesPerElement + c]; }
    // boundary element be, local node n -> VOLUME node index (be.node(n).volumeNode().index())
    int boundaryElementVolumeNode(size_t be, size_t n) const { return m_bdryElemNodes[be * nodesPerBoundaryElement + n]; }
    int boundaryElementVolumeVertex(size_t be, size_t c) const { return m_bdryElemVerts[be * _K + c]; }
    // boundary node bn -> volume node (m_volNodeForBdryNode, FEMMesh.hh:421-428)
    int volumeNodeForBoundaryNode(size_t bn) const {
        return bn < m_bV.size() ? m_bV[bn] : int(m_nV) + m_volEdgeForBdryEdge[bn - m_bV.size()];
    }
    // volume node -> boundary node or -1 (m_bdryNodeForVolNode, FEMMesh.hh:408-419)
    int boundaryNodeForVolumeNode(size_t n) const { return m_bdryNodeForNode[n]; }
    Real boundaryElementVolume(size_t be) const { return m_bdryVol[be]; }
    Point boundaryElementNormal(size_t be) const { return m_bdryNormal[be]; }

    // host-side element embedding (EmbeddedElement.hh:170-190, 211-231); the device computes the
    // same quantities in K1 -- this copy serves host-only consumers (mesh volume, tests)
    Real elementVolume(size_t e) const {
        Point p[_K + 1];
        for (size_t c = 0; c <= _K; ++c) p[c] = nodePosition(elementVertex(e, c));
        if (_K == 3) {
            Vector3D a, b, d;
            for (size_t r = 0; r < 3; ++r) { a[r] = p[3][r] - p[1][r]; b[r] = p[2][r] - p[1][r]; d[r] = p[0][r] - p[1][r]; }
            return d.dot(cross(a, b)) / 6.0;
        }
        const Real e1x = p[0][0] - p[2][0], e1y = p[0][1] - p[2][1], e2x = p[1][0] - p[0][0], e2y = p[1][1] - p[0][1];
        return (e1x * e2y - e1y * e2x) / 2.0;
    }
    // gradients of the barycentric coordinates of element e: g[a][r] = d lambda_a / d x_r
    // (LinearlyEmbeddedSimplex::embed, EmbeddedElement.hh:170-190, 211-231)
    void elementGradLambda(size_t e, Real g[_K + 1][_K]) const {
        Point p[_K + 1];
        for (size_t c = 0; c <= _K; ++c) p[c] = nodePosition(elementVertex(e, c));
        if (_K == 3) {
            auto v3 = [&](size_t i) { return Vector3D{p[i][0], p[i][1], p[i][2]}; };
            const Vector3D n0 = cross(v3(3) - v3(1), v3(2) - v3(1)), n1 = cross(v3(2) - v3(0), v3(3) - v3(0)),
                           n2 = cross(v3(3) - v3(0), v3(1) - v3(0)), n3 = cross(v3(1) - v3(0), v3(2) - v3(0));
            const Real V6 = (v3(0) - v3(1)).dot(n0);
            const Vector3D *n[4] = {&n0, &n1, &n2, &n3};
            for (size_t a = 0; a < 4; ++a) for (size_t r = 0; r < 3; ++r) g[a][r] = (*n[a])[r] / V6;
        } else {
            const Real ex[3] = {p[2][0] - p[1][0], p[0][0] - p[2][0], p[1][0] - p[0][0]};
            const Real ey[3] = {p[2][1] - p[1][1], p[0][1] - p[2][1], p[1][1] - p[0][1]};
            const Real dblA = ex[1] * ey[2] - ey[1] * ex[2];
            for (size_t a = 0; a < 3; ++a) { g[a][0] = -ey[a] / dblA; g[a][1] = ex[a] / dblA; }
        }
    }
    Real volume() const {
        Real v = 0;
        for (size_t e = 0; e < m_ne; ++e) v += elementVolume(e);
        return v;
    }

    const BBox<Point> &boundingBox() const { return m_bbox; }

    // (re-)embed: vertex nodes from the passed positions, edge nodes at edge midpoints
    // (FEMMesh.hh:221-237) -- elements are always straight-sided.
    template <typename Vertices>
    void setNodePositions(const Vertices &vertices) {
        if (vertices.size() != m_nV) throw std::runtime_error("setNodePositions: wrong vertex count");
        for (size_t v = 0; v < m_nV; ++v)
            for (size_t c = 0; c < _K; ++c) m_nodes[v * _K + c] = vertices[v][c];
        for (size_t k = 0; k < m_nEdgeNodes; ++k)
            for (size_t c = 0; c < _K; ++c)
                m_nodes[(m_nV + k) * _K + c] = 0.5 * (m_nodes[size_t(m_edgeEnds[2 * k]) * _K + c] +
                                                      m_nodes[size_t(m_edgeEnds[2 * k + 1]) * _K + c]);
        m_embedBoundary();
        m_computeBBox();
    }

private:
    size_t m_nV = 0, m_ne = 0, m_nEdgeNodes = 0, m_nbe = 0;
    std::vector<Real> m_nodes;
    std::vector<int32_t> m_elemNodes;
    std::vector<int32_t> m_edgeEnds;            // [2*numEdgeNodes]: (tip, tail) vertices, tip/tail as stored (min,max)
    std::vector<int32_t> m_bdryElemVerts, m_bdryElemNodes;
    std::vector<int32_t> m_bV;                  // boundary vertex -> volume vertex
    std::vector<int32_t> m_volEdgeForBdryEdge;  // boundary edge node -> volume edge node index
    std::vector<int32_t> m_bdryNodeForNode;
    std::vector<Real> m_bdryVol;
    std::vector<Point> m_bdryNormal;
    BBox<Point> m_bbox;

    template <typename Elements, typename Vertices>
    void m_build(const Elements &elems, const Vertices &vertices, bool suppressNonmanifoldWarning) {
        using namespace femmesh_detail;
        m_nV = vertices.size();
        m_ne = elems.size();
        constexpr size_t nv = _K + 1, nedge = Simplex::numEdges(_K), npe = nodesPerElement;
        m_elemNodes.assign(m_ne * npe, 0);
        for (size_t e = 0; e < m_ne; ++e) {
            if (elems[e].size() != nv) throw std::runtime_error(_K == 3 ? "Mesh must be pure tet" : "Mesh must be pure triangle");
            for (size_t c = 0; c < nv; ++c) {
                const size_t v = elems[e][c];
                if (v >= m_nV) throw std::runtime_error("Bad vertex index encountered.");
                m_elemNodes[e * npe + c] = (int32_t)v;
            }
        }
        // ---- volume edge nodes, first encounter (FEMMesh.inl:22-37)
        std::unique_ptr<Hash96> edgeTable;
        if (_Deg == 2) {
            edgeTable.reset(new Hash96(m_ne * nedge / (_K == 3 ? 4 : 1) + m_nV));
            m_edgeEnds.clear();
            constexpr size_t kAhead = 12;
            const unsigned T = buildThreads(m_ne);
            if (T > 1) {
                // Parallel first-encounter numbering: every chunk of elements numbers ITS edges in its own sweep
                // order (chunk-local table), then the chunks' distinct edges are merged in chunk order -- an edge
                // first met in chunk c is numbered in c's turn, after every edge of the earlier chunks and in c's
                // local order, which is exactly the order of the plain sweep.  The sequential part probes each
                // chunk's distinct edges once (~1.2 per element) instead of every element edge (6 per tet).
                std::vector<std::vector<uint64_t>> keys(T);
                parallelChunks(m_ne, T, [&](unsigned c, size_t e0, size_t e1) {
                    Hash96 local((e1 - e0) * nedge / (_K == 3 ? 4 : 1) + (e1 - e0) / 2 + 64);
                    std::vector<uint64_t> lk;                 // thread-private while it grows (no false sharing)
                    lk.reserve((e1 - e0) * 5 / 4 + 64);
                    for (size_t e = e0; e < e1; ++e) {
                        if (e + kAhead < e1)
                            for (size_t ei = 0; ei < nedge; ++ei)
                                local.prefetch(pairKey(m_elemNodes[(e + kAhead) * npe + Simplex::edgeStartNode(ei)],
                                                       m_elemNodes[(e + kAhead) * npe + Simplex::edgeEndNode(ei)]), 0);
                        for (size_t ei = 0; ei < nedge; ++ei) {
                            const uint64_t key = pairKey(m_elemNodes[e * npe + Simplex::edgeStartNode(ei)],
                                                         m_elemNodes[e * npe + Simplex::edgeEndNode(ei)]);
                            bool ins;
                            auto &slot = local.findOrInsert(key, 0, ins);
                            if (ins) { slot.val = (int32_t)lk.size(); lk.push_back(key); }
                            m_elemNodes[e * npe + nv + ei] = slot.val;        // chunk-local id until the merge
                        }
                    }
                    keys[c] = std::move(lk);
                });
                std::vector<std::vector<int32_t>> remap(T);
                for (unsigned c = 0; c < T; ++c) {
                    const std::vector<uint64_t> &lk = keys[c];
                    remap[c].resize(lk.size());
                    for (size_t i = 0; i < lk.size(); ++i) {
                        if (i + kAhead < lk.size()) edgeTable->prefetch(lk[i + kAhead], 0);
                        bool ins;
                        auto &slot = edgeTable->findOrInsert(lk[i], 0, ins);
                        if (ins) {
                            slot.val = (int32_t)m_nEdgeNodes++;
                            m_edgeEnds.push_back((int32_t)(lk[i] >> 32));            // pairKey: (min, max)
                            m_edgeEnds.push_back((int32_t)(lk[i] & 0xffffffffu));
                        }
                        remap[c][i] = slot.val;
                    }
                }
                parallelChunks(m_ne, T, [&](unsigned c, size_t e0, size_t e1) {
                    for (size_t e = e0; e < e1; ++e)
                        for (size_t ei = 0; ei < nedge; ++ei) {
                            int32_t &en = m_elemNodes[e * npe + nv + ei];
                            en = (int32_t)(m_nV + remap[c][en]);
                        }
                });
            } else
            for (size_t e = 0; e < m_ne; ++e) {
                if (e + kAhead < m_ne)
                    for (size_t ei = 0; ei < nedge; ++ei)
                        edgeTable->prefetch(pairKey(m_elemNodes[(e + kAhead) * npe + Simplex::edgeStartNode(ei)],
                                                    m_elemNodes[(e + kAhead) * npe + Simplex::edgeEndNode(ei)]), 0);
                for (size_t ei = 0; ei < nedge; ++ei) {
                    const int a = m_elemNodes[e * npe + Simplex::edgeStartNode(ei)];
                    const int b = m_elemNodes[e * npe + Simplex::edgeEndNode(ei)];
                    bool ins;
                    auto &slot = edgeTable->findOrInsert(pairKey(a, b), 0, ins);
                    if (ins) {
                        slot.val = (int32_t)m_nEdgeNodes++;
                        m_edgeEnds.push_back(std::min(a, b));
                        m_edgeEnds.push_back(std::max(a, b));
                    }
                    m_elemNodes[e * npe + nv + ei] = (int32_t)(m_nV + slot.val);
                }
            }
        }
        // ---- boundary extraction
        struct BFace { int v[3]; int hf; };
        std::vector<BFace> bfaces;
        if (_K == 3) {
            static const int fc[4][3] = {{1, 3, 2}, {0, 2, 3}, {0, 3, 1}, {0, 1, 2}};   // TetMesh.hh:221-226
            auto faceKey = [&](size_t t, int f, uint64_t &k, uint32_t &tag) {      // sorted vertex triple of face f of tet t
                int v[3] = {m_elemNodes[t * npe + fc[f][0]], m_elemNodes[t * npe + fc[f][1]], m_elemNodes[t * npe + fc[f][2]]};
                if (v[0] > v[1]) std::swap(v[0], v[1]);
                if (v[1] > v[2]) std::swap(v[1], v[2]);
                if (v[0] > v[1]) std::swap(v[0], v[1]);
                k = ((uint64_t)(uint32_t)v[0] << 32) | (uint32_t)v[1];
                tag = (uint32_t)v[2];
            };
            constexpr size_t kAhead = 12;
            // count the occurrences of every face of tets [t0, t1) in `table` (val = first half-face 4t+f)
            auto sweepFaces = [&](Hash96 &table, size_t t0, size_t t1) {
                for (size_t t = t0; t < t1; ++t) {
                    if (t + kAhead < t1)
                        for (int f = 0; f < 4; ++f) {
                            uint64_t k; uint32_t tag;
                            faceKey(t + kAhead, f, k, tag);
                            table.prefetch(k, tag);
                        }
                    for (int f = 0; f < 4; ++f) {
                        uint64_t k; uint32_t tag;
                        faceKey(t, f, k, tag);
                        bool ins;
                        auto &slot = table.findOrInsert(k, tag, ins);
                        if (ins) { slot.val = (int32_t)(4 * t + f); slot.aux = 1; }
                        else if (++slot.aux > 2) throw std::runtime_error("Non-manifold input detected.");
                    }
                }
            };
            const unsigned T = buildThreads(m_ne);
            std::vector<std::unique_ptr<Hash96>> chunkFaces(T > 1 ? T : 0);
            size_t unpaired = 0;
            if (T > 1) {
                // every chunk pairs the faces interior to it; only its unpaired faces (true boundary + the faces
                // it shares with other chunks) go through the sequential table below
                std::vector<size_t> cnt(T, 0);
                parallelChunks(m_ne, T, [&](unsigned c, size_t t0, size_t t1) {
                    chunkFaces[c].reset(new Hash96(2 * (t1 - t0) + 1024));
                    sweepFaces(*chunkFaces[c], t0, t1);
                    size_t n = 0;
                    for (const auto &sl : chunkFaces[c]->slots) n += (sl.val >= 0 && sl.aux == 1);
                    cnt[c] = n;
                });
                for (size_t n : cnt) unpaired += n;
            }
            Hash96 faces(T > 1 ? unpaired + 16 : 2 * m_ne + 16);
            if (T > 1) {
                for (unsigned c = 0; c < T; ++c)
                    for (const auto &sl : chunkFaces[c]->slots) {
                        if (!(sl.val >= 0 && sl.aux == 1)) continue;
                        bool ins;
                        auto &slot = faces.findOrInsert(sl.k, sl.t, ins);
                        if (ins) { slot.val = sl.val; slot.aux = 1; }
                        else if (++slot.aux > 2) throw std::runtime_error("Non-manifold input detected.");
                    }
                // a face still unpaired here must not be hidden as an interior pair of another chunk (three tets on one face)
                for (const auto &sl : faces.slots) {
                    if (!(sl.val >= 0 && sl.aux == 1)) continue;
                    const size_t t = size_t(sl.val) / 4;
                    for (unsigned c = 0; c < T; ++c) {
                        if (t >= m_ne * c / T && t < m_ne * (c + 1) / T) continue;       // its own chunk
                        if (chunkFaces[c]->find(sl.k, sl.t)) throw std::runtime_error("Non-manifold input detected.");
                    }
                }
                chunkFaces.clear();
            } else {
                sweepFaces(faces, 0, m_ne);
            }
            for (const auto &s : faces.slots)
                if (s.val >= 0 && s.aux == 1)
                    bfaces.push_back(BFace{{(int)(s.k >> 32), (int)(s.k & 0xffffffffu), (int)s.t}, s.val});
            std::sort(bfaces.begin(), bfaces.end(), [](const BFace &a, const BFace &b) {
                if (a.v[0] != b.v[0]) return a.v[0] < b.v[0];
                if (a.v[1] != b.v[1]) return a.v[1] < b.v[1];
                return a.v[2] < b.v[2];
            });
            m_nbe = bfaces.size();
            m_bdryElemVerts.assign(m_nbe * 3, 0);
            std::vector<int32_t> Vb(m_nV, -1);
            for (size_t i = 0; i < m_nbe; ++i) {
                const int hf = bfaces[i].hf, t = hf / 4, f = hf % 4;
                for (int c = 0; c < 3; ++c) {
                    const int v = m_elemNodes[size_t(t) * npe + fc[f][c]];     // m_vertexOfHalfFace(c, bhf)
                    if (Vb[v] == -1) { Vb[v] = (int32_t)m_bV.size(); m_bV.push_back(v); }   // TetMesh.inl:76-85
                    m_bdryElemVerts[i * 3 + (2 - c)] = v;                       // TetMesh.hh:462-468
                }
            }
        } else {
            // half-edge he = 3t + c: TIP = corner (c+2)%3, TAIL = corner (c+1)%3 (TriMesh.hh:286-298)
            Hash96 edges(2 * m_ne + 16);
            for (size_t t = 0; t < m_ne; ++t)
                for (int c = 0; c < 3; ++c) {
                    const int tip = m_elemNodes[t * npe + (c + 2) % 3], tail = m_elemNodes[t * npe + (c + 1) % 3];
                    bool ins;
                    auto &slot = edges.findOrInsert(pairKey(tip, tail), 0, ins);
                    if (ins) { slot.val = (int32_t)(3 * t + c); slot.aux = 1; }
                    else {
                        if (++slot.aux > 2) throw std::runtime_error("Non-manifold edge detected");
                        const int he0 = slot.val, t0 = he0 / 3, c0 = he0 % 3;
                        const int tip0 = m_elemNodes[size_t(t0) * npe + (c0 + 2) % 3], tail0 = m_elemNodes[size_t(t0) * npe + (c0 + 1) % 3];
                        if (tip != tail0 || tail != tip0) throw std::runtime_error("Inconsistent triangle orientations.");
                    }
                }
            for (const auto &s : edges.slots)
                if (s.val >= 0 && s.aux == 1) bfaces.push_back(BFace{{(int)(s.k >> 32), (int)(s.k & 0xffffffffu), 0}, s.val});
            std::sort(bfaces.begin(), bfaces.end(), [](const BFace &a, const BFace &b) {
                if (a.v[0] != b.v[0]) return a.v[0] < b.v[0];
                return a.v[1] < b.v[1];
            });
            m_nbe = bfaces.size();
            m_bdryElemVerts.assign(m_nbe * 2, 0);
            std::vector<int32_t> Vb(m_nV, -1);
            for (size_t i = 0; i < m_nbe; ++i) {
                const int vhe = bfaces[i].hf, t = vhe / 3, c = vhe % 3;
                // boundary edge tip = volume half-edge tail and vice versa (TriMesh.inl:104-106)
                const int tipVV = m_elemNodes[size_t(t) * npe + (c + 1) % 3], tailVV = m_elemNodes[size_t(t) * npe + (c + 2) % 3];
                if (Vb[tipVV] == -1) { Vb[tipVV] = (int32_t)m_bV.size(); m_bV.push_back(tipVV); }
                if (Vb[tailVV] == -1) { Vb[tailVV] = (int32_t)m_bV.size(); m_bV.push_back(tailVV); }
                m_bdryElemVerts[i * 2 + 0] = tailVV;     // vertex(0) = tail(), vertex(1) = tip()
                m_bdryElemVerts[i * 2 + 1] = tipVV;
            }
            if (m_bV.size() != m_nbe && !suppressNonmanifoldWarning)
                std::cerr << "WARNING: Boundary is non-manifold; this will break certain traversal operations" << std::endl;
        }
        // ---- boundary nodes (FEMMesh.inl:39-59)
        constexpr size_t npbe = nodesPerBoundaryElement, nbedge = Simplex::numEdges(_K - 1);
        m_bdryElemNodes.assign(m_nbe * npbe, 0);
        std::vector<int32_t> bdryEdgeForVolEdge(m_nEdgeNodes, -1);
        std::vector<uint8_t> coinciding(_Deg == 2 ? m_nEdgeNodes : 0, 0);
        for (size_t be = 0; be < m_nbe; ++be) {
            for (size_t c = 0; c < _K; ++c) m_bdryElemNodes[be * npbe + c] = m_bdryElemVerts[be * _K + c];
            if (_Deg == 2)
                for (size_t ei = 0; ei < nbedge; ++ei) {
                    const int a = m_bdryElemVerts[be * _K + Simplex::edgeStartNode(ei)];
                    const int b = m_bdryElemVerts[be * _K + Simplex::edgeEndNode(ei)];
                    const auto *slot = edgeTable->find(pairKey(a, b), 0);
                    if (!slot) throw std::runtime_error("boundary edge without volume edge");
                    const int volNode = slot->val;
                    if (coinciding[volNode] < 255) ++coinciding[volNode];
                    if (bdryEdgeForVolEdge[volNode] == -1) {
                        bdryEdgeForVolEdge[volNode] = (int32_t)m_volEdgeForBdryEdge.size();
                        m_volEdgeForBdryEdge.push_back(volNode);
                    }
                    m_bdryElemNodes[be * npbe + _K + ei] = (int32_t)(m_nV + volNode);
                }
        }
        if (_Deg == 2 && !suppressNonmanifoldWarning) {
            size_t nm = 0;
            for (uint8_t c : coinciding) nm += c > 2;
            if (nm > 0) std::cerr << "WARNING: " << nm << " non-manifold tetmesh edge(s) detected." << std::endl;
        }
        m_bdryNodeForNode.assign(numNodes(), -1);
        for (size_t bn = 0; bn < numBoundaryNodes(); ++bn) m_bdryNodeForNode[volumeNodeForBoundaryNode(bn)] = (int32_t)bn;

        m_nodes.assign(numNodes() * _K, 0.0);
        setNodePositions(vertices);
    }

    void m_embedBoundary() {
        m_bdryVol.assign(m_nbe, 0.0);
        m_bdryNormal.assign(m_nbe, Point());
        for (size_t be = 0; be < m_nbe; ++be) {
            if (_K == 3) {   // EmbeddedElement.hh:128-149
                Vector3D p[3];
                for (int c = 0; c < 3; ++c)
                    for (int r = 0; r < 3; ++r) p[c][r] = m_nodes[size_t(m_bdryElemVerts[be * 3 + c]) * 3 + r];
                Vector3D e1 = p[0] - p[2], e2 = p[1] - p[0];
                Vector3D n = cross(e1, e2);
                const Real dblA = n.norm();
                n /= dblA;
                m_bdryVol[be] = dblA / 2.0;
                for (int r = 0; r < 3; ++r) m_bdryNormal[be][r] = n[r];
            } else {         // EmbeddedElement.hh:87-104
                Real e[2];
                for (int r = 0; r < 2; ++r)
                    e[r] = m_nodes[size_t(m_bdryElemVerts[be * 2 + 1]) * 2 + r] - m_nodes[size_t(m_bdryElemVerts[be * 2 + 0]) * 2 + r];
                const Real L = std::sqrt(e[0] * e[0] + e[1] * e[1]);
                m_bdryVol[be] = L;
                m_bdryNormal[be][0] = -e[1] / L;
                m_bdryNormal[be][1] = e[0] / L;
            }
        }
    }

    void m_computeBBox() {     // FEMMesh.hh:438-446: over ALL nodes
        if (numNodes() == 0) { m_bbox = BBox<Point>(); return; }
        m_bbox = BBox<Point>(nodePosition(0), nodePosition(0));
        for (size_t i = 1; i < numNodes(); ++i) m_bbox.unionPoint(nodePosition(i));
    }
};

#endif

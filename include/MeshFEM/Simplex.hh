// Simplex node counts and Gmsh-consistent local orderings (mirrors Simplex.hh:16-46).
#ifndef MESHFEM_B200_SIMPLEX_HH
#define MESHFEM_B200_SIMPLEX_HH
#include <cstddef>
#include <stdexcept>

namespace Simplex {
constexpr size_t numVertices(size_t K) { return K + 1; }
constexpr size_t numEdges(size_t K) { return (K * (K + 1)) / 2; }
constexpr size_t numNodes(size_t K, size_t deg) {
    return K == 1 ? deg + 1
                  : (K == 2 ? ((deg + 1) * (deg + 2)) / 2
                            : (K == 3 ? ((deg + 1) * (deg + 2) * (deg + 3)) / 6
                                      : throw std::logic_error("Simplex dimension must be 1, 2, or 3")));
}
enum { Edge = 1, Triangle = 2, Tetrahedron = 3 };
// edge k joins local vertices edgeStartNode(k), edgeEndNode(k): {0,1,2,0,2,1} / {1,2,0,3,3,3}
constexpr size_t edgeStartNode(size_t i) { return (i < 3) ? i : (6 - i) % 3; }
constexpr size_t edgeEndNode(size_t i) { return (i < 3) ? (i + 1) % 3 : 3; }
}  // namespace Simplex
#endif

// Drop the vertices no element references and renumber the elements, in place
// (same contract as filters/remove_dangling_vertices.hh of the reference: relative vertex order kept).
#ifndef MESHFEM_B200_REMOVE_DANGLING_VERTICES_HH
#define MESHFEM_B200_REMOVE_DANGLING_VERTICES_HH
#include <cstddef>
#include <stdexcept>
#include <vector>

template <class Vertex, class Element>
void remove_dangling_vertices(std::vector<Vertex> &vertices, std::vector<Element> &elements) {
    const size_t unused = size_t(-1);
    std::vector<size_t> newIndex(vertices.size(), unused);
    for (const auto &e : elements)
        for (size_t c = 0; c < e.size(); ++c) {
            if (e[c] >= vertices.size()) throw std::out_of_range("remove_dangling_vertices: element references a missing vertex");
            newIndex[e[c]] = 0;
        }
    size_t kept = 0;
    for (size_t v = 0; v < vertices.size(); ++v) {
        if (newIndex[v] == unused) continue;
        if (kept != v) vertices[kept] = vertices[v];
        newIndex[v] = kept++;
    }
    vertices.resize(kept);
    for (auto &e : elements)
        for (size_t c = 0; c < e.size(); ++c) e[c] = newIndex[e[c]];
}
#endif

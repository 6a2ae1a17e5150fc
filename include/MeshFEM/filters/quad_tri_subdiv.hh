// Quad -> 4 triangles around a centre vertex (filters/quad_tri_subdiv.hh of the reference).
#ifndef MESHFEM_B200_QUAD_TRI_SUBDIV_HH
#define MESHFEM_B200_QUAD_TRI_SUBDIV_HH
#include <MeshFEM/Types.hh>

#include <stdexcept>
#include <vector>

template <class Vertex, class Element>
void quad_tri_subdiv(const std::vector<Vertex> &inVertices, const std::vector<Element> &inElements,
                     std::vector<Vertex> &outVertices, std::vector<Element> &outElements, std::vector<size_t> &quadIdx,
                     bool ignoreNonQuads = true) {
    outVertices.reserve(inVertices.size() + inElements.size());
    outVertices = inVertices;
    outElements.clear(), outElements.reserve(4 * inElements.size());
    std::vector<size_t> oldQuadIdx(quadIdx);
    if (oldQuadIdx.empty()) for (size_t i = 0; i < inElements.size(); ++i) oldQuadIdx.push_back(i);
    if (oldQuadIdx.size() != inElements.size()) throw std::runtime_error("Invalid quadIdx");
    quadIdx.clear(), quadIdx.reserve(4 * inElements.size());
    for (size_t i = 0; i < inElements.size(); ++i) {
        const auto &e = inElements[i];
        if (e.size() != 4) {
            if (ignoreNonQuads) { quadIdx.push_back(oldQuadIdx[i]); outElements.push_back(e); continue; }
            throw std::runtime_error("Non-quad encountered.");
        }
        Point3D center = Point3D(inVertices[e[0]]);
        center += Point3D(inVertices[e[1]]);
        center += Point3D(inVertices[e[2]]);
        center += Point3D(inVertices[e[3]]);
        center *= 0.25;
        const size_t ci = outVertices.size();
        outVertices.emplace_back(center);
        for (size_t v = 0; v < 4; ++v) {
            outElements.emplace_back(e[v], e[(v + 1) % 4], ci);
            quadIdx.push_back(oldQuadIdx[i]);
        }
    }
}
#endif

// Synthetic lattice meshes: restates the reference's `grid` tool inputs
// (filters/gen_grid.hh:14-92: integer-lattice vertices, quads / hexes in Gmsh order).
#ifndef MESHFEM_B200_GEN_GRID_HH
#define MESHFEM_B200_GEN_GRID_HH
#include <stdexcept>
#include <vector>

template <class Vertex, class Element>
void gen_grid(size_t sx, size_t sy, std::vector<Vertex> &vertices, std::vector<Element> &elements) {
    const size_t nCols = sx, nRows = sy;
    vertices.clear(), elements.clear();
    vertices.reserve((nCols + 1) * (nRows + 1));
    elements.reserve(nCols * nRows);
    auto idx = [=](size_t r, size_t c) { return (nCols + 1) * r + c; };
    for (size_t r = 0; r <= nRows; ++r)
        for (size_t c = 0; c <= nCols; ++c) vertices.emplace_back(Real(c), Real(r), Real(0));
    for (size_t r = 0; r < nRows; ++r)
        for (size_t c = 0; c < nCols; ++c)
            elements.emplace_back(idx(r, c), idx(r, c + 1), idx(r + 1, c + 1), idx(r + 1, c));
}

template <class Vertex, class Element>
void gen_grid(size_t sx, size_t sy, size_t sz, std::vector<Vertex> &vertices, std::vector<Element> &elements) {
    const size_t nCols = sx, nRows = sy, nSlices = sz;
    vertices.clear(), elements.clear();
    vertices.reserve((nCols + 1) * (nRows + 1) * (nSlices + 1));
    elements.reserve(nCols * nRows * nSlices);
    auto idx = [=](size_t s, size_t r, size_t c) { return (nCols + 1) * ((nRows + 1) * s + r) + c; };
    for (size_t s = 0; s <= nSlices; ++s)
        for (size_t r = 0; r <= nRows; ++r)
            for (size_t c = 0; c <= nCols; ++c) vertices.emplace_back(Real(c), Real(r), Real(s));
    for (size_t s = 0; s < nSlices; ++s)
        for (size_t r = 0; r < nRows; ++r)
            for (size_t c = 0; c < nCols; ++c)
                elements.emplace_back(idx(s, r, c), idx(s, r, c + 1), idx(s, r + 1, c + 1), idx(s, r + 1, c),
                                      idx(s + 1, r, c), idx(s + 1, r, c + 1), idx(s + 1, r + 1, c + 1),
                                      idx(s + 1, r + 1, c));
}

template <class Vertex, class Element>
void gen_grid(const std::vector<size_t> &sizes, std::vector<Vertex> &vertices, std::vector<Element> &elements) {
    switch (sizes.size()) {
        case 2: gen_grid(sizes[0], sizes[1], vertices, elements); break;
        case 3: gen_grid(sizes[0], sizes[1], sizes[2], vertices, elements); break;
        default: throw std::runtime_error("Only 2D and 3D grids are supported.");
    }
}
#endif

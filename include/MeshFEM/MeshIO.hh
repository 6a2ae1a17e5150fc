// Mesh file I/O: the formats the assemble-and-solve configs use.
// Mirrors src/lib/MeshFEM/MeshIO.hh (IOVertex / IOElement / load / save) for
// Gmsh MSH 2.2 ascii+binary (MeshIO.cc:527-760), OFF, OBJ (triangles) and MEDIT tets.
#ifndef MESHFEM_B200_MESHIO_HH
#define MESHFEM_B200_MESHIO_HH
#include <MeshFEM/Types.hh>

#include <iostream>
#include <string>
#include <vector>

namespace MeshIO {

struct IOVertex {
    Point3D point;
    IOVertex() {}
    IOVertex(Real x, Real y, Real z = 0) : point{x, y, z} {}
    IOVertex(const Point3D &p) : point(p) {}
    IOVertex(const Point2D &p) : point{p[0], p[1], 0.0} {}
    Real &operator[](size_t i) { return point[i]; }
    Real operator[](size_t i) const { return point[i]; }
    void set(Real x, Real y, Real z) { point = Point3D{x, y, z}; }
    operator Point3D() const { return point; }
};

struct IOElement {
    std::vector<size_t> corners;
    IOElement() {}
    explicit IOElement(size_t n) : corners(n) {}
    IOElement(size_t a, size_t b, size_t c) : corners{a, b, c} {}
    IOElement(size_t a, size_t b, size_t c, size_t d) : corners{a, b, c, d} {}
    IOElement(size_t a, size_t b, size_t c, size_t d, size_t e, size_t f, size_t g, size_t h)
        : corners{a, b, c, d, e, f, g, h} {}
    size_t size() const { return corners.size(); }
    void resize(size_t n) { corners.resize(n); }
    void push_back(size_t i) { corners.push_back(i); }
    size_t &operator[](size_t i) { return corners[i]; }
    size_t operator[](size_t i) const { return corners[i]; }
    std::vector<size_t>::const_iterator begin() const { return corners.begin(); }
    std::vector<size_t>::const_iterator end() const { return corners.end(); }
};

enum Format { FMT_OFF, FMT_OBJ, FMT_MSH, FMT_MSH_ASCII, FMT_MEDIT, FMT_GUESS, FMT_INVALID };
enum MeshType { MESH_LINE, MESH_TRI, MESH_QUAD, MESH_TET, MESH_HEX, MESH_LINE_DEG2, MESH_TRI_DEG2, MESH_TET_DEG2,
                MESH_TRI_QUAD, MESH_GUESS, MESH_INVALID };

Format guessFormat(const std::string &path);
size_t meshDegree(MeshType type);
size_t meshDimension(MeshType type);

MeshType load(const std::string &path, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements,
              Format format = FMT_GUESS, MeshType type = MESH_GUESS);
MeshType load(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements, Format format,
              MeshType type = MESH_GUESS);
void save(const std::string &path, const std::vector<IOVertex> &nodes, const std::vector<IOElement> &elements,
          Format format = FMT_GUESS, MeshType type = MESH_GUESS);
void save(std::ostream &os, const std::vector<IOVertex> &nodes, const std::vector<IOElement> &elements,
          Format format, MeshType type = MESH_GUESS);

// Gmsh element-type table (MeshIO.cc:527-531)
struct MSHElementInfo { MeshType meshType; int elementType; size_t nodesPerElem; };
MSHElementInfo mshInfoForMeshType(MeshType t);
MSHElementInfo mshInfoForElementType(int et);
MSHElementInfo mshInfoForNodeCount(size_t n);

}  // namespace MeshIO
#endif

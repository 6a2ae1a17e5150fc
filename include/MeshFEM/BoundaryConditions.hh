// Boundary-condition objects, the .bc JSON reader and the periodic condition
// (mirrors BoundaryConditions.hh / BoundaryConditions.cc:217-388 and
// PeriodicBoundaryMatcher.hh:38-258 of the reference).
//
// .bc top-level keys: no_rigid_motion, fix_periodic_pair_{x,y,z}, pin_translation, regions.
// Region types: dirichlet[xyz], target[xyz], force, traction, pressure, delta force,
// dirichlet nodes, target nodes, delta force nodes, {traction,pressure,force} elements,
// dirichlet elements (+ "element vertices").  Region geometry: box, box% (relative to the mesh
// bounding box).  Values: 2- or 3-vectors of numbers, or of tinyexpr-style expressions.
// `path` / `polygon` regions, `contact` and `fracture` are parsed as unsupported (they belong
// to products outside the assemble-and-solve path).
#ifndef MESHFEM_B200_BOUNDARYCONDITIONS_HH
#define MESHFEM_B200_BOUNDARYCONDITIONS_HH
#include <MeshFEM/ComponentMask.hh>
#include <MeshFEM/ExpressionVector.hh>
#include <MeshFEM/Geometry.hh>
#include <MeshFEM/JSON.hh>
#include <MeshFEM/Types.hh>

#include <algorithm>
#include <array>
#include <bitset>
#include <fstream>
#include <iostream>
#include <limits>
#include <map>
#include <memory>
#include <sstream>
#include <unordered_map>

enum class NeumannType { Pressure, Traction, Force };

template <size_t _N>
struct BoundaryCondition {
    std::shared_ptr<Region<VectorND<_N>>> region;
    BoundaryCondition() : region(new BBox<VectorND<_N>>()) {}
    explicit BoundaryCondition(std::shared_ptr<Region<VectorND<_N>>> r) : region(std::move(r)) {}
    bool containsPoint(const VectorND<_N> &p) const { return region->containsPoint(p); }
    virtual ~BoundaryCondition() {}
};
template <size_t _N> using CondPtr = std::shared_ptr<BoundaryCondition<_N>>;
template <size_t _N> using ConstCondPtr = std::shared_ptr<const BoundaryCondition<_N>>;

// value that is either a constant vector or an expression vector
template <size_t _N>
struct ValueOrExpression {
    VectorND<_N> value;
    ExpressionVector expr;
    bool isExpression() const { return expr.size() > 0; }
    VectorND<_N> operator()(const ExpressionEnvironment &env) const {
        return isExpression() ? expr.eval<_N>(env) : value;
    }
};

template <size_t _N>
struct NeumannCondition : public BoundaryCondition<_N> {
    typedef std::shared_ptr<Region<VectorND<_N>>> R;
    NeumannCondition(R r, Real p) : BoundaryCondition<_N>(r), type(NeumannType::Pressure), m_pressure(p) {}
    NeumannCondition(R r, const VectorND<_N> &t, NeumannType ty) : BoundaryCondition<_N>(r), type(ty) { m_val.value = t; }
    NeumannCondition(R r, const ExpressionVector &e, NeumannType ty) : BoundaryCondition<_N>(r), type(ty) { m_val.expr = e; }
    NeumannType type;
    Real pressure(const ExpressionEnvironment &) const { return m_pressure; }
    VectorND<_N> traction(const ExpressionEnvironment &env) const { return m_val(env); }
private:
    Real m_pressure = 0;
    ValueOrExpression<_N> m_val;
};

template <size_t _N>
struct DirichletCondition : public BoundaryCondition<_N> {
    typedef std::shared_ptr<Region<VectorND<_N>>> R;
    DirichletCondition(R r, const VectorND<_N> &v, ComponentMask m) : BoundaryCondition<_N>(r), componentMask(m) { m_val.value = v; }
    DirichletCondition(R r, const ExpressionVector &e, ComponentMask m) : BoundaryCondition<_N>(r), componentMask(m) { m_val.expr = e; }
    ComponentMask componentMask;
    VectorND<_N> displacement(const ExpressionEnvironment &env) const { return m_val(env); }
private:
    ValueOrExpression<_N> m_val;
};

template <size_t _N>
struct TargetCondition : public DirichletCondition<_N> { using DirichletCondition<_N>::DirichletCondition; };

template <size_t _N>
struct DirichletElementsCondition : public BoundaryCondition<_N> {
    DirichletElementsCondition(const std::vector<IVectorND<_N>> &ev, const VectorND<_N> &v, ComponentMask m)
        : elementVertices(ev), componentMask(m) { m_val.value = v; }
    DirichletElementsCondition(const std::vector<IVectorND<_N>> &ev, const ExpressionVector &e, ComponentMask m)
        : elementVertices(ev), componentMask(m) { m_val.expr = e; }
    std::vector<IVectorND<_N>> elementVertices;
    ComponentMask componentMask;
    bool containsElement(const IVectorND<_N> &idx) const {
        for (const auto &e : elementVertices) if (e == idx) return true;
        return false;
    }
    VectorND<_N> displacement(const ExpressionEnvironment &env) const { return m_val(env); }
private:
    ValueOrExpression<_N> m_val;
};

template <size_t _N>
struct NeumannElementsCondition : public BoundaryCondition<_N> {
    struct Value {
        NeumannType type;
        VectorND<_N> v;
        Real pressure() const { return v[0]; }
        VectorND<_N> traction() const { return v; }
        VectorND<_N> force() const { return v; }
    };
    NeumannElementsCondition(NeumannType t, const std::vector<UnorderedTriplet> &corners, const std::vector<VectorND<_N>> &values) {
        for (size_t i = 0; i < corners.size(); ++i) m_values[corners[i]] = Value{t, values[i]};
    }
    bool hasValueForElement(const UnorderedTriplet &e) const { return m_values.count(e) > 0; }
    const Value &getValue(const UnorderedTriplet &e) const { return m_values.at(e); }
    size_t numElements() const { return m_values.size(); }
private:
    std::map<UnorderedTriplet, Value> m_values;
};

template <size_t _N>
struct DirichletNodesCondition : public BoundaryCondition<_N> {
    DirichletNodesCondition(const std::vector<size_t> &idx, const std::vector<VectorND<_N>> &d, ComponentMask m)
        : indices(idx), displacements(d), componentMask(m) {}
    std::vector<size_t> indices;
    std::vector<VectorND<_N>> displacements;
    ComponentMask componentMask;
};
template <size_t _N>
struct TargetNodesCondition : public DirichletNodesCondition<_N> { using DirichletNodesCondition<_N>::DirichletNodesCondition; };

template <size_t _N>
struct DeltaForceCondition : public BoundaryCondition<_N> {
    typedef std::shared_ptr<Region<VectorND<_N>>> R;
    DeltaForceCondition(R r, const VectorND<_N> &v) : BoundaryCondition<_N>(r) { m_val.value = v; }
    DeltaForceCondition(R r, const ExpressionVector &e) : BoundaryCondition<_N>(r) { m_val.expr = e; }
    VectorND<_N> force(const ExpressionEnvironment &env) const { return m_val(env); }
private:
    ValueOrExpression<_N> m_val;
};

template <size_t _N>
struct DeltaForceNodesCondition : public BoundaryCondition<_N> {
    DeltaForceNodesCondition(const std::vector<size_t> &idx, const std::vector<VectorND<_N>> &f) : indices(idx), forces(f) {}
    std::vector<size_t> indices;
    std::vector<VectorND<_N>> forces;
};

// fix_periodic_pair_<component>: "<face axis>"
template <size_t _N>
struct PeriodicPairDirichletCondition {
    PeriodicPairDirichletCondition(size_t component, size_t face) : m_component(component), m_face(face) {}
    ComponentMask component() const { ComponentMask m; m.set(m_component); return m; }
    size_t faceAxis() const { return m_face; }
    // first pair of boundary nodes lying on the min / max faces of axis `face` with equal
    // remaining coordinates (tolerance 1e-7)
    template <class Mesh>
    std::pair<size_t, size_t> pair(const Mesh &mesh) const {
        const auto &bb = mesh.boundingBox();
        const Real eps = 1e-7;
        for (size_t i = 0; i < mesh.numBoundaryNodes(); ++i) {
            auto p = mesh.nodePosition(mesh.volumeNodeForBoundaryNode(i));
            if (std::abs(p[m_face] - bb.minCorner[m_face]) > eps) continue;
            for (size_t j = 0; j < mesh.numBoundaryNodes(); ++j) {
                auto q = mesh.nodePosition(mesh.volumeNodeForBoundaryNode(j));
                if (std::abs(q[m_face] - bb.maxCorner[m_face]) > eps) continue;
                bool match = true;
                for (size_t d = 0; d < _N; ++d) if (d != m_face && std::abs(p[d] - q[d]) > eps) match = false;
                if (match) return {i, j};
            }
        }
        throw std::runtime_error("Couldn't find periodic pair");
    }
private:
    size_t m_component, m_face;
};

// ---------------------------------------------------------------------------------------------
// .bc reader (BoundaryConditions.cc:28-45, 47-61, 217-388)
// ---------------------------------------------------------------------------------------------
namespace bc_detail {
inline Vector3D parseVectorLenient(const mjson::json &params) {
    Vector3D v;
    int n = 0;
    bool ok = params.is_array();
    if (ok)
        for (const auto &val : params) {
            if (!val.is_number()) { ok = false; break; }
            if (n < 3) v[n] = val.number();
            ++n;
        }
    if (!ok) n = -1;
    if (n != 2 && n != 3) throw std::runtime_error("Error parsing vector; read " + std::to_string(n) + " components");
    return v;
}
inline std::vector<std::string> parseExpressionVector(const mjson::json &params) {
    std::vector<std::string> result;
    for (const auto &val : params) {
        if (val.is_string()) result.push_back(val.str());
        else if (val.is_number()) result.push_back(val.dump());
        else throw std::runtime_error("Failed to parse expression vector");
    }
    return result;
}
}  // namespace bc_detail

template <size_t _N>
std::vector<CondPtr<_N>> readBoundaryConditions(std::istream &is, const BBox<VectorND<_N>> &bbox, bool &noRigidMotion,
                                                std::vector<PeriodicPairDirichletCondition<_N>> &pps,
                                                ComponentMask &pinTranslation) {
    using namespace bc_detail;
    typedef VectorND<_N> V;
    const mjson::json params = mjson::json::parse(is);
    std::vector<CondPtr<_N>> conds;
    noRigidMotion = params.value("no_rigid_motion", false);
    static const std::vector<std::string> componentStrings = {"x", "y", "z"};
    for (size_t c = 0; c < _N; ++c) {
        const std::string pairCondition("fix_periodic_pair_" + componentStrings[c]);
        if (params.count(pairCondition)) {
            const std::string faceSpecifier = params[pairCondition].str();
            size_t face = _N;
            for (size_t c2 = 0; c2 < _N; ++c2) {
                if (c2 == c) continue;
                if (faceSpecifier == componentStrings[c2]) face = c2;
            }
            if (face == _N) throw std::runtime_error("invalid " + pairCondition);
            pps.emplace_back(c, face);
        }
    }
    pinTranslation.setComponentString(params.value("pin_translation", ""));

    for (const auto &tcond : params["regions"]) {
        std::string type = tcond["type"].str();
        std::vector<size_t> node_indices;
        std::vector<V> node_values, element_values;
        std::vector<IVectorND<_N>> element_vertices;
        std::vector<UnorderedTriplet> element_corners;
        std::shared_ptr<Region<V>> region(new BBox<V>());
        V value;
        ExpressionVector exprVec;
        ComponentMask cmask("xyz");
        std::string prefix;
        if (type.substr(0, 9) == "dirichlet") { prefix = "dirichlet"; type = type.substr(9); }
        else if (type.substr(0, 6) == "target") { prefix = "target"; type = type.substr(6); }
        if (prefix.size()) {
            size_t len = 0;
            for (char ch : type) { if (ch < 'x' || ch > 'z') break; ++len; }
            if (len > 3) throw std::runtime_error("invalid mask");
            if (len > 0) cmask.setComponentString(type.substr(0, len));
            type = prefix + type.substr(len);
        }
        auto parseNodeValues = [&](const mjson::json &vals) {
            for (const auto &val : vals) {
                const Vector3D disp = parseVectorLenient(val[0]);
                for (const auto &nd : val[1]) {
                    if (!nd.is_number()) throw std::runtime_error("Error parsing node condition values.");
                    node_indices.push_back((size_t)nd.number());
                    node_values.push_back(truncateFrom3D<V>(disp));
                }
            }
        };
        auto parseElementValues = [&](const mjson::json &vals) {
            std::runtime_error err("Error parsing element condition values.");
            for (const auto &val : vals) {
                const Vector3D vec = parseVectorLenient(val[0]);
                for (const auto &elem : val[1]) {
                    std::vector<size_t> idx;
                    for (const auto &cidx : elem) { if (!cidx.is_number()) throw err; idx.push_back((size_t)cidx.number()); }
                    if (idx.size() == 2) idx.push_back(0);
                    if (idx.size() != 3) throw err;
                    element_values.push_back(truncateFrom3D<V>(vec));
                    element_corners.emplace_back((int)idx[0], (int)idx[1], (int)idx[2]);
                }
            }
        };
        if (type.find("nodes") != std::string::npos) parseNodeValues(tcond["values"]);
        else if (type == "traction elements" || type == "pressure elements" || type == "force elements") parseElementValues(tcond["values"]);
        else {
            if (tcond.count("box")) {
                region->minCorner = truncateFrom3D<V>(parseVectorLenient(tcond["box"]["minCorner"]));
                region->maxCorner = truncateFrom3D<V>(parseVectorLenient(tcond["box"]["maxCorner"]));
            } else if (tcond.count("box%")) {
                region->minCorner = bbox.interpolatePoint(truncateFrom3D<V>(parseVectorLenient(tcond["box%"]["minCorner"])));
                region->maxCorner = bbox.interpolatePoint(truncateFrom3D<V>(parseVectorLenient(tcond["box%"]["maxCorner"])));
            } else if (tcond.count("element vertices")) {
                for (const auto &val : tcond["element vertices"]) {
                    IVectorND<_N> corners;
                    size_t i = 0;
                    for (const auto &x : val) { if (i < _N) corners[i] = (int)x.number(); ++i; }
                    if (i != _N) throw std::runtime_error("Error parsing element vertices.");
                    element_vertices.push_back(corners);
                }
            } else if (tcond.count("path") || tcond.count("polygon")) {
                throw std::runtime_error("path/polygon regions are not supported by this build");
            }
            try {
                value = truncateFrom3D<V>(parseVectorLenient(tcond["value"]));
            } catch (...) {
                auto expressions = parseExpressionVector(tcond["value"]);
                if ((_N == 2) && (expressions.size() == 3) && (std::stod(expressions[2]) == 0)) expressions.pop_back();
                if (expressions.size() != _N) throw std::runtime_error("Incorrect expression vector size");
                for (const auto &expr : expressions) exprVec.add(expr);
            }
        }
        BoundaryCondition<_N> *c;
        if (exprVec.size() > 0) {
            if (type == "traction") c = new NeumannCondition<_N>(region, exprVec, NeumannType::Traction);
            else if (type == "dirichlet") c = new DirichletCondition<_N>(region, exprVec, cmask);
            else if (type == "dirichlet elements") c = new DirichletElementsCondition<_N>(element_vertices, exprVec, cmask);
            else if (type == "target") c = new TargetCondition<_N>(region, exprVec, cmask);
            else if (type == "delta force") c = new DeltaForceCondition<_N>(region, exprVec);
            else throw std::runtime_error("Only region-based traction, dirichlet, target, and delta force support expression vectors");
        } else {
            if (type == "pressure") c = new NeumannCondition<_N>(region, value[0]);
            else if (type == "traction") c = new NeumannCondition<_N>(region, value, NeumannType::Traction);
            else if (type == "force") c = new NeumannCondition<_N>(region, value, NeumannType::Force);
            else if (type == "dirichlet") c = new DirichletCondition<_N>(region, value, cmask);
            else if (type == "dirichlet elements") c = new DirichletElementsCondition<_N>(element_vertices, value, cmask);
            else if (type == "target") c = new TargetCondition<_N>(region, value, cmask);
            else if (type == "dirichlet nodes") c = new DirichletNodesCondition<_N>(node_indices, node_values, cmask);
            else if (type == "target nodes") c = new TargetNodesCondition<_N>(node_indices, node_values, cmask);
            else if (type == "traction elements") c = new NeumannElementsCondition<_N>(NeumannType::Traction, element_corners, element_values);
            else if (type == "pressure elements") c = new NeumannElementsCondition<_N>(NeumannType::Pressure, element_corners, element_values);
            else if (type == "force elements") c = new NeumannElementsCondition<_N>(NeumannType::Force, element_corners, element_values);
            else if (type == "delta force") c = new DeltaForceCondition<_N>(region, value);
            else if (type == "delta force nodes") c = new DeltaForceNodesCondition<_N>(node_indices, node_values);
            else if (type == "contact" || type == "fracture") throw std::runtime_error("'" + type + "' conditions are not supported by this build");
            else throw std::runtime_error("Invalid type '" + type + "'");
        }
        conds.push_back(CondPtr<_N>(c));
    }
    return conds;
}

template <size_t _N>
std::vector<CondPtr<_N>> readBoundaryConditions(const std::string &cpath, const BBox<VectorND<_N>> &bbox, bool &noRigidMotion,
                                                std::vector<PeriodicPairDirichletCondition<_N>> &pps,
                                                ComponentMask &pinTranslation) {
    std::ifstream inFile(cpath);
    if (!inFile.is_open()) throw std::runtime_error("Couldn't open BC file");
    return readBoundaryConditions<_N>(inFile, bbox, noRigidMotion, pps, pinTranslation);
}

template <size_t _N>
std::vector<CondPtr<_N>> readBoundaryConditions(const std::string &cpath, const BBox<VectorND<_N>> &bbox, bool &noRigidMotion) {
    std::vector<PeriodicPairDirichletCondition<_N>> pps;
    ComponentMask pin;
    auto result = readBoundaryConditions<_N>(cpath, bbox, noRigidMotion, pps, pin);
    if (pps.size()) throw std::runtime_error("Didn't expect PeriodicPairDirichletCondition");
    return result;
}

// ---------------------------------------------------------------------------------------------
// Periodic boundary matching (PeriodicBoundaryMatcher.hh:38-258) and PeriodicCondition
// (BoundaryConditions.hh:457-561)
// ---------------------------------------------------------------------------------------------
namespace PeriodicBoundaryMatcher {

template <size_t N>
struct FaceMembership {
    std::bitset<2 * N> membership;
    FaceMembership() {}
    template <class Point>
    FaceMembership(const Point &p, const BBox<VectorND<N>> &cell, Real epsilon = 1e-5) {
        for (size_t d = 0; d < N; ++d) {
            membership[d] = std::abs(p[d] - cell.minCorner[d]) <= epsilon;
            membership[N + d] = std::abs(p[d] - cell.maxCorner[d]) <= epsilon;
        }
    }
    static FaceMembership AllFaces() { FaceMembership r; r.membership.set(); return r; }
    bool onMinFace(size_t d) const { return membership[d]; }
    bool onMaxFace(size_t d) const { return membership[N + d]; }
    bool onMinOrMaxFace(size_t d) const { return onMinFace(d) || onMaxFace(d); }
    size_t count() const { return membership.count(); }
    bool onAnyMaxFace() const { return (membership >> N).any(); }
    bool isMinimalNode() const { return !onAnyMaxFace(); }
    FaceMembership &operator&=(const FaceMembership &b) { membership &= b.membership; return *this; }
};

// hashed grid with cell size max(eps, 1e-7) (CollisionGrid.hh:60-90): exact-cell + neighbour lookup
template <size_t N>
class CollisionGrid {
public:
    explicit CollisionGrid(Real cellSize) : m_cellSize(cellSize) {}
    void addPoint(const VectorND<N> &p, size_t idx) { m_cells[key(cell(p))].push_back({p, idx}); }
    // closest stored point within eps, or (-1, inf)
    std::pair<long, Real> getClosestPoint(const VectorND<N> &q, Real eps) const {
        long best = -1;
        Real bestDist = std::numeric_limits<Real>::max();
        const auto c = cell(q);
        std::array<long, N> o;
        const int total = N == 3 ? 27 : 9;
        for (int n = 0; n < total; ++n) {
            int r = n;
            for (size_t d = 0; d < N; ++d) { o[d] = c[d] + (r % 3) - 1; r /= 3; }
            auto it = m_cells.find(key(o));
            if (it == m_cells.end()) continue;
            for (const auto &e : it->second) {
                const Real dist = (e.p - q).norm();
                if (dist <= eps && dist < bestDist) { bestDist = dist; best = (long)e.idx; }
            }
        }
        return {best, bestDist};
    }
private:
    struct Entry { VectorND<N> p; size_t idx; };
    Real m_cellSize;
    std::unordered_map<uint64_t, std::vector<Entry>> m_cells;
    std::array<long, N> cell(const VectorND<N> &p) const {
        std::array<long, N> c;
        for (size_t d = 0; d < N; ++d) c[d] = (long)std::floor(p[d] / m_cellSize);
        return c;
    }
    static uint64_t key(const std::array<long, N> &c) {
        uint64_t h = 1469598103934665603ULL;
        for (size_t d = 0; d < N; ++d) { h ^= (uint64_t)c[d] + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2); h *= 1099511628211ULL; }
        return h;
    }
};

static constexpr size_t NONE = std::numeric_limits<size_t>::max();

// PeriodicBoundaryMatcher::match (:149-258): one identified node set per "minimal" node
template <size_t N>
void match(const std::vector<VectorND<N>> &bdryPoints, const BBox<VectorND<N>> &cell,
           const std::vector<FaceMembership<N>> &faceMembership, std::vector<std::vector<size_t>> &nodeSets,
           std::vector<size_t> &nodeSetForNode, Real epsilon = 1e-7) {
    CollisionGrid<N> cgrid(std::max(epsilon, 1.0e-7));
    for (size_t i = 0; i < bdryPoints.size(); ++i)
        if (!faceMembership[i].isMinimalNode()) cgrid.addPoint(bdryPoints[i], i);
    nodeSetForNode.assign(bdryPoints.size(), NONE);
    nodeSets.clear();
    for (size_t i = 0; i < bdryPoints.size(); ++i) {
        const auto &fm = faceMembership[i];
        if (!fm.isMinimalNode()) continue;
        nodeSetForNode[i] = nodeSets.size();
        const size_t numPeriodicFaces = fm.count();
        const size_t numIdentifiedNodes = size_t(1) << numPeriodicFaces;
        nodeSets.push_back(std::vector<size_t>(numIdentifiedNodes, NONE));
        auto &ns = nodeSets.back();
        ns[0] = i;
        for (size_t n = 1; n < numIdentifiedNodes; ++n) {
            auto query = bdryPoints[i];
            size_t idx = 0;
            for (size_t d = 0; d < N; ++d)
                if (fm.onMinFace(d))
                    if (n & (size_t(1) << idx++)) query[d] = cell.maxCorner[d];
            auto result = cgrid.getClosestPoint(query, epsilon);
            if (result.first < 0) {
                std::stringstream ss;
                ss << "Couldn't find " << n << "th periodic-identified node for minimal boundary node " << i;
                throw std::runtime_error(ss.str());
            }
            const size_t pair = (size_t)result.first;
            if (nodeSetForNode.at(pair) != NONE) throw std::runtime_error("Non bijective node set assignment.");
            nodeSetForNode.at(pair) = nodeSetForNode[i];
            ns[n] = pair;
        }
    }
    for (size_t i = 0; i < bdryPoints.size(); ++i)
        if (nodeSetForNode[i] == NONE) throw std::runtime_error("Unmatched non-minimal boundary node " + std::to_string(i));
}

// PeriodicBoundaryMatcher::matchPermittingMismatch (:268-372): for regular voxel grids whose opposite faces do not
// carry the same nodes.  Per axis, every max-face node is paired with the min-face node at its projected position
// if there is one; the identified node sets are the connected components of that pairing graph, created in
// boundary-node order.  Unpaired face memberships are counted and reported, not errors.
template <size_t N>
void matchPermittingMismatch(const std::vector<VectorND<N>> &bdryPoints, const BBox<VectorND<N>> &cell,
                             const std::vector<FaceMembership<N>> &faceMembership, std::vector<std::vector<size_t>> &nodeSets,
                             std::vector<size_t> &nodeSetForNode, Real epsilon = 1e-7) {
    const size_t numBdryPts = bdryPoints.size();
    std::vector<std::array<size_t, N>> pair(numBdryPts);
    for (auto &p : pair) p.fill(NONE);
    for (size_t d = 0; d < N; ++d) {
        CollisionGrid<N> cgrid(std::max(epsilon, 1.0e-7));
        for (size_t i = 0; i < numBdryPts; ++i)
            if (faceMembership[i].onMinFace(d)) cgrid.addPoint(bdryPoints[i], i);
        for (size_t i = 0; i < numBdryPts; ++i) {
            if (!faceMembership[i].onMaxFace(d)) continue;
            auto query = bdryPoints[i];
            query[d] = cell.minCorner[d];
            const auto result = cgrid.getClosestPoint(query, epsilon);
            if (result.first < 0) continue;                     // mismatch
            const size_t pi = (size_t)result.first;
            if (pair[i][d] != NONE || pair[pi][d] != NONE) throw std::runtime_error("Non-bijective boundary matching");
            pair[i][d] = pi;
            pair[pi][d] = i;
        }
    }
    nodeSetForNode.assign(numBdryPts, NONE);
    nodeSets.clear();
    size_t numMismatches = 0;
    std::vector<size_t> queue;
    for (size_t i = 0; i < numBdryPts; ++i) {
        if (nodeSetForNode[i] != NONE) continue;
        const size_t nsi = nodeSets.size();
        nodeSetForNode[i] = nsi;
        nodeSets.emplace_back(1, i);
        queue.assign(1, i);
        for (size_t head = 0; head < queue.size(); ++head) {
            const size_t u = queue[head];
            for (size_t d = 0; d < N; ++d) {
                if (!faceMembership[u].onMinOrMaxFace(d)) continue;
                const size_t v = pair[u][d];
                if (v == NONE) { ++numMismatches; continue; }
                if (nodeSetForNode[v] != NONE) {
                    if (nodeSetForNode[v] != nsi) throw std::runtime_error("node set conflict in periodic matching");
                    continue;
                }
                nodeSetForNode[v] = nsi;
                nodeSets[nsi].push_back(v);
                queue.push_back(v);
            }
        }
    }
    if (numMismatches > 0)
        std::cerr << "WARNING: detected " << numMismatches << " mismatches in periodic node identification" << std::endl;
}

}  // namespace PeriodicBoundaryMatcher

template <size_t _N>
class PeriodicCondition {
public:
    static constexpr size_t NO_DOF = std::numeric_limits<size_t>::max();
    // ignoreDims: axes along which the cell is NOT periodic (BoundaryConditions.hh:457-505 of the reference)
    template <typename Mesh>
    PeriodicCondition(const Mesh &mesh, Real epsilon = 1e-7, bool ignoreMismatch = false,
                      const std::vector<size_t> &ignoreDims = std::vector<size_t>()) : m_ignoreDims(ignoreDims) {
        using namespace PeriodicBoundaryMatcher;
        const BBox<VectorND<_N>> cell = mesh.boundingBox();
        std::vector<VectorND<_N>> bdryPts;
        bdryPts.reserve(mesh.numBoundaryNodes());
        for (size_t bn = 0; bn < mesh.numBoundaryNodes(); ++bn) bdryPts.push_back(mesh.nodePosition(mesh.volumeNodeForBoundaryNode(bn)));
        std::vector<FaceMembership<_N>> fm;
        fm.reserve(bdryPts.size());
        for (const auto &p : bdryPts) fm.emplace_back(p, cell, epsilon);
        if (!ignoreDims.empty()) {
            // nodes on a periodic face lose their membership of the ignored faces; all others lose every membership
            std::vector<size_t> periodicDims;
            for (size_t d = 0; d < _N; ++d)
                if (std::find(ignoreDims.begin(), ignoreDims.end(), d) == ignoreDims.end()) periodicDims.push_back(d);
            for (auto &m : fm) {
                bool onSignificantDim = false;
                for (size_t d : periodicDims) onSignificantDim = onSignificantDim || m.onMinOrMaxFace(d);
                for (size_t d = 0; d < _N; ++d) {
                    const bool ignored = std::find(ignoreDims.begin(), ignoreDims.end(), d) != ignoreDims.end();
                    if (!onSignificantDim || ignored) { m.membership[d] = false; m.membership[d + _N] = false; }
                }
            }
        }
        std::vector<std::vector<size_t>> bdryNodeSets;
        std::vector<size_t> bdryNodeSetForBdryNode;
        if (ignoreMismatch) matchPermittingMismatch<_N>(bdryPts, cell, fm, bdryNodeSets, bdryNodeSetForBdryNode, epsilon);
        else match<_N>(bdryPts, cell, fm, bdryNodeSets, bdryNodeSetForBdryNode, epsilon);

        // boundary elements whose nodes all share one cell face (determineCellFaceBoundaryElements :126-146)
        m_isPeriodicBoundaryElement.assign(mesh.numBoundaryElements(), false);
        for (size_t be = 0; be < mesh.numBoundaryElements(); ++be) {
            auto pb = FaceMembership<_N>::AllFaces();
            for (size_t n = 0; n < Mesh::nodesPerBoundaryElement; ++n)
                pb &= fm.at(mesh.boundaryNodeForVolumeNode(mesh.boundaryElementVolumeNode(be, n)));
            if (pb.count() > 1) throw std::runtime_error("Boundary element on more than one cell face.");
            m_isPeriodicBoundaryElement[be] = pb.count() > 0;
        }
        // DoF ids in node order, each new boundary node pulling in its identified copies (:533-554)
        m_dofForNode.assign(mesh.numNodes(), NO_DOF);
        m_nodesForDoF.clear();
        for (size_t n = 0; n < mesh.numNodes(); ++n) {
            if (m_dofForNode[n] != NO_DOF) continue;
            const int bn = mesh.boundaryNodeForVolumeNode(n);
            if (bn >= 0) {
                auto &ns = bdryNodeSets[bdryNodeSetForBdryNode[bn]];
                for (size_t &ni : ns) {
                    ni = mesh.volumeNodeForBoundaryNode(ni);
                    m_dofForNode[ni] = m_nodesForDoF.size();
                }
                m_nodesForDoF.emplace_back(std::move(ns));
            } else {
                m_dofForNode[n] = m_nodesForDoF.size();
                m_nodesForDoF.emplace_back(1, n);
            }
        }
    }
    // Identified node pairs read from a file, one "a b" pair of volume node indices per line (:563-610; the
    // reference calls this format a temporary hack): DoFs are the connected components of the pair graph, in order
    // of their lowest node.  No boundary element is marked periodic.
    template <typename Mesh>
    PeriodicCondition(const Mesh &mesh, const std::string &pcFile) {
        std::cerr << "WARNING: periodic boundary condition files are a temporary hack." << std::endl;
        std::ifstream file(pcFile);
        if (!file.is_open()) throw std::runtime_error("Couldn't open " + pcFile);
        std::vector<std::vector<size_t>> adj(mesh.numNodes());
        std::string line;
        while (std::getline(file, line)) {
            std::istringstream ls(line);
            size_t a, b;
            if (!(ls >> a >> b)) continue;
            if (a >= mesh.numNodes() || b >= mesh.numNodes()) throw std::runtime_error("Periodic pair node index out of bounds in " + pcFile);
            adj[a].push_back(b);
            adj[b].push_back(a);
        }
        m_dofForNode.assign(mesh.numNodes(), NO_DOF);
        m_nodesForDoF.clear();
        std::vector<size_t> queue;
        for (size_t n = 0; n < mesh.numNodes(); ++n) {
            if (m_dofForNode[n] != NO_DOF) continue;
            const size_t dof = m_nodesForDoF.size();
            m_dofForNode[n] = dof;
            queue.assign(1, n);
            for (size_t head = 0; head < queue.size(); ++head)
                for (size_t v : adj[queue[head]]) {
                    if (m_dofForNode[v] != NO_DOF) continue;
                    m_dofForNode[v] = dof;
                    queue.push_back(v);
                }
            m_nodesForDoF.push_back(queue);
        }
        m_isPeriodicBoundaryElement.assign(mesh.numBoundaryElements(), false);
    }
    const std::vector<size_t> &getIgnoreDims() const { return m_ignoreDims; }
    const std::vector<size_t> &periodicDoFsForNodes() const { return m_dofForNode; }
    size_t numPeriodicDoFs() const { return m_nodesForDoF.size(); }
    bool isPeriodicBE(size_t be) const { return m_isPeriodicBoundaryElement.at(be); }
    bool isPeriodicNode(size_t vni) const { return identifiedNodes(vni).size() > 1; }
    const std::vector<size_t> &identifiedNodes(size_t vni) const { return m_nodesForDoF.at(m_dofForNode.at(vni)); }

private:
    std::vector<size_t> m_dofForNode;
    std::vector<std::vector<size_t>> m_nodesForDoF;
    std::vector<bool> m_isPeriodicBoundaryElement;
    std::vector<size_t> m_ignoreDims;
};
#endif

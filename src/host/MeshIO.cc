// MeshIO implementation (see include/MeshFEM/MeshIO.hh).  Format semantics follow
// src/lib/MeshFEM/MeshIO.cc of the reference: :527-531 element table, :533-616 MSH writer,
// :625-760 MSH reader (consecutive 1-based node ids, one element type per file, 8-byte reals).
#include <MeshFEM/MeshIO.hh>

#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <stdexcept>

namespace MeshIO {

static const MSHElementInfo kInfo[] = {{MESH_TRI, 2, 3},      {MESH_TET, 4, 4},       {MESH_QUAD, 3, 4},
                                       {MESH_HEX, 5, 8},      {MESH_TRI_DEG2, 9, 6},  {MESH_TET_DEG2, 11, 10},
                                       {MESH_LINE, 1, 2},     {MESH_LINE_DEG2, 8, 3}};

MSHElementInfo mshInfoForMeshType(MeshType t) {
    for (const auto &i : kInfo) if (i.meshType == t) return i;
    throw std::runtime_error("Unsupported MSH mesh type");
}
MSHElementInfo mshInfoForElementType(int et) {
    for (const auto &i : kInfo) if (i.elementType == et) return i;
    throw std::runtime_error("Unsupported MSH element type " + std::to_string(et));
}
MSHElementInfo mshInfoForNodeCount(size_t n) {
    for (const auto &i : kInfo) if (i.nodesPerElem == n) return i;
    throw std::runtime_error("Unsupported MSH element size " + std::to_string(n));
}

size_t meshDegree(MeshType t) {
    return (t == MESH_LINE_DEG2 || t == MESH_TRI_DEG2 || t == MESH_TET_DEG2) ? 2 : 1;
}
size_t meshDimension(MeshType t) {
    switch (t) {
        case MESH_LINE: case MESH_LINE_DEG2: return 1;
        case MESH_TRI: case MESH_QUAD: case MESH_TRI_DEG2: case MESH_TRI_QUAD: return 2;
        case MESH_TET: case MESH_HEX: case MESH_TET_DEG2: return 3;
        default: throw std::runtime_error("Invalid mesh type");
    }
}

Format guessFormat(const std::string &path) {
    auto pos = path.rfind('.');
    if (pos == std::string::npos) return FMT_INVALID;
    std::string ext = path.substr(pos);
    std::transform(ext.begin(), ext.end(), ext.begin(), ::tolower);
    if (ext == ".off") return FMT_OFF;
    if (ext == ".obj") return FMT_OBJ;
    if (ext == ".msh") return FMT_MSH;
    if (ext == ".mesh") return FMT_MEDIT;
    return FMT_INVALID;
}

// Next non-blank, non-comment line (MeshIO.cc getDataLine)
static bool getDataLine(std::istream &is, std::string &line) {
    while (std::getline(is, line)) {
        if (!line.empty() && line.back() == '\r') line.pop_back();
        size_t b = line.find_first_not_of(" \t");
        if (b == std::string::npos) continue;
        if (line[b] == '#') continue;
        size_t e = line.find_last_not_of(" \t");
        line = line.substr(b, e - b + 1);
        return true;
    }
    return false;
}

static void skipNewline(std::istream &is) {
    char c;
    is.read(&c, 1);
    if (c != '\n') throw std::runtime_error("Newline expected, got ascii " + std::to_string(int(c)) + " instead");
}

static MeshType guessType(const std::vector<IOElement> &elements) {
    if (elements.empty()) return MESH_INVALID;
    size_t n = elements[0].size();
    bool same = true, triquad = true;
    for (const auto &e : elements) {
        if (e.size() != n) same = false;
        if (e.size() != 3 && e.size() != 4) triquad = false;
    }
    if (same) {
        if (n == 3) return MESH_TRI;
        if (n == 4) return MESH_TET;     // reference convention: 4 nodes => tet
        if (n == 8) return MESH_HEX;
        if (n == 2) return MESH_LINE;
        if (n == 6) return MESH_TRI_DEG2;
        if (n == 10) return MESH_TET_DEG2;
    }
    if (triquad) return MESH_TRI_QUAD;
    return MESH_INVALID;
}

static MeshType loadMSH(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements,
                        MeshType type) {
    MSHElementInfo ei{MESH_INVALID, -1, 0};
    if (type != MESH_GUESS) ei = mshInfoForMeshType(type);
    std::runtime_error badFmt("Bad MSH file format"), unsFmt("Unsupported MSH file format");
    std::string line;
    if (!getDataLine(is, line) || line != "$MeshFormat") throw badFmt;
    double version; int file_type, data_size;
    is >> version >> file_type >> data_size;
    if (size_t(file_type) > 1 || data_size != (int)sizeof(double)) throw unsFmt;
    const bool binary = file_type == 1;
    if (binary) {
        skipNewline(is);
        int one; is.read((char *)&one, sizeof(int));
        if (one != 1) throw unsFmt;
    }
    if (!getDataLine(is, line) || line != "$EndMeshFormat") throw badFmt;
    if (!getDataLine(is, line) || line != "$Nodes") throw badFmt;
    size_t numNodes; is >> numNodes;
    nodes.resize(numNodes);
    if (binary) {
        skipNewline(is);
        int idx = 0;
        for (size_t i = 0; i < numNodes; ++i) {
            int newIdx; is.read((char *)&newIdx, sizeof(int));
            if (newIdx != ++idx) throw unsFmt;
            double v[3]; is.read((char *)v, sizeof(v));
            if (is.fail()) throw badFmt;
            nodes[i].set(v[0], v[1], v[2]);
        }
    } else {
        int idx = 0;
        for (size_t i = 0; i < numNodes; ++i) {
            if (!getDataLine(is, line)) throw badFmt;
            std::istringstream iss(line);
            int newIdx; iss >> newIdx;
            if (newIdx != ++idx) throw unsFmt;
            iss >> nodes[i][0] >> nodes[i][1] >> nodes[i][2];
            if (iss.fail()) throw badFmt;
        }
    }
    if (!getDataLine(is, line) || line != "$EndNodes") throw badFmt;
    if (!getDataLine(is, line) || line != "$Elements") throw badFmt;
    size_t numElements; is >> numElements;
    elements.assign(numElements, IOElement());
    if (binary) {
        skipNewline(is);
        size_t readElements = 0;
        std::vector<int> data;
        while (readElements < numElements) {
            int header[3]; is.read((char *)header, sizeof(header));
            if (ei.elementType == -1) ei = mshInfoForElementType(header[0]);
            if (header[0] != ei.elementType) throw badFmt;
            const size_t newSize = readElements + header[1];
            if (newSize > numElements) throw badFmt;
            const int intCount = 1 + header[2] + (int)ei.nodesPerElem;
            data.resize(intCount);
            for (size_t e = readElements; e < newSize; ++e) {
                is.read((char *)data.data(), intCount * sizeof(int));
                elements[e].resize(ei.nodesPerElem);
                for (size_t c = 0; c < ei.nodesPerElem; ++c) elements[e][c] = data[1 + header[2] + c] - 1;
            }
            readElements = newSize;   // (the reference's `+= newSize`, MeshIO.cc:721, is only right for one block)
            if (!is) throw badFmt;
        }
    } else {
        for (size_t i = 0; i < numElements; ++i) {
            if (!getDataLine(is, line)) throw badFmt;
            std::istringstream iss(line);
            int idx, etype; size_t numTags;
            iss >> idx >> etype >> numTags;
            while (numTags-- > 0) { int dummy; iss >> dummy; }
            if (ei.elementType == -1) ei = mshInfoForElementType(etype);
            if (etype != ei.elementType) throw badFmt;
            elements[i].resize(ei.nodesPerElem);
            for (size_t c = 0; c < ei.nodesPerElem; ++c) { iss >> idx; elements[i][c] = idx - 1; }
            if (iss.fail()) throw badFmt;
        }
    }
    if (!getDataLine(is, line) || line != "$EndElements") throw badFmt;
    if (numElements == 0) return type == MESH_GUESS ? MESH_INVALID : type;
    return ei.meshType;
}

static void saveMSH(std::ostream &os, const std::vector<IOVertex> &nodes, const std::vector<IOElement> &elements,
                    MeshType type, bool binary) {
    if (nodes.empty()) throw std::runtime_error("Empty mesh.");
    if (elements.empty() && type == MESH_GUESS) type = MESH_TRI;
    MSHElementInfo ei = (type == MESH_GUESS) ? mshInfoForNodeCount(elements.back().size()) : mshInfoForMeshType(type);
    os << "$MeshFormat\n" << 2.2 << " " << (binary ? 1 : 0) << " " << sizeof(double) << '\n';
    if (binary) { int one = 1; os.write((char *)&one, sizeof(int)); os << '\n'; }
    os << "$EndMeshFormat\n$Nodes\n" << nodes.size() << '\n';
    if (binary) {
        for (size_t i = 1; i <= nodes.size(); ++i) {
            int id = (int)i; os.write((char *)&id, sizeof(int));
            double xyz[3] = {nodes[i - 1][0], nodes[i - 1][1], nodes[i - 1][2]};
            os.write((char *)xyz, sizeof(xyz));
        }
        os << '\n';
    } else {
        os << std::setprecision(17);
        for (size_t i = 0; i < nodes.size(); ++i)
            os << i + 1 << " " << nodes[i][0] << " " << nodes[i][1] << " " << nodes[i][2] << '\n';
    }
    os << "$EndNodes\n$Elements\n" << elements.size() << '\n';
    if (binary) {
        if (!elements.empty()) {
            int hdr[3] = {ei.elementType, (int)elements.size(), 0};
            os.write((char *)hdr, sizeof(hdr));
        }
        for (size_t i = 1; i <= elements.size(); ++i) {
            int id = (int)i; os.write((char *)&id, sizeof(int));
            if (elements[i - 1].size() != ei.nodesPerElem) throw std::runtime_error("Illegal sized element");
            for (size_t c = 0; c < ei.nodesPerElem; ++c) { int ci = (int)(elements[i - 1][c] + 1); os.write((char *)&ci, sizeof(int)); }
        }
        os << '\n';
    } else {
        for (size_t i = 0; i < elements.size(); ++i) {
            if (elements[i].size() != ei.nodesPerElem) throw std::runtime_error("Illegal sized element");
            os << i + 1 << " " << ei.elementType << " " << 0;
            for (size_t c = 0; c < ei.nodesPerElem; ++c) os << " " << elements[i][c] + 1;
            os << '\n';
        }
    }
    os << "$EndElements\n";
    os.flush();
}

static MeshType loadOFF(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements) {
    std::string line;
    if (!getDataLine(is, line) || line.substr(0, 3) != "OFF") throw std::runtime_error("Bad OFF header");
    std::string rest = line.substr(3);
    size_t nv = 0, nf = 0, ne = 0;
    {
        std::istringstream iss(rest);
        if (!(iss >> nv >> nf)) {
            if (!getDataLine(is, line)) throw std::runtime_error("Bad OFF header");
            std::istringstream iss2(line);
            if (!(iss2 >> nv >> nf)) throw std::runtime_error("Bad OFF header");
            iss2 >> ne;
        }
    }
    nodes.resize(nv); elements.assign(nf, IOElement());
    for (size_t i = 0; i < nv; ++i) {
        if (!getDataLine(is, line)) throw std::runtime_error("Bad OFF vertex");
        std::istringstream iss(line);
        iss >> nodes[i][0] >> nodes[i][1] >> nodes[i][2];
        if (iss.fail()) throw std::runtime_error("Bad OFF vertex");
    }
    for (size_t i = 0; i < nf; ++i) {
        if (!getDataLine(is, line)) throw std::runtime_error("Bad OFF face");
        std::istringstream iss(line);
        size_t n; iss >> n;
        elements[i].resize(n);
        for (size_t c = 0; c < n; ++c) iss >> elements[i][c];
        if (iss.fail()) throw std::runtime_error("Bad OFF face");
    }
    // OFF faces are polygons: a 4-gon is a quad, never a tet
    MeshType t = guessType(elements);
    if (t == MESH_TET) t = MESH_QUAD;
    return t;
}

static MeshType loadOBJ(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements) {
    std::string line;
    nodes.clear(); elements.clear();
    while (getDataLine(is, line)) {
        std::istringstream iss(line);
        std::string tok; iss >> tok;
        if (tok == "v") { IOVertex v; iss >> v[0] >> v[1] >> v[2]; nodes.push_back(v); }
        else if (tok == "f") {
            IOElement e; std::string item;
            while (iss >> item) e.push_back(std::stoul(item.substr(0, item.find('/'))) - 1);
            elements.push_back(e);
        }
    }
    MeshType t = guessType(elements);
    if (t == MESH_TET) t = MESH_QUAD;
    return t;
}

static MeshType loadMedit(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements) {
    std::string tok;
    nodes.clear(); elements.clear();
    while (is >> tok) {
        if (tok == "Vertices") {
            size_t n; is >> n; nodes.resize(n);
            for (size_t i = 0; i < n; ++i) { int ref; is >> nodes[i][0] >> nodes[i][1] >> nodes[i][2] >> ref; }
        } else if (tok == "Tetrahedra") {
            size_t n; is >> n; elements.assign(n, IOElement(4));
            for (size_t i = 0; i < n; ++i) {
                int ref; size_t a, b, c, d; is >> a >> b >> c >> d >> ref;
                elements[i][0] = a - 1; elements[i][1] = b - 1; elements[i][2] = c - 1; elements[i][3] = d - 1;
            }
        } else if (tok == "End") break;
    }
    if (elements.empty()) throw std::runtime_error("Only support linear tets.");
    return MESH_TET;
}

MeshType load(std::istream &is, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements, Format format,
              MeshType type) {
    switch (format) {
        case FMT_MSH: case FMT_MSH_ASCII: return loadMSH(is, nodes, elements, type);
        case FMT_OFF: return loadOFF(is, nodes, elements);
        case FMT_OBJ: return loadOBJ(is, nodes, elements);
        case FMT_MEDIT: return loadMedit(is, nodes, elements);
        default: throw std::runtime_error("Unsupported mesh format");
    }
}

MeshType load(const std::string &path, std::vector<IOVertex> &nodes, std::vector<IOElement> &elements, Format format,
              MeshType type) {
    if (format == FMT_GUESS) format = guessFormat(path);
    if (format == FMT_INVALID) throw std::runtime_error("Unrecognized mesh file extension: " + path);
    std::ifstream is(path, std::ios::binary);
    if (!is.is_open()) throw std::runtime_error("Couldn't open input file " + path);
    return load(is, nodes, elements, format, type);
}

void save(std::ostream &os, const std::vector<IOVertex> &nodes, const std::vector<IOElement> &elements,
          Format format, MeshType type) {
    switch (format) {
        case FMT_MSH: saveMSH(os, nodes, elements, type, true); break;
        case FMT_MSH_ASCII: saveMSH(os, nodes, elements, type, false); break;
        case FMT_OFF: {
            os << "OFF\n" << nodes.size() << " " << elements.size() << " 0\n" << std::setprecision(17);
            for (const auto &v : nodes) os << v[0] << " " << v[1] << " " << v[2] << '\n';
            for (const auto &e : elements) { os << e.size(); for (size_t c : e) os << " " << c; os << '\n'; }
            break;
        }
        default: throw std::runtime_error("Unsupported output mesh format");
    }
}

void save(const std::string &path, const std::vector<IOVertex> &nodes, const std::vector<IOElement> &elements,
          Format format, MeshType type) {
    if (format == FMT_GUESS) format = guessFormat(path);
    if (format == FMT_INVALID) throw std::runtime_error("Unrecognized mesh file extension: " + path);
    std::ofstream os(path, std::ios::binary);
    if (!os.is_open()) throw std::runtime_error("Couldn't open output file " + path);
    save(os, nodes, elements, format, type);
}

}  // namespace MeshIO

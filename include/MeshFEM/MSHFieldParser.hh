// MSHFieldParser -- reads the mesh plus the $NodeData / $ElementData sections of a Gmsh 2.2
// .msh (ascii or binary), the inverse of MSHFieldWriter.  Mirrors the query surface of the
// reference's MSHFieldParser (MSHFieldParser.hh:33-130, MSHFieldParser.cc:60-264) that the CLIs on
// the assemble-and-solve path use: scalarField / vectorField / symmetricMatrixField by name and
// domain type, "Field query unmatched." when absent.  $ElementNodeData sections are skipped
// (per-element interpolant fields are not consumed anywhere on this path).
#ifndef MESHFEM_B200_MSHFIELDPARSER_HH
#define MESHFEM_B200_MSHFIELDPARSER_HH
#include <MeshFEM/Fields.hh>
#include <MeshFEM/MeshIO.hh>

#include <fstream>
#include <map>
#include <sstream>

template <size_t N>
class MSHFieldParser {
public:
    typedef VectorField<Real, N> VField;
    typedef ScalarField<Real> SField;
    typedef SymmetricMatrixField<Real, N> SMField;

    explicit MSHFieldParser(const std::string &mshPath, bool permitDimMismatch = false) {
        std::ifstream is(mshPath, std::ios::binary);
        if (!is.is_open()) throw std::runtime_error("Couldn't open input file " + mshPath);
        std::string l1, l2;
        std::getline(is, l1);
        std::getline(is, l2);
        bool binary = false;
        {
            std::istringstream fmt(l2);
            double version = 0;
            int fileType = 0;
            fmt >> version >> fileType;
            binary = fileType == 1;
        }
        is.clear();
        is.seekg(0);
        m_type = MeshIO::load(is, m_vertices, m_elements, MeshIO::FMT_MSH);
        if (!permitDimMismatch && MeshIO::meshDimension(m_type) != N) throw std::runtime_error("Mesh/parser dimension mismatch.");
        m_parseFields(is, binary);
    }

    const std::vector<MeshIO::IOElement> &elements() const { return m_elements; }
    const std::vector<MeshIO::IOVertex> &vertices() const { return m_vertices; }
    MeshIO::MeshType meshType() const { return m_type; }
    size_t meshDegree() const { return MeshIO::meshDegree(m_type); }
    size_t meshDimension() const { return MeshIO::meshDimension(m_type); }
    size_t numElements() const { return m_elements.size(); }
    size_t numVertices() const { return m_vertices.size(); }

    const SField &scalarField(const std::string &name, DomainType reqType = DomainType::ANY) const { return m_getField(m_scalarFields, name, reqType); }
    const VField &vectorField(const std::string &name, DomainType reqType = DomainType::ANY) const { return m_getField(m_vectorFields, name, reqType); }
    const SMField &symmetricMatrixField(const std::string &name, DomainType reqType = DomainType::ANY) const { return m_getField(m_symmetricMatrixFields, name, reqType); }
    const SField &scalarField(const std::string &name, DomainType reqType, DomainType &actualType) const { actualType = reqType; return m_getField(m_scalarFields, name, actualType); }
    const VField &vectorField(const std::string &name, DomainType reqType, DomainType &actualType) const { actualType = reqType; return m_getField(m_vectorFields, name, actualType); }

    std::vector<std::string> scalarFieldNames(DomainType type = DomainType::ANY) const { return m_keys(m_scalarFields, type); }
    std::vector<std::string> vectorFieldNames(DomainType type = DomainType::ANY) const { return m_keys(m_vectorFields, type); }
    std::vector<std::string> symmetricMatrixFieldNames(DomainType type = DomainType::ANY) const { return m_keys(m_symmetricMatrixFields, type); }

private:
    std::vector<MeshIO::IOElement> m_elements;
    std::vector<MeshIO::IOVertex> m_vertices;
    MeshIO::MeshType m_type;
    std::map<std::string, std::pair<DomainType, VField>> m_vectorFields;
    std::map<std::string, std::pair<DomainType, SField>> m_scalarFields;
    std::map<std::string, std::pair<DomainType, SMField>> m_symmetricMatrixFields;

    template <class F>
    static const F &m_getField(const std::map<std::string, std::pair<DomainType, F>> &fields, const std::string &name, DomainType &type) {
        auto it = fields.find(name);
        if (it != fields.end() && (type == DomainType::ANY || it->second.first == type)) {
            type = it->second.first;
            return it->second.second;
        }
        throw std::runtime_error("Field query unmatched.");
    }
    template <class F>
    static std::vector<std::string> m_keys(const std::map<std::string, std::pair<DomainType, F>> &fields, DomainType type) {
        std::vector<std::string> r;
        for (const auto &kv : fields) if (type == DomainType::ANY || kv.second.first == type) r.push_back(kv.first);
        return r;
    }

    static bool m_nextLine(std::istream &is, std::string &line) {
        while (std::getline(is, line)) {
            while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
            size_t b = 0;
            while (b < line.size() && (line[b] == ' ' || line[b] == '\t')) ++b;
            line = line.substr(b);
            if (!line.empty()) return true;
        }
        return false;
    }
    static long m_intLine(std::istream &is) {
        std::string l;
        if (!m_nextLine(is, l)) throw std::runtime_error("Bad MSH field format");
        return std::stol(l);
    }

    void m_parseFields(std::istream &is, bool binary) {
        const std::runtime_error badFMT("Bad MSH field format");
        std::string header;
        while (m_nextLine(is, header)) {
            DomainType type;
            size_t expected;
            std::string footer;
            bool elementNodeData = false;
            if (header == "$ElementData") { type = DomainType::PER_ELEMENT; expected = numElements(); footer = "$EndElementData"; }
            else if (header == "$NodeData") { type = DomainType::PER_NODE; expected = numVertices(); footer = "$EndNodeData"; }
            else if (header == "$ElementNodeData") { type = DomainType::PER_ELEMENT; expected = numElements(); footer = "$EndElementNodeData"; elementNodeData = true; }
            else throw std::runtime_error("Unrecognized MSH section: " + header);
            if (m_intLine(is) != 1) throw badFMT;
            std::string name;
            if (!m_nextLine(is, name) || name.size() < 3 || name.front() != '"' || name.back() != '"') throw badFMT;
            name = name.substr(1, name.size() - 2);
            const long nRealTags = m_intLine(is);
            for (long i = 0; i < nRealTags; ++i) { std::string l; m_nextLine(is, l); }
            if (m_intLine(is) != 3) throw badFMT;
            m_intLine(is);   // timestep
            size_t dim = (size_t)m_intLine(is);
            const size_t numEntries = (size_t)m_intLine(is);
            if (numEntries != expected) throw std::runtime_error("Illegal number of field values");
            const size_t npe = m_elements.empty() ? 0 : m_elements[0].size();
            if (elementNodeData) dim *= npe;
            std::vector<double> data(dim * numEntries);
            for (size_t i = 0; i < numEntries; ++i) {
                if (binary) {
                    int idx, n = 1;
                    is.read((char *)&idx, sizeof(int));
                    if (elementNodeData) { is.read((char *)&n, sizeof(int)); if ((size_t)n != npe) throw std::runtime_error("Unexpected number-of-nodes-per-element"); }
                    is.read((char *)&data[i * dim], dim * sizeof(double));
                    if (!is) throw badFMT;
                } else {
                    std::string l;
                    if (!m_nextLine(is, l)) throw badFMT;
                    std::istringstream ls(l);
                    long idx;
                    ls >> idx;
                    if (elementNodeData) { long n; ls >> n; if ((size_t)n != npe) throw std::runtime_error("Unexpected number-of-nodes-per-element"); }
                    for (size_t d = 0; d < dim; ++d) if (!(ls >> data[i * dim + d])) throw badFMT;
                }
            }
            std::string f;
            if (!m_nextLine(is, f) || f != footer) throw badFMT;
            if (elementNodeData) continue;
            if (dim == 1) {
                SField field(numEntries);
                for (size_t i = 0; i < numEntries; ++i) field[i] = data[i];
                m_scalarFields.emplace(name, std::make_pair(type, std::move(field)));
            } else if (dim == 3) {
                VField field(numEntries);
                for (size_t i = 0; i < numEntries; ++i) for (size_t c = 0; c < N; ++c) field[N * i + c] = data[3 * i + c];
                m_vectorFields.emplace(name, std::make_pair(type, std::move(field)));
            } else if (dim == 9) {
                SMField field(numEntries);
                for (size_t i = 0; i < numEntries; ++i)
                    for (size_t r = 0; r < N; ++r) for (size_t c = r; c < N; ++c)
                        field.data()[SMField::F * i + flattenIndices<N>(r, c)] = data[9 * i + 3 * r + c];
                m_symmetricMatrixFields.emplace(name, std::make_pair(type, std::move(field)));
            } else throw std::runtime_error("Bad field dimension");
        }
    }
};
#endif

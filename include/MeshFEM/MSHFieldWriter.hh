// Gmsh .msh field output (mirrors MSHFieldWriter.hh:41-196, 329-351): mesh (vertex-subsampled by
// default) followed by $NodeData / $ElementData sections with one string tag (the quoted field
// name), no real tags, three int tags (0, dim, count); vectors padded to 3, symmetric matrices
// expanded to 9 row-major doubles; binary records (int32 id, doubles...).
#ifndef MESHFEM_B200_MSHFIELDWRITER_HH
#define MESHFEM_B200_MSHFIELDWRITER_HH
#include <MeshFEM/Fields.hh>
#include <MeshFEM/MeshIO.hh>

#include <fstream>

class MSHFieldWriter {
public:
    MSHFieldWriter(const std::string &mshPath, const std::vector<MeshIO::IOVertex> &nodes,
                   const std::vector<MeshIO::IOElement> &elements, MeshIO::MeshType meshType = MeshIO::MESH_GUESS,
                   bool binary = true)
        : m_linearSubsample(false), m_outStream(mshPath, std::ios::binary), m_numVertices(nodes.size()),
          m_numNodes(nodes.size()), m_numElements(elements.size()), m_binary(binary) {
        if (!m_outStream.is_open()) throw std::runtime_error("Failed to open output file '" + mshPath + "'");
        MeshIO::save(m_outStream, nodes, elements, binary ? MeshIO::FMT_MSH : MeshIO::FMT_MSH_ASCII, meshType);
    }

    template <typename Mesh>
    MSHFieldWriter(const std::string &mshPath, const Mesh &mesh, bool linearSubsample = true,
                   MeshIO::MeshType meshType = MeshIO::MESH_GUESS, bool binary = true)
        : m_linearSubsample(linearSubsample), m_outStream(mshPath, std::ios::binary), m_numVertices(mesh.numVertices()),
          m_numNodes(mesh.numNodes()), m_numElements(mesh.numElements()), m_binary(binary) {
        if (!m_outStream.is_open()) { std::cout << "Failed to open output file '" << mshPath << '\'' << std::endl; return; }
        std::vector<MeshIO::IOVertex> outNodes;
        std::vector<MeshIO::IOElement> outElements;
        const size_t nOut = linearSubsample ? mesh.numVertices() : mesh.numNodes();
        const size_t perElem = linearSubsample ? Mesh::verticesPerElement : Mesh::nodesPerElement;
        m_outNodesPerElement = perElem;
        outNodes.reserve(nOut);
        for (size_t n = 0; n < nOut; ++n) outNodes.emplace_back(padTo3D(mesh.nodePosition(n)));
        outElements.reserve(m_numElements);
        for (size_t e = 0; e < m_numElements; ++e) {
            outElements.emplace_back(perElem);
            for (size_t c = 0; c < perElem; ++c) outElements.back()[c] = mesh.elementNode(e, c);
        }
        MeshIO::save(m_outStream, outNodes, outElements, binary ? MeshIO::FMT_MSH : MeshIO::FMT_MSH_ASCII, meshType);
    }

    template <typename Field>
    void addField(const std::string &name, const Field &f, DomainType type = DomainType::GUESS) {
        std::runtime_error invalidDim("Invalid field dimension.");
        size_t numEntries = 0;
        m_determineDomainTypeAndNumEntries(f.domainSize(), type, numEntries);
        const std::string sectionHeader = (type == DomainType::PER_ELEMENT) ? "ElementData" : "NodeData";
        const size_t dim = f.dim();
        size_t paddedDim = dim;
        switch (f.fieldType()) {
            case FIELD_SCALAR: if (dim != 1) throw invalidDim; break;
            case FIELD_VECTOR: if (dim == 2) paddedDim = 3; if (paddedDim != 3) throw invalidDim; break;
            case FIELD_MATRIX: if (f.N() != 2 && f.N() != 3) throw invalidDim; paddedDim = 9; break;
        }
        m_outStream << '$' << sectionHeader << std::endl << '1' << std::endl << '"' << name << '"' << std::endl
                    << '0' << std::endl << '3' << std::endl << '0' << std::endl << paddedDim << std::endl << numEntries << std::endl;
        for (size_t i = 1; i <= numEntries; ++i) {
            auto val = f(i - 1);
            if (m_binary) { int out = int(i); m_outStream.write((char *)&out, sizeof(int)); }
            else m_outStream << i;
            if (f.fieldType() == FIELD_MATRIX) {
                for (size_t k = 0; k < 3; ++k)
                    for (size_t l = 0; l < 3; ++l) {
                        double value = ((k < f.N()) && (l < f.N())) ? val[flattenIndices(f.N(), k, l)] : 0;
                        if (m_binary) m_outStream.write((char *)&value, sizeof(double)); else m_outStream << ' ' << value;
                    }
            } else {
                for (size_t c = 0; c < paddedDim; ++c) {
                    double value = (c < dim) ? val[c] : 0;
                    if (m_binary) m_outStream.write((char *)&value, sizeof(double)); else m_outStream << ' ' << value;
                }
            }
            if (!m_binary) m_outStream << std::endl;
        }
        m_outStream << "$End" << sectionHeader << std::endl;
    }

    // Vector-of-interpolants field -> $ElementNodeData (MSHFieldWriter.hh:262-306): per element its index, the
    // number of output nodes per element, then 9 (padded 3x3) values per node.  With linear subsampling only
    // the vertex values of a higher-degree interpolant are written.
    template <typename _Real, size_t _N>
    void addField(const std::string &name, const SymmetricMatrixInterpolantField<_Real, _N> &f, DomainType type = DomainType::PER_ELEMENT) {
        if (type != DomainType::PER_ELEMENT || f.domainSize() != m_numElements)
            throw std::runtime_error("Vector-of-interpolants must be per-element.");
        const size_t nOut = m_outNodesPerElement;
        if (nOut == 0) throw std::runtime_error("ElementNodeData needs a writer constructed from a mesh");
        if (f.nodesPerElement() < nOut) throw std::runtime_error("Interpolant has too few nodes");
        m_outStream << "$ElementNodeData" << std::endl << '1' << std::endl << '"' << name << '"' << std::endl << '0' << std::endl
                    << '3' << std::endl << '0' << std::endl << 9 << std::endl << m_numElements << std::endl;
        for (size_t i = 1; i <= m_numElements; ++i) {
            if (m_binary) { int out[2] = {int(i), int(nOut)}; m_outStream.write((char *)out, 2 * sizeof(int)); }
            else m_outStream << i << ' ' << nOut;
            for (size_t n = 0; n < nOut; ++n)
                for (size_t k = 0; k < 3; ++k)
                    for (size_t l = 0; l < 3; ++l) {
                        const double value = ((k < _N) && (l < _N)) ? f(i - 1, n, flattenIndices(_N, k, l)) : 0.0;
                        if (m_binary) m_outStream.write((const char *)&value, sizeof(double)); else m_outStream << ' ' << value;
                    }
            if (!m_binary) m_outStream << std::endl;
        }
        m_outStream << "$EndElementNodeData" << std::endl;
    }

private:
    void m_determineDomainTypeAndNumEntries(size_t domainSize, DomainType &type, size_t &numEntries) const {
        std::runtime_error invalidSize("Invalid field domain size.");
        if (type == DomainType::GUESS) {
            if (domainSize == m_numElements) type = DomainType::PER_ELEMENT;
            else if (domainSize == m_numVertices || domainSize == m_numNodes) type = DomainType::PER_NODE;
            else throw invalidSize;
        }
        if (type == DomainType::PER_ELEMENT) { numEntries = m_numElements; if (domainSize != m_numElements) throw invalidSize; }
        else {
            numEntries = m_linearSubsample ? m_numVertices : m_numNodes;
            if (!(domainSize == m_numNodes || (m_linearSubsample && domainSize == m_numVertices))) throw invalidSize;
        }
    }
    bool m_linearSubsample;
    size_t m_outNodesPerElement = 0;     // nodes per element of the written mesh (ElementNodeData)
    std::ofstream m_outStream;
    size_t m_numVertices, m_numNodes, m_numElements;
    bool m_binary;
};
#endif

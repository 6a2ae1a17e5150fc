// DeformedCells_cli -- generates and homogenizes base cells that have been deformed linearly, with the
// reference's command line (src/bin/DeformedCells_cli.cc:35-418 of MeshFEM):
//   --homogenize                   periodic conditions are matched on the UNDEFORMED cell, the nodes are then moved by
//                                  x -> J (x - centre) and the cell problems are solved on the deformed geometry
//                                  (Eh over the deformed cell volume |Y| det J);
//   --homogenize --transformVersion  the equivalent computation on the undeformed cell with the base material pulled
//                                  back, E' = E.transform(J^-1), and the result pushed forward, Eh = Eh'.transform(J);
//   -p                             2D only: "theta lambda" lines on stdin, J = R(theta) diag(lambda, 1) R(theta)^T;
//   -t 'nx ny [nz]' -o out         writes nx x ny x nz deformed copies of the cell (not glued, as in the reference).
// Assembly and solves run on the GPU through libmfem_b200; tiling needs no GPU.
#include <MeshFEM/CmdLine.hh>
#include <MeshFEM/GlobalBenchmark.hh>
#include <MeshFEM/JSON.hh>
#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/MeshIO.hh>
#include <MeshFEM/PeriodicHomogenization.hh>
#include <MeshFEM/filters/remove_dangling_vertices.hh>

#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <vector>

using namespace std;
using namespace PeriodicHomogenization;

[[noreturn]] static void usage(int exitVal, const CmdLine &opts) {
    cerr << "Usage: DeformedCells_cli [options] in.msh -j 'u_x,x u_x,y ...' out.msh" << endl;
    opts.printOptions(cerr);
    cerr << endl;
    exit(exitVal);
}

static CmdLine parseCmdLine(int argc, const char *argv[]) {
    CmdLine args;
    args.positional("mesh");
    args.flag("help", 0, "Produce this help message")
        .flag("homogenize", 0, "run homogenization")
        .flag("transformVersion", 0, "use transform version of homogenization")
        .value("material", 'm', "base material")
        .value("jacobian", 'j', "linear deformation jacobian")
        .value("displacedMesh", 0, "file prefix containing displaced mesh")
        .value("displacementScale", 0, "used to scale displacements obtained from constant plus periodic strain")
        .flag("parametrizedTransform", 'p', "read a list of parameterized deformations from stdin")
        .value("degree", 'd', "degree of finite elements", "2")
        .value("tile", 't', "tilings 'nx ny nz' (default: 1)")
        .value("out", 'o', "output file of deformed geometry (and w_ij fields if homogenization is run)")
        .value("dumpJson", 0, "dump info into a json file)")
        .value("device", 0, "CUDA device ordinal", "0")
        .value("rtol", 0, "PCG relative residual tolerance", "1e-10")
        .value("maxIters", 0, "PCG iteration cap", "200000");
    try {
        args.parse(argc, argv);
    } catch (std::exception &e) {
        cerr << "Error: " << e.what() << endl << endl;
        usage(1, args);
    }
    bool fail = false;
    if (args.count("mesh") == 0) {
        cerr << "Error: must specify input mesh" << endl;
        fail = true;
    }
    if (args.count("tile") + args.count("homogenize") == 2) {
        cerr << "Error: do not specify both tiling and homogenization" << endl;
        fail = true;
    }
    if (args.count("out") + args.count("homogenize") == 0) {
        cerr << "Error: no operation requested." << endl;
        fail = true;
    }
    if (args.count("jacobian") + args.count("parametrizedTransform") != 1) {
        cerr << "Error: must specify either deformation jacobian or parametrizedTransform" << endl;
        fail = true;
    }
    if (fail || args.count("help")) usage(fail, args);
    return args;
}

static vector<string> splitWhitespace(const string &s) {
    istringstream is(s);
    vector<string> parts;
    string tok;
    while (is >> tok) parts.push_back(tok);
    return parts;
}

template <class ETensor>
static void dumpJson(const ETensor &EhDefo, const string &filename) {
    mjson::json data = mjson::json::object();
    data["elasticity_tensor"] = mjson::json(EhDefo.getCoefficients());
    data["homogenized_moduli"] = mjson::json(EhDefo.getOrthotropicParameters());
    ofstream out(filename);
    if (!out) throw runtime_error("Couldn't open " + filename);
    out << data.dump();
}

template <size_t _N>
static Real determinant(const Real (&J)[_N][_N]) {
    if (_N == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
    return J[0][0] * (J[1][1] * J[2 % _N][2 % _N] - J[1][2 % _N] * J[2 % _N][1]) - J[0][1] * (J[1][0] * J[2 % _N][2 % _N] - J[1][2 % _N] * J[2 % _N][0]) +
           J[0][2 % _N] * (J[1][0] * J[2 % _N][1] - J[1][1] * J[2 % _N][0]);
}

template <size_t _N>
static void inverse(const Real (&J)[_N][_N], Real (&Jinv)[_N][_N]) {
    const Real det = determinant<_N>(J);
    if (det == 0.0) throw runtime_error("Singular deformation jacobian");
    if (_N == 2) {
        Jinv[0][0] = J[1][1] / det; Jinv[0][1] = -J[0][1] / det;
        Jinv[1][0] = -J[1][0] / det; Jinv[1][1] = J[0][0] / det;
        return;
    }
    for (size_t i = 0; i < _N; ++i)
        for (size_t j = 0; j < _N; ++j) {       // cofactor of (j, i)
            const size_t r0 = (j + 1) % _N, r1 = (j + 2) % _N, c0 = (i + 1) % _N, c1 = (i + 2) % _N;
            Jinv[i][j] = (J[r0][c0] * J[r1][c1] - J[r0][c1] * J[r1][c0]) / det;
        }
}

template <size_t _N, size_t _FEMDegree>
void execute(const CmdLine &args, const vector<MeshIO::IOVertex> &inVertices, const vector<MeshIO::IOElement> &inElements) {
    typedef LinearElasticity::Mesh<_N, _FEMDegree> Mesh;
    typedef LinearElasticity::Simulator<Mesh> Simulator;
    typedef typename Simulator::VField VField;
    typedef typename Simulator::ETensor ETensor;
    typedef typename Simulator::SMatrix SMatrix;
    Materials::Constant<_N> mat;
    if (args.count("material")) mat.setFromFile(args.str("material"));
    const ETensor EBase = mat.getTensor();

    cout << setprecision(16);

    // the simulator (and with it the GPU handle) is only needed for homogenization
    unique_ptr<Simulator> simPtr;
    auto makeSim = [&]() -> Simulator & {
        simPtr.reset(new Simulator(inElements, inVertices, args.integer("device")));
        simPtr->setMaterial(EBase);
        simPtr->setSolverTolerance(stod(args.str("rtol")), args.integer("maxIters"));
        return *simPtr;
    };

    if (args.count("parametrizedTransform")) {
        if (!args.count("homogenize") || args.count("tile")) throw runtime_error("parametrizedTransform only supports homogenization");
        if (args.count("transformVersion") == 0) cerr << "WARNING: running transformVersion" << endl;
        if (_N != 2) throw runtime_error("parametrizedTransform only supports 2D");
        Simulator &sim = makeSim();
        string line;
        while (getline(cin, line)) {
            const size_t first = line.find_first_not_of(" \t\r");
            if (first == string::npos || line[first] == '#') continue;      // getDataLine: skip blanks and comments
            const auto comps = splitWhitespace(line);
            if (comps.size() != 2) throw runtime_error("invalid input transformation: " + line);
            const Real theta = stod(comps[0]), lambda = stod(comps[1]);
            // rot * stretch * rot^T
            const Real c = cos(theta), s = sin(theta);
            Real jacobian[_N][_N], jinv[_N][_N];
            jacobian[0][0] = lambda * c * c + s * s;
            jacobian[0][1 % _N] = (lambda - 1.0) * c * s;
            jacobian[1 % _N][0] = (lambda - 1.0) * c * s;
            jacobian[1 % _N][1 % _N] = lambda * s * s + c * c;
            for (size_t i = 0; i < _N; ++i) { for (size_t j = 0; j < _N; ++j) cout << (j ? " " : "") << jacobian[i][j]; cout << endl; }
            inverse<_N>(jacobian, jinv);
            sim.setMaterial(EBase.transform(jinv));
            vector<VField> w_ij;
            solveCellProblems(w_ij, sim);
            const ETensor EhDefo = homogenizedElasticityTensorDisplacementForm(w_ij, sim).transform(jacobian);
            const ETensor ShDefo = EhDefo.inverse();
            cout << theta << '\t' << lambda;
            for (const ETensor *T : {&EhDefo, &ShDefo})
                for (size_t i = 0; i < flatLen(_N); ++i) for (size_t j = i; j < flatLen(_N); ++j) cout << '\t' << T->D(i, j);
            cout << endl;
            if (args.count("dumpJson")) dumpJson(EhDefo, args.str("dumpJson"));
        }
        return;
    }

    // Parse jacobian.
    Real jacobian[_N][_N], jinv[_N][_N];
    const auto jacobianComponents = splitWhitespace(args.str("jacobian"));
    if (jacobianComponents.size() != _N * _N) throw runtime_error("Invalid deformation jacobian");
    for (size_t i = 0; i < _N; ++i)
        for (size_t j = 0; j < _N; ++j) jacobian[i][j] = stod(jacobianComponents[_N * i + j]);

    // bounding box of the vertices (== the mesh's: edge nodes are midpoints)
    VectorND<_N> minC, maxC;
    for (size_t c = 0; c < _N; ++c) { minC[c] = inVertices.at(0)[c]; maxC[c] = inVertices.at(0)[c]; }
    for (const auto &v : inVertices)
        for (size_t c = 0; c < _N; ++c) { minC[c] = std::min(minC[c], (Real)v[c]); maxC[c] = std::max(maxC[c], (Real)v[c]); }
    const VectorND<_N> center = 0.5 * (minC + maxC), dims = maxC - minC;
    Real bboxVolume = 1.0;
    for (size_t c = 0; c < _N; ++c) bboxVolume *= dims[c];

    vector<MeshIO::IOVertex> deformedVertices;
    for (const auto &v : inVertices) {
        MeshIO::IOVertex d;
        for (size_t i = 0; i < _N; ++i) {
            Real s = 0.0;
            for (size_t j = 0; j < _N; ++j) s += jacobian[i][j] * (v[j] - center[j]);
            d[i] = s;
        }
        deformedVertices.push_back(d);
    }
    const Real deformedCellVolume = bboxVolume * determinant<_N>(jacobian);

    auto report = [&](const ETensor &EhDefo) {
        cout << "Elasticity tensor:" << endl;
        cout << EhDefo << endl << endl;
        cout << "Homogenized Moduli: ";
        EhDefo.printOrthotropic(cout);
        if (args.count("dumpJson")) dumpJson(EhDefo, args.str("dumpJson"));
    };

    if (args.count("homogenize") && args.count("transformVersion")) {
        Simulator &sim = makeSim();
        inverse<_N>(jacobian, jinv);
        sim.setMaterial(EBase.transform(jinv));
        vector<VField> w_ij;
        solveCellProblems(w_ij, sim);
        report(homogenizedElasticityTensorDisplacementForm(w_ij, sim).transform(jacobian));
    } else if (args.count("homogenize")) {
        Simulator &sim = makeSim();
        sim.applyPeriodicConditions();
        sim.applyNoRigidMotionConstraint();
        sim.setUsePinNoRigidTranslationConstraint(true);
        sim.updateMeshNodePositions(deformedVertices);
        shared_ptr<MSHFieldWriter> writer;
        if (args.count("out")) writer = make_shared<MSHFieldWriter>(args.str("out"), sim.mesh());
        constexpr size_t numStrains = flatLen(_N);
        vector<VField> rhs;
        for (size_t i = 0; i < numStrains; ++i) rhs.push_back(sim.constantStrainLoad(-SMatrix::CanonicalBasis(i)));
        vector<VField> w_ij = sim.solve(rhs);              // one batched PCG for all cell problems
        if (writer)
            for (size_t i = 0; i < numStrains; ++i) {
                writer->addField("load_ij " + to_string(i), sim.dofToNodeField(rhs[i]), DomainType::PER_NODE);
                writer->addField("w_ij" + to_string(i), w_ij[i], DomainType::PER_NODE);
                writer->addField("strain w_ij " + to_string(i), sim.averageStrainField(w_ij[i]), DomainType::PER_ELEMENT);
            }
        report(homogenizedElasticityTensorDisplacementForm(w_ij, sim, deformedCellVolume));

        if (args.count("displacedMesh")) {
            if (!writer) throw runtime_error("--displacedMesh needs --out (the u_cstrain fields are written there)");
            const string outMesh = args.str("displacedMesh");
            const auto &mesh = sim.mesh();
            const auto displacedCenter = mesh.boundingBox().center();
            const Real displacementScale = args.count("displacementScale") ? stod(args.str("displacementScale")) : 0.1;
            vector<VField> cstrainDisp_ij;
            for (size_t index = 0; index < 3; ++index) {           // the reference probes the first three strains
                VField u(mesh.numNodes());
                const SMatrix strain = SMatrix::CanonicalBasis(index);
                for (size_t n = 0; n < mesh.numNodes(); ++n) u.set(n, strain.contract(mesh.nodePosition(n) - displacedCenter) + w_ij[index](n));
                writer->addField("u_cstrain_ij" + to_string(index), u, DomainType::PER_NODE);
                cstrainDisp_ij.push_back(u);
            }
            vector<MeshIO::IOVertex> baseVertices;
            for (size_t vi = 0; vi < mesh.numVertices(); ++vi) {
                MeshIO::IOVertex p;
                for (size_t c = 0; c < _N; ++c) p[c] = mesh.nodePosition(vi)[c];
                baseVertices.push_back(p);
            }
            writer.reset();                                          // flush the field file before the geometry changes
            for (size_t index = 0; index < 3; ++index) {
                // as in the reference each displaced mesh starts from the previous one's geometry
                vector<MeshIO::IOVertex> displacedVertices;
                for (size_t vi = 0; vi < mesh.numVertices(); ++vi) {
                    MeshIO::IOVertex p;
                    for (size_t c = 0; c < _N; ++c) p[c] = mesh.nodePosition(vi)[c] + displacementScale * cstrainDisp_ij[index](vi)[c];
                    displacedVertices.push_back(p);
                }
                sim.updateMeshNodePositions(displacedVertices);
                MSHFieldWriter writerDisplacedMesh(outMesh + to_string(index) + ".msh", sim.mesh());
            }
        }
        const auto &info = sim.lastSolveInfo();
        cout << "PCG iterations (last cell problem):\t" << info.iterations << "\trelative residual:\t" << info.rel_residual << endl;
    } else if (args.count("tile")) {
        const auto tileComponents = splitWhitespace(args.str("tile"));
        if (tileComponents.size() != _N) throw runtime_error("Invalid number of tiling dimensions");
        vector<size_t> tilings;
        for (const auto &c : tileComponents) {
            const int ci = stoi(c);
            if (ci <= 0) throw runtime_error("Invalid number of tilings");
            tilings.push_back((size_t)ci);
        }
        if (_N == 2) tilings.push_back(1);
        vector<MeshIO::IOVertex> tiledVertices;
        vector<MeshIO::IOElement> tiledElements;
        const size_t numCellVertices = inVertices.size();
        size_t copy = 0;
        for (size_t i = 0; i < tilings[0]; ++i)
            for (size_t j = 0; j < tilings[1]; ++j)
                for (size_t k = 0; k < tilings[2]; ++k, ++copy) {
                    const Real delta[3] = {i * dims[0], j * dims[1], (_N > 2) ? k * dims[_N - 1] : 0.0};
                    Real offset[_N];
                    for (size_t a = 0; a < _N; ++a) {
                        offset[a] = 0.0;
                        for (size_t b = 0; b < _N; ++b) offset[a] += jacobian[a][b] * delta[b];
                    }
                    for (const auto &v : deformedVertices) {
                        MeshIO::IOVertex p;
                        for (size_t a = 0; a < _N; ++a) p[a] = v[a] + offset[a];
                        tiledVertices.push_back(p);
                    }
                    for (auto e : inElements) {
                        for (size_t ei = 0; ei < e.size(); ++ei) e[ei] += numCellVertices * copy;
                        tiledElements.push_back(e);
                    }
                }
        // duplicated vertices on the glued faces are kept (the reference's "TODO: merge duplicated vertices")
        remove_dangling_vertices(tiledVertices, tiledElements);
        MeshIO::save(args.str("out"), tiledVertices, tiledElements);
    }
}

int main(int argc, const char *argv[]) {
    try {
        CmdLine args = parseCmdLine(argc, argv);
        vector<MeshIO::IOVertex> inVertices;
        vector<MeshIO::IOElement> inElements;
        auto type = MeshIO::load(args.str("mesh"), inVertices, inElements, MeshIO::FMT_GUESS, MeshIO::MESH_GUESS);
        size_t dim;
        if (type == MeshIO::MESH_TET) dim = 3;
        else if (type == MeshIO::MESH_TRI) dim = 2;
        else throw std::runtime_error("Mesh must be triangle or tet.");
        const int deg = args.integer("degree");
        auto exec = (dim == 3) ? ((deg == 2) ? execute<3, 2> : execute<3, 1>) : ((deg == 2) ? execute<2, 2> : execute<2, 1>);
        exec(args, inVertices, inElements);
    } catch (const std::exception &e) {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

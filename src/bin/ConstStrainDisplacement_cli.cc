// ConstStrainDisplacement_cli -- constant-strain displacement of a base cell, optionally with the
// fluctuation displacements of periodic homogenization added, with the reference's command line
// (src/bin/ConstStrainDisplacement_cli.cc:44-300 of MeshFEM):
//   u(x) = strain . (x - bbox centre)   [+ sum_ij doubler_ij strain_ij w_ij(x)]
// A macroscopic STRESS probe (-S) is converted with the homogenized compliance.  Fields written:
// "w_ij k" (with -f), "u_cstrain", "f_cstrain" = K u (with -l), "stress" (element averages).
// The cell problems, K u and the stress field run on the GPU through libmfem_b200.
#include <MeshFEM/CmdLine.hh>
#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/MeshIO.hh>
#include <MeshFEM/OrthotropicHomogenization.hh>
#include <MeshFEM/PeriodicHomogenization.hh>

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <vector>

using namespace std;
using namespace PeriodicHomogenization;

[[noreturn]] static void usage(int exitVal, const CmdLine &opts) {
    cout << "Usage: ConstStrainDisplacement_cli [options] in.msh -s 'e_00 e_11 ...' out.msh" << endl;
    opts.printOptions(cout);
    cout << endl;
    exit(exitVal);
}

static CmdLine parseCmdLine(int argc, const char *argv[]) {
    CmdLine args;
    args.positional("mesh").positional("outMesh");
    args.flag("help", 0, "Produce this help message")
        .value("material", 'm', "base material")
        .value("strain", 's', "macroscopic strain tensor")
        .value("stress", 'S', "macroscopic stress tensor")
        .value("degree", 'd', "degree of finite elements", "2")
        .flag("nodalLoad", 'l', "compute the effective force on each node.")
        .flag("addFluctuation", 'f', "add fluctuation strains to the displacement")
        .value("macroOut", 0, "also output the unit cell deformation")
        .value("manualPeriodicVertices", 0, "Manually specify identified periodic vertices")
        .flag("orthotropicCell", 'O', "Analyze the orthotropic symmetry base cell only")
        .value("device", 0, "CUDA device ordinal", "0")
        .value("rtol", 0, "PCG relative residual tolerance", "1e-10");
    try {
        args.parse(argc, argv);
    } catch (std::exception &e) {
        cout << "Error: " << e.what() << endl << endl;
        usage(1, args);
    }
    bool fail = false;
    if (args.count("outMesh") == 0) {
        cout << "Error: must specify input and output mesh" << endl;
        fail = true;
    }
    if (args.count("strain") + args.count("stress") != 1) {
        cout << "Error: must specify macro strain or stress tensor" << endl;
        fail = true;
    }
    int d = 0;
    try { d = args.integer("degree"); } catch (...) {}
    if (d < 1 || d > 2) {
        cout << "Error: FEM Degree must be 1 or 2" << endl;
        fail = true;
    }
    if (fail || args.count("help")) usage(fail, args);
    return args;
}

template <size_t _N, size_t _FEMDegree>
void execute(const CmdLine &args, const vector<MeshIO::IOVertex> &inVertices, const vector<MeshIO::IOElement> &inElements) {
    Materials::Constant<_N> mat;
    if (args.count("material")) mat.setFromFile(args.str("material"));
    typedef LinearElasticity::Simulator<LinearElasticity::Mesh<_N, _FEMDegree>> Simulator;
    typedef typename Simulator::VField VField;
    Simulator sim(inElements, inVertices, args.integer("device"));
    sim.setMaterial(mat.getTensor());
    std::unique_ptr<PeriodicCondition<_N>> pc;
    if (args.count("manualPeriodicVertices")) pc.reset(new PeriodicCondition<_N>(sim.mesh(), args.str("manualPeriodicVertices")));
    sim.setSolverTolerance(std::stod(args.str("rtol")));
    const auto &mesh = sim.mesh();
    MSHFieldWriter writer(args.str("outMesh"), mesh);

    // strain (or stress) probe
    vector<Real> comps;
    {
        istringstream ss(args.count("strain") ? args.str("strain") : args.str("stress"));
        Real v;
        while (ss >> v) comps.push_back(v);
    }
    if (comps.size() != flatLen(_N)) throw runtime_error("Invalid strain tensor");
    SymmetricMatrixValue<Real, _N> strain;
    for (size_t i = 0; i < comps.size(); ++i) strain[i] = comps[i];

    std::vector<VField> w_ij;
    const bool ortho = args.count("orthotropicCell") != 0;
    auto doCellProblemSolve = [&]() {
        if (!ortho) solveCellProblems(w_ij, sim, 1e-7, false, std::move(pc));
        else PeriodicHomogenization::Orthotropic::solveCellProblems(w_ij, sim, 1e-7);
    };
    auto getHomogenizedTensor = [&]() {
        return ortho ? PeriodicHomogenization::Orthotropic::homogenizedElasticityTensorDisplacementForm(w_ij, sim)
                     : homogenizedElasticityTensorDisplacementForm(w_ij, sim);
    };
    if (args.count("stress")) {       // convert the stress probe to the corresponding strain probe
        doCellProblemSolve();
        strain = getHomogenizedTensor().inverse().doubleContract(strain);
    }

    const auto bbox = mesh.boundingBox();
    const VectorND<_N> center = bbox.center();

    if (args.count("macroOut")) {
        if (_N != 2) throw std::runtime_error("macro displacement output currently only supported in 2D");
        std::vector<MeshIO::IOVertex> squareVertices;
        std::vector<MeshIO::IOElement> squareElems;
        VField uMacro(4);
        SymmetricMatrixField<Real, _N> stressMacro(2);
        if (w_ij.size() == 0) doCellProblemSolve();
        const auto sm = getHomogenizedTensor().doubleContract(strain);
        for (size_t e = 0; e < 2; ++e) for (size_t k = 0; k < flatLen(_N); ++k) stressMacro.data()[flatLen(_N) * e + k] = sm[k];
        size_t i = 0;           // 2 3 / 0 1
        for (Real y : {bbox.minCorner[1], bbox.maxCorner[1]})
            for (Real x : {bbox.minCorner[0], bbox.maxCorner[0]}) {
                VectorND<_N> p;
                p[0] = x, p[1] = y;
                squareVertices.emplace_back(padTo3D(p));
                uMacro.set(i++, strain.contract(p - center));
            }
        squareElems.emplace_back(0, 1, 3);
        squareElems.emplace_back(0, 3, 2);
        MSHFieldWriter mwriter(args.str("macroOut"), squareVertices, squareElems);
        mwriter.addField("u_cstrain", uMacro, DomainType::PER_NODE);
        mwriter.addField("stress", stressMacro, DomainType::PER_ELEMENT);
    }

    VField cstrainDisp(mesh.numNodes());
    for (size_t n = 0; n < mesh.numNodes(); ++n) cstrainDisp.set(n, strain.contract(mesh.nodePosition(n) - center));

    if (args.count("addFluctuation")) {
        if (w_ij.size() == 0) doCellProblemSolve();
        // keep the periodic boundary on the cell faces in an average sense: per component, subtract the mean
        // over the boundary nodes on the matching minimum face (:205-232)
        for (auto &w : w_ij) {
            VectorND<_N> translation;
            vector<int> numAveraged(_N);
            for (size_t bni = 0; bni < mesh.numBoundaryNodes(); ++bni) {
                const size_t n = (size_t)mesh.volumeNodeForBoundaryNode(bni);
                const auto p = mesh.nodePosition(n);
                for (size_t d = 0; d < _N; ++d)
                    if (std::abs(p[d] - bbox.minCorner[d]) < 1e-9) { translation[d] += w(n)[d]; ++numAveraged[d]; }
            }
            for (size_t d = 0; d < _N; ++d) translation[d] /= numAveraged[d];
            for (size_t n = 0; n < w.domainSize(); ++n) w.set(n, w(n) - translation);
        }
        for (size_t i = 0; i < w_ij.size(); ++i) {
            const Real c = ((i < _N) ? 1.0 : 2.0) * strain[i];
            for (size_t n = 0; n < mesh.numNodes(); ++n) cstrainDisp.add(n, c * w_ij[i](n));
            writer.addField("w_ij " + to_string(i), w_ij[i], DomainType::PER_NODE);
        }
    }

    writer.addField("u_cstrain", cstrainDisp, DomainType::PER_NODE);
    if (args.count("nodalLoad")) writer.addField("f_cstrain", sim.applyStiffnessMatrix(cstrainDisp), DomainType::PER_NODE);
    writer.addField("stress", sim.averageStressField(cstrainDisp), DomainType::PER_ELEMENT);
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

// Simulate_cli -- linear-elasticity simulation driver with the reference's command line
// (src/bin/Simulate_cli.cc:22-321 of MeshFEM): same options, same flow
//   load mesh -> Simulator -> material -> [dump K] -> read .bc -> pins / BCs / pair conditions ->
//   solve -> strain, stress, load -> out.msh fields "u", "load", "strain", "stress", "Ku" ->
//   region force printout -> BENCHMARK_REPORT
// with the assemble-and-solve path running on the GPU through libmfem_b200 (no CPU fallback).
//
// Differences from the reference, all outside the hot path:
//  * heterogeneous material from a .msh (-m x.msh -f name) is read by MSHFieldParser (E/nu or the
//    orthotropic parameter fields) and uploaded as a per-element D array;
//  * extra options: --device, --rtol, --maxIters (PCG controls; the reference's direct solver has none).
#include <MeshFEM/CmdLine.hh>
#include <MeshFEM/GlobalBenchmark.hh>
#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/MSHFieldParser.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/MeshIO.hh>

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <vector>

using namespace std;

[[noreturn]] static void usage(int exitVal, const CmdLine &opts) {
    cout << "Usage: Simulate_cli [options] mesh" << endl;
    opts.printOptions(cout);
    cout << endl;
    exit(exitVal);
}

static CmdLine parseCmdLine(int argc, const char *argv[]) {
    CmdLine args;
    args.positional("mesh");
    args.flag("help", 0, "Produce this help message")
        .value("material", 'm', "simulation material material", "")
        .value("matFieldName", 'f', "name of material field to load from .msh passed as --material", "")
        .value("boundaryConditions", 'b', "boundary conditions")
        .value("outputMSH", 'o', "output mesh")
        .value("dumpMatrix", 0, "dump system matrix in triplet format", "")
        .value("degree", 'd', "FEM degree (1 or 2)", "2")
        .flag("fullDegreeFieldOutput", 'D', "Output full-degree nodal fields (don't do piecewise linear subsample)")
        .value("extraMesh", 'e', "adds another independent input mesh to problem")
        .value("device", 0, "CUDA device ordinal", "0")
        .value("rtol", 0, "PCG relative residual tolerance", "1e-10")
        .value("maxIters", 0, "PCG iteration cap", "200000");
    try {
        args.parse(argc, argv);
    } catch (std::exception &e) {
        cout << "Error: " << e.what() << endl << endl;
        usage(1, args);
    }
    bool fail = false;
    if (args.count("mesh") == 0) {
        cout << "Error: must specify input mesh" << endl;
        fail = true;
    }
    if (args.str("dumpMatrix").size() == 0 && args.count("outputMSH") == 0) {
        cout << "Error: must specify output msh file (unless dumping a stiffness matrix)" << endl;
        fail = true;
    }
    if (args.count("outputMSH") && (args.count("boundaryConditions") == 0)) {
        cout << "Error: must specify boundary conditions to run a simulation" << endl;
        fail = true;
    }
    if (fail || args.count("help")) usage(fail, args);
    return args;
}

template <size_t _N, size_t _Deg>
void execute(const CmdLine &args, const vector<MeshIO::IOVertex> &inVertices, const vector<MeshIO::IOElement> &inElements) {
    const size_t numElements = inElements.size();
    typedef LinearElasticity::Mesh<_N, _Deg> Mesh;
    using Simulator = LinearElasticity::Simulator<Mesh>;
    typedef typename Simulator::ETensor ETensor;
    Simulator sim(inElements, inVertices, args.integer("device"));
    sim.setSolverTolerance(std::stod(args.str("rtol")), args.integer("maxIters"));

    typedef ScalarField<Real> SField;
    const string &materialPath = args.str("material"), &matFieldName = args.str("matFieldName"),
                 &matrixPath = args.str("dumpMatrix");
    string bcPath, outMSH;
    if (args.count("boundaryConditions")) bcPath = args.str("boundaryConditions");
    if (args.count("outputMSH")) outMSH = args.str("outputMSH");

    if (fileExtension(materialPath) == ".msh") {
        // heterogeneous material: isotropic names E nu; orthotropic names
        // E_x E_y [E_z] nu_yx [nu_zx nu_zy] [mu_yz mu_zx] mu[_xy]   (Simulate_cli.cc:104-165)
        MSHFieldParser<_N> fieldParser(materialPath);
        auto domainSizeChecker = [=](const vector<SField> &fs) -> bool {
            return all_of(fs.begin(), fs.end(), [=](const SField &f) { return f.domainSize() == numElements; });
        };
        runtime_error sizeErr("Material parameter fields of incorrect size.");
        runtime_error notFound("No complete material parameter field was found.");
        vector<SField> paramFields;
        for (string name : {"E", "nu"}) {
            name = matFieldName + name;
            try { paramFields.push_back(fieldParser.scalarField(name, DomainType::PER_ELEMENT)); } catch (...) { /* try orthotropic */ }
        }
        vector<ETensor> Es(numElements);
        if (paramFields.size() == 2) {
            if (!domainSizeChecker(paramFields)) throw sizeErr;
            for (size_t i = 0; i < numElements; ++i) Es[i].setIsotropic(paramFields[0][i], paramFields[1][i]);
            sim.setPerElementMaterial(Es);
            cout << "Loaded " << _N << "D isotropic material" << endl;
        } else {
            paramFields.clear();
            const vector<vector<string>> orthotropicNames = {{"E_x", "E_y", "nu_yx", "mu"},
                                                             {"E_x", "E_y", "E_z", "nu_yx", "nu_zx", "nu_zy", "mu_yz", "mu_zx", "mu_xy"}};
            for (string name : orthotropicNames.at(_N - 2)) {
                name = matFieldName + name;
                try { paramFields.push_back(fieldParser.scalarField(name, DomainType::PER_ELEMENT)); } catch (...) { throw notFound; }
            }
            if (!domainSizeChecker(paramFields)) throw sizeErr;
            for (size_t i = 0; i < numElements; ++i) {
                const auto &p = paramFields;
                if (_N == 2) Es[i].setOrthotropic2D(p[0][i], p[1][i], p[2][i], p[3][i]);
                else Es[i].setOrthotropic3D(p[0][i], p[1][i], p[2][i], p[3][i], p[4][i], p[5][i], p[6][i], p[7][i], p[8][i]);
            }
            sim.setPerElementMaterial(Es);
            cout << "Loaded " << _N << "D Orthotropic material" << endl;
        }
    } else {
        Materials::Constant<_N> mat;
        if (materialPath != "") mat.setFromFile(materialPath);
        sim.setMaterial(mat.getTensor());
    }

    // just dumping the stiffness matrix without simulating (:178-184)
    if ((matrixPath != "") && (bcPath == "")) {
        typename Simulator::TMatrix K;
        sim.m_assembleStiffnessMatrix(K);
        K.sumRepeated();
        K.dumpBinary(matrixPath);
        exit(0);
    }

    bool noRigidMotion;
    vector<PeriodicPairDirichletCondition<_N>> pps;
    ComponentMask pinTranslationComponents;
    auto bconds = readBoundaryConditions<_N>(bcPath, sim.mesh().boundingBox(), noRigidMotion, pps, pinTranslationComponents);
    sim.applyTranslationPins(pinTranslationComponents);
    sim.applyBoundaryConditions(bconds);
    sim.applyPeriodicPairDirichletConditions(pps);
    if (noRigidMotion) sim.applyNoRigidMotionConstraint();

    if (matrixPath != "") sim.dumpSystem(matrixPath);

    BENCHMARK_START_TIMER_SECTION("Simulation");
    auto u = sim.solve();
    auto e = sim.averageStrainField(u);
    auto s = sim.averageStressField(u);
    auto f = sim.dofToNodeField(sim.neumannLoad());
    BENCHMARK_STOP_TIMER_SECTION("Simulation");

    const bool linearSubsampleFields = args.count("fullDegreeFieldOutput") == 0;
    MSHFieldWriter writer(outMSH, sim.mesh(), linearSubsampleFields);
    writer.addField("u", u, DomainType::PER_NODE);
    writer.addField("load", f, DomainType::PER_NODE);
    if ((_Deg == 1) || linearSubsampleFields) {
        // constant (average) strain/stress for piecewise linear u (:203-207)
        writer.addField("strain", e, DomainType::PER_ELEMENT);
        writer.addField("stress", s, DomainType::PER_ELEMENT);
    } else {
        // full-degree per-element strain/stress as ElementNodeData (:208-224)
        writer.addField("strain", sim.strainField(u), DomainType::PER_ELEMENT);
        writer.addField("stress", sim.stressField(u), DomainType::PER_ELEMENT);
    }

    sim.reportRegionSurfaceForces(u);
    writer.addField("Ku", sim.applyStiffnessMatrix(u), DomainType::PER_NODE);

    const auto &info = sim.lastSolveInfo();
    cout << "PCG iterations:\t" << info.iterations << "\trelative residual:\t" << info.rel_residual << endl;
    BENCHMARK_REPORT();
}

int main(int argc, const char *argv[]) {
    try {
        CmdLine args = parseCmdLine(argc, argv);
        vector<MeshIO::IOVertex> inVertices;
        vector<MeshIO::IOElement> inElements;
        const string meshPath = args.str("mesh");
        auto type = MeshIO::load(meshPath, inVertices, inElements, MeshIO::FMT_GUESS, MeshIO::MESH_GUESS);

        size_t dim;
        if (type == MeshIO::MESH_TET) dim = 3;
        else if (type == MeshIO::MESH_TRI) dim = 2;
        else throw std::runtime_error("Mesh must be pure triangle or tet.");

        if (args.count("extraMesh") > 0) {   // second independent mesh appended to the problem (:271-311)
            vector<MeshIO::IOVertex> inExtraVertices;
            vector<MeshIO::IOElement> inExtraElements;
            auto typeExtra = MeshIO::load(args.str("extraMesh"), inExtraVertices, inExtraElements, MeshIO::FMT_GUESS, MeshIO::MESH_GUESS);
            if (type != typeExtra) {
                std::cerr << "Extra mesh of different type." << std::endl;
                throw std::runtime_error("Extra mesh of different type.");
            }
            for (auto &e : inExtraElements)
                for (size_t i = 0; i < dim + 1; ++i) e[i] += inVertices.size();
            inVertices.insert(inVertices.end(), inExtraVertices.begin(), inExtraVertices.end());
            inElements.insert(inElements.end(), inExtraElements.begin(), inExtraElements.end());
        }

        const int deg = args.integer("degree");
        auto exec = (dim == 3) ? ((deg == 2) ? execute<3, 2> : execute<3, 1>) : ((deg == 2) ? execute<2, 2> : execute<2, 1>);
        exec(args, inVertices, inElements);
    } catch (const std::exception &e) {
        // the reference lets the exception escape (terminate prints what()); we print it and fail
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

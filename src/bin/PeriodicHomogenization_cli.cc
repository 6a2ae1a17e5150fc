// PeriodicHomogenization_cli -- homogenized elasticity tensor of a periodic base cell, with the
// reference's command line and printout (src/bin/PeriodicHomogenization_cli.cc:22-275 of MeshFEM):
//   load cell -> Simulator -> solveCellProblems (periodic DoFs, pinned translation, one assembled K,
//   flatLen(N) solves) -> homogenizedElasticityTensorDisplacementForm -> Eh, eigenstrains,
//   compliance, approximate moduli / Poisson ratios, anisotropy -> optional field output.
// The assembly and the solves run on the GPU through libmfem_b200 (no CPU fallback).
//
// Differences from the reference, all outside the hot path: extra options --device / --rtol / --maxIters control the PCG.
#include <MeshFEM/CmdLine.hh>
#include <MeshFEM/GlobalBenchmark.hh>
#include <MeshFEM/LinearElasticity.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/Materials.hh>
#include <MeshFEM/MeshIO.hh>
#include <MeshFEM/OrthotropicHomogenization.hh>
#include <MeshFEM/PeriodicHomogenization.hh>
#include <MeshFEM/TensorProjection.hh>

#include <cstdlib>
#include <fstream>
#include <iomanip>
#include <iostream>
#include <vector>

using namespace std;
using namespace PeriodicHomogenization;

[[noreturn]] static void usage(int exitVal, const CmdLine &opts) {
    cout << "Usage: PeriodicHomogenization_cli [options] mesh" << endl;
    opts.printOptions(cout);
    cout << endl;
    exit(exitVal);
}

static CmdLine parseCmdLine(int argc, const char *argv[]) {
    CmdLine args;
    args.positional("mesh");
    args.flag("help", 0, "Produce this help message")
        .value("material", 'm', "base material")
        .value("degree", 'd', "degree of finite elements", "2")
        .value("m2mstress", 'M', "Dump macroscopic to microscopic stress tensors to specified file")
        .value("fieldOutput", 'o', "Dump fluctuation stress and strain fields to specified msh file")
        .flag("centerFluctuationDisplacements", 'c', "Shift each fluctuation displacement so that it averages to zero")
        .flag("fullDegreeFieldOutput", 'D', "Output full-degree nodal fields (don't do piecewise linear subsample)")
        .flag("distanceToIsotropy", 0, "Output the distance to the closest isotropic tensor")
        .value("distanceToMaterial", 0, "Output the distance to a particular material")
        .flag("ignorePeriodicMismatch", 0, "Ignore mismatched nodes on the periodic faces (useful for voxel grids)")
        .value("manualPeriodicVertices", 0, "Manually specify identified periodic vertices")
        .flag("orthotropicCell", 'O', "Analyze the orthotropic symmetry base cell only")
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

    typedef LinearElasticity::Mesh<_N, _FEMDegree> Mesh;
    typedef LinearElasticity::Simulator<Mesh> Simulator;
    Simulator sim(inElements, inVertices, args.integer("device"));
    sim.setMaterial(mat.getTensor());
    sim.setSolverTolerance(std::stod(args.str("rtol")), args.integer("maxIters"));
    typedef typename Simulator::ETensor ETensor;
    typedef typename Simulator::VField VField;

    std::unique_ptr<PeriodicCondition<_N>> pc;
    if (args.count("manualPeriodicVertices")) pc.reset(new PeriodicCondition<_N>(sim.mesh(), args.str("manualPeriodicVertices")));

    BENCHMARK_START_TIMER_SECTION("Cell Problems");
    std::vector<VField> w_ij;
    const bool orthotropicCell = args.count("orthotropicCell") != 0;
    if (!orthotropicCell) solveCellProblems(w_ij, sim, 1e-7, args.count("ignorePeriodicMismatch") != 0, std::move(pc));
    else PeriodicHomogenization::Orthotropic::solveCellProblems(w_ij, sim, 1e-7);
    BENCHMARK_STOP_TIMER_SECTION("Cell Problems");

    BENCHMARK_START_TIMER_SECTION("Compute Tensor");
    ETensor Eh = orthotropicCell ? PeriodicHomogenization::Orthotropic::homogenizedElasticityTensorDisplacementForm(w_ij, sim)
                                 : homogenizedElasticityTensorDisplacementForm(w_ij, sim);
    BENCHMARK_STOP_TIMER_SECTION("Compute Tensor");

    cout << setprecision(16);
    cout << "Homogenized elasticity tensor:" << endl;
    cout << Eh << endl << endl;

    auto eigs = Eh.computeEigenstrains();
    static const char *labels[3] = {"Minimum", "Intermediate", "Max"};
    for (size_t k = 0; k < 3; ++k) {
        cout << labels[k] << " Eh eigenvalue " << eigs.lambdas[k] << " for eigenstrain:";
        for (size_t i = 0; i < flatLen(_N); ++i) cout << (i ? " " : " ") << eigs.strains[k][i];
        cout << endl;
    }

    ETensor S = Eh.inverse();
    cout << "Homogenized compliance tensor:" << endl;
    cout << S << endl;
    vector<Real> moduli(flatLen(_N));
    // shear moduli are multiplied by 4 in the flattened compliance tensor
    for (size_t i = 0; i < flatLen(_N); ++i) moduli[i] = ((i < _N) ? 1.0 : 0.25) / S.D(i, i);

    vector<Real> poisson;
    if (_N == 2) poisson = {-S.D(0, 1) / S.D(1, 1), -S.D(1, 0) / S.D(0, 0)};
    else poisson = {-S.D(0, 1) / S.D(1, 1), -S.D(0, 2) / S.D(2, 2), -S.D(1, 2) / S.D(2, 2),
                    -S.D(1, 0) / S.D(0, 0), -S.D(2, 0) / S.D(0, 0), -S.D(2, 1) / S.D(1, 1)};

    if (_N == 2) {
        cout << "Approximate Young moduli:\t" << moduli[0] << "\t" << moduli[1] << endl;
        cout << "Approximate shear modulus:\t" << moduli[2] << endl;
        cout << "v_yx, v_xy:\t" << poisson[0] << "\t" << poisson[1] << endl;
    } else {
        cout << "Approximate Young moduli:\t" << moduli[0] << "\t" << moduli[1] << "\t" << moduli[2] << endl;
        cout << "Approximate shear moduli:\t" << moduli[3] << "\t" << moduli[4] << "\t" << moduli[5] << endl;
        cout << "v_yx, v_zx, v_zy:\t" << poisson[0] << "\t" << poisson[1] << "\t" << poisson[2] << endl;
        cout << "v_xy, v_xz, v_yz:\t" << poisson[3] << "\t" << poisson[4] << "\t" << poisson[5] << endl;
    }
    cout << "Anisotropy:\t" << Eh.anisotropy() << endl;

    if (args.count("m2mstress")) {
        // macroscopic-to-microscopic stress tensors E : G_e : Eh^-1 per element, and the G_e in gtensors.txt (:173-186)
        const string mpath = args.str("m2mstress");
        ofstream mfile(mpath);
        ofstream gfile("gtensors.txt");
        mfile << setprecision(16);
        gfile << setprecision(16);
        if (!mfile.is_open()) throw runtime_error("Failed to open output file " + mpath);
        const auto G = macroStrainToMicroStrainTensors(w_ij, sim);
        for (size_t ei = 0; ei < sim.mesh().numElements(); ++ei) {
            G[ei].writeUnflattened(gfile);
            gfile << endl;
            sim.elementTensor(ei).doubleContractTensor(G[ei].doubleContract(S)).writeUnflattened(mfile);
            mfile << endl;
        }
    }

    if (args.count("fieldOutput")) {
        const bool linearSubsampleFields = args.count("fullDegreeFieldOutput") == 0;
        MSHFieldWriter writer(args.str("fieldOutput"), sim.mesh(), linearSubsampleFields);
        if (args.count("centerFluctuationDisplacements")) {
            for (auto &w : w_ij) {
                VectorND<_N> total;
                for (size_t ii = 0; ii < w.domainSize(); ++ii) total += w(ii);
                total *= 1.0 / w.domainSize();
                for (size_t ii = 0; ii < w.domainSize(); ++ii) w.set(ii, w(ii) - total);
            }
        }
        for (size_t i = 0; i < w_ij.size(); ++i) {
            writer.addField("load_ij " + to_string(i), sim.dofToNodeField(sim.constantStrainLoad(-Simulator::SMatrix::CanonicalBasis(i))), DomainType::PER_NODE);
            writer.addField("w_ij " + to_string(i), w_ij[i], DomainType::PER_NODE);
            if ((_FEMDegree == 1) || linearSubsampleFields)
                writer.addField("strain w_ij " + to_string(i), sim.averageStrainField(w_ij[i]), DomainType::PER_ELEMENT);
            else        // full-degree per-element strain as ElementNodeData (:214-226)
                writer.addField("strain w_ij " + to_string(i), sim.strainField(w_ij[i]), DomainType::PER_ELEMENT);
        }
    }

    if (args.count("distanceToIsotropy")) {
        const ETensor isoFit = closestIsotropicTensor(Eh);
        ETensor diff = Eh;
        diff *= -1.0;
        diff += isoFit;
        cout << endl;
        cout << "(Sq Rel Frob) Distance to Isotropy:\t" << diff.frobeniusNormSq() / isoFit.frobeniusNormSq() << endl;
        cout << "Closest isotropic tensor:" << endl << isoFit << endl;
        cout << endl;
    }

    if (args.count("distanceToMaterial")) {
        Materials::Constant<_N> targetMat(args.str("distanceToMaterial"));
        const auto &tgtE = targetMat.getTensor();
        ETensor diff = tgtE;
        diff *= -1.0;
        diff += Eh;
        cout << "(Sq Rel Frob) Distance to Specified Tensor:\t" << diff.frobeniusNormSq() / tgtE.frobeniusNormSq() << endl;
    }

    const auto &info = sim.lastSolveInfo();
    cout << "PCG iterations (last cell problem):\t" << info.iterations << "\trelative residual:\t" << info.rel_residual << endl;
    BENCHMARK_REPORT();
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

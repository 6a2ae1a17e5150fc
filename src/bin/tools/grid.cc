// grid -- create grid meshes, optionally tesselated into triangles / tetrahedra, with the
// reference tool's command line (src/bin/tools/grid.cc:24-146 of MeshFEM):
//   grid CxR[xS] out.msh [-t] [-m minCorner -M maxCorner]
// -t writes the symmetric simplex tesselation (4 triangles per quad, 24 tets per hex) together
// with the per-simplex "cell_index" element field; without it the quad / hex grid itself.
#include <MeshFEM/CmdLine.hh>
#include <MeshFEM/MSHFieldWriter.hh>
#include <MeshFEM/filters/gen_grid.hh>
#include <MeshFEM/filters/hex_tet_subdiv.hh>
#include <MeshFEM/filters/quad_tri_subdiv.hh>

#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

using namespace std;

[[noreturn]] static void usage(int exitVal, const CmdLine &opts) {
    cout << "Usage: grid CxR[xS] out.msh [options]" << endl;
    opts.printOptions(cout);
    cout << endl;
    exit(exitVal);
}

static vector<string> split(const string &s, char sep) {
    vector<string> parts;
    string item;
    istringstream ss(s);
    while (getline(ss, item, sep)) parts.push_back(item);
    return parts;
}

static Point3D parseVector(size_t expectedSize, const string &cstring) {
    const runtime_error parseError("Invalid minCorner (must be comma-separated components)");
    const auto parts = split(cstring, ',');
    if (parts.size() != expectedSize || expectedSize > 3) throw parseError;
    Point3D result;
    try {
        for (size_t i = 0; i < expectedSize; ++i) result[i] = stof(parts[i]);   // the reference parses with stof
    } catch (...) { throw parseError; }
    return result;
}

int main(int argc, const char *argv[]) {
    CmdLine args;
    args.positional("gridSize").positional("outFile");
    args.flag("help", 0, "Produce this help message")
        .flag("tesselate", 't', "tesselate into tetrahedra or triangles")
        .value("minCorner", 'm', "minCorner of the grid bounding box (defaults to 0,0,0)")
        .value("maxCorner", 'M', "maxCorner of the grid bounding box (defaults to sx,sy,sz)");
    try {
        args.parse(argc, argv);
    } catch (std::exception &e) {
        cout << "Error: " << e.what() << endl << endl;
        usage(1, args);
    }
    if (args.count("help")) usage(0, args);
    if ((args.count("gridSize") == 0) || (args.count("outFile") == 0)) {
        cout << "Must specify grid size and output path";
        usage(1, args);
    }
    if (args.count("minCorner") != args.count("maxCorner")) {
        cout << "Must specify full bounding box" << endl;
        usage(1, args);
    }

    try {
        vector<MeshIO::IOVertex> gridVertices, simplexVertices;
        vector<MeshIO::IOElement> gridElements, simplices;
        vector<size_t> sizes;
        for (const auto &s : split(args.str("gridSize"), 'x')) sizes.push_back(stoul(s));
        gen_grid(sizes, gridVertices, gridElements);

        if (args.count("minCorner")) {
            const Point3D minCorner = parseVector(sizes.size(), args.str("minCorner"));
            const Point3D maxCorner = parseVector(sizes.size(), args.str("maxCorner"));
            Point3D scale = maxCorner - minCorner;   // current grid is [0, sx] x ...
            for (size_t i = 0; i < sizes.size(); ++i) scale[i] /= sizes[i];
            for (auto &v : gridVertices)
                for (size_t i = 0; i < 3; ++i) v.point[i] = scale[i] * v.point[i] + minCorner[i];
        }

        const string outPath = args.str("outFile");
        if (args.count("tesselate")) {
            vector<size_t> cellIdx;
            if (sizes.size() == 2) quad_tri_subdiv(gridVertices, gridElements, simplexVertices, simplices, cellIdx);
            else hex_tet_subdiv(gridVertices, gridElements, simplexVertices, simplices, cellIdx);
            MSHFieldWriter writer(outPath, simplexVertices, simplices);
            cout << "Writing mesh file..." << endl;
            ScalarField<double> cell_index(cellIdx.size());
            for (size_t i = 0; i < cellIdx.size(); ++i) cell_index[i] = cellIdx[i];
            writer.addField("cell_index", cell_index, DomainType::PER_ELEMENT);
        } else {
            const MeshIO::MeshType type = (sizes.size() == 2) ? MeshIO::MESH_QUAD : MeshIO::MESH_HEX;
            MeshIO::save(outPath, gridVertices, gridElements, MeshIO::FMT_GUESS, type);
        }
    } catch (const std::exception &e) {
        std::cerr << "terminate called after throwing an instance of 'std::runtime_error'\n  what():  " << e.what() << std::endl;
        return 1;
    }
    return 0;
}

// hopeDgToVTK - legacy-VTK export of DG fields, the step after the solver (counterpart of HopeFOAM-0.1/applications/utilities/DG/dgToVTK,
// dgToVTK.C:27-50: every high-order element is sub-triangulated on its node lattice).  Host-only: needs no GPU.
//   hopeDgToVTK -case <caseDir> -time <timeName> [field ...]      ->  <caseDir>/VTK/<case>_<timeName>.vtk
// Points = the K*Np nodal locations (element-contiguous, discontinuous across elements), cells = N^2 linear triangles per element,
// POINT_DATA = the requested dgScalarField / dgVectorField files of the time directory (default: every field file found there).
#include <dirent.h>

#include "dgCFD.H"

using namespace Foam;

static bool readField(const fileName& path, label nDof, std::vector<double>& data, int& nCmpt)
{
    std::ifstream probe(path);
    if (!probe) return false;
    const dictionary d = dictionary::fromFile(path);
    if (!d.found("internalField")) return false;
    const ITstream& in = d.lookup("internalField");
    size_t i = 1;
    auto isVec = [&](size_t k) { return k < in.size() && in[k] == "("; };
    if (in[0] == "uniform") {
        nCmpt = isVec(1) ? 3 : 1;
        std::vector<double> v;
        if (nCmpt == 3) { for (int c = 0; c < 3; ++c) v.push_back(std::strtod(in[2 + c].c_str(), nullptr)); }
        else v.push_back(std::strtod(in[1].c_str(), nullptr));
        data.resize((size_t)nDof * nCmpt);
        for (label k = 0; k < nDof; ++k) for (int c = 0; c < nCmpt; ++c) data[(size_t)k * nCmpt + c] = v[c];
        return true;
    }
    while (i < in.size() && in[i] != "(") ++i;      // nonuniform List<...> N (
    ++i;
    nCmpt = isVec(i) ? 3 : 1;
    data.resize((size_t)nDof * nCmpt);
    for (label k = 0; k < nDof; ++k) {
        if (nCmpt == 3) ++i;
        for (int c = 0; c < nCmpt; ++c) data[(size_t)k * nCmpt + c] = std::strtod(in[i++].c_str(), nullptr);
        if (nCmpt == 3) ++i;
    }
    return true;
}

int main(int argc, char* argv[])
{
    argList args(argc, argv);
    word timeName = "0";
    std::vector<word> fields;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-case") ++i;
        else if (a == "-time" && i + 1 < argc) timeName = argv[++i];
        else if (a[0] != '-') fields.push_back(a);
    }
    Time runTime(args);
    dgMesh mesh(runTime, true);
    const label N = mesh.baseOrder(), Np = mesh.nDofPerCell(), K = mesh.nCells();
    const fileName tdir = runTime.path() + "/" + timeName;
    if (fields.empty()) {
        if (DIR* dp = opendir(tdir.c_str())) {
            while (dirent* e = readdir(dp)) if (e->d_name[0] != '.') fields.push_back(e->d_name);
            closedir(dp);
        } else
            FatalErrorInFunction << "cannot open time directory " << tdir << abort(FatalError);
        std::sort(fields.begin(), fields.end());
    }
    // node (n, m) of the lattice: rows of constant s (n), r increasing (m)  (triangleBaseFunction.C:134-142)
    std::vector<label> rowStart(N + 2, 0);
    for (label n = 0; n <= N; ++n) rowStart[n + 1] = rowStart[n] + (N + 1 - n);
    std::vector<label> sub;
    for (label n = 0; n < N; ++n)
        for (label m = 0; m < N - n; ++m) {
            const label a = rowStart[n] + m, b = a + 1, c = rowStart[n + 1] + m;
            sub.insert(sub.end(), {a, b, c});
            if (m < N - n - 1) sub.insert(sub.end(), {b, c + 1, c});
        }
    const label nSub = (label)sub.size() / 3;     // = N^2
    const List<vector> px = mesh.dofLocation();

    const fileName vdir = runTime.path() + "/VTK";
    ::mkdir(vdir.c_str(), 0777);
    std::string caseName = runTime.path();
    while (!caseName.empty() && caseName.back() == '/') caseName.pop_back();
    caseName = caseName.substr(caseName.find_last_of('/') + 1);
    const fileName out = vdir + "/" + caseName + "_" + timeName + ".vtk";
    std::ofstream os(out);
    os << std::setprecision(12);
    os << "# vtk DataFile Version 2.0\nhopeDgToVTK " << caseName << " time " << timeName << " order " << N << "\nASCII\nDATASET UNSTRUCTURED_GRID\n";
    os << "POINTS " << px.size() << " double\n";
    forAll(px, i) os << px[i].x() << ' ' << px[i].y() << " 0\n";
    os << "CELLS " << K * nSub << ' ' << (long)K * nSub * 4 << "\n";
    for (label k = 0; k < K; ++k)
        for (label t = 0; t < nSub; ++t) os << "3 " << k * Np + sub[3 * t] << ' ' << k * Np + sub[3 * t + 1] << ' ' << k * Np + sub[3 * t + 2] << "\n";
    os << "CELL_TYPES " << K * nSub << "\n";
    for (label k = 0; k < K * nSub; ++k) os << "5\n";
    os << "POINT_DATA " << px.size() << "\n";
    label nWritten = 0;
    for (const word& f : fields) {
        std::vector<double> data;
        int nCmpt = 0;
        if (!readField(tdir + "/" + f, K * Np, data, nCmpt)) continue;
        if (nCmpt == 1) {
            os << "SCALARS " << f << " double 1\nLOOKUP_TABLE default\n";
            for (double v : data) os << v << "\n";
        } else {
            os << "VECTORS " << f << " double\n";
            for (size_t i = 0; i < data.size(); i += 3) os << data[i] << ' ' << data[i + 1] << ' ' << data[i + 2] << "\n";
        }
        ++nWritten;
    }
    Info << "wrote " << out << " : " << px.size() << " points, " << K * nSub << " triangles, " << nWritten << " fields" << endl;
    return 0;
}

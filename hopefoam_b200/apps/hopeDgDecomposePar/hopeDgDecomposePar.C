// hopeDgDecomposePar - splits a case for a `-parallel` run (counterpart of HopeFOAM-0.1/applications/utilities/DG/dgDecomposePar for the
// explicit 2-D path).  Host-only: needs no GPU.
//   hopeDgDecomposePar -case <caseDir> [-time <timeName>]
// reads system/decomposeParDict (numberOfSubdomains; method simple | manual) and writes, for every rank r,
//   processor<r>/constant/polyMesh/{points,faces,owner,neighbour,boundary,cellProcAddressing,pointProcAddressing,boundaryProcAddressing}
//   processor<r>/<time>/<every field file of <case>/<time>>
// The cell / point / patch ordering rules are dgDecomposePar's (domainDecompositionMesh.C:102-511, implemented in csrc/mesh.cpp:
// cells and points in ascending global id, original patches kept in order, one processor patch per neighbour in ascending rank, its faces
// in ascending global face id on both sides).  The polyMesh is written as one layer of prisms over the rank's triangles: internal faces
// in upper-triangular order, then the patches, `frontAndBackPlanes` (empty), then the processor patches (myProcNo / neighbProcNo).
// faceProcAddressing is not written (the lateral faces are regenerated, not copied from the global polyMesh).
#include <dirent.h>

#include <algorithm>

#include "dgCFD.H"

using namespace Foam;

namespace
{

const char* banner =
    "/*--------------------------------*- C++ -*----------------------------------*\\\n"
    "| hopeDgDecomposePar                                                          |\n"
    "\\*---------------------------------------------------------------------------*/\n";

void header(std::ostream& os, const word& cls, const word& location, const word& object)
{
    os << banner << "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       " << cls << ";\n    location    \"" << location
       << "\";\n    object      " << object << ";\n}\n// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //\n\n";
}

void mkdirs(const fileName& path)
{
    for (size_t i = 1; i <= path.size(); ++i)
        if (i == path.size() || path[i] == '/') ::mkdir(path.substr(0, i).c_str(), 0777);
}

void writeLabelList(const fileName& dir, const word& name, const std::vector<int32_t>& v)
{
    std::ofstream os(dir + "/" + name);
    header(os, "labelList", "constant/polyMesh", name);
    os << v.size() << "\n(\n";
    for (int32_t x : v) os << x << "\n";
    os << ")\n";
}

struct ProcMesh
{
    int64_t K = 0, F = 0, P = 0;
    int32_t nPatches = 0;
    std::vector<double> xy;
    std::vector<int32_t> tris, own, nbr, locO, locN, cellAddr, pointAddr, patchNbr, patchFaceGlobal;
    std::vector<word> names, types;
    std::vector<std::vector<int32_t>> patchFaces;
};

ProcMesh query(hdg_context* c)
{
    ProcMesh m;
    int64_t ng;
    hdg_mesh_counts(c, &m.K, &m.F, &m.nPatches, &ng);
    m.P = hdg_mesh_num_points(c);
    m.xy.resize((size_t)m.P * 2);
    hdg_mesh_get_points(c, m.xy.data());
    m.tris.resize((size_t)m.K * 3);
    hdg_mesh_get_cell_vertices(c, m.tris.data());
    m.own.resize(m.F); m.nbr.resize(m.F); m.locO.resize(m.F); m.locN.resize(m.F);
    hdg_mesh_get_faces(c, m.own.data(), m.nbr.data(), m.locO.data(), m.locN.data(), nullptr);
    size_t nPatchFaces = 0;
    for (int32_t p = 0; p < m.nPatches; ++p) {
        char nm[128], ty[64];
        int32_t nf;
        hdg_mesh_patch_info(c, p, nm, 128, ty, 64, &nf);
        m.names.push_back(nm);
        m.types.push_back(ty);
        m.patchFaces.emplace_back((size_t)nf);
        if (nf) hdg_mesh_patch_faces(c, p, m.patchFaces.back().data());
        nPatchFaces += nf;
    }
    m.cellAddr.resize(m.K); m.pointAddr.resize(m.P); m.patchNbr.resize(m.nPatches); m.patchFaceGlobal.resize(nPatchFaces);
    hdg_mesh_proc_addressing(c, m.cellAddr.data(), m.pointAddr.data(), m.patchNbr.data(), m.patchFaceGlobal.data());
    return m;
}

// the lateral quad over the edge that cell `c` traverses counter-clockwise as its local face `loc`: outward normal for that cell
void quad(std::ostream& os, const ProcMesh& m, int32_t c, int32_t loc)
{
    const int32_t a = m.tris[3 * c + loc], b = m.tris[3 * c + (loc + 1) % 3];
    os << "4(" << a << ' ' << b << ' ' << b + m.P << ' ' << a + m.P << ")\n";
}

void writePolyMesh(const fileName& dir, const ProcMesh& m, int32_t rank, int32_t nGlobalPoints, int32_t nGlobalPatches)
{
    mkdirs(dir);
    {
        std::ofstream os(dir + "/points");
        header(os, "vectorField", "constant/polyMesh", "points");
        os << std::setprecision(17) << 2 * m.P << "\n(\n";
        for (int z = 0; z < 2; ++z)
            for (int64_t i = 0; i < m.P; ++i) os << '(' << m.xy[2 * i] << ' ' << m.xy[2 * i + 1] << ' ' << z << ")\n";
        os << ")\n";
    }
    // internal faces, upper-triangular order: owner = the lower cell, sorted by (owner, neighbour)
    struct Int { int32_t o, n, cell, loc; };
    std::vector<Int> internal;
    for (int64_t f = 0; f < m.F; ++f)
        if (m.nbr[f] >= 0) {
            if (m.own[f] < m.nbr[f]) internal.push_back({m.own[f], m.nbr[f], m.own[f], m.locO[f]});
            else internal.push_back({m.nbr[f], m.own[f], m.nbr[f], m.locN[f]});
        }
    std::sort(internal.begin(), internal.end(), [](const Int& a, const Int& b) { return a.o != b.o ? a.o < b.o : a.n < b.n; });
    std::vector<int32_t> owner, neighbour;
    struct Block { word name, type; int32_t n, start, nbrProc; };
    std::vector<Block> blocks, procBlocks;
    std::ofstream fs(dir + "/faces");
    header(fs, "faceList", "constant/polyMesh", "faces");
    size_t nPatchFaces = 0;
    for (const auto& pf : m.patchFaces) nPatchFaces += pf.size();
    fs << internal.size() + nPatchFaces + 2 * m.K << "\n(\n";
    for (const Int& i : internal) { quad(fs, m, i.cell, i.loc); owner.push_back(i.o); neighbour.push_back(i.n); }
    auto patchBlock = [&](int32_t p) {
        const int32_t start = (int32_t)owner.size();
        for (int32_t f : m.patchFaces[p]) { quad(fs, m, m.own[f], m.locO[f]); owner.push_back(m.own[f]); }
        return Block{m.names[p], m.types[p], (int32_t)m.patchFaces[p].size(), start, m.patchNbr[p]};
    };
    for (int32_t p = 0; p < m.nPatches; ++p)
        if (m.patchNbr[p] < 0) blocks.push_back(patchBlock(p));
    {   // base plane z == 0 (outward normal -z: v0 v2 v1), then the top plane
        const int32_t start = (int32_t)owner.size();
        for (int64_t c = 0; c < m.K; ++c) {
            fs << "3(" << m.tris[3 * c] << ' ' << m.tris[3 * c + 2] << ' ' << m.tris[3 * c + 1] << ")\n";
            owner.push_back((int32_t)c);
        }
        for (int64_t c = 0; c < m.K; ++c) {
            fs << "3(" << m.tris[3 * c] + m.P << ' ' << m.tris[3 * c + 1] + m.P << ' ' << m.tris[3 * c + 2] + m.P << ")\n";
            owner.push_back((int32_t)c);
        }
        blocks.push_back(Block{"frontAndBackPlanes", "empty", (int32_t)(2 * m.K), start, -1});
    }
    for (int32_t p = 0; p < m.nPatches; ++p)
        if (m.patchNbr[p] >= 0) blocks.push_back(patchBlock(p));
    fs << ")\n";
    writeLabelList(dir, "owner", owner);
    writeLabelList(dir, "neighbour", neighbour);
    {
        std::ofstream os(dir + "/boundary");
        header(os, "polyBoundaryMesh", "constant/polyMesh", "boundary");
        os << blocks.size() << "\n(\n";
        for (const Block& b : blocks) {
            os << "    " << b.name << "\n    {\n        type            " << b.type << ";\n        nFaces          " << b.n
               << ";\n        startFace       " << b.start << ";\n";
            if (b.nbrProc >= 0) os << "        myProcNo        " << rank << ";\n        neighbProcNo    " << b.nbrProc << ";\n";
            os << "    }\n";
        }
        os << ")\n";
    }
    writeLabelList(dir, "cellProcAddressing", m.cellAddr);
    std::vector<int32_t> pa(2 * m.P);
    for (int64_t i = 0; i < m.P; ++i) { pa[i] = m.pointAddr[i]; pa[i + m.P] = m.pointAddr[i] + nGlobalPoints; }
    writeLabelList(dir, "pointProcAddressing", pa);
    std::vector<int32_t> ba;      // original patches keep their index, the empty patch follows them, processor patches map to -1
    for (int32_t p = 0; p < m.nPatches; ++p) if (m.patchNbr[p] < 0) ba.push_back(p);
    ba.push_back(nGlobalPatches);
    for (int32_t p = 0; p < m.nPatches; ++p) if (m.patchNbr[p] >= 0) ba.push_back(-1);
    writeLabelList(dir, "boundaryProcAddressing", ba);
}

// values of a `uniform v` / `nonuniform List<..> n ( ... )` stream as text tokens per entity (a vector keeps its parentheses)
bool listTokens(const ITstream& in, std::vector<std::string>& items, bool& uniform)
{
    uniform = !in.empty() && in[0] == "uniform";
    size_t i = 1;
    if (!uniform) {
        while (i < in.size() && in[i] != "(") ++i;
        ++i;
    }
    while (i < in.size()) {
        if (in[i] == ")") break;
        if (in[i] == "(") {
            std::string v = "(";
            for (++i; i < in.size() && in[i] != ")"; ++i) v += (v.size() > 1 ? " " : "") + in[i];
            items.push_back(v + ")");
            ++i;
        } else
            items.push_back(in[i++]);
        if (uniform) break;
    }
    return !items.empty();
}

}  // namespace

int main(int argc, char* argv[])
{
    argList args(argc, argv);
    word timeName = "0";
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-case") ++i;
        else if (a == "-time" && i + 1 < argc) timeName = argv[++i];
    }
    Time runTime(args);
    dgMesh mesh(runTime, true);
    const fileName root = runTime.rootPath();
    const label K = mesh.nCells(), Np = mesh.nDofPerCell(), Nfp = mesh.nDofPerFace();
    std::vector<int32_t> cellToProc((size_t)K);
    int32_t nProcs = 0;
    mesh.check(hdg_decompose_from_dict(mesh.ctx(), root.c_str(), &nProcs, cellToProc.data()), "decompositionMethod::New");
    Info << "Decomposing mesh " << dgMesh::defaultRegion << nl << nl << "Number of processors: " << nProcs << nl << endl;

    // global patch face lists: position of a dgFace inside its patch (for nonuniform boundary values)
    std::vector<std::vector<int32_t>> gPatchFaces((size_t)mesh.nPatches());
    for (label p = 0; p < mesh.nPatches(); ++p) {
        gPatchFaces[p].resize((size_t)mesh.patchNFaces(p));
        if (mesh.patchNFaces(p)) hdg_mesh_patch_faces(mesh.ctx(), p, gPatchFaces[p].data());
    }
    std::vector<word> fields;
    if (DIR* dp = opendir((root + "/" + timeName).c_str())) {
        while (dirent* e = readdir(dp)) if (e->d_name[0] != '.') fields.push_back(e->d_name);
        closedir(dp);
    }
    std::sort(fields.begin(), fields.end());
    const int32_t nGlobalPoints = (int32_t)hdg_mesh_num_points(mesh.ctx());

    for (int32_t r = 0; r < nProcs; ++r) {
        hdg_context* local = nullptr;
        if (hdg_create(-1, &local) != 0) FatalErrorInFunction << "cannot create a host context" << abort(FatalError);
        mesh.check(hdg_set_order(local, mesh.baseOrder()), "hopeDgDecomposePar");
        if (hdg_mesh_decompose(mesh.ctx(), nProcs, cellToProc.data(), r, local) != 0)
            FatalErrorInFunction << hdg_last_error(local) << abort(FatalError);
        const ProcMesh pm = query(local);
        const fileName pdir = root + "/processor" + std::to_string(r);
        writePolyMesh(pdir + "/constant/polyMesh", pm, r, nGlobalPoints, mesh.nPatches());
        Info << "Processor " << r << nl << "    Number of cells = " << pm.K << nl;
        for (int32_t p = 0; p < pm.nPatches; ++p)
            if (pm.patchNbr[p] >= 0) Info << "    Number of faces shared with processor " << pm.patchNbr[p] << " = " << pm.patchFaces[p].size() << nl;
        Info << endl;

        // ---- fields (dgFieldDecomposer: internal field through cellProcAddressing, patch fields through the face addressing) ----
        mkdirs(pdir + "/" + timeName);
        for (const word& fname : fields) {
            const dictionary d = dictionary::fromFile(root + "/" + timeName + "/" + fname);
            if (!d.found("internalField")) continue;
            std::vector<std::string> items;
            bool uni;
            listTokens(d.lookup("internalField"), items, uni);
            const bool isVec = !items.empty() && items[0][0] == '(';
            std::ofstream os(pdir + "/" + timeName + "/" + fname);
            header(os, isVec ? "dgVectorField" : "dgScalarField", timeName, fname);
            os << "dimensions      " << (d.found("dimensions") ? d.lookup("dimensions").str() : std::string("[0 0 0 0 0 0 0]")) << ";\n\n";
            if (uni) os << "internalField   uniform " << items[0] << ";\n\n";
            else {
                if ((label)items.size() != K * Np) FatalErrorInFunction << "internalField of " << fname << " has " << items.size() << " values, expected " << K * Np << abort(FatalError);
                os << "internalField   nonuniform List<" << (isVec ? "vector" : "scalar") << "> \n" << pm.K * Np << "\n(\n";
                for (int64_t c = 0; c < pm.K; ++c)
                    for (label i = 0; i < Np; ++i) os << items[(size_t)pm.cellAddr[c] * Np + i] << "\n";
                os << ")\n;\n\n";
            }
            os << "boundaryField\n{\n";
            const dictionary& bf = d.subDict("boundaryField");
            size_t off = 0;
            auto patchEntry = [&](int32_t p) {
                os << "    " << pm.names[p] << "\n    {\n";
                if (pm.patchNbr[p] >= 0) {
                    os << "        type            processor;\n        value           uniform " << (uni ? items[0] : (isVec ? std::string("(0 0 0)") : std::string("0"))) << ";\n";
                } else {
                    const dictionary& pd = bf.subDict(pm.names[p]);
                    os << "        type            " << pd.lookup("type")[0] << ";\n";
                    if (pd.found("value")) {
                        std::vector<std::string> pv;
                        bool puni;
                        listTokens(pd.lookup("value"), pv, puni);
                        if (puni) os << "        value           uniform " << pv[0] << ";\n";
                        else if (!pm.patchFaces[p].empty()) {
                            os << "        value           nonuniform List<" << (isVec ? "vector" : "scalar") << "> " << pm.patchFaces[p].size() * Nfp << "(";
                            for (size_t k = 0; k < pm.patchFaces[p].size(); ++k) {
                                const int32_t g = pm.patchFaceGlobal[off + k];
                                const auto& gl = gPatchFaces[p];
                                const size_t pos = std::find(gl.begin(), gl.end(), g) - gl.begin();
                                for (label i = 0; i < Nfp; ++i) os << (k || i ? " " : "") << pv[pos * Nfp + i];
                            }
                            os << ");\n";
                        }
                    }
                }
                os << "    }\n";
            };
            std::vector<size_t> offs((size_t)pm.nPatches);
            for (int32_t p = 0; p < pm.nPatches; ++p) { offs[p] = off; off += pm.patchFaces[p].size(); }
            for (int32_t p = 0; p < pm.nPatches; ++p)
                if (pm.patchNbr[p] < 0) { off = offs[p]; patchEntry(p); }
            os << "    frontAndBackPlanes\n    {\n        type            empty;\n    }\n";
            for (int32_t p = 0; p < pm.nPatches; ++p)
                if (pm.patchNbr[p] >= 0) { off = offs[p]; patchEntry(p); }
            os << "}\n";
        }
        hdg_destroy(local);
    }
    Info << "End" << nl << endl;
    return 0;
}

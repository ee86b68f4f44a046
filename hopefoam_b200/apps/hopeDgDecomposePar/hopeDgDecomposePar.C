// hopeDgDecomposePar - splits a case for a `-parallel` run (counterpart of HopeFOAM-0.1/applications/utilities/DG/dgDecomposePar for the
// explicit 2-D path).  Host-only: needs no GPU.
//   hopeDgDecomposePar -case <caseDir> [-time <timeName>]
// reads system/decomposeParDict (numberOfSubdomains; method simple | manual) and writes, for every rank r,
//   processor<r>/constant/polyMesh/{points,faces,owner,neighbour,boundary,cellProcAddressing,pointProcAddressing,boundaryProcAddressing}
//   processor<r>/<time>/<every field file of <case>/<time>>
// The polyMesh of a processor is cut out of the global polyMesh exactly as the reference does (domainDecompositionMesh.C:102-511,
// domainDecomposition.C:270-420): cells and points in ascending global id; faces = internal faces of the processor in ascending global
// face id, then every original patch in order (with the faces whose cell lives here, `empty` front/back planes included), then one
// processor patch per neighbour in ascending rank with the cut faces in ascending global face id, reversed on the side that holds the
// neighbour cell; points, faces and patch entries are copies of the global ones.  cell / point / face / boundaryProcAddressing are
// written in the reference's format (faceProcAddressing with the turning index: +-(global face + 1), domainDecomposition.C:1014-1027).
// The DG view of the same decomposition (dgFace order of the patches, used to split the boundary fields) comes from the library
// (hdg_mesh_decompose, csrc/mesh.cpp) and agrees with it by construction.
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

// ---- the global polyMesh as the reference's domainDecomposition sees it: 3-D points, point lists of the faces, owner / neighbour, patches
struct PolyPatch { word name; std::string body; int32_t nFaces = 0, startFace = 0; };      // body: the entry's text without nFaces / startFace
struct PolyMesh
{
    std::vector<std::string> points;                 // "(x y z)" kept as text: processor points are bit-identical copies
    std::vector<std::vector<int32_t>> faces;
    std::vector<int32_t> owner, neighbour;
    std::vector<PolyPatch> patches;
    int32_t nCells = 0;
};

std::string slurp(const fileName& path)
{
    std::ifstream in(path);
    if (!in) FatalErrorInFunction << "cannot open file " << path << abort(FatalError);
    std::stringstream ss;
    ss << in.rdbuf();
    std::string t = ss.str(), clean;
    for (size_t i = 0; i < t.size();) {      // strip comments
        if (t.compare(i, 2, "/*") == 0) { const size_t e = t.find("*/", i + 2); i = e == std::string::npos ? t.size() : e + 2; }
        else if (t.compare(i, 2, "//") == 0) { const size_t e = t.find('\n', i); i = e == std::string::npos ? t.size() : e; }
        else clean.push_back(t[i++]);
    }
    const size_t hdr = clean.find("FoamFile");
    if (hdr != std::string::npos) clean.erase(hdr, clean.find('}', hdr) - hdr + 1);
    return clean;
}

// position just after the '(' opening the top-level list; n = its declared size
size_t listOpen(const std::string& t, int64_t& n)
{
    size_t i = 0;
    while (i < t.size() && !std::isdigit((unsigned char)t[i])) ++i;
    char* end = nullptr;
    n = std::strtoll(t.c_str() + i, &end, 10);
    const size_t p = t.find('(', (size_t)(end - t.c_str()));
    if (p == std::string::npos) FatalErrorInFunction << "malformed list" << abort(FatalError);
    return p + 1;
}

std::vector<int32_t> readLabels(const fileName& path)
{
    const std::string t = slurp(path);
    int64_t n;
    const char* c = t.c_str() + listOpen(t, n);
    std::vector<int32_t> v((size_t)n);
    for (int64_t k = 0; k < n; ++k) { char* e; v[(size_t)k] = (int32_t)std::strtol(c, &e, 10); c = e; }
    return v;
}

PolyMesh readPolyMesh(const fileName& dir)
{
    PolyMesh m;
    {
        const std::string t = slurp(dir + "/points");
        int64_t n;
        size_t i = listOpen(t, n);
        for (int64_t k = 0; k < n; ++k) {
            const size_t a = t.find('(', i), b = t.find(')', a);
            m.points.push_back(t.substr(a, b - a + 1));
            i = b + 1;
        }
    }
    {
        const std::string t = slurp(dir + "/faces");
        int64_t n;
        const char* c = t.c_str() + listOpen(t, n);
        m.faces.resize((size_t)n);
        for (int64_t k = 0; k < n; ++k) {
            char* e;
            const long np = std::strtol(c, &e, 10);
            c = e;
            while (*c && *c != '(') ++c;
            ++c;
            for (long d = 0; d < np; ++d) { m.faces[(size_t)k].push_back((int32_t)std::strtol(c, &e, 10)); c = e; }
            while (*c && *c != ')') ++c;
            ++c;
        }
    }
    m.owner = readLabels(dir + "/owner");
    m.neighbour = readLabels(dir + "/neighbour");
    for (int32_t o : m.owner) m.nCells = std::max(m.nCells, o + 1);
    {
        const std::string t = slurp(dir + "/boundary");
        int64_t n;
        size_t i = listOpen(t, n);
        for (int64_t k = 0; k < n; ++k) {
            while (i < t.size() && std::isspace((unsigned char)t[i])) ++i;
            size_t j = i;
            while (j < t.size() && !std::isspace((unsigned char)t[j]) && t[j] != '{') ++j;
            PolyPatch P;
            P.name = t.substr(i, j - i);
            const size_t ob = t.find('{', j);
            size_t cb = ob + 1;      // matching brace; `#{ ... #}` code blocks of arc patches may hold braces of their own
            for (int depth = 1; cb < t.size(); ++cb) {
                if (t.compare(cb, 2, "#{") == 0) { cb = t.find("#}", cb) + 1; continue; }
                if (t[cb] == '{') ++depth;
                else if (t[cb] == '}' && --depth == 0) break;
            }
            const std::string body = t.substr(ob + 1, cb - ob - 1);
            // statements: keep everything but nFaces / startFace verbatim
            for (size_t q = 0; q < body.size();) {
                size_t e;
                const size_t code = body.find("#{", q), semi = body.find(';', q);
                if (semi == std::string::npos) break;
                if (code != std::string::npos && code < semi) e = body.find(';', body.find("#}", code));
                else e = semi;
                std::string stmt = body.substr(q, e - q + 1);
                q = e + 1;
                std::stringstream s2(stmt);
                std::string key, val;
                s2 >> key >> val;
                if (key == "nFaces") P.nFaces = std::atoi(val.c_str());
                else if (key == "startFace") P.startFace = std::atoi(val.c_str());
                else {
                    const size_t f = stmt.find_first_not_of(" \t\r\n");
                    if (f != std::string::npos) P.body += "        " + stmt.substr(f) + "\n";
                }
            }
            m.patches.push_back(P);
            i = cb + 1;
        }
    }
    return m;
}

// processor polyMesh of rank `rank`, exactly as domainDecomposition::decomposeMesh / writeDecomposition build it
// (domainDecompositionMesh.C:102-511, domainDecomposition.C:270-420, 1000-1060):
//   cells    ascending global id
//   faces    internal faces with both cells here (ascending global face id, no turning index) | for every patch in order the faces whose
//            cell lives here | per neighbour processor in ascending order the cut faces in ascending global face id, +(f+1) where this
//            processor holds the face's owner cell, -(f+1) (face reversed) where it holds the neighbour cell
//   points   the points of those faces in ascending global id
void writeProcPolyMesh(const fileName& dir, const PolyMesh& g, const std::vector<int32_t>& c2p, int32_t rank, const std::vector<int32_t>& cellAddr)
{
    mkdirs(dir);
    const size_t nInternal = g.neighbour.size();
    std::vector<int32_t> faceAddr;                       // faceProcAddressing: +-(global face + 1)
    for (size_t f = 0; f < nInternal; ++f)
        if (c2p[g.owner[f]] == rank && c2p[g.neighbour[f]] == rank) faceAddr.push_back((int32_t)f + 1);
    const int32_t nProcInternal = (int32_t)faceAddr.size();
    struct Block { word name; std::string body; int32_t n, start; };
    std::vector<Block> blocks;
    for (const PolyPatch& P : g.patches) {
        const int32_t start = (int32_t)faceAddr.size();
        for (int32_t f = P.startFace; f < P.startFace + P.nFaces; ++f)
            if (c2p[g.owner[f]] == rank) faceAddr.push_back(f + 1);
        blocks.push_back({P.name, P.body, (int32_t)faceAddr.size() - start, start});
    }
    std::vector<std::pair<int32_t, int32_t>> cuts;      // (neighbour processor, signed face index)
    for (size_t f = 0; f < nInternal; ++f) {
        const int32_t po = c2p[g.owner[f]], pn = c2p[g.neighbour[f]];
        if (po == pn) continue;
        if (po == rank) cuts.push_back({pn, (int32_t)f + 1});
        else if (pn == rank) cuts.push_back({po, -((int32_t)f + 1)});
    }
    std::stable_sort(cuts.begin(), cuts.end(), [](const std::pair<int32_t, int32_t>& a, const std::pair<int32_t, int32_t>& b) { return a.first < b.first; });
    for (size_t i = 0; i < cuts.size();) {
        size_t j = i;
        const int32_t start = (int32_t)faceAddr.size();
        for (; j < cuts.size() && cuts[j].first == cuts[i].first; ++j) faceAddr.push_back(cuts[j].second);
        const std::string q = std::to_string(cuts[i].first);
        blocks.push_back({"procBoundary" + std::to_string(rank) + "to" + q,
                          "        type            processor;\n        myProcNo        " + std::to_string(rank) + ";\n        neighbProcNo    " + q + ";\n",
                          (int32_t)(j - i), start});
        i = j;
    }
    // points
    std::vector<char> used(g.points.size(), 0);
    for (int32_t fa : faceAddr) for (int32_t pt : g.faces[(size_t)std::abs(fa) - 1]) used[(size_t)pt] = 1;
    std::vector<int32_t> pointAddr, lookup(g.points.size(), -1);
    for (size_t pt = 0; pt < g.points.size(); ++pt)
        if (used[pt]) { lookup[pt] = (int32_t)pointAddr.size(); pointAddr.push_back((int32_t)pt); }
    std::vector<int32_t> cellLookup((size_t)g.nCells, -1);
    for (size_t c = 0; c < cellAddr.size(); ++c) cellLookup[(size_t)cellAddr[c]] = (int32_t)c;
    {
        std::ofstream os(dir + "/points");
        header(os, "vectorField", "constant/polyMesh", "points");
        os << pointAddr.size() << "\n(\n";
        for (int32_t pt : pointAddr) os << g.points[(size_t)pt] << "\n";
        os << ")\n";
    }
    std::vector<int32_t> owner, neighbour;
    {
        std::ofstream os(dir + "/faces");
        header(os, "faceList", "constant/polyMesh", "faces");
        os << faceAddr.size() << "\n(\n";
        for (size_t i = 0; i < faceAddr.size(); ++i) {
            const int32_t fa = faceAddr[i];
            const size_t f = (size_t)std::abs(fa) - 1;
            std::vector<int32_t> pts = g.faces[f];
            if (fa < 0) std::reverse(pts.begin() + 1, pts.end());      // face::reverseFace keeps the first point
            os << pts.size() << '(';
            for (size_t k = 0; k < pts.size(); ++k) os << (k ? " " : "") << lookup[(size_t)pts[k]];
            os << ")\n";
            owner.push_back(cellLookup[(size_t)(fa > 0 ? g.owner[f] : g.neighbour[f])]);
            if ((int32_t)i < nProcInternal) neighbour.push_back(cellLookup[(size_t)g.neighbour[f]]);
        }
        os << ")\n";
    }
    writeLabelList(dir, "owner", owner);
    writeLabelList(dir, "neighbour", neighbour);
    {
        std::ofstream os(dir + "/boundary");
        header(os, "polyBoundaryMesh", "constant/polyMesh", "boundary");
        os << blocks.size() << "\n(\n";
        for (const Block& b : blocks)
            os << "    " << b.name << "\n    {\n" << b.body << "        nFaces          " << b.n << ";\n        startFace       " << b.start << ";\n    }\n";
        os << ")\n";
    }
    writeLabelList(dir, "cellProcAddressing", cellAddr);
    writeLabelList(dir, "pointProcAddressing", pointAddr);
    writeLabelList(dir, "faceProcAddressing", faceAddr);
    std::vector<int32_t> ba;      // original patches keep their index, processor patches map to -1
    for (size_t p = 0; p < g.patches.size(); ++p) ba.push_back((int32_t)p);
    for (size_t p = g.patches.size(); p < blocks.size(); ++p) ba.push_back(-1);
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
    const PolyMesh poly = readPolyMesh(runTime.constant() + "/polyMesh");
    if (poly.nCells != K) FatalErrorInFunction << "polyMesh has " << poly.nCells << " cells, the DG mesh " << K << abort(FatalError);

    for (int32_t r = 0; r < nProcs; ++r) {
        hdg_context* local = nullptr;
        if (hdg_create(-1, &local) != 0) FatalErrorInFunction << "cannot create a host context" << abort(FatalError);
        mesh.check(hdg_set_order(local, mesh.baseOrder()), "hopeDgDecomposePar");
        if (hdg_mesh_decompose(mesh.ctx(), nProcs, cellToProc.data(), r, local) != 0)
            FatalErrorInFunction << hdg_last_error(local) << abort(FatalError);
        const ProcMesh pm = query(local);
        const fileName pdir = root + "/processor" + std::to_string(r);
        writeProcPolyMesh(pdir + "/constant/polyMesh", poly, cellToProc, r, pm.cellAddr);
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
            for (int32_t p = 0; p < pm.nPatches; ++p)
                if (pm.patchNbr[p] >= 0) { off = offs[p]; patchEntry(p); }
            os << "}\n";
        }
        hdg_destroy(local);
    }
    Info << "End" << nl << endl;
    return 0;
}

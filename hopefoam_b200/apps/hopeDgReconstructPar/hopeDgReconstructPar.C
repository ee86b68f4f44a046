// hopeDgReconstructPar - merges the fields a `-parallel` run wrote into processorN/<time>/ back into <case>/<time>/, the step after the
// run (counterpart of HopeFOAM-0.1/applications/utilities/DG/dgReconstructPar: dgFieldReconstructor maps the internal field of every
// processor through cellProcAddressing).  Host-only: needs no GPU.
//   hopeDgReconstructPar -case <caseDir> [-time <timeName>] [field ...]
// A cell keeps its nodal values; because the vertex order of a cell may differ between the processor polyMesh and the global one, the Np
// nodes of a cell are matched by their coordinates.  Boundary fields: the patch types are carried over and the `value` lists of the
// original patches are mapped back from the processors that own their faces (dgFieldReconstructor maps patch fields through
// faceProcAddressing; here the patch nodes are matched by coordinates like the cell nodes), so that a restart from the reconstructed
// time directory keeps its inflow / fixed boundary data.  A patch some processor wrote no value for falls back to type `calculated`.
#include <dirent.h>

#include "dgCFD.H"

using namespace Foam;

static labelList readLabelList(const fileName& path)
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
    size_t pos = 0;
    if (hdr != std::string::npos) pos = clean.find('}', hdr) + 1;
    const size_t open = clean.find('(', pos);
    if (open == std::string::npos) FatalErrorInFunction << path << " holds no list" << abort(FatalError);
    const label n = std::atoi(clean.substr(pos, open - pos).c_str());
    labelList out(n);
    std::istringstream is(clean.substr(open + 1));
    for (label i = 0; i < n; ++i) is >> out[i];
    return out;
}

static bool readInternal(const dictionary& d, label nDof, std::vector<double>& data, int& nCmpt, const char* key = "internalField")
{
    if (!d.found(key)) return false;
    const ITstream& in = d.lookup(key);
    size_t i = 1;
    auto isVec = [&](size_t k) { return k < in.size() && in[k] == "("; };
    if (in[0] == "uniform") {
        nCmpt = isVec(1) ? 3 : 1;
        std::vector<double> v;
        for (int c = 0; c < nCmpt; ++c) v.push_back(std::strtod(in[(nCmpt == 3 ? 2 : 1) + c].c_str(), nullptr));
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
    word timeName;
    std::vector<word> fields;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "-case") ++i;
        else if (a == "-time" && i + 1 < argc) timeName = argv[++i];
        else if (a[0] != '-') fields.push_back(a);
    }
    Time runTime(args);
    const fileName root = runTime.rootPath();
    label nProcs = 0;
    for (;; ++nProcs) {
        struct stat st;
        if (::stat((root + "/processor" + std::to_string(nProcs)).c_str(), &st) != 0) break;
    }
    if (nProcs == 0) FatalErrorInFunction << "No processor directories found in " << root << abort(FatalError);
    if (timeName.empty()) {      // latest time of processor0
        scalar best = -1;
        if (DIR* dp = opendir((root + "/processor0").c_str())) {
            while (dirent* e = readdir(dp)) {
                char* end = nullptr;
                const scalar v = std::strtod(e->d_name, &end);
                if (end != e->d_name && *end == '\0' && v > best) { best = v; timeName = e->d_name; }
            }
            closedir(dp);
        }
        if (timeName.empty()) FatalErrorInFunction << "no time directory in " << root << "/processor0" << abort(FatalError);
    }
    Info << "Reconstructing fields for time " << timeName << " from " << nProcs << " processors" << nl << endl;

    dgMesh mesh(runTime, true);
    const label Np = mesh.nDofPerCell(), K = mesh.nCells();
    const List<vector> px = mesh.dofLocation();
    if (fields.empty()) {
        if (DIR* dp = opendir((root + "/processor0/" + timeName).c_str())) {
            while (dirent* e = readdir(dp)) if (e->d_name[0] != '.') fields.push_back(e->d_name);
            closedir(dp);
        }
        std::sort(fields.begin(), fields.end());
    }
    std::vector<std::vector<double>> glob(fields.size());
    // boundary values of the original patches: [field][patch] -> values in the global patch-dof order, filled[...] counts the nodes set
    const label nPatch = mesh.nPatches();
    std::vector<List<vector>> gpx((size_t)nPatch);
    for (label p = 0; p < nPatch; ++p) gpx[p] = mesh.patchDofLocation(p);
    std::vector<std::vector<std::vector<double>>> bval(fields.size(), std::vector<std::vector<double>>((size_t)nPatch));
    std::vector<std::vector<label>> bfilled(fields.size(), std::vector<label>((size_t)nPatch, 0));
    std::vector<std::vector<char>> bmissing(fields.size(), std::vector<char>((size_t)nPatch, 0));
    std::vector<int> nCmpt(fields.size(), 0);
    std::vector<dictionary> first(fields.size());
    std::vector<char> seen((size_t)K, 0);
    for (label r = 0; r < nProcs; ++r) {
        Pstream::setParallel(r, nProcs, 0, root, "reconstruct");      // Time::path() now points into processor<r>
        dgMesh pm(runTime, true);
        const labelList addr = readLabelList(runTime.constant() + "/polyMesh/cellProcAddressing");
        if (addr.size() != pm.nCells()) FatalErrorInFunction << "cellProcAddressing of processor" << r << " does not match its mesh" << abort(FatalError);
        const List<vector> lx = pm.dofLocation();
        // node permutation of every local cell onto its global cell
        std::vector<label> perm((size_t)addr.size() * Np);
        for (label k = 0; k < addr.size(); ++k) {
            const label g = addr[k];
            if (g < 0 || g >= K) FatalErrorInFunction << "cellProcAddressing entry " << g << " out of range" << abort(FatalError);
            seen[g] = 1;
            for (label i = 0; i < Np; ++i) {
                label bestJ = 0;
                scalar bestD = GREAT;
                for (label j = 0; j < Np; ++j) {
                    const scalar dd = magSqr(lx[k * Np + i] - px[g * Np + j]);
                    if (dd < bestD) { bestD = dd; bestJ = j; }
                }
                perm[(size_t)k * Np + i] = g * Np + bestJ;
            }
        }
        // local patch node -> global patch node (original patches keep their index and name on every processor).  Face end nodes
        // coincide with those of the next face, so FACES are matched first (by the mean of their Nfp nodes, unique per face) and
        // the nodes inside a matched face by their coordinates.
        std::vector<std::vector<label>> pperm((size_t)nPatch);
        const label Nfp = mesh.nDofPerFace();
        auto faceCentre = [&](const List<vector>& x, label fc) { vector c = x[fc * Nfp]; for (label i = 1; i < Nfp; ++i) c = c + x[fc * Nfp + i]; return c * (1.0 / Nfp); };
        for (label p = 0; p < nPatch && p < pm.nPatches(); ++p) {
            if (pm.patchName(p) != mesh.patchName(p)) continue;
            const List<vector> lpx = pm.patchDofLocation(p);
            const label nGF = gpx[p].size() / Nfp, nLF = lpx.size() / Nfp;
            std::vector<vector> gc((size_t)nGF);
            for (label fc = 0; fc < nGF; ++fc) gc[fc] = faceCentre(gpx[p], fc);
            std::vector<label> order((size_t)nGF);
            for (label i = 0; i < nGF; ++i) order[i] = i;
            std::sort(order.begin(), order.end(), [&](label a, label b) { return gc[a].x() < gc[b].x(); });
            pperm[p].assign((size_t)lpx.size(), -1);
            for (label fc = 0; fc < nLF; ++fc) {
                const vector c = faceCentre(lpx, fc);
                const scalar tol = 1e-9 * (1.0 + mag(c));
                auto lo = std::lower_bound(order.begin(), order.end(), c.x() - tol, [&](label a, scalar v) { return gc[a].x() < v; });
                label gf = -1;
                scalar bestD = GREAT;
                for (auto it = lo; it != order.end() && gc[*it].x() <= c.x() + tol; ++it) {
                    const scalar dd = magSqr(c - gc[*it]);
                    if (dd < bestD) { bestD = dd; gf = *it; }
                }
                if (gf < 0 || bestD > tol * tol)
                    FatalErrorInFunction << "patch " << mesh.patchName(p) << ": a face of processor" << r << " is not on the undecomposed patch" << abort(FatalError);
                for (label i = 0; i < Nfp; ++i) {
                    label bestJ = 0;
                    scalar bd = GREAT;
                    for (label jn = 0; jn < Nfp; ++jn) {
                        const scalar dd = magSqr(lpx[fc * Nfp + i] - gpx[p][gf * Nfp + jn]);
                        if (dd < bd) { bd = dd; bestJ = jn; }
                    }
                    pperm[p][(size_t)fc * Nfp + i] = gf * Nfp + bestJ;
                }
            }
        }
        for (size_t f = 0; f < fields.size(); ++f) {
            const fileName path = runTime.path() + "/" + timeName + "/" + fields[f];
            std::ifstream probe(path);
            if (!probe) FatalErrorInFunction << "cannot open file " << path << abort(FatalError);
            const dictionary d = dictionary::fromFile(path);
            std::vector<double> loc;
            int nc = 0;
            if (!readInternal(d, addr.size() * Np, loc, nc)) FatalErrorInFunction << path << " has no internalField" << abort(FatalError);
            if (r == 0) { nCmpt[f] = nc; glob[f].assign((size_t)K * Np * nc, 0.0); first[f] = d; }
            if (nc != nCmpt[f]) FatalErrorInFunction << "field " << fields[f] << " changes its type between processors" << abort(FatalError);
            for (size_t q = 0; q < perm.size(); ++q)
                for (int c = 0; c < nc; ++c) glob[f][(size_t)perm[q] * nc + c] = loc[q * nc + c];
            const dictionary* lbf = d.isDict("boundaryField") ? &d.subDict("boundaryField") : nullptr;
            for (label p = 0; p < nPatch; ++p) {
                if (pperm[p].empty()) continue;                       // this processor owns no face of the patch
                std::vector<double> pv;
                int pnc = 0;
                if (!lbf || !lbf->isDict(mesh.patchName(p)) ||
                    !readInternal(lbf->subDict(mesh.patchName(p)), (label)pperm[p].size(), pv, pnc, "value") || pnc != nc) {
                    bmissing[f][p] = 1;
                    continue;
                }
                if (bval[f][p].empty()) bval[f][p].assign((size_t)gpx[p].size() * nc, 0.0);
                for (size_t q = 0; q < pperm[p].size(); ++q)
                    for (int c = 0; c < nc; ++c) bval[f][p][(size_t)pperm[p][q] * nc + c] = pv[q * nc + c];
                bfilled[f][p] += (label)pperm[p].size();
            }
        }
    }
    Pstream::setSerial();
    for (label g = 0; g < K; ++g)
        if (!seen[g]) FatalErrorInFunction << "cell " << g << " belongs to no processor" << abort(FatalError);

    const fileName tdir = root + "/" + timeName;
    ::mkdir(tdir.c_str(), 0777);
    for (size_t f = 0; f < fields.size(); ++f) {
        const int nc = nCmpt[f];
        std::ofstream os(tdir + "/" + fields[f]);
        // merged data must survive the round trip: never fewer digits than a double holds unless the case asks for more
        os << std::setprecision(std::max<label>(17, runTime.controlDict().lookupOrDefault<label>("writePrecision", 6)));
        os << "FoamFile\n{\n    version     2.0;\n    format      ascii;\n    class       " << (nc == 3 ? "dgVectorField" : "dgScalarField")
           << ";\n    location    \"" << timeName << "\";\n    object      " << fields[f] << ";\n}\n\n";
        os << "dimensions      " << (first[f].found("dimensions") ? first[f].lookup("dimensions").str() : std::string("[0 0 0 0 0 0 0]")) << ";\n\n";
        os << "internalField   nonuniform List<" << (nc == 3 ? "vector" : "scalar") << "> \n" << K * Np << "\n(\n";
        for (label q = 0; q < K * Np; ++q) {
            if (nc == 3) os << '(' << glob[f][(size_t)q * 3] << ' ' << glob[f][(size_t)q * 3 + 1] << ' ' << glob[f][(size_t)q * 3 + 2] << ")\n";
            else os << glob[f][q] << "\n";
        }
        os << ")\n;\n\nboundaryField\n{\n";
        const dictionary* bf = first[f].isDict("boundaryField") ? &first[f].subDict("boundaryField") : nullptr;
        for (label p = 0; p < mesh.nPatches(); ++p) {
            word type = "calculated";
            if (bf && bf->isDict(mesh.patchName(p))) type = bf->subDict(mesh.patchName(p)).lookup("type")[0];
            const bool haveValues = !bmissing[f][p] && bfilled[f][p] == gpx[p].size() && gpx[p].size() > 0;
            if (type == "fixedValue" && !haveValues && gpx[p].size() > 0) {
                Info << "--> hopeDgReconstructPar Warning: field " << fields[f] << ", patch " << mesh.patchName(p)
                     << ": the processors hold no complete value list; written as type calculated" << endl;
                type = "calculated";
            }
            os << "    " << mesh.patchName(p) << "\n    {\n        type            " << type << ";\n";
            if (haveValues) {
                os << "        value           nonuniform List<" << (nc == 3 ? "vector" : "scalar") << "> " << gpx[p].size() << "(";
                for (label q = 0; q < gpx[p].size(); ++q) {
                    if (nc == 3) os << '(' << bval[f][p][(size_t)q * 3] << ' ' << bval[f][p][(size_t)q * 3 + 1] << ' ' << bval[f][p][(size_t)q * 3 + 2] << ") ";
                    else os << bval[f][p][q] << ' ';
                }
                os << ");\n";
            }
            os << "    }\n";
        }
        os << "}\n";
        Info << "    " << fields[f] << nl;
    }
    Info << nl << "End" << nl << endl;
    return 0;
}

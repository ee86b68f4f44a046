// See mesh.hpp for the reference citations.
#include "mesh.hpp"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace hdg {

namespace {
struct EdgeRec {
    int32_t a, b;      // canonical point ids, a <= b
    int32_t cell, face;
};
}  // namespace

void Mesh::build(int64_t nPts, const double* pxy, int64_t nK, const int32_t* ptris, const int32_t* pointEquiv,
                 int nPatches, const int32_t* patchStart, const int32_t* edgeCell, const int32_t* edgePoints,
                 const std::vector<std::string>* names, const std::vector<std::string>* types)
{
    if (nK <= 0 || nPts <= 0) throw std::runtime_error("empty mesh");
    K = nK;
    nPoints = nPts;
    periodicGlue = pointEquiv != nullptr;
    if (pointEquiv) this->pointEquiv.assign(pointEquiv, pointEquiv + nPts);
    else this->pointEquiv.clear();
    polyFace.clear();
    xy.assign(pxy, pxy + 2 * nPts);
    tris.assign(ptris, ptris + 3 * nK);
    auto canon = [&](int32_t p) { return pointEquiv ? pointEquiv[p] : p; };

    for (int64_t c = 0; c < K; ++c) {
        for (int v = 0; v < 3; ++v)
            if (tris[3 * c + v] < 0 || tris[3 * c + v] >= nPts) throw std::runtime_error("triangle vertex id out of range");
        const double* p0 = &xy[2 * (size_t)tris[3 * c]];
        const double* p1 = &xy[2 * (size_t)tris[3 * c + 1]];
        const double* p2 = &xy[2 * (size_t)tris[3 * c + 2]];
        const double cross = (p1[0] - p0[0]) * (p2[1] - p0[1]) - (p1[1] - p0[1]) * (p2[0] - p0[0]);
        if (cross < 0) std::swap(tris[3 * c + 1], tris[3 * c + 2]);      // dgPolyMesh.C:490-509
        if (cross == 0) throw std::runtime_error("degenerate triangle " + std::to_string(c));
    }

    // match edges through a sort of (min,max) canonical point pairs
    std::vector<EdgeRec> edges((size_t)3 * K);
    for (int64_t c = 0; c < K; ++c)
        for (int f = 0; f < 3; ++f) {
            const int32_t a = canon(tris[3 * c + f]), b = canon(tris[3 * c + (f + 1) % 3]);
            edges[(size_t)3 * c + f] = {std::min(a, b), std::max(a, b), (int32_t)c, (int32_t)f};
        }
    std::sort(edges.begin(), edges.end(), [](const EdgeRec& x, const EdgeRec& y) {
        if (x.a != y.a) return x.a < y.a;
        if (x.b != y.b) return x.b < y.b;
        if (x.cell != y.cell) return x.cell < y.cell;
        return x.face < y.face;
    });
    std::vector<int32_t> nbr((size_t)3 * K, -1), nbrFace((size_t)3 * K, -1);
    for (size_t i = 0; i < edges.size();) {
        size_t j = i + 1;
        while (j < edges.size() && edges[j].a == edges[i].a && edges[j].b == edges[i].b) ++j;
        if (j - i == 2) {
            const EdgeRec &e0 = edges[i], &e1 = edges[i + 1];
            if (e0.cell == e1.cell) throw std::runtime_error("cell is its own neighbour (mesh too coarse for periodic gluing)");
            nbr[(size_t)3 * e0.cell + e0.face] = e1.cell;
            nbrFace[(size_t)3 * e0.cell + e0.face] = e1.face;
            nbr[(size_t)3 * e1.cell + e1.face] = e0.cell;
            nbrFace[(size_t)3 * e1.cell + e1.face] = e0.face;
        } else if (j - i > 2) {
            throw std::runtime_error("non-manifold edge");
        }
        i = j;
    }

    // dgFaces: created by the lower-numbered cell (= poly owner), cell-major / local-face-minor
    faceOwner.clear(); faceNbr.clear(); faceLocO.clear(); faceLocN.clear(); faceRot.clear();
    faceOwner.reserve((size_t)(3 * K / 2 + 16));
    cellFace.assign((size_t)3 * K, -1);
    for (int64_t c = 0; c < K; ++c)
        for (int f = 0; f < 3; ++f) {
            const int32_t nb = nbr[(size_t)3 * c + f];
            if (nb >= 0 && nb < c) continue;
            const int32_t fid = (int32_t)faceOwner.size();
            faceOwner.push_back((int32_t)c);
            faceNbr.push_back(nb);
            faceLocO.push_back(f);
            cellFace[(size_t)3 * c + f] = fid;
            if (nb >= 0) {
                const int32_t nf = nbrFace[(size_t)3 * c + f];
                faceLocN.push_back(nf);
                cellFace[(size_t)3 * nb + nf] = fid;
                const int32_t first = canon(tris[3 * c + f]);                     // dgPolyMesh.C:868-896
                faceRot.push_back(canon(tris[(size_t)3 * nb + nf]) == first ? 0 : 1);
            } else {
                faceLocN.push_back(-1);
                faceRot.push_back(-1);
            }
        }
    F = (int64_t)faceOwner.size();

    // patches (dgPatch.C:70-100)
    patches.clear();
    faceGhost.assign((size_t)F, -1);
    facePatch.assign((size_t)F, -1);
    nGhost = 0;
    for (int p = 0; p < nPatches; ++p) {
        Patch P;
        P.name = names ? (*names)[p] : ("patch" + std::to_string(p));
        P.type = types ? (*types)[p] : std::string("patch");
        P.ghostStart = nGhost;
        for (int32_t e = patchStart[p]; e < patchStart[p + 1]; ++e) {
            const int32_t c = edgeCell[e], pa = edgePoints[2 * e], pb = edgePoints[2 * e + 1];
            if (c < 0 || c >= K) throw std::runtime_error("patch edge cell out of range");
            int32_t fid = -1;
            for (int f = 0; f < 3; ++f) {
                const int32_t a = tris[3 * (size_t)c + f], b = tris[3 * (size_t)c + (f + 1) % 3];
                if ((a == pa && b == pb) || (a == pb && b == pa)) { fid = cellFace[(size_t)3 * c + f]; break; }
            }
            if (fid < 0) throw std::runtime_error("patch edge not found in its owner cell");
            if (faceNbr[fid] >= 0) throw std::runtime_error("patch edge is an interior face");
            if (faceGhost[fid] >= 0) throw std::runtime_error("boundary edge listed twice");
            P.faces.push_back(fid);
            faceGhost[fid] = (int32_t)nGhost++;
            facePatch[fid] = p;
        }
        patches.push_back(std::move(P));
    }
    for (int64_t f = 0; f < F; ++f)
        if (faceNbr[f] < 0 && faceGhost[f] < 0)
            throw std::runtime_error("boundary face " + std::to_string(f) + " (cell " + std::to_string(faceOwner[f]) +
                                     ") belongs to no patch");
}

void Mesh::connCodes(const int* patchKind, int32_t* out) const
{
    // code byte: bits 0-1 neighbour's local face, 0x4 reversed trace, 0x8 ghost region, 0x10 reflective, 0x20 dgFace owner
    // (kCode* in dg_kernels.cuh; hopedg.cu asserts that the two agree).  Patch kinds: 0 fixedValue, 1 zeroGradient, 2 reflective,
    // 3 processor (HDG_BC_* in include/hopedg.h).  Kind | 0x100 on a zeroGradient / reflective patch = "frozen": the exterior trace
    // is read from the patch's ghost slots, where hdg_state_freeze_traces stored the interior trace of an earlier field (the
    // reflective mirror is still applied by the kernel)
    for (int64_t k = 0; k < K; ++k) {
        unsigned codes = 0;
        for (int f = 0; f < 3; ++f) {
            const int32_t fid = cellFace[(size_t)3 * k + f];
            const bool owner = faceOwner[fid] == k && faceLocO[fid] == f;
            unsigned code = owner ? 0x20u : 0u;
            int32_t nb;
            if (faceNbr[fid] >= 0) {
                if (owner) { nb = faceNbr[fid]; code |= (unsigned)faceLocN[fid]; }
                else       { nb = faceOwner[fid]; code |= (unsigned)faceLocO[fid]; }
                if (faceRot[fid] == 1) code |= 0x4u;
            } else {
                const int kind = patchKind[facePatch[fid]] & 0xff;
                const bool frozen = patchKind[facePatch[fid]] & 0x100;
                if (kind == 0 || kind == 3) {
                    nb = faceGhost[fid];
                    code |= 0x8u;
                } else if ((kind == 1 || kind == 2) && frozen) {
                    nb = faceGhost[fid];
                    code |= 0x8u;
                    if (kind == 2) code |= 0x10u;
                } else if (kind == 1 || kind == 2) {
                    nb = (int32_t)k;
                    code |= (unsigned)f;
                    if (kind == 2) code |= 0x10u;
                } else
                    throw std::runtime_error("patch " + patches[facePatch[fid]].name + ": unsupported boundary kind on a patch that owns faces");
            }
            out[4 * k + f] = nb;
            codes |= code << (8 * f);
        }
        out[4 * k + 3] = (int32_t)codes;
    }
}

void Mesh::boundarySlots(int32_t* bslot, int32_t* ghostFirst) const
{
    for (int64_t k = 0; k < K; ++k)
        for (int f = 0; f < 3; ++f) {
            const int32_t fid = cellFace[(size_t)3 * k + f];
            bslot[3 * k + f] = faceNbr[fid] < 0 ? faceGhost[fid] : -1;
        }
    for (const Patch& P : patches)
        for (size_t i = 0; i < P.faces.size(); ++i) ghostFirst[P.ghostStart + (int64_t)i] = (int32_t)P.ghostStart;
}

void Mesh::elementGeometry(int64_t k, double g[16]) const
{
    const double* v0 = &xy[2 * (size_t)tris[3 * k]];
    const double* v1 = &xy[2 * (size_t)tris[3 * k + 1]];
    const double* v2 = &xy[2 * (size_t)tris[3 * k + 2]];
    // x = -(r+s)/2 v0 + (r+1)/2 v1 + (s+1)/2 v2   (triangleBaseFunction.C:303-313)
    const double xr = 0.5 * (v1[0] - v0[0]), yr = 0.5 * (v1[1] - v0[1]);
    const double xs = 0.5 * (v2[0] - v0[0]), ys = 0.5 * (v2[1] - v0[1]);
    const double J = xr * ys - yr * xs;
    g[0] = ys / J;    // rx
    g[1] = -xs / J;   // ry
    g[2] = -yr / J;   // sx
    g[3] = xr / J;    // sy
    const double nx[3] = {yr, ys - yr, -ys};        // triangleBaseFunction.C:354-391
    const double ny[3] = {-xr, xr - xs, xs};
    for (int f = 0; f < 3; ++f) {
        const double sJ = std::sqrt(nx[f] * nx[f] + ny[f] * ny[f]);
        g[4 + 3 * f] = nx[f] / sJ;
        g[5 + 3 * f] = ny[f] / sJ;
        g[6 + 3 * f] = sJ / J;
    }
    g[13] = J;
    g[14] = g[15] = 0.0;
}

// ------------------------------------------------------------------------------------------------
// ASCII polyMesh reader (only what dgPolyMesh consumes)
// ------------------------------------------------------------------------------------------------
namespace {

std::string slurpFoam(const std::string& path)
{
    std::ifstream in(path, std::ios::binary);
    if (!in) throw std::runtime_error("cannot open " + path);
    std::stringstream ss;
    ss << in.rdbuf();
    std::string t = ss.str(), out;
    out.reserve(t.size());
    for (size_t i = 0; i < t.size();) {          // strip comments
        if (t.compare(i, 2, "/*") == 0) {
            const size_t e = t.find("*/", i + 2);
            i = (e == std::string::npos) ? t.size() : e + 2;
        } else if (t.compare(i, 2, "//") == 0) {
            const size_t e = t.find('\n', i);
            i = (e == std::string::npos) ? t.size() : e;
        } else
            out.push_back(t[i++]);
    }
    const size_t h = out.find("FoamFile");       // drop the header dictionary
    if (h != std::string::npos) {
        const size_t e = out.find('}', h);
        if (e != std::string::npos) out.erase(h, e - h + 1);
    }
    return out;
}

// position just after the '(' that opens the top-level list; n = declared size
size_t listStart(const std::string& t, int64_t& n)
{
    size_t i = 0;
    while (i < t.size() && !std::isdigit((unsigned char)t[i])) ++i;
    char* end = nullptr;
    n = std::strtoll(t.c_str() + i, &end, 10);
    size_t p = (size_t)(end - t.c_str());
    p = t.find('(', p);
    if (p == std::string::npos) throw std::runtime_error("malformed list");
    return p + 1;
}

}  // namespace

void Mesh::readPolyMesh(const std::string& dir)
{
    // points
    std::vector<double> P;
    {
        const std::string t = slurpFoam(dir + "/points");
        int64_t n;
        size_t i = listStart(t, n);
        P.reserve((size_t)3 * n);
        const char* c = t.c_str() + i;
        for (int64_t k = 0; k < n; ++k) {
            while (*c && *c != '(') ++c;
            ++c;
            for (int d = 0; d < 3; ++d) {
                char* e;
                P.push_back(std::strtod(c, &e));
                c = e;
            }
            while (*c && *c != ')') ++c;
            ++c;
        }
    }
    // faces
    std::vector<int32_t> fStart(1, 0), fPts;
    {
        const std::string t = slurpFoam(dir + "/faces");
        int64_t n;
        size_t i = listStart(t, n);
        const char* c = t.c_str() + i;
        for (int64_t k = 0; k < n; ++k) {
            char* e;
            const long m = std::strtol(c, &e, 10);
            c = e;
            while (*c && *c != '(') ++c;
            ++c;
            for (long d = 0; d < m; ++d) {
                fPts.push_back((int32_t)std::strtol(c, &e, 10));
                c = e;
            }
            while (*c && *c != ')') ++c;
            ++c;
            fStart.push_back((int32_t)fPts.size());
        }
    }
    auto readLabels = [&](const std::string& name) {
        const std::string t = slurpFoam(dir + "/" + name);
        int64_t n;
        size_t i = listStart(t, n);
        std::vector<int32_t> v;
        v.reserve((size_t)n);
        const char* c = t.c_str() + i;
        for (int64_t k = 0; k < n; ++k) {
            char* e;
            v.push_back((int32_t)std::strtol(c, &e, 10));
            c = e;
        }
        return v;
    };
    const std::vector<int32_t> owner = readLabels("owner"), neighbour = readLabels("neighbour");
    const int64_t nFaces = (int64_t)fStart.size() - 1;
    if ((int64_t)owner.size() != nFaces) throw std::runtime_error("owner size != faces size");

    // boundary: name { type X; nFaces N; startFace S; ... }
    struct PP { std::string name, type; int32_t nFaces = -1, startFace = -1, nbrProc = -1; };
    std::vector<PP> pps;
    {
        std::string t = slurpFoam(dir + "/boundary");
        for (size_t a; (a = t.find("#{")) != std::string::npos;) {      // drop codeStream bodies of `arc` patches
            const size_t b = t.find("#}", a);
            t.erase(a, (b == std::string::npos ? t.size() : b + 2) - a);
        }
        int64_t n;
        size_t i = listStart(t, n);
        for (int64_t k = 0; k < n; ++k) {
            while (i < t.size() && std::isspace((unsigned char)t[i])) ++i;
            size_t j = i;
            while (j < t.size() && !std::isspace((unsigned char)t[j]) && t[j] != '{') ++j;
            PP pp;
            pp.name = t.substr(i, j - i);
            const size_t ob = t.find('{', j), cb = t.find('}', ob);
            if (ob == std::string::npos || cb == std::string::npos) throw std::runtime_error("malformed boundary file");
            std::stringstream body(t.substr(ob + 1, cb - ob - 1));
            std::string stmt;
            while (std::getline(body, stmt, ';')) {
                std::stringstream s2(stmt);
                std::string key, val;
                s2 >> key >> val;
                if (key == "type") pp.type = val;
                else if (key == "nFaces") pp.nFaces = std::atoi(val.c_str());
                else if (key == "startFace") pp.startFace = std::atoi(val.c_str());
                else if (key == "neighbProcNo") pp.nbrProc = std::atoi(val.c_str());      // processor patches of a processorN/ mesh
            }
            if (pp.type.empty() || pp.nFaces < 0 || pp.startFace < 0) throw std::runtime_error("patch " + pp.name + ": missing type/nFaces/startFace");
            pps.push_back(pp);
            i = cb + 1;
        }
    }

    int32_t nCells = 0;
    for (int32_t o : owner) nCells = std::max(nCells, o + 1);
    // base (z == 0) face of every prism (dgPolyMesh.C:154-190)
    std::vector<int32_t> T((size_t)3 * nCells, -1);
    auto tryBase = [&](int32_t cell, int64_t f) {
        const int32_t n = fStart[f + 1] - fStart[f];
        for (int32_t k = 0; k < n; ++k)
            if (P[3 * (size_t)fPts[fStart[f] + k] + 2] != 0.0) return;
        if (n != 3) throw std::runtime_error("only prism cells (triangles) are supported on this path");
        for (int k = 0; k < 3; ++k) T[(size_t)3 * cell + k] = fPts[fStart[f] + k];
    };
    for (int64_t f = 0; f < nFaces; ++f) {
        tryBase(owner[f], f);
        if (f < (int64_t)neighbour.size()) tryBase(neighbour[f], f);
    }
    for (int32_t c = 0; c < nCells; ++c)
        if (T[(size_t)3 * c] < 0) throw std::runtime_error("cell " + std::to_string(c) + " has no face in the plane z == 0 (dgPolyMesh requires the base plane at exactly z = 0)");

    std::vector<double> pxy((size_t)2 * (P.size() / 3));
    for (size_t p = 0; p < P.size() / 3; ++p) { pxy[2 * p] = P[3 * p]; pxy[2 * p + 1] = P[3 * p + 1]; }

    std::vector<int32_t> patchStart(1, 0), edgeCell, edgePts;
    std::vector<std::string> names, types;
    for (const PP& pp : pps) {
        names.push_back(pp.name);
        types.push_back(pp.type);
        if (pp.type != "empty") {
            for (int32_t f = pp.startFace; f < pp.startFace + pp.nFaces; ++f) {
                int32_t e[2], ne = 0;
                for (int32_t k = fStart[f]; k < fStart[f + 1]; ++k)
                    if (P[3 * (size_t)fPts[k] + 2] == 0.0 && ne < 2) e[ne++] = fPts[k];
                if (ne != 2) throw std::runtime_error("patch face without a z == 0 edge");
                edgeCell.push_back(owner[f]);
                edgePts.push_back(e[0]);
                edgePts.push_back(e[1]);
            }
        }
        patchStart.push_back((int32_t)edgeCell.size());
    }
    build((int64_t)pxy.size() / 2, pxy.data(), nCells, T.data(), nullptr, (int)pps.size(), patchStart.data(), edgeCell.data(),
          edgePts.data(), &names, &types);
    for (size_t p = 0; p < pps.size(); ++p) patches[p].nbrProc = pps[p].type == "processor" ? pps[p].nbrProc : -1;
    // polyMesh id of the lateral face behind every dgFace (orders the inter-processor faces in decompose())
    std::vector<EdgeRec> lat;
    for (int64_t f = 0; f < nFaces; ++f) {
        int32_t e[2], ne = 0;
        bool up = false;
        for (int32_t k = fStart[f]; k < fStart[f + 1]; ++k) {
            if (P[3 * (size_t)fPts[k] + 2] == 0.0) { if (ne < 2) e[ne] = fPts[k]; ++ne; }
            else up = true;
        }
        if (ne == 2 && up) lat.push_back({std::min(e[0], e[1]), std::max(e[0], e[1]), (int32_t)f, 0});
    }
    std::sort(lat.begin(), lat.end(), [](const EdgeRec& x, const EdgeRec& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
    polyFace.assign((size_t)F, -1);
    for (int64_t f = 0; f < F; ++f) {
        const int64_t c = faceOwner[f];
        const int32_t a = tris[3 * c + faceLocO[f]], b = tris[3 * c + (faceLocO[f] + 1) % 3];
        const EdgeRec key{std::min(a, b), std::max(a, b), 0, 0};
        auto it = std::lower_bound(lat.begin(), lat.end(), key, [](const EdgeRec& x, const EdgeRec& y) { return x.a != y.a ? x.a < y.a : x.b < y.b; });
        if (it == lat.end() || it->a != key.a || it->b != key.b) throw std::runtime_error("dgFace without a lateral polyMesh face");
        polyFace[(size_t)f] = it->cell;
    }
}

// ------------------------------------------------------------------------------------------------
// decomposition
// ------------------------------------------------------------------------------------------------
std::vector<int32_t> Mesh::decomposeSimple(int nx, int ny, int nz, double delta) const
{
    if (nx < 1 || ny < 1 || nz < 1) throw std::runtime_error("simpleCoeffs n must be positive");
    const double d = 1 - 0.5 * delta * delta, d2 = d * d, a = delta, a2 = a * a;      // geomDecomp.C:53-64
    const double R[3][3] = {{d2, -a * d, a}, {a * d - a2 * d, a * a2 + d2, -2 * a * d}, {a * d2 + a2, a * d - a2 * d, d2 - a2}};
    std::vector<double> rc((size_t)3 * K);
    for (int64_t c = 0; c < K; ++c) {
        double x = 0, y = 0;
        for (int v = 0; v < 3; ++v) { x += xy[2 * (size_t)tris[3 * c + v]]; y += xy[2 * (size_t)tris[3 * c + v] + 1]; }
        x /= 3.0; y /= 3.0;
        const double z = 0.0;     // constant over a one-layer mesh: shifts every rotated coordinate equally
        for (int i = 0; i < 3; ++i) rc[(size_t)3 * c + i] = R[i][0] * x + R[i][1] * y + R[i][2] * z;
    }
    std::vector<int32_t> finalDecomp((size_t)K, 0), idx((size_t)K), group((size_t)K);
    const int n[3] = {nx, ny, nz};
    int mult = 1;
    for (int dir = 0; dir < 3; ++dir) {
        if (n[dir] == 1) continue;      // one group: every cell gets 0 in this direction whatever the order (skips a sort of K keys)
        for (int64_t c = 0; c < K; ++c) idx[c] = (int32_t)c;
        std::stable_sort(idx.begin(), idx.end(), [&](int32_t p, int32_t q) { return rc[(size_t)3 * p + dir] < rc[(size_t)3 * q + dir]; });
        // assignToProcessorGroup (simpleGeomDecomp.C:55-84): the first (size - jump*n) groups get one extra cell
        const int64_t jump = K / n[dir], fst = K - jump * n[dir];
        int64_t ind = 0;
        int j = 0;
        for (; j < fst; ++j) for (int64_t k = 0; k < jump + 1; ++k) group[ind++] = j;
        for (; j < n[dir]; ++j) for (int64_t k = 0; k < jump; ++k) group[ind++] = j;
        for (int64_t i = 0; i < K; ++i) finalDecomp[idx[i]] += mult * group[i];
        mult *= n[dir];
    }
    return finalDecomp;
}

// ------------------------------------------------------------------------------------------------
// graph partitioner (stand-in for scotch / metis)
// ------------------------------------------------------------------------------------------------
namespace {

struct DualGraph {
    std::vector<int32_t> adj;      // 3 per cell, -1 = no neighbour
    const int32_t* of(int32_t c) const { return &adj[(size_t)3 * c]; }
};

// breadth-first order of the cells of `sub` (flag[c] == id) from `seed`; returns the order (unreached components are appended from
// their lowest cell, so disconnected sub-graphs are handled)
void bfsOrder(const DualGraph& g, const std::vector<int32_t>& sub, const std::vector<int32_t>& flag, int32_t id, int32_t seed,
              std::vector<int32_t>& order, std::vector<int32_t>& mark, int32_t stamp)
{
    order.clear();
    size_t next = 0, scan = 0;
    auto push = [&](int32_t c) { mark[(size_t)c] = stamp; order.push_back(c); };
    push(seed);
    while (order.size() < sub.size()) {
        if (next == order.size()) {      // another component
            while (mark[(size_t)sub[scan]] == stamp) ++scan;
            push(sub[scan]);
        }
        const int32_t c = order[next++];
        for (int k = 0; k < 3; ++k) {
            const int32_t n = g.of(c)[k];
            if (n >= 0 && flag[(size_t)n] == id && mark[(size_t)n] != stamp) push(n);
        }
    }
}

}  // namespace

std::vector<int32_t> Mesh::decomposeGraph(int nProcs) const
{
    if (nProcs < 1) throw std::runtime_error("numberOfSubdomains must be positive");
    DualGraph g;
    g.adj.assign((size_t)3 * K, -1);
    for (int64_t f = 0; f < F; ++f)
        if (faceNbr[f] >= 0) {
            g.adj[(size_t)3 * faceOwner[f] + faceLocO[f]] = faceNbr[f];
            g.adj[(size_t)3 * faceNbr[f] + faceLocN[f]] = faceOwner[f];
        }
    std::vector<int32_t> part((size_t)K, 0), flag((size_t)K, 0), mark((size_t)K, -1), order, order2;
    int32_t stamp = 0, nextId = 1;
    struct Job { std::vector<int32_t> cells; int32_t id, firstPart, nParts; };
    std::vector<Job> stack;
    {
        Job j;
        j.cells.resize((size_t)K);
        for (int64_t c = 0; c < K; ++c) j.cells[(size_t)c] = (int32_t)c;
        j.id = 0; j.firstPart = 0; j.nParts = nProcs;
        stack.push_back(std::move(j));
    }
    std::vector<int8_t> side((size_t)K, 0), locked((size_t)K, 0);
    while (!stack.empty()) {
        Job job = std::move(stack.back());
        stack.pop_back();
        if (job.nParts == 1) {
            for (int32_t c : job.cells) part[(size_t)c] = job.firstPart;
            continue;
        }
        const int32_t k1 = job.nParts / 2, k2 = job.nParts - k1;
        const int64_t n = (int64_t)job.cells.size();
        const int64_t n1 = (n * k1 + job.nParts / 2) / job.nParts;      // target size of side 0
        // three starting cuts, each refined on the graph, the smallest cut wins: (0) breadth-first growth from a pseudo-peripheral cell
        // (the last cell of a sweep from the lowest cell, twice), (1) / (2) the cells sorted by the x / y coordinate of their centroid
        std::vector<int8_t> bestSide;
        int64_t bestCut = -1;
        for (int cand = 0; cand < 3; ++cand) {
            if (cand == 0) {
                bfsOrder(g, job.cells, flag, job.id, job.cells[0], order, mark, ++stamp);
                bfsOrder(g, job.cells, flag, job.id, order.back(), order2, mark, ++stamp);
                bfsOrder(g, job.cells, flag, job.id, order2.back(), order, mark, ++stamp);
            } else {
                order = job.cells;
                const int d = cand - 1;
                auto cen = [&](int32_t c) { return xy[2 * (size_t)tris[3 * (size_t)c] + d] + xy[2 * (size_t)tris[3 * (size_t)c + 1] + d] + xy[2 * (size_t)tris[3 * (size_t)c + 2] + d]; };
                std::stable_sort(order.begin(), order.end(), [&](int32_t p, int32_t q) { return cen(p) < cen(q); });
            }
            for (int64_t i = 0; i < n; ++i) { side[(size_t)order[(size_t)i]] = i < n1 ? 0 : 1; }
            // Fiduccia-Mattheyses refinement: move the best-gain unlocked boundary cell of the side that is not below its target, keep the
            // best prefix.  gain = (neighbours on the other side) - (neighbours on the own side); degrees <= 3 -> 7 buckets
            auto gainOf = [&](int32_t c) {
                int gsum = 0;
                for (int k = 0; k < 3; ++k) {
                    const int32_t m = g.of(c)[k];
                    if (m >= 0 && flag[(size_t)m] == job.id) gsum += side[(size_t)m] != side[(size_t)c] ? 1 : -1;
                }
                return gsum;
            };
            const int64_t tol = std::max<int64_t>(1, n / 200);      // 0.5 % imbalance allowed while moving, exact balance restored at the end
            for (int pass = 0; pass < 6; ++pass) {
                std::vector<std::vector<int32_t>> bucket[2];
                bucket[0].assign(7, {});
                bucket[1].assign(7, {});
                for (int32_t c : job.cells) {
                    locked[(size_t)c] = 0;
                    bool boundary = false;
                    for (int k = 0; k < 3; ++k) {
                        const int32_t m = g.of(c)[k];
                        if (m >= 0 && flag[(size_t)m] == job.id && side[(size_t)m] != side[(size_t)c]) boundary = true;
                    }
                    if (boundary) bucket[side[(size_t)c]][(size_t)(gainOf(c) + 3)].push_back(c);
                }
                int64_t size0 = 0;
                for (int32_t c : job.cells) size0 += side[(size_t)c] == 0;
                std::vector<int32_t> moved;
                int64_t cur = 0, best = 0;
                size_t bestLen = 0;
                const int64_t limit = std::max<int64_t>(64, n / 8);
                for (int64_t it = 0; it < limit; ++it) {
                    // candidate side: the larger one relative to its target (ties: side 0)
                    const int from0 = (size0 - n1) >= 0 ? 0 : 1;
                    int32_t pick = -1;
                    int pickSide = -1;
                    for (int attempt = 0; attempt < 2 && pick < 0; ++attempt) {
                        const int sd = attempt == 0 ? from0 : 1 - from0;
                        const int64_t newDiff = std::llabs((size0 + (sd == 0 ? -1 : 1)) - n1);
                        if (newDiff > tol) continue;
                        for (int b = 6; b >= 0 && pick < 0; --b) {
                            auto& v = bucket[sd][(size_t)b];
                            while (!v.empty()) {
                                const int32_t c = v.back();
                                v.pop_back();
                                if (locked[(size_t)c] || side[(size_t)c] != sd || gainOf(c) + 3 != b) continue;      // stale entry
                                pick = c;
                                pickSide = sd;
                                break;
                            }
                        }
                    }
                    if (pick < 0) break;
                    cur += gainOf(pick);
                    side[(size_t)pick] = (int8_t)(1 - pickSide);
                    locked[(size_t)pick] = 1;
                    size0 += pickSide == 0 ? -1 : 1;
                    moved.push_back(pick);
                    for (int k = 0; k < 3; ++k) {
                        const int32_t m = g.of(pick)[k];
                        if (m >= 0 && flag[(size_t)m] == job.id && !locked[(size_t)m]) bucket[side[(size_t)m]][(size_t)(gainOf(m) + 3)].push_back(m);
                    }
                    if (cur > best || (cur == best && std::llabs(size0 - n1) == 0 && bestLen == 0 && cur > 0)) { best = cur; bestLen = moved.size(); }
                    if (cur < best - 50) break;
                }
                for (size_t i = moved.size(); i > bestLen; --i) side[(size_t)moved[i - 1]] = (int8_t)(1 - side[(size_t)moved[i - 1]]);      // roll back
                if (best <= 0) break;
            }
            // exact balance: move zero-cost-first boundary cells from the larger side until side 0 holds n1 cells
            {
                int64_t size0 = 0;
                for (int32_t c : job.cells) size0 += side[(size_t)c] == 0;
                while (size0 != n1) {
                    const int from = size0 > n1 ? 0 : 1;
                    int32_t pick = -1;
                    int bestGain = -4;
                    for (int32_t c : job.cells) {
                        if (side[(size_t)c] != from) continue;
                        const int gn = gainOf(c);
                        bool boundary = false;
                        for (int k = 0; k < 3; ++k) {
                            const int32_t m = g.of(c)[k];
                            if (m >= 0 && flag[(size_t)m] == job.id && side[(size_t)m] != from) boundary = true;
                        }
                        if (boundary && gn > bestGain) { bestGain = gn; pick = c; }
                    }
                    if (pick < 0) {      // no boundary cell (disconnected): take any
                        for (int32_t c : job.cells) if (side[(size_t)c] == from) { pick = c; break; }
                    }
                    side[(size_t)pick] = (int8_t)(1 - from);
                    size0 += from == 0 ? -1 : 1;
                }
            }

            int64_t cut = 0;
            for (int32_t c : job.cells)
                for (int k = 0; k < 3; ++k) {
                    const int32_t m = g.of(c)[k];
                    if (m > c && flag[(size_t)m] == job.id && side[(size_t)m] != side[(size_t)c]) ++cut;
                }
            if (bestCut < 0 || cut < bestCut) {
                bestCut = cut;
                bestSide.resize((size_t)n);
                for (int64_t i = 0; i < n; ++i) bestSide[(size_t)i] = side[(size_t)job.cells[(size_t)i]];
            }
        }
        for (int64_t i = 0; i < n; ++i) side[(size_t)job.cells[(size_t)i]] = bestSide[(size_t)i];
        Job a, b;
        a.id = nextId++; a.firstPart = job.firstPart; a.nParts = k1;
        b.id = nextId++; b.firstPart = job.firstPart + k1; b.nParts = k2;
        for (int32_t c : job.cells) {
            if (side[(size_t)c] == 0) { a.cells.push_back(c); flag[(size_t)c] = a.id; }
            else { b.cells.push_back(c); flag[(size_t)c] = b.id; }
        }
        if (a.cells.empty() || b.cells.empty()) throw std::runtime_error("graph partitioner: more sub-domains than cells");
        stack.push_back(std::move(b));
        stack.push_back(std::move(a));
    }
    return part;
}

Mesh::LocalMesh Mesh::decompose(const std::vector<int32_t>& cellToProc, int nProcs, int rank) const
{
    if ((int64_t)cellToProc.size() != K) throw std::runtime_error("cellToProc size != number of cells");
    if (rank < 0 || rank >= nProcs) throw std::runtime_error("rank out of range");
    for (int32_t p : cellToProc) if (p < 0 || p >= nProcs) throw std::runtime_error("cellToProc entry out of range");
    LocalMesh L;
    std::vector<int32_t> g2l((size_t)K, -1);
    for (int64_t c = 0; c < K; ++c)
        if (cellToProc[c] == rank) { g2l[c] = (int32_t)L.cellAddr.size(); L.cellAddr.push_back((int32_t)c); }
    // points used by the local cells, ascending global id
    std::vector<char> used((size_t)nPoints, 0);
    for (int32_t c : L.cellAddr) for (int v = 0; v < 3; ++v) used[tris[3 * (size_t)c + v]] = 1;
    std::vector<int32_t> pg2l((size_t)nPoints, -1);
    for (int64_t p = 0; p < nPoints; ++p)
        if (used[p]) { pg2l[p] = (int32_t)L.pointAddr.size(); L.pointAddr.push_back((int32_t)p); }
    for (int32_t p : L.pointAddr) { L.xy.push_back(xy[2 * (size_t)p]); L.xy.push_back(xy[2 * (size_t)p + 1]); }
    for (int32_t c : L.cellAddr) for (int v = 0; v < 3; ++v) L.tris.push_back(pg2l[tris[3 * (size_t)c + v]]);
    if (periodicGlue) {
        // periodic gluing (an extension: the reference has no compiled cyclic patch).  Faces glued across the wrap are ordinary interior
        // dgFaces of the global mesh; inside one processor they stay glued through the restriction of the point map (the canonical
        // representative of a class = its lowest local point), across processors they become processor-patch faces like any other cut.
        std::vector<int32_t> rep((size_t)nPoints, -1);      // global canonical id -> lowest local point of the class
        L.pointEquiv.resize(L.pointAddr.size());
        for (size_t lp = 0; lp < L.pointAddr.size(); ++lp) {
            const int32_t cg = pointEquiv[(size_t)L.pointAddr[lp]];
            if (rep[(size_t)cg] < 0) rep[(size_t)cg] = (int32_t)lp;
            L.pointEquiv[lp] = rep[(size_t)cg];
        }
    }
    auto addEdge = [&](int32_t fid, int32_t globalCell, int localFace) {
        L.edgeCell.push_back(g2l[globalCell]);
        L.edgePts.push_back(pg2l[tris[3 * (size_t)globalCell + localFace]]);
        L.edgePts.push_back(pg2l[tris[3 * (size_t)globalCell + (localFace + 1) % 3]]);
        L.patchFaceGlobal.push_back(fid);
    };
    // original patches: faces whose cell lives here, in patch order (domainDecompositionMesh.C:160-185)
    L.patchStart.push_back(0);
    for (const Patch& P : patches) {
        for (int32_t fid : P.faces)
            if (cellToProc[faceOwner[fid]] == rank) addEdge(fid, faceOwner[fid], faceLocO[fid]);
        L.patchStart.push_back((int32_t)L.edgeCell.size());
        L.names.push_back(P.name);
        L.types.push_back(P.type);
        L.patchNbrProc.push_back(-1);
    }
    // inter-processor faces: ascending global (poly) face id, grouped by neighbour processor in ascending order (:215-240, :355-400)
    struct Cut { int64_t key; int32_t fid, nbrProc; };
    std::vector<Cut> cuts;
    for (int64_t f = 0; f < F; ++f) {
        if (faceNbr[f] < 0) continue;
        const int32_t po = cellToProc[faceOwner[f]], pn = cellToProc[faceNbr[f]];
        if (po == pn || (po != rank && pn != rank)) continue;
        // polyMesh face id when known, else the upper-triangular rank (owner, neighbour) a valid polyMesh would have
        const int64_t key = polyFace.empty() ? (int64_t)faceOwner[f] * K + faceNbr[f] : polyFace[(size_t)f];
        cuts.push_back({key, (int32_t)f, po == rank ? pn : po});
    }
    std::sort(cuts.begin(), cuts.end(), [](const Cut& x, const Cut& y) { return x.nbrProc != y.nbrProc ? x.nbrProc < y.nbrProc : x.key < y.key; });
    for (size_t i = 0; i < cuts.size();) {
        size_t j = i;
        const int32_t q = cuts[i].nbrProc;
        for (; j < cuts.size() && cuts[j].nbrProc == q; ++j) {
            const int32_t f = cuts[j].fid;
            if (cellToProc[faceOwner[f]] == rank) addEdge(f, faceOwner[f], faceLocO[f]);
            else                                  addEdge(f, faceNbr[f], faceLocN[f]);
        }
        L.patchStart.push_back((int32_t)L.edgeCell.size());
        L.names.push_back("procBoundary" + std::to_string(rank) + "to" + std::to_string(q));
        L.types.push_back("processor");
        L.patchNbrProc.push_back(q);
        i = j;
    }
    return L;
}

}  // namespace hdg

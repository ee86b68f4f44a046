// 2-D DG mesh: triangles, dgFace connectivity, patches, affine geometric factors.
//
// Reference behaviour restated (paths relative to HopeFOAM-0.1/src/DG/):
//   dgMesh/dgPolyMesh.C:154-190   z==0 face of each prism gives the 3 vertex labels
//   dgMesh/dgPolyMesh.C:490-509   CCW swap of v1,v2
//   dgMesh/dgPolyMesh.C:346-396   dgFace creation by the poly owner, cell-major / local-face-minor
//   dgMesh/dgPolyMesh.C:868-896   firstPointIndex -> faceRotate
//   dgMesh/dgPolyMesh.C:999-1038  faceIndexInOwner / faceIndexInNeighbour / dgCellFaceNewID
//   dgMesh/dgPatches/dgPatch/dgPatch.C:70-100  patch face -> dgFace id, polyPatch order
//   element/baseFunctions/.../triangleBaseFunction.C:303-391  node map, dxdr, drdx, normals, fscale
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace hdg {

struct Patch {
    std::string name, type;          // as in constant/polyMesh/boundary
    std::vector<int32_t> faces;      // dgFace ids in polyPatch face order (dgFaceIndex_)
    int64_t ghostStart = 0;          // first ghost-face slot of this patch
    int32_t nbrProc = -1;            // neighbProcNo of a `processor` patch read from a processorN/constant/polyMesh/boundary
};

struct Mesh {
    int64_t K = 0, nPoints = 0, F = 0, nGhost = 0;
    std::vector<double> xy;          // nPoints*2
    std::vector<int32_t> tris;       // K*3, CCW, v0 = first vertex of the base face as stored
    std::vector<int32_t> faceOwner, faceNbr, faceLocO, faceLocN, faceRot;   // F each
    std::vector<int32_t> cellFace;   // K*3 dgCellFaceNewID_
    std::vector<int32_t> faceGhost;  // F: ghost slot of a patch face, -1 for interior faces
    std::vector<int32_t> facePatch;  // F: patch id, -1 interior
    std::vector<Patch> patches;
    std::vector<int32_t> polyFace;   // F: polyMesh face id of the lateral face behind a dgFace (readPolyMesh), empty otherwise
    bool periodicGlue = false;       // built with a pointEquiv map
    std::vector<int32_t> pointEquiv; // the map itself (canonical point id per point; empty = identity): decompose() restricts it

    // builds the connectivity; pointEquiv (optional) identifies points for periodic gluing
    void build(int64_t nPoints, const double* xy, int64_t K, const int32_t* tris, const int32_t* pointEquiv,
               int nPatches, const int32_t* patchStart, const int32_t* edgeCell, const int32_t* edgePoints,
               const std::vector<std::string>* names, const std::vector<std::string>* types);
    // reads an ASCII constant/polyMesh directory (one layer of prisms, base plane at z == 0)
    void readPolyMesh(const std::string& dir);

    // ---- domain decomposition (dgDecomposePar rules, applications/utilities/DG/dgDecomposePar/domainDecompositionMesh.C) ----
    struct LocalMesh {
        std::vector<int32_t> cellAddr;        // cellProcAddressing: local cell -> global cell (ascending, :124)
        std::vector<int32_t> pointAddr;       // pointProcAddressing: local point -> global point (ascending, :463-511)
        std::vector<double> xy;
        std::vector<int32_t> tris;            // local vertex ids, same vertex order as the global cell
        std::vector<int32_t> pointEquiv;      // local canonical point ids when the global mesh is glued periodically (else empty)
        std::vector<int32_t> patchStart, edgeCell, edgePts;   // original patches (all of them, possibly empty) + processor patches
        std::vector<int32_t> patchNbrProc;    // -1 for an original patch, else the neighbour processor (ascending, :355-372)
        std::vector<int32_t> patchFaceGlobal; // per local patch edge: global dgFace id (faceProcAddressing restricted to patch faces)
        std::vector<std::string> names, types;
    };
    // `simple` geometric decomposition of the cell centres (src/parallel/decompose/decompositionMethods/simpleGeomDecomp/simpleGeomDecomp.C:129-197)
    std::vector<int32_t> decomposeSimple(int nx, int ny, int nz, double delta) const;
    // graph partition of the cell-adjacency (dual) graph: recursive bisection, each cut grown breadth-first from a pseudo-peripheral
    // cell and refined by Fiduccia-Mattheyses passes.  Stands in for `method scotch | metis` (the reference links the real libraries,
    // decompositionMethods/scotchDecomp; neither can be built here): same input (the cell-cell graph of
    // decompositionMethod::calcCellCells), same contract (balanced parts, small cut), NOT the same cellToProc as scotch's.
    std::vector<int32_t> decomposeGraph(int nProcs) const;
    LocalMesh decompose(const std::vector<int32_t>& cellToProc, int nProcs, int rank) const;

    // per-state connectivity (one int4 per element: neighbour element / ghost slot per face + 3 packed code bytes, dg_kernels.cuh
    // kCode*) for the patch kinds of one field (HDG_BC_*); out = K*4 ints.  Used by the stage kernels and by the slope limiter.
    void connCodes(const int* patchKind, int32_t* out) const;
    // bslot[3k+f] = ghost slot of a boundary face of element k (-1 interior); ghostFirst[slot] = first slot of the slot's patch
    void boundarySlots(int32_t* bslot, int32_t* ghostFirst) const;

    // affine geometric factors of element k: g[0..3] = rx, ry, sx, sy; g[4+3f..] = nx, ny, Fscale of face f; g[13] = J
    void elementGeometry(int64_t k, double g[16]) const;
};

}  // namespace hdg

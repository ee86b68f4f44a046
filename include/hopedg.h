/* hopedg.h - C ABI of the B200-native explicit nodal-DG stage library (libhopedg.so)
 *
 * Drop-in boundary for ONE hot path of HopeFOAM-0.1: the explicit DG right-hand side + RK stage that
 * dgEulerFoam executes (tutorials/DG/2D/isentropicVortex/dgEulerFoam/dgEulerFoam.C:64-131) and its
 * scalar-advection sibling (dgc::div(U,T) with LF flux).  The reference has no C ABI (it is a C++
 * template DSL on OpenFOAM run-time selection); each entry point below names the reference interface
 * it replaces.  Paths are relative to HopeFOAM-0.1/ ; DG/ = src/DG/.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; hdg_last_error(ctx) gives the message
 *     (the C++ facade turns it into the reference's FatalErrorInFunction ... abort(FatalError) style).
 *   - the library never aborts and never falls back to the CPU: without a CUDA device hdg_create fails.
 *   - host arrays are owned by the caller; device buffers by the context.
 *   - scalar = double, label = int32 (HopeFOAM etc/bashrc:80-84).
 *   - nodal host layout is the reference's: element-contiguous AoS, Field<scalar>[K*Np],
 *     Field<vector>[K*Np][3] (z ignored/zero) - DG/element/physicalElementData/physicalElementData.C:388-401.
 *   - one context per GPU / rank; calls on a context are serialised by the caller.
 */
#ifndef HOPEDG_H
#define HOPEDG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hdg_context hdg_context;

/* ---- patch (boundary-condition) kinds: DG/fields/dgPatchFields/{basic,constraint}/ ------------------ */
enum {
    HDG_BC_FIXED_VALUE   = 0,  /* basic/fixedValue/fixedValueDgPatchField.C:108-119                     */
    HDG_BC_ZERO_GRADIENT = 1,  /* basic/zeroGradient/zeroGradientDgPatchField.C:101-110                 */
    HDG_BC_REFLECTIVE    = 2,  /* basic/reflective/reflectiveDgPatchField.C:111-154                     */
    HDG_BC_PROCESSOR     = 3,  /* constraint/processor/processorDgPatchField.C:235-331 (halo)           */
    HDG_BC_EMPTY         = 4   /* front/back planes of the one-layer 3-D polyMesh: carry no dgFaces     */
};

/* ---- numerical flux kinds -------------------------------------------------------------------------- */
enum {
    HDG_FLUX_ROE     = 0,      /* DG/DG/godunovFlux/fluxSchemes/scheme/RoeFlux/RoeFlux.C:46-191         */
    HDG_FLUX_LF      = 1,      /* DG/DG/simpleFlux/schemes/LFFlux/LFFlux.C:105-211 (advection); on the Euler entry points: point-wise
                                * local Lax-Friedrichs (Rusanov) - an extension, the reference's godunovScheme knows Roe only */
    HDG_FLUX_AVERAGE = 2,      /* DG/DG/simpleFlux/schemes/averageFlux/averageFlux.C:95-190             */
    HDG_FLUX_NONE    = 3       /* DG/DG/simpleFlux/schemes/noneFlux/noneFlux.C:45-97                    */
};

/* ---- lifetime ----------------------------------------------------------------------------------------
 * replaces: Foam::dgMesh construction (DG/dgMesh/dgMesh.C:71-116) + PetscInitialize.                   */
int  hdg_create(int device, hdg_context** out);
void hdg_destroy(hdg_context* ctx);
const char* hdg_last_error(const hdg_context* ctx);      /* valid until the next call on ctx             */
const char* hdg_version(void);
int  hdg_sync(hdg_context* ctx);                          /* joins the context's streams                  */

/* ---- order + reference element -------------------------------------------------------------------------
 * replaces: system/dgSolution DG{baseOrder} -> stdElementSets::getElement(N,"tri")
 * (DG/element/stdElementSets/stdElementSets.C:43-56).  N = 1..8 use the reference's cubature table (its limit,
 * gaussTriangleIntegration.C:50-64: the reference aborts beyond); N = 9, 10 (BASELINE configs[3] sweep) use an own collapsed
 * Gauss-Legendre x Gauss-Jacobi(1,0) rule of degree 3(N+1) - no reference counterpart, parity unpinned.  Before hdg_set_mesh*.          */
int hdg_set_order(hdg_context* ctx, int N);
int hdg_get_sizes(const hdg_context* ctx, int32_t* Np, int32_t* Nfp, int32_t* Ng, int32_t* Nfg);
/* reference-element operators, row-major doubles (for parity tests against the oracle):
 *   what = "r","s" (Np) | "V","invV","Dr","Ds" (Np x Np) | "gr","gs","gw" (Ng) | "Vg","Dgr","Dgs" (Ng x Np)
 *          | "fx","fw" (Nfg) | "If" (Nfg x Nfp) | "Mref" (Np x Np) | "Pr","Ps" (Np x Ng) | "LIFT" (Np x 3*Nfg)
 *          | "faceShift" (3 x Np x Nfp: cell-node displacement per face-node displacement, curved patches)
 *   returns the number of doubles written (<= cap) or -1.                                               */
int64_t hdg_get_operator(const hdg_context* ctx, const char* what, double* out, int64_t cap);
/* faceToCellIndex_[face][rotate][i] as int32[3*2*Nfp] (triangleBaseFunction.C:75-94)                    */
int hdg_get_face_to_cell_index(const hdg_context* ctx, int32_t* out);

/* ---- mesh ------------------------------------------------------------------------------------------------
 * replaces: dgPolyMesh ctor (DG/dgMesh/dgPolyMesh.C:36-234) + physicalElementData::initElements
 * (DG/element/physicalElementData/physicalElementData.C:60-161) + dgPatch ctor (dgPatch.C:50-113).
 *
 * hdg_set_mesh_triangles: K CCW triangles over nPoints 2-D points.  Connectivity follows the dgPolyMesh
 * rules: local face f joins vertices f,(f+1)%3; dgFaces are created cell-major/local-face-minor by the
 * lower-numbered cell; faceRotate = position of the owner's first face vertex in the neighbour's face.
 *   pointEquiv   optional int32[nPoints]: canonical point id used for edge matching (periodic gluing -
 *                extension, the reference has no compiled cyclic patch); NULL = identity.
 *   patches      nPatches patches; patch p owns boundary edges patchStart[p]..patchStart[p+1]-1, each
 *                given as (cell, pointA, pointB) in polyPatch face order.                               */
int hdg_set_mesh_triangles(hdg_context* ctx, int64_t nPoints, const double* xy, int64_t K, const int32_t* tris,
                           const int32_t* pointEquiv, int32_t nPatches, const int32_t* patchStart,
                           const int32_t* edgeCell, const int32_t* edgePoints);
/* reads <dir>/{points,faces,owner,neighbour,boundary} (ASCII polyMesh, one layer of prisms with the base
 * plane at z==0 exactly, dgPolyMesh.C:154-190) and calls hdg_set_mesh_triangles.                        */
int hdg_set_mesh_polymesh(hdg_context* ctx, const char* polyMeshDir);

/* ---- domain decomposition -----------------------------------------------------------------------------------
 * replaces dgDecomposePar for the explicit path: applications/utilities/DG/dgDecomposePar/domainDecompositionMesh.C:102-511
 * (cells per processor in ascending global id :124; original patches kept in order, faces where the cell lives :160-185;
 * inter-processor faces in ascending global face id grouped by neighbour processor in ascending order :215-240,355-400;
 * points in ascending global id :463-511) and the `simple` method (src/parallel/decompose/decompositionMethods/
 * simpleGeomDecomp/simpleGeomDecomp.C:129-197: bands of sorted, delta-rotated cell centres; decomposeParDict
 * `method simple; simpleCoeffs{ n (nx ny nz); delta 0.001; }`).  scotch is vendored in the reference but needs flex/bison;
 * a cellToProc list produced elsewhere (the reference's `manual` method / cellDecomposition file) is accepted as is.
 *   hdg_decompose_simple : cellToProc[K] of the mesh held by ctx (a host-only context is enough)
 *   hdg_mesh_decompose   : builds rank's processor mesh inside `local` (order already set); its patches are the original
 *                          patches (same indices, possibly empty) followed by one processor patch per neighbour
 *   hdg_mesh_proc_addressing : cellProcAddressing[K], pointProcAddressing[nPoints], neighbour processor per patch (-1 for
 *                          original patches), global dgFace id per patch face (patch-major)
 *   hdg_decompose_from_dict : reads <case>/system/decomposeParDict (numberOfSubdomains; method simple|manual; simpleCoeffs{n;delta};
 *                          manualCoeffs{dataFile} -> <case>/constant/<dataFile> labelList) as dgDecomposePar does; `method scotch |
 *                          metis` map to hdg_decompose_graph
 *   hdg_decompose_graph  : graph partition of the cell-cell graph (the input decompositionMethod::calcCellCells gives scotch / metis in the
 *                          reference, scotchDecomp.C): native recursive bisection (breadth-first growth from a pseudo-peripheral cell +
 *                          Fiduccia-Mattheyses refinement), parts balanced to one cell.  The cellToProc is NOT the one scotch would give
 *                          (a scotch cellDecomposition still drops in through `method manual`).                                     */
int hdg_decompose_simple(const hdg_context* ctx, int32_t nx, int32_t ny, int32_t nz, double delta, int32_t* cellToProc);
int hdg_decompose_graph(const hdg_context* ctx, int32_t nProcs, int32_t* cellToProc);
int hdg_decompose_from_dict(const hdg_context* ctx, const char* caseDir, int32_t* nProcs, int32_t* cellToProc);
int hdg_mesh_decompose(const hdg_context* global, int32_t nProcs, const int32_t* cellToProc, int32_t rank, hdg_context* local);
int hdg_mesh_proc_addressing(const hdg_context* ctx, int32_t* cellProcAddressing, int32_t* pointProcAddressing,
                             int32_t* patchNbrProc, int32_t* patchFaceGlobal);
int64_t hdg_mesh_num_points(const hdg_context* ctx);
int hdg_mesh_get_points(const hdg_context* ctx, double* xy /* nPoints*2: the vertices of the base plane */);

int hdg_mesh_counts(const hdg_context* ctx, int64_t* K, int64_t* F, int32_t* nPatches, int64_t* nGhostFaces);
/* connectivity as the reference holds it (int32[F] each; neighbour / faceLocN / faceRot = -1 on patches) */
int hdg_mesh_get_faces(const hdg_context* ctx, int32_t* faceOwner, int32_t* faceNbr, int32_t* faceLocO,
                       int32_t* faceLocN, int32_t* faceRot);
int hdg_mesh_get_cell_vertices(const hdg_context* ctx, int32_t* tris /* K*3, after dgPolyMesh CCW rules */);
/* patch p: name/type strings (type as in polyMesh/boundary), its dgFace ids in polyPatch order          */
int hdg_mesh_patch_info(const hdg_context* ctx, int32_t p, char* name, int32_t nameCap, char* type, int32_t typeCap,
                        int32_t* nFaces);
int hdg_mesh_patch_faces(const hdg_context* ctx, int32_t p, int32_t* dgFaceIndex);
/* device-side topology as the kernels read it (queries; they also work on a host-only context).  conn: per element 4 ints =
 * neighbour element / ghost slot per local face + three packed code bytes (dg_kernels.cuh kCode*) for the given per-patch
 * HDG_BC_* kinds; bslot: ghost slot of each boundary face (K*3, -1 interior), ghostFirst: first slot of the slot's patch;
 * node table: faceToCellIndex padded to [3][2][NfpPad] (returns its length)                                              */
int hdg_mesh_conn_codes(const hdg_context* ctx, const int32_t* patchKind, int32_t nPatches, int32_t* conn /* K*4 */);
int hdg_mesh_boundary_slots(const hdg_context* ctx, int32_t* bslot /* K*3 */, int32_t* ghostFirst /* nGhost */);
int hdg_get_node_table(const hdg_context* ctx, int32_t* out, int32_t cap);
/* physical node coordinates dofLocation_ (K*Np*2 doubles, AoS x,y) - triangleBaseFunction.C:303-313     */
int hdg_mesh_node_coords(const hdg_context* ctx, double* xy);
/* patch node coordinates in patch-dof order (sum over patch faces of Nfp; owner face-node order)        */
int hdg_mesh_patch_node_coords(const hdg_context* ctx, int32_t p, double* xy);
/* Curved boundary patches (polyMesh/boundary `type arc`, dgMesh/dgPatches/constraint/arc/arcDgPatch.C): `positions` = the Nfp nodes of
 * every face of the patch on the curve, patch-dof order, (x, y) per node (what arcDgPatch::positions returns: interior face nodes moved
 * onto the parametric curve, end points kept).  The displacement is blended into the owner cells' node locations as
 * physicalElementData::updatePatchDofIndexMapping does (physicalElementData.C:185-224, triangleBaseFunction::addFaceShiftToCell,
 * triangleBaseFunction.C:421-466); hdg_mesh_node_coords / hdg_mesh_patch_node_coords return the displaced locations afterwards.  As in the
 * reference, the operators keep the straight-sided metrics: dgMesh.C:110-113 builds them (initElements) before the displacement and
 * never rebuilds them.                                                                                                                */
int hdg_mesh_set_curved_patch(hdg_context* ctx, int32_t patch, const double* positions);

/* ---- state (a group of nodal scalar planes advanced together) -----------------------------------------
 * replaces: GeometricDofField<Type,dgPatchField,dgGeoMesh> storage + its boundary field
 * (DG/fields/GeometricDofField/GeometricDofField.H:80-405).  A dgScalarField is 1 plane, a dgVectorField 2
 * planes (x,y); the Euler state is 4 planes (rho, rhoU.x, rhoU.y, Ener).  Each state owns two device
 * copies (current / stage) so SSP-RK2 and LSERK run without host traffic.                               */
int hdg_state_create(hdg_context* ctx, int32_t nPlanes, int32_t* stateId);
int hdg_state_destroy(hdg_context* ctx, int32_t stateId);
/* AoS <-> device SoA of the CURRENT copy.  `host` holds K*Np nodes with hostStride doubles per node (1 for a
 * Field<scalar>, 3 for a Field<vector>); component c of every node goes to / comes from plane plane0+c,
 * c < nPlanes <= hostStride.  Download zero-fills the components >= nPlanes (the z of a 2-D vector).
 * One H2D / D2H copy per call; the AoS<->SoA transposition runs on the device.                           */
int hdg_state_upload(hdg_context* ctx, int32_t stateId, int32_t plane0, int32_t nPlanes, const double* host, int32_t hostStride);
int hdg_state_download(hdg_context* ctx, int32_t stateId, int32_t plane0, int32_t nPlanes, double* host, int32_t hostStride);
/* The same transfers without waiting, on two dedicated copy streams (PCIe is full duplex: the upload of one state overlaps the
 * download of another and the stage kernels of a third).  `host` must be pinned (cudaHostAlloc / cudaHostRegister) and stay
 * untouched until hdg_sync.  Ordering is kept by the library: compute calls on a state wait for its pending upload, a download
 * waits for the compute work enqueued before it, an upload waits until a pending download of the same state has read the planes
 * and - if its host buffer overlaps the download's - has written the host buffer.  This is how a driver that streams many independent cases (or time slabs) through one GPU keeps the link busy in
 * both directions; the reference has no counterpart (its fields live in host memory, GeometricDofField.H:116-119).           */
int hdg_state_upload_async(hdg_context* ctx, int32_t stateId, int32_t plane0, int32_t nPlanes, const double* host, int32_t hostStride);
int hdg_state_download_async(hdg_context* ctx, int32_t stateId, int32_t plane0, int32_t nPlanes, double* host, int32_t hostStride);
/* boundary field of plane `plane` on patch p: kind + (for fixedValue) the patch dof values, nFaces*Nfp
 * doubles in patch-dof order (physicalElementData.C:163-183).  replaces dgPatchField<Type> construction
 * and the solver hook setBoundaryValues (TUT/isentropicVortex/dgEulerFoam/setBoundaryValues.H:27-49).   */
int hdg_state_set_patch_kind(hdg_context* ctx, int32_t stateId, int32_t patch, int32_t bcKind);
int hdg_state_set_patch_values(hdg_context* ctx, int32_t stateId, int32_t plane0, int32_t nPlanes, int32_t patch,
                               const double* values, int32_t hostStride);

/* ---- the hot path ------------------------------------------------------------------------------------------
 * hdg_euler_stage: one fused explicit stage of the compressible Euler system on state `s`
 *     q_out = a * q_n + b * ( q_in + dt * L(q_in) )
 * where L is updateGaussField + Roe flux + the three dg::solveEquation calls of dgEulerFoam.C:77-90
 * (GeometricDofField.C:648-655, RoeFlux.C:46-191, defaultConvectionScheme.C:48-129, defaultGrad.C:87-166,
 *  EulerDdtScheme.C:118-144, dgMatrixSolve.C:90-213).
 *   stageIndex 0: q_in = q_n (current copy), result -> stage copy            (use a=0, b=1)
 *   stageIndex 1: q_in = stage copy, result -> current copy (in place on q_n) (SSP-RK2: a=b=0.5)
 * hdg_euler_step_ssprk2 = both stages (the whole time step of dgEulerFoam.C:67-123).                    */
int hdg_euler_stage(hdg_context* ctx, int32_t stateId, double gamma, double dt, int32_t fluxKind,
                    int32_t stageIndex, double a, double b);
int hdg_euler_step_ssprk2(hdg_context* ctx, int32_t stateId, double gamma, double dt, int32_t fluxKind);
/* the same stage restricted to elements [elemBegin, elemEnd) plus an optional second range [elemBegin2, elemEnd2) (pass 0,0 for
 * none; ranges octet-aligned: multiples of 8, or end == K): lets a multi-GPU driver advance the partition-boundary elements
 * and the interior in separate launches so that the halo exchange of the next stage overlaps the interior launch. */
int hdg_euler_stage_range(hdg_context* ctx, int32_t stateId, double gamma, double dt, int32_t fluxKind, int32_t stageIndex,
                          double a, double b, int64_t elemBegin, int64_t elemEnd, int64_t elemBegin2, int64_t elemEnd2);
/* low-storage RK(5,4) (coefficients rk4a/rk4b of TUT/isentropicVortex/dgEulerFoam/createFields.H:119-138,
 * declared but unused by the reference solver):  res = A_s*res + dt*L(q) ; q = q + B_s*res               */
int hdg_euler_step_lserk45(hdg_context* ctx, int32_t stateId, double gamma, double dt, int32_t fluxKind);

/* hdg_advect_stage: dg::solveEquation(dgm::ddt(T) + dgc::div(U,T)) with `div(U,T) default LF|average`
 * (dgcDiv.C:88-107, EquationConvectionScheme.H:109-125, defaultConvectionScheme.C:216-303, LFFlux.C:105-211);
 * T = 1-plane state, U = 2-plane state (nodal velocity, boundary field included).  Same a/b/stage meaning. */
int hdg_advect_stage(hdg_context* ctx, int32_t stateT, int32_t stateU, double dt, int32_t fluxKind,
                     int32_t stageIndex, double a, double b);
int hdg_advect_step_ssprk2(hdg_context* ctx, int32_t stateT, int32_t stateU, double dt, int32_t fluxKind);
/* low-storage RK(5,4) on the same operator (rk4a/rk4b of createFields.H:119-138): res = A_s*res + dt*L(T) ; T = T + B_s*res */
int hdg_advect_step_lserk45(hdg_context* ctx, int32_t stateT, int32_t stateU, double dt, int32_t fluxKind);

/* The same stage for a solver that keeps rho, rhoU, Ener as three separate fields, as the reference does
 * (1-, 2- and 1-plane states): reads the CURRENT copies, writes a*aux + b*(q + dt*L(q)) into the STAGE copies; the
 * caller commits each field with hdg_state_swap - this is what the C++ facade's dg::solveEquation does.           */
int hdg_euler_stage_fields(hdg_context* ctx, int32_t stateRho, int32_t stateRhoU, int32_t stateEner, double gamma, double dt,
                           int32_t fluxKind, double a, double b, int32_t auxRho, int32_t auxRhoU, int32_t auxEner);
/* The general form the facade's lazy evaluation uses (dgEulerFoam.C:70-117 without the field copies and the separate SSP combination):
 *   s[3]    the fields advanced (rho, rhoU, Ener: 1-, 2-, 1-plane states): their patch kinds and boundary (ghost) data; the result
 *           a*aux + b*(q + dt*L(q)) goes to their STAGE copies
 *   src[3]  -1, or the states whose CURRENT copies hold the nodal data q: `rho1 = rho` followed by the first stage never materialises
 *           rho1's copy (rho's nodes, rho1's boundary data)
 *   out2[3] -1, or a second result a2*aux2 + b2*(q + dt*L(q)) into the STAGE copies of these states: `rho = 0.5*rho + 0.5*rho1` right after
 *           the second stage (dgEulerFoam.C:115-117) comes out of the same launch (out2 = aux2 = rho, a2 = b2 = 0.5), no axpby
 *   exchange  != 0 on a decomposed mesh (hdg_comm_init done): the processor-patch halo of the three fields' CURRENT values is exchanged
 *           by this call (processorDgPatchField::initEvaluate / evaluate, processorDgPatchField.C:235-331) and hidden behind compute:
 *           pack, ncclSend/Recv and unpack run on the halo stream while the octets without a processor face are advanced; the octets
 *           next to processor patches follow when the ghosts have landed.  Not combined with src[].
 * The caller commits every output field with hdg_state_swap.  hdg_state_copy_ghosts: dst takes src's boundary (ghost) region only.     */
typedef struct hdg_euler_fields_stage {
    int32_t s[3], src[3], aux[3], out2[3], aux2[3];
    double gamma, dt, a, b, a2, b2;
    int32_t fluxKind;
    int32_t exchange;
} hdg_euler_fields_stage;
int hdg_euler_stage_fields_ex(hdg_context* ctx, const hdg_euler_fields_stage* stage);
int hdg_state_copy_ghosts(hdg_context* ctx, int32_t dstState, int32_t srcState);
int hdg_state_swap(hdg_context* ctx, int32_t stateId);                 /* current <-> stage copy                 */
/* Godunov.limite(rho, rhoU, Ener) with `limiteScheme Triangle` (DG/godunovFlux/limiteSchemes/scheme/Trianglelimite/
 * Trianglelimite.C:61-864): area-weighted gradient limiter on (rho, u, v, p), P1 reconstruction about the cell averages, in
 * place on the current copies of the three fields.  The reference hard-wires gamma = 1.4 (:74), eps = 1e-10 (:716) and
 * tol = 1e-2 (:803); a cell whose mean density is below tol makes the reference loop forever (:823-827) - here the density
 * slope of such a cell becomes zero.  Single rank only (the reference's coupled branch is empty, :170-172).            */
int hdg_euler_limit(hdg_context* ctx, int32_t stateRho, int32_t stateRhoU, int32_t stateEner, double gamma, double eps, double tol);
/* Boundary data that lag the field, as in the reference: zeroGradient / reflective patch fields are evaluated from the interior only by
 * correctBoundaryConditions (after each solve, dgMatrixSolve.C:209), and Godunov.limite changes the interior WITHOUT re-evaluating them
 * (doubleMach/dgEulerFoam/dgEulerFoam.C:92-107: the second stage sees the wall data of the unlimited field).  freeze stores the current
 * interior trace of those patches in their ghost slots and makes the next hdg_euler_stage_fields read them from there (the reflective
 * mirror is still applied); that stage, or hdg_state_thaw (= correctBoundaryConditions), returns the state to evaluating from the field.
 * STATUS: written after the round's GPU time was spent - host logic (connectivity codes) tested, device path not yet run.             */
int hdg_state_freeze_traces(hdg_context* ctx, int32_t stateId);
int hdg_state_thaw(hdg_context* ctx, int32_t stateId);
/* the limiter's cell-average weights: column sums of the reference mass matrix / 2 (:109-116), Np doubles                 */
int hdg_limiter_weights(const hdg_context* ctx, double* mpp);
/* field assignment rho1 = rho (internal + boundary field, dgEulerFoam.C:70-72)                                  */
int hdg_state_copy(hdg_context* ctx, int32_t dstState, int32_t srcState);
/* field algebra on the current copies: dst = a*x + b*y (rho = 0.5*rho + 0.5*rho1, dgEulerFoam.C:115-117)       */
int hdg_state_axpby(hdg_context* ctx, int32_t dstState, double a, int32_t xState, double b, int32_t yState);

/* sum_i |q_i - ref_i| over the nodal dofs of a plane against a host reference (eulerError.H:32-38 uses
 * gSum(mag(diff))/nDof); ref in AoS with stride.  Device reduction, deterministic order.                */
int hdg_state_l1_diff(hdg_context* ctx, int32_t stateId, int32_t plane, const double* ref, int32_t hostStride,
                      double* out);

/* ---- multi-GPU halo (processor patches) ------------------------------------------------------------------
 * replaces processorDgPatchField::initEvaluate/evaluate (processorDgPatchField.C:235-331): the owner-side
 * nodal trace of every plane on a processor patch, reversed per face (faceRotate), lands in the peer's
 * ghost slots.  The library packs/unpacks on the device; the transport (NCCL send/recv, or P2P copy) is
 * driven by the caller with the device pointers returned here, so that the library has no MPI/NCCL
 * link dependency.                                                                                      */
int hdg_halo_counts(const hdg_context* ctx, int32_t patch, int64_t* nDoublesPerPlane /* nFaces * NfpPad */);
/* optional: let the caller own the send/recv buffers (e.g. tensors registered with a communicator)       */
int hdg_halo_bind(hdg_context* ctx, int32_t patch, void* devSendBuf, void* devRecvBuf, int64_t capDoubles);
int hdg_halo_pack(hdg_context* ctx, int32_t stateId, int32_t which /*0 current,1 stage*/, int32_t patch,
                  void** devSendBuf, int64_t* nDoubles);
int hdg_halo_recv_buffer(hdg_context* ctx, int32_t stateId, int32_t patch, void** devRecvBuf, int64_t* nDoubles);
int hdg_halo_unpack(hdg_context* ctx, int32_t stateId, int32_t which, int32_t patch);
/* One process per GPU without a Python driver (the facade's `-parallel`): an NCCL communicator owned by the context.  libnccl is
 * opened with dlopen when hdg_comm_init is called; rank 0 publishes the NCCL id through `idFile` (a path all ranks see, unique per run).
 * hdg_halo_exchange = pack + grouped send/recv + unpack of ALL processor patches of a state copy (the neighbour rank of a patch is the
 * neighbProcNo of its polyMesh boundary entry), ordered against the compute stream by events.  Replaces the MPI calls of
 * processorDgPatchField.C:235-331 and Pstream's gSum / the PETSc ownership range of dgMesh.C:194-219. */
int hdg_comm_init(hdg_context* ctx, int32_t rank, int32_t worldSize, const char* idFile);
int hdg_comm_rank_size(const hdg_context* ctx, int32_t* rank, int32_t* size);
int hdg_halo_exchange(hdg_context* ctx, int32_t stateId, int32_t which /*0 current, 1 stage*/);
/* Overlapped exchange for ANY decomposition, owned by the library (the reference exchanges before every evaluation and hides nothing,
 * processorDgPatchField.C:235-331; its call sites are the three updateGaussField() of dgEulerFoam.C:77-79).  The octets (groups of 8
 * elements) that own a processor face are advanced first; their traces are packed, sent and unpacked on the halo stream while the launch
 * over all other octets runs.  hdg_euler_step_ssprk2_parallel = the whole SSP-RK2 step of dgEulerFoam.C:67-117 on one rank (NCCL transport,
 * collective); hdg_group_euler_step_ssprk2 = the same step for n contexts of ONE process (index = processor number; peer copies instead of
 * NCCL) - a single-process multi-GPU driver, and the form in which the exchange is tested on one GPU.  The ghosts of the result are
 * current on return, so consecutive steps need no extra exchange; any other write to the state triggers one blocking exchange first.
 * hdg_mesh_set_patch_neighbour declares a patch of a caller-built mesh as a processor patch towards `nbrRank` (meshes from
 * hdg_mesh_decompose / processorN directories carry that already); `tag` orders several patches between the same pair of ranks and must
 * agree on both sides.  hdg_par_counts: processor faces, octets with / without a processor face, processor patches of this rank.        */
int hdg_euler_step_ssprk2_parallel(hdg_context* ctx, int32_t stateId, double gamma, double dt, int32_t fluxKind);
int hdg_group_euler_step_ssprk2(hdg_context** ctxs, const int32_t* stateIds, int32_t n, double gamma, double dt, int32_t fluxKind);
int hdg_mesh_set_patch_neighbour(hdg_context* ctx, int32_t patch, int32_t nbrRank, int32_t tag);
int hdg_par_counts(hdg_context* ctx, int64_t* nProcFaces, int64_t* nBoundaryOctets, int64_t* nInteriorOctets, int32_t* nNeighbours);
int hdg_comm_allreduce_sum(hdg_context* ctx, double* hostValues, int32_t n);
int hdg_comm_allgather_i64(hdg_context* ctx, int64_t value, int64_t* out /* worldSize entries */);
/* streams: 0 = compute (stage kernels, uploads), 1 = halo (pack/unpack run here).  hdg_stream returns the cudaStream_t as void*;
 * hdg_stream_wait makes stream `waiter` wait for everything enqueued so far on stream `signaler` (event record + wait).   */
void* hdg_stream(hdg_context* ctx, int32_t which /*0 compute, 1 halo*/);
int hdg_stream_wait(hdg_context* ctx, int32_t waiter, int32_t signaler);

/* ---- measurement hooks -------------------------------------------------------------------------------------- */
/* number of kernel launches issued by this context so far (bench.py's gpu_launches)                     */
int64_t hdg_launch_count(const hdg_context* ctx);
/* names of the kernels one full-mesh Euler stage launches at the context's order, '+'-separated, into out[cap]: the split stage
 * "eulerFaceFluxKernel<N>+eulerElemKernel<N>" (one Roe flux per dgFace, as defaultConvectionScheme.C:114-127 hands one flux to both
 * cells; N = 1, 2: "eulerFacePairFluxKernel<N>") or the fused "eulerStageKernel<N>" (thin partition-boundary launches, HDG_EULER_SPLIT=0).
 * Returns the number of kernels. */
int hdg_euler_stage_kernels(hdg_context* ctx, char* out, int32_t cap);
/* FP64 pipe peak of this GPU, measured live (back-to-back DMMA.8x8x4 chains for about `seconds`; DMMA and DFMA share one pipe on B200):
 * the denominator of bench.py's `fp64` object, taken at the clocks of the same run                          */
int hdg_measure_fp64_peak(hdg_context* ctx, double seconds, double* tflops);
/* device pointer of plane 0 of a state copy (for external CUDA-event timing / debugging)                 */
void* hdg_state_device_ptr(hdg_context* ctx, int32_t stateId, int32_t which);
/* device layout of a state plane: [Kpad][NpPad] element nodes, then [nGhost][NfpPad] ghost traces;
 * persistent grid sizes of the two stage kernels (blocks of 128 threads)                                 */
int hdg_layout(const hdg_context* ctx, int64_t* Kpad, int32_t* NpPad, int32_t* NfpPad, int64_t* planeStride,
               int64_t* ghostBase, int32_t* eulerGrid, int32_t* advectGrid);

#ifdef __cplusplus
}
#endif
#endif /* HOPEDG_H */

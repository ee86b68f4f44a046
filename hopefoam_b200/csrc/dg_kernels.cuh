// Hand-written sm_100a FP64 kernels for the fused explicit DG stage (Euler / scalar advection).
//
// Design: one warp advances an "octet" of 8 elements.  Every dense contraction of the stage
// (interpolation to cubature points, weak-form projection, face-trace interpolation, lift) is issued as
// DMMA.8x8x4 (mma.sync.m8n8k4.f64) with the ELEMENTS on the M axis and the reference-element operator as the
// B operand, so the accumulator layout of one contraction (row = element, col = 2*(lane%4)+h) is exactly the
// A-operand layout of the next one after a fixed permutation of the operator's K index.  Fluxes are evaluated
// point-wise on the accumulator registers in between; nothing but the operator fragments touches shared
// memory and the state is streamed from HBM once per stage.
//
// Reference arithmetic restated here (paths relative to HopeFOAM-0.1/src/DG/):
//   fields/dgGaussField/dgGaussField.C:188-269                        interpolation + face-trace gather
//   DG/godunovFlux/fluxSchemes/scheme/RoeFlux/RoeFlux.C:46-191         Roe flux
//   DG/simpleFlux/schemes/LFFlux/LFFlux.C:105-211                      LF flux (nodal variant)
//   DG/convectionSchemes/defaultConvectionScheme/defaultConvectionScheme.C:48-303  volume + surface integrals
//   DG/gradSchemes/defaultGrad/defaultGrad.C:87-166                    pressure gradient volume term
//   DG/ddtSchemes/EulerDdtScheme/EulerDdtScheme.C:118-144 + dgMatrices/dgMatrix/dgMatrixSolve.C:90-213  explicit update / mass solve
#pragma once
#include <cstdint>

namespace hdg {

constexpr int kNgTable[11] = {0, 12, 19, 36, 54, 73, 93, 118, 145, 256, 289};   // cubature points of order 3(N+1): the reference's table (N <= 8),
                                                                                  // own collapsed 16x16 / 17x17 rules for N = 9, 10

template <int N>
struct Dims {
    static constexpr int Np = (N + 1) * (N + 2) / 2;
    static constexpr int Nfp = N + 1;
    static constexpr int Ng = kNgTable[N];
    static constexpr int Nfg = N + 2;
    static constexpr int NpPad = (Np + 7) / 8 * 8;     // nodal storage stride of one element (doubles)
    static constexpr int KT = (Np + 3) / 4;            // k-tiles over nodes
    static constexpr int NT = NpPad / 8;               // n-tiles over nodes
    static constexpr int GT = (Ng + 7) / 8;            // n-tiles over cell cubature points
    static constexpr int FGT = (Nfg + 7) / 8;          // n-tiles over the Gauss points of one face
    static constexpr int FKT = (Nfp + 3) / 4;          // k-tiles over the nodes of one face trace
    static constexpr int NfpPad = FKT * 4;             // ghost-trace stride (doubles)
    // operator fragment tables, in doubles ([tile...][lane])
    static constexpr int oVg = 0;
    static constexpr int oPr = oVg + GT * KT * 32;
    static constexpr int oPs = oPr + GT * 2 * NT * 32;
    static constexpr int oIf = oPs + GT * 2 * NT * 32;
    static constexpr int oLift = oIf + FGT * FKT * 32;
    static constexpr int tableDoubles = oLift + 3 * FGT * 2 * NT * 32;
    // advection (nodal collapse) tables
    // volume k-tiles follow the double2 load layout: k-tile (2*nt'+h), slot j <-> node 8*nt' + 2*j + h  => 2*NT k-tiles
    static constexpr int oDwr = 0;                     // [2*NT][NT][32]
    static constexpr int oDws = oDwr + 2 * NT * NT * 32;
    static constexpr int oLiftN = oDws + 2 * NT * NT * 32; // [3][FKT][NT][32]
    static constexpr int advTableDoubles = oLiftN + 3 * FKT * NT * 32;   // part staged into shared memory by advectStageKernel
    // TMA-pipelined advection kernel (dg_advect_tma.cu): the three face traces share ONE K axis (slot = face*Nfp + i), so the
    // nodal lift is a single Np x 3Nfp operator: KTC k-tiles instead of 3*FKT (N=4: 4 instead of 6)
    static constexpr int KTC = (3 * Nfp + 3) / 4;
    static constexpr int oLiftC = advTableDoubles;     // [KTC][NT][32]
    static constexpr int advTableDoublesAll = oLiftC + KTC * NT * 32;
    static constexpr int nodeTabInts = 3 * 2 * NfpPad; // faceToCellIndex padded
    // element kernel of the split stage (dg_euler_split.cu): [Vg][Pr][Ps] as above, then the weak nodal derivative in the A layout of the
    // nodal fragments (k-tile kt, slot j <-> node 4 kt + j) for the density equation - its flux rhoU is linear in the nodal data, so
    // Pr Vg (rx q1 + ry q2) needs no quadrature - and the lift over ONE K axis of all three faces (slot = face * Nfg + point)
    static constexpr int KTL = (3 * Nfg + 3) / 4;
    static constexpr int sVg = 0;
    static constexpr int sPr = sVg + GT * KT * 32;
    static constexpr int sPs = sPr + GT * 2 * NT * 32;
    static constexpr int sDwr = sPs + GT * 2 * NT * 32;       // [KT][NT][32]
    static constexpr int sDws = sDwr + KT * NT * 32;
    static constexpr int sLiftC = sDws + KT * NT * 32;        // [KTL][NT][32]
    static constexpr int splitTableDoubles = sLiftC + KTL * NT * 32;
    static constexpr int fluxSlots = (Nfg + 1) / 2 * 2; // split stage: Gauss-point slots per (face, field) record, Nfg rounded up to even (16-B pairs)
    // N >= 9 (beyond the reference's cubature table): the operator fragments (367 / 530 KB) no longer fit in shared memory and the nodal
    // A fragments no longer fit in registers: the Euler stage kernel reads the fragments through L1 from global memory and parks the A
    // fragments of a warp's octet in shared memory ([f][kt][lane], each lane reads back its own slots)
    static constexpr bool big = N >= 9;
    static constexpr int eulerThreads = N <= 6 ? 128 : 256;
    static constexpr int eulerSmemDoubles = big ? (eulerThreads / 32) * 4 * KT * 32 : tableDoubles;
};

inline int npPadOf(int N) { const int Np = (N + 1) * (N + 2) / 2; return (Np + 7) / 8 * 8; }
inline int nfpPadOf(int N) { return (N + 1 + 3) / 4 * 4; }

// per-face connectivity byte (4 packed into conn[e].w)
enum : unsigned {
    kCodeFaceMask = 0x3,   // neighbour's local face id
    kCodeRev = 0x4,        // read the neighbour trace reversed (faceRotate == 1)
    kCodeGhost = 0x8,      // exterior trace lives in the ghost region (fixedValue / processor patch)
    kCodeReflect = 0x10,   // reflective wall: mirror the momentum of the exterior state
    kCodeOwner = 0x20      // this element is the dgFace owner (flux evaluated in the owner's orientation)
};

struct StageParams {
    const double* qin[4];  // one pointer per conserved plane (rho, rhoU.x, rhoU.y, Ener): each [Kpad*NpPad | ghosts]
    const double* qghost[4]; // plane bases the GHOST traces (fixedValue / processor / frozen) are read from: = qin unless the element data
                             // come from another field's planes (the facade's lazy `rho1 = rho`: rho's nodes, rho1's boundary data)
    const double* qaux[4]; // SSP: q_n ; LSRK: unused
    double* qout[4];
    double* res[4];        // LSRK residual (in/out), else nullptr
    // optional second result of the same stage (mode 0): q_out2 = A2*q_aux2 + B2*(q_in + dt*L).  The facade's SSP-RK2 loop ends with
    // `rho = 0.5*rho + 0.5*rho1` right after the second stage (dgEulerFoam.C:115-117): one launch writes rho1 AND the combination
    // (the stage is FP64-bound at 16 % of the HBM roofline: the extra streams are free)
    double* qout2[4];
    const double* qaux2[4];
    double A2, B2;
    const double* geo;     // [Kpad][16]
    const int4* conn;      // [Kpad] : x,y,z = neighbour element / ghost slot per face, w = 3 packed code bytes
    const double* tables;  // operator fragments
    const int* nodeTab;    // [3][2][NfpPad]
    int64_t K;
    int64_t octBegin, octEnd;  // range of 8-element octets this launch advances (interior / partition-boundary split)
    int64_t octBegin2, octEnd2; // optional second range handled by the same launch (the other partition-boundary row)
    const int* octList;    // optional explicit list of octets (overrides the ranges): the processor-adjacent / interior octets of an
    int64_t nList;         // arbitrary decomposition, so that the halo exchange overlaps the interior launch (hdg_euler_step_ssprk2_parallel)
    int64_t ghostBase;     // offset of the ghost region inside a plane (= Kpad*NpPad)
    double gamma, dt, A, B;
    int mode;              // 0: q_out = A*q_aux + B*(q_in + dt*L)   1: res = A*res + dt*L ; q_out = q_in + B*res
    int fluxKind;          // 0 Roe (RoeFlux.C:46-191), 1 point-wise local Lax-Friedrichs (extension, dg_device.cuh)
    // split stage (dg_euler_split.cu): the Roe flux of every dgFace is evaluated ONCE, at the owner's Gauss points and in the owner's
    // orientation (as the reference does, defaultConvectionScheme.C:114-127), by eulerFaceFluxKernel into flux[F][4][fluxSlots<N>];
    // eulerElemKernel then lifts it on both sides (negated and read in reverse point order by the neighbour)
    double* flux;
    const int* faceOwner;  // [F] : owner element * 4 + owner's local face
    const int4* elemFace;  // [Kpad] : dgFace id of the element's local faces 0..2
    int64_t F;
    const double* splitTables;  // element kernel's fragments: [Vg][Pr][Ps][Dwr][Dws][combined lift] (Dims<N>::s*)
};

// Gauss-point slots per (face, field) record of the split stage's flux array: Nfg rounded up to even (16-B pairs)
inline int fluxSlotsOf(int N) { return (N + 2 + 1) / 2 * 2; }      // = Dims<N>::fluxSlots

struct HaloPlanes { double* p[4]; };      // plane pointers of a halo pack / unpack over all processor faces

// device geometry record [16 doubles]: rx ry sx sy | (nx,ny) x 3 faces | Fscale x 3 faces | J
constexpr int kGeoN = 4;    // nx of face f at kGeoN + 2f, ny at kGeoN + 2f + 1
constexpr int kGeoFs = 10;  // Fscale of face f at kGeoFs + f

struct AdvectParams {
    const double* Tin; const double* Taux; double* Tout; double* res;
    const double* U;       // 2 planes [planeStrideU]
    const double* UZ;      // the same values as (x,y) pairs [planeStrideU][2] (TMA kernel only; nullptr otherwise)
    const double* geo; const int4* connT; const int4* connU;
    const double* tables; const int* nodeTab;
    int64_t K, planeStrideT, planeStrideU, ghostBase;
    double dt, A, B;
    int mode, fluxKind;
    int sameConn;          // connU == connT (same boundary kinds on T and U)
    int anyReflect;        // some U patch is reflective (otherwise the mirror step is skipped)
};

}  // namespace hdg

// libhopedg.so - context, device memory, operator-fragment tables and the C ABI (include/hopedg.h).
// No CPU fallback: every compute entry point launches the sm_100a kernels in dg_kernels.cu.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>       // types only: the library is opened with dlopen when hdg_comm_init is called (never in a process that does not ask)
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <fstream>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hopedg.h"
#include "../include/hopedg/foamLite.H"      // dictionary grammar only (header-only, no OpenFOAM)
#include "dg_kernels.cuh"
#include "dg_limiter_core.hpp"
#include "mesh.hpp"
#include "ref_element.hpp"

namespace hdg {
void launchEulerStage(int N, const StageParams& p, int grid, cudaStream_t st);
bool eulerSplitAvailable(int N);
void launchEulerSplit(int N, const StageParams& p, bool faces, int smCount, cudaStream_t st);
void launchAdvectStage(int N, const AdvectParams& p, int grid, cudaStream_t st);
bool advectUsesTma(int N);
void launchZipPlanes(const double* x, const double* y, double* out, int64_t n, cudaStream_t st);
void stageOccupancy(int N, int* eulerBlocks, size_t* eulerSmem, int* advBlocks, size_t* advSmem);
int eulerWarpsPerBlock(int N);
void launchAosToPlane(const double* src, int hostStride, double* dst, int64_t K, int Np, int NpPad, cudaStream_t st);
void launchPlaneToAos(const double* src, double* dst, int hostStride, int64_t K, int Np, int NpPad, cudaStream_t st);
void launchPatchToGhost(const double* src, int hostStride, double* ghost, int64_t nFaces, int Nfp, int NfpPad, cudaStream_t st);
void launchPatchToGhostAll(const double* src, int hostStride, double* ghost0, double* ghost1, int64_t planeStride, int nPlanes, int64_t nFaces, int Nfp,
                           int NfpPad, cudaStream_t st);
void launchAxpby(double* dst, double a, const double* x, double b, const double* y, int64_t n, cudaStream_t st);
void launchL1Diff(const double* q, const double* ref, int64_t K, int Np, int NpPad, double* partial, int nBlocks, cudaStream_t st);
void launchHaloPack(const double* q, int64_t planeStride, int nPlanes, const int* faceElem, const int* faceLoc, const int* nodeTab,
                    int64_t nFaces, int Nfp, int NfpPad, int NpPad, double* buf, cudaStream_t st, int rev = 1);
void launchHaloUnpack(const double* buf, double* q, int64_t planeStride, int nPlanes, int64_t ghostOff, int64_t nFaces, int NfpPad,
                      cudaStream_t st);
void launchHaloPackAll(const HaloPlanes& q, int nPlanes, const int* faceElem, const int* faceLoc, const int* nodeTab, int64_t nFaces, int Nfp,
                       int NfpPad, int NpPad, double* buf, cudaStream_t st);
void launchHaloUnpackAll(const double* buf, const HaloPlanes& q, int nPlanes, const int* faceGhost, int64_t ghostBase, int64_t nFaces, int NfpPad,
                         cudaStream_t st);
int launchTriangleLimiter(const LimiterView& v, cudaStream_t st);
double launchFp64Peak(double* out, int blocks, int iters, cudaStream_t st);
}  // namespace hdg

using namespace hdg;

#define CUDA_OK(call)                                                                                             \
    do {                                                                                                          \
        cudaError_t e_ = (call);                                                                                  \
        if (e_ != cudaSuccess)                                                                                    \
            throw std::runtime_error(std::string(#call) + " failed: " + cudaGetErrorString(e_) + " (" + __FILE__ + ":" + \
                                     std::to_string(__LINE__) + ")");                                             \
    } while (0)

namespace {

struct State {
    int nPlanes = 0;
    double* d[2] = {nullptr, nullptr};   // 0 = current (q_n), 1 = stage copy
    double* res = nullptr;               // LSERK residual (lazily allocated)
    int4* conn = nullptr;
    std::vector<int> patchKind;
    bool connDirty = true;
    // frozen traces (hdg_state_freeze_traces): zeroGradient / reflective patches read their exterior trace from the ghost slots
    int4* connFrozen = nullptr;
    bool frozen = false, connFrozenDirty = true;
    // asynchronous transfer pipeline (hdg_state_upload_async / hdg_state_download_async)
    cudaEvent_t evUp = nullptr, evRead = nullptr, evDown = nullptr;
    bool upPending = false, readPending = false, downPending = false;
    const char *downLo = nullptr, *downHi = nullptr;      // host range the pending asynchronous downloads of this state write to
    bool usedSinceRead = true;
    // interleaved copy (x,y) of a 2-plane state for the TMA advection kernel: rebuilt when the state may have changed
    double* zip = nullptr;
    const double* zipOf = nullptr;
    uint64_t version = 1, zipVersion = 0;
    // processor-patch ghosts: the copy whose ghost slots hold the neighbours' traces of its CURRENT contents (set by the in-library
    // exchange; valid while haloVersion == version)
    const double* haloFreshBuf = nullptr;
    uint64_t haloVersion = 0;
    bool external = false;          // a raw device pointer was handed out: contents may change behind the library's back      // compute calls touched the planes after the last asynchronous download was enqueued
};

struct HaloPatch {
    int* faceElem = nullptr;
    int* faceLoc = nullptr;
    double* send = nullptr;
    double* recv = nullptr;
    int64_t capDoubles = 0;
    bool ownsBuffers = false;
};

// Everything the overlapped processor-patch exchange of a context needs (built on first use per mesh): the processor patches in
// message order (ascending neighbour, then tag), their faces concatenated, the octets that own a processor face and the rest.
struct ParPlan {
    bool built = false;
    std::vector<int> patches, nbr, tag;
    std::vector<int64_t> faceOff;             // first face of entry i in the concatenated list; back() = nPF
    int64_t nPF = 0, nB = 0, nI = 0;
    int *dFaceElem = nullptr, *dFaceLoc = nullptr, *dFaceGhost = nullptr, *dOctB = nullptr, *dOctI = nullptr;
    double *send = nullptr, *recv = nullptr;  // nPF * 4 * NfpPad doubles each, layout [face][plane][NfpPad]
    cudaEvent_t evBoundary = nullptr, evHalo = nullptr, evPacked = nullptr, evCopied = nullptr;
    bool haloPending = false, copiedPending = false;
    void release()
    {
        cudaFree(dFaceElem); cudaFree(dFaceLoc); cudaFree(dFaceGhost); cudaFree(dOctB); cudaFree(dOctI); cudaFree(send); cudaFree(recv);
        if (evBoundary) { cudaEventDestroy(evBoundary); cudaEventDestroy(evHalo); cudaEventDestroy(evPacked); cudaEventDestroy(evCopied); }
        *this = ParPlan();
    }
};

}  // namespace

struct hdg_context {
    int device = 0;
    bool hostOnly = false;   // device == -1: mesh / operator queries only (CPU tests of the host logic); compute calls fail
    cudaStream_t stream = nullptr, haloStream = nullptr;
    cudaStream_t inStream = nullptr, outStream = nullptr;      // created on first use by the asynchronous transfers
    cudaEvent_t evCompute = nullptr;
    double* dRing[2] = {nullptr, nullptr};                     // device staging rings of the in / out transfer streams
    size_t ringCap[2] = {0, 0}, ringOff[2] = {0, 0};
    std::string err;
    // one process per GPU: NCCL communicator of the processor-patch halo exchange (hdg_comm_init)
    ncclComm_t comm = nullptr;
    ncclResult_t (*commDestroy)(ncclComm_t) = nullptr;
    int commRank = 0, commSize = 1;
    cudaEvent_t evHalo = nullptr;
    double* dReduce = nullptr;
    int N = 0;
    bool hasRef = false, hasMesh = false;
    RefElement ref;
    Mesh mesh;
    Mesh::LocalMesh procAddr;      // filled by hdg_mesh_decompose (cell/point addressing, neighbour processor per patch)
    int64_t Kpad = 0, planeStride = 0, ghostBase = 0;
    int NpPad = 0, NfpPad = 0;
    double* dGeo = nullptr;
    // split Euler stage (dg_euler_split.cu): one flux record per dgFace + the face / element index arrays, built on first use
    double* dFlux = nullptr;
    int* dFaceOwner = nullptr;
    int4* dElemFace = nullptr;
    int splitOrders = -1;           // bit N set: order N runs the split stage (HDG_EULER_SPLIT overrides the default)
    double* dTables = nullptr;
    double* dAdvTables = nullptr;
    double* dSplitTables = nullptr;  // element kernel of the split Euler stage: [Vg][Pr][Ps][Dwr][Dws][combined lift]
    int* dNodeTab = nullptr;
    std::vector<int> nodeTabHost;   // [3][2][NfpPad] as uploaded (hdg_get_node_table)
    double* dStage = nullptr;
    size_t stageDoubles = 0;
    double* dPartial = nullptr;
    // boundary values on their way to the device: pinned host slots + device slots used round-robin, so that hdg_state_set_patch_values
    // returns without waiting for the stream (a copy from pageable memory makes the host wait for everything enqueued before it)
    struct UpSlot { double* h = nullptr; double* d = nullptr; size_t cap = 0; cudaEvent_t ev = nullptr; bool used = false; };
    UpSlot upSlot[8];
    int upNext = 0;
    UpSlot& acquireUpSlot(size_t n)
    {
        UpSlot& s = upSlot[upNext];
        upNext = (upNext + 1) % 8;
        if (s.used) CUDA_OK(cudaEventSynchronize(s.ev));      // the slot's previous transfer (8 calls ago) has long finished
        if (s.cap < n) {
            if (s.h) cudaFreeHost(s.h);
            if (s.d) cudaFree(s.d);
            s.h = s.d = nullptr;
            s.cap = 0;
            const size_t cap = std::max<size_t>(n, 4096);
            CUDA_OK(cudaHostAlloc(&s.h, cap * sizeof(double), cudaHostAllocDefault));
            CUDA_OK(cudaMalloc(&s.d, cap * sizeof(double)));
            s.cap = cap;
        }
        if (!s.ev) CUDA_OK(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
        return s;
    }
    // slope limiter (hdg_euler_limit): topology / reference-node arrays and the work arrays, built on first use per mesh
    int* dLimInts = nullptr;
    double* dLimDoubles = nullptr;
    double limCabc[4] = {0, 0, 0, 0};      // LimiterView::cabc, summed when the arrays above are built
    std::vector<std::unique_ptr<State>> states;
    std::vector<HaloPatch> halo;
    std::map<int64_t, std::vector<double>> nodeShift;      // curved (`arc`) patches: displaced dofLocation of their owner cells, Np x 2 per cell
    std::vector<int> patchTag;      // hdg_mesh_set_patch_neighbour: orders several patches towards the same neighbour
    ParPlan par;
    int smCount = 0, eulerGrid = 0, advGrid = 0;
    int64_t launches = 0;

    void ensureStage(size_t doubles)
    {
        if (doubles <= stageDoubles) return;
        if (dStage) CUDA_OK(cudaFree(dStage));
        dStage = nullptr;
        stageDoubles = 0;
        CUDA_OK(cudaMalloc(&dStage, doubles * sizeof(double)));
        stageDoubles = doubles;
    }
    State& state(int id)
    {
        requireDevice();
        if (id < 0 || id >= (int)states.size() || !states[id]) throw std::runtime_error("invalid state id " + std::to_string(id));
        State& s = *states[id];
        if (s.upPending) {      // an asynchronous upload into this state is in flight: the compute stream orders itself after it
            cudaStreamWaitEvent(stream, s.evUp, 0);
            s.upPending = false;
        }
        if (s.readPending) {    // an asynchronous download still reads the planes: compute calls (which may overwrite them) wait for it
            cudaStreamWaitEvent(stream, s.evRead, 0);
            s.readPending = false;
        }
        s.usedSinceRead = true;
        ++s.version;                // every accessor but peekState() may be followed by a write
        return s;
    }
    State& peekState(int id)    // read-only use by a compute call: ordered after a pending upload, does not invalidate derived copies
    {
        State& s = state(id);
        --s.version;
        return s;
    }
    State& rawState(int id)     // no ordering against the transfer streams (used by the asynchronous transfers themselves)
    {
        requireDevice();
        if (id < 0 || id >= (int)states.size() || !states[id]) throw std::runtime_error("invalid state id " + std::to_string(id));
        return *states[id];
    }
    // staging space of n doubles on transfer ring `which` (0 = in, 1 = out).  Regions are reused in stream order, so a bump
    // allocator that wraps around is safe as long as one request fits.
    double* ringAlloc(int which, size_t n)
    {
        if (n > ringCap[which]) {
            if (inStream) cudaStreamSynchronize(which == 0 ? inStream : outStream);
            if (dRing[which]) cudaFree(dRing[which]);
            dRing[which] = nullptr;
            ringCap[which] = 0;
            if (cudaMalloc(&dRing[which], 2 * n * sizeof(double)) != cudaSuccess) throw std::runtime_error("cudaMalloc of the transfer staging ring failed");
            ringCap[which] = 2 * n;
            ringOff[which] = 0;
        }
        if (ringOff[which] + n > ringCap[which]) ringOff[which] = 0;
        double* ptr = dRing[which] + ringOff[which];
        ringOff[which] += (n + 15) / 16 * 16;
        return ptr;
    }
    void ensureTransferStreams()
    {
        if (inStream) return;
        if (cudaStreamCreateWithFlags(&inStream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&outStream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaEventCreateWithFlags(&evCompute, cudaEventDisableTiming) != cudaSuccess)
            throw std::runtime_error("cannot create the transfer streams");
    }
    void requireDevice() const
    {
        if (hostOnly) throw std::runtime_error("host-only context (device -1) cannot compute: there is no CPU fallback, create the context on a CUDA device");
    }
    void requireMesh() const
    {
        requireDevice();
        if (!hasRef) throw std::runtime_error("hdg_set_order has not been called");
        if (!hasMesh) throw std::runtime_error("no mesh set");
    }
    void freeMeshDevice()
    {
        for (auto& s : states)
            if (s) {
                cudaFree(s->d[0]); cudaFree(s->d[1]); cudaFree(s->res); cudaFree(s->conn); cudaFree(s->connFrozen); cudaFree(s->zip);
                if (s->evUp) { cudaEventDestroy(s->evUp); cudaEventDestroy(s->evRead); cudaEventDestroy(s->evDown); }
            }
        states.clear();
        for (auto& h : halo) {
            cudaFree(h.faceElem); cudaFree(h.faceLoc);
            if (h.ownsBuffers) { cudaFree(h.send); cudaFree(h.recv); }
        }
        halo.clear();
        nodeShift.clear();
        par.release();
        patchTag.clear();
        cudaFree(dGeo);
        dGeo = nullptr;
        cudaFree(dFlux); cudaFree(dFaceOwner); cudaFree(dElemFace);
        dFlux = nullptr; dFaceOwner = nullptr; dElemFace = nullptr;
        cudaFree(dLimInts); cudaFree(dLimDoubles);
        dLimInts = nullptr;
        dLimDoubles = nullptr;
    }
    ~hdg_context()
    {
        if (hostOnly) return;
        cudaSetDevice(device);
        freeMeshDevice();
        cudaFree(dTables); cudaFree(dAdvTables); cudaFree(dSplitTables); cudaFree(dNodeTab); cudaFree(dStage); cudaFree(dPartial);
        if (comm) { if (commDestroy) commDestroy(comm); cudaEventDestroy(evHalo); cudaFree(dReduce); }
        for (UpSlot& u : upSlot) { if (u.h) cudaFreeHost(u.h); cudaFree(u.d); if (u.ev) cudaEventDestroy(u.ev); }
        cudaFree(dRing[0]); cudaFree(dRing[1]);
        if (evCompute) cudaEventDestroy(evCompute);
        if (inStream) cudaStreamDestroy(inStream);
        if (outStream) cudaStreamDestroy(outStream);
        if (stream) cudaStreamDestroy(stream);
        if (haloStream) cudaStreamDestroy(haloStream);
    }
};

namespace {

// ---- operator fragment tables (layout documented in dg_kernels.cuh / DESIGN.md §4) -----------------------
template <int N>
void buildTablesT(const RefElement& r, std::vector<double>& tab, std::vector<double>& adv, std::vector<int>& nodeTab, std::vector<double>& split)
{
    using D = Dims<N>;
    tab.assign(D::tableDoubles, 0.0);
    split.assign(D::splitTableDoubles, 0.0);
    adv.assign(D::advTableDoublesAll, 0.0);
    nodeTab.assign(D::nodeTabInts, 0);
    const int Np = r.Np, Ng = r.Ng, Nfp = r.Nfp, Nfg = r.Nfg;
    for (int lane = 0; lane < 32; ++lane) {
        const int e = lane >> 2, j = lane & 3;
        // Vg: B[k=j][n=e] = Vg[g = 8gt+e][node = 4kt+j]; padded cubature points replicate point 0 (finite fluxes)
        for (int gt = 0; gt < D::GT; ++gt)
            for (int kt = 0; kt < D::KT; ++kt) {
                int g = gt * 8 + e;
                if (g >= Ng) g = 0;
                const int node = kt * 4 + j;
                tab[D::oVg + (gt * D::KT + kt) * 32 + lane] = node < Np ? r.Vg[(size_t)g * Np + node] : 0.0;
            }
        // Pr/Ps: B[k=j][n=e] = P[node = 8nt+e][g = 8gt+2j+h]
        for (int gt = 0; gt < D::GT; ++gt)
            for (int h = 0; h < 2; ++h)
                for (int nt = 0; nt < D::NT; ++nt) {
                    const int g = gt * 8 + 2 * j + h, node = nt * 8 + e;
                    const bool in = g < Ng && node < Np;
                    const size_t o = ((size_t)(gt * 2 + h) * D::NT + nt) * 32 + lane;
                    tab[D::oPr + o] = in ? r.Pr[(size_t)node * Ng + g] : 0.0;
                    tab[D::oPs + o] = in ? r.Ps[(size_t)node * Ng + g] : 0.0;
                }
        // trace interpolation: B[k=j][n=e] = If[p = 8fgt+e][i = 4fkt+j]
        for (int fgt = 0; fgt < D::FGT; ++fgt)
            for (int fkt = 0; fkt < D::FKT; ++fkt) {
                int p = fgt * 8 + e;
                if (p >= Nfg) p = 0;
                const int i = fkt * 4 + j;
                tab[D::oIf + (fgt * D::FKT + fkt) * 32 + lane] = i < Nfp ? r.If[(size_t)p * Nfp + i] : 0.0;
            }
        // lift (sign folded in): B[k=j][n=e] = -LIFT[node = 8nt+e][face f, p = 8fgt+2j+h]
        for (int f = 0; f < 3; ++f)
            for (int fgt = 0; fgt < D::FGT; ++fgt)
                for (int h = 0; h < 2; ++h)
                    for (int nt = 0; nt < D::NT; ++nt) {
                        const int p = fgt * 8 + 2 * j + h, node = nt * 8 + e;
                        const bool in = p < Nfg && node < Np;
                        tab[D::oLift + (((f * D::FGT + fgt) * 2 + h) * D::NT + nt) * 32 + lane] =
                            in ? -r.LIFT[(size_t)node * 3 * Nfg + f * Nfg + p] : 0.0;
                    }
        // advection: weak nodal derivative and nodal lift
        for (int kt = 0; kt < 2 * D::NT; ++kt)
            for (int nt = 0; nt < D::NT; ++nt) {
                const int in_ = 8 * (kt / 2) + 2 * j + (kt & 1), out = nt * 8 + e;     // permuted K: matches the double2 loads
                const bool in = in_ < Np && out < Np;
                adv[D::oDwr + (kt * D::NT + nt) * 32 + lane] = in ? r.Dwr[(size_t)out * Np + in_] : 0.0;
                adv[D::oDws + (kt * D::NT + nt) * 32 + lane] = in ? r.Dws[(size_t)out * Np + in_] : 0.0;
            }
        for (int f = 0; f < 3; ++f)
            for (int fkt = 0; fkt < D::FKT; ++fkt)
                for (int nt = 0; nt < D::NT; ++nt) {
                    const int i = fkt * 4 + j, out = nt * 8 + e;
                    const bool in = i < Nfp && out < Np;
                    adv[D::oLiftN + ((f * D::FKT + fkt) * D::NT + nt) * 32 + lane] = in ? -r.LIFTn[(size_t)out * 3 * Nfp + f * Nfp + i] : 0.0;
                }
        // split stage, element kernel: weak nodal derivative on the plain node order and the lift over the combined Gauss-point axis
        for (int kt = 0; kt < D::KT; ++kt)
            for (int nt = 0; nt < D::NT; ++nt) {
                const int in_ = kt * 4 + j, out = nt * 8 + e;
                const bool in = in_ < Np && out < Np;
                split[D::sDwr + (kt * D::NT + nt) * 32 + lane] = in ? r.Dwr[(size_t)out * Np + in_] : 0.0;
                split[D::sDws + (kt * D::NT + nt) * 32 + lane] = in ? r.Dws[(size_t)out * Np + in_] : 0.0;
            }
        for (int kt = 0; kt < D::KTL; ++kt)
            for (int nt = 0; nt < D::NT; ++nt) {
                const int slot = kt * 4 + j, out = nt * 8 + e;
                const bool in = slot < 3 * Nfg && out < Np;
                split[D::sLiftC + (kt * D::NT + nt) * 32 + lane] = in ? -r.LIFT[(size_t)out * 3 * Nfg + slot] : 0.0;
            }
        // combined-face nodal lift: B[k=j][n=e] = -LIFTn[node = 8nt+e][slot = 4kt+j],  slot = face*Nfp + i
        for (int kt = 0; kt < D::KTC; ++kt)
            for (int nt = 0; nt < D::NT; ++nt) {
                const int slot = kt * 4 + j, out = nt * 8 + e;
                const bool in = slot < 3 * Nfp && out < Np;
                adv[D::oLiftC + (kt * D::NT + nt) * 32 + lane] = in ? -r.LIFTn[(size_t)out * 3 * Nfp + slot] : 0.0;
            }
    }
    for (int f = 0; f < 3; ++f)
        for (int rot = 0; rot < 2; ++rot)
            for (int i = 0; i < D::NfpPad; ++i) nodeTab[(f * 2 + rot) * D::NfpPad + i] = i < Nfp ? r.f2cIdx(f, rot, i) : 0;
    // [Vg][Pr][Ps] of the split table are the fused kernel's
    std::copy(tab.begin() + D::oVg, tab.begin() + D::oIf, split.begin() + D::sVg);
    static_assert(D::sDwr - D::sVg == D::oIf - D::oVg, "split table prefix = [Vg][Pr][Ps]");
}

void buildTables(int N, const RefElement& r, std::vector<double>& tab, std::vector<double>& adv, std::vector<int>& nodeTab, std::vector<double>& split)
{
    switch (N) {
        case 1: buildTablesT<1>(r, tab, adv, nodeTab, split); break;
        case 2: buildTablesT<2>(r, tab, adv, nodeTab, split); break;
        case 3: buildTablesT<3>(r, tab, adv, nodeTab, split); break;
        case 4: buildTablesT<4>(r, tab, adv, nodeTab, split); break;
        case 5: buildTablesT<5>(r, tab, adv, nodeTab, split); break;
        case 6: buildTablesT<6>(r, tab, adv, nodeTab, split); break;
        case 7: buildTablesT<7>(r, tab, adv, nodeTab, split); break;
        case 8: buildTablesT<8>(r, tab, adv, nodeTab, split); break;
        case 9: buildTablesT<9>(r, tab, adv, nodeTab, split); break;
        case 10: buildTablesT<10>(r, tab, adv, nodeTab, split); break;
        default: throw std::runtime_error("baseOrder " + std::to_string(N) + " is not supported (1..10)");
    }
}

void uploadMesh(hdg_context* c)
{
    const Mesh& m = c->mesh;
    c->Kpad = (m.K + 7) / 8 * 8;
    c->ghostBase = c->Kpad * c->NpPad;
    c->planeStride = (c->ghostBase + m.nGhost * c->NfpPad + 15) / 16 * 16;
    if (c->hostOnly) { c->hasMesh = true; return; }
    c->freeMeshDevice();
    std::vector<double> geo((size_t)c->Kpad * 16, 0.0);
    // device record (16 doubles): rx ry sx sy | nx0 ny0 | nx1 ny1 | nx2 ny2 | Fscale0 Fscale1 Fscale2 | J | 0 0
    // (the (nx,ny) pair of a face is one aligned 16-B vector; Mesh::elementGeometry keeps the face-major host order)
    for (int64_t k = 0; k < m.K; ++k) {
        double g[16];
        m.elementGeometry(k, g);
        double* d = &geo[(size_t)k * 16];
        for (int i = 0; i < 4; ++i) d[i] = g[i];
        for (int f = 0; f < 3; ++f) {
            d[kGeoN + 2 * f] = g[4 + 3 * f];
            d[kGeoN + 2 * f + 1] = g[5 + 3 * f];
            d[kGeoFs + f] = g[6 + 3 * f];
        }
        d[13] = g[13];
    }
    for (int64_t k = m.K; k < c->Kpad; ++k) std::memcpy(&geo[(size_t)k * 16], &geo[(size_t)(m.K - 1) * 16], 16 * sizeof(double));
    CUDA_OK(cudaMalloc(&c->dGeo, geo.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(c->dGeo, geo.data(), geo.size() * sizeof(double), cudaMemcpyHostToDevice));
    // halo descriptors for every patch (used only for processor patches)
    c->halo.resize(m.patches.size());
    for (size_t p = 0; p < m.patches.size(); ++p) {
        const auto& faces = m.patches[p].faces;
        if (faces.empty()) continue;
        std::vector<int> fe(faces.size()), fl(faces.size());
        for (size_t i = 0; i < faces.size(); ++i) { fe[i] = m.faceOwner[faces[i]]; fl[i] = m.faceLocO[faces[i]]; }
        CUDA_OK(cudaMalloc(&c->halo[p].faceElem, fe.size() * sizeof(int)));
        CUDA_OK(cudaMalloc(&c->halo[p].faceLoc, fl.size() * sizeof(int)));
        CUDA_OK(cudaMemcpy(c->halo[p].faceElem, fe.data(), fe.size() * sizeof(int), cudaMemcpyHostToDevice));
        CUDA_OK(cudaMemcpy(c->halo[p].faceLoc, fl.data(), fl.size() * sizeof(int), cudaMemcpyHostToDevice));
    }
    c->hasMesh = true;
}

// per-state connectivity: depends on the state's patch kinds
void refreshConn(hdg_context* c, State& s)
{
    const Mesh& m = c->mesh;
    static_assert(kCodeFaceMask == 0x3 && kCodeRev == 0x4 && kCodeGhost == 0x8 && kCodeReflect == 0x10 && kCodeOwner == 0x20,
                  "Mesh::connCodes writes these bytes");
    static_assert(HDG_BC_FIXED_VALUE == 0 && HDG_BC_ZERO_GRADIENT == 1 && HDG_BC_REFLECTIVE == 2 && HDG_BC_PROCESSOR == 3, "Mesh::connCodes");
    auto build = [&](const std::vector<int>& kinds, int4*& dev) {
        std::vector<int4> conn((size_t)c->Kpad);
        m.connCodes(kinds.data(), reinterpret_cast<int32_t*>(conn.data()));
        for (int64_t k = m.K; k < c->Kpad; ++k) conn[(size_t)k] = conn[(size_t)m.K - 1];      // padding elements repeat the last one
        if (!dev) CUDA_OK(cudaMalloc(&dev, conn.size() * sizeof(int4)));
        CUDA_OK(cudaMemcpyAsync(dev, conn.data(), conn.size() * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
        CUDA_OK(cudaStreamSynchronize(c->stream));
    };
    if (s.connDirty) {
        build(s.patchKind, s.conn);
        s.connDirty = false;
        s.connFrozenDirty = true;
    }
    if (s.frozen && s.connFrozenDirty) {
        std::vector<int> kinds(s.patchKind);
        for (int& k : kinds)
            if (k == HDG_BC_ZERO_GRADIENT || k == HDG_BC_REFLECTIVE) k |= 0x100;
        build(kinds, s.connFrozen);
        s.connFrozenDirty = false;
    }
}
inline const int4* activeConn(const State& s) { return s.frozen ? s.connFrozen : s.conn; }

// Orders that run the split stage by default (face-flux kernel + element kernel, dg_euler_split.cu); HDG_EULER_SPLIT=0 / 1 forces the
// fused kernel / the split stage for every order that has one, HDG_EULER_SPLIT=0x.. gives the order mask itself.
constexpr int kSplitDefaultOrders = 0x7fe;      // N = 1..10

bool useSplitStage(hdg_context* c)
{
    if (c->splitOrders < 0) {
        int mask = kSplitDefaultOrders;
        if (const char* e = std::getenv("HDG_EULER_SPLIT")) {
            const long v = std::strtol(e, nullptr, 0);
            mask = v == 0 ? 0 : (v == 1 ? 0x7fe : (int)v);
        }
        c->splitOrders = mask;
    }
    return ((c->splitOrders >> c->N) & 1) && eulerSplitAvailable(c->N);
}

void ensureSplit(hdg_context* c)
{
    if (c->dFlux) return;
    const Mesh& m = c->mesh;
    std::vector<int> fo((size_t)m.F);
    for (int64_t f = 0; f < m.F; ++f) fo[(size_t)f] = m.faceOwner[(size_t)f] * 4 + m.faceLocO[(size_t)f];
    std::vector<int4> ef((size_t)c->Kpad);
    for (int64_t k = 0; k < m.K; ++k) ef[(size_t)k] = make_int4(m.cellFace[3 * (size_t)k], m.cellFace[3 * (size_t)k + 1], m.cellFace[3 * (size_t)k + 2], 0);
    for (int64_t k = m.K; k < c->Kpad; ++k) ef[(size_t)k] = ef[(size_t)m.K - 1];
    CUDA_OK(cudaMalloc(&c->dFaceOwner, fo.size() * sizeof(int)));
    CUDA_OK(cudaMalloc(&c->dElemFace, ef.size() * sizeof(int4)));
    CUDA_OK(cudaMemcpy(c->dFaceOwner, fo.data(), fo.size() * sizeof(int), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(c->dElemFace, ef.data(), ef.size() * sizeof(int4), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&c->dFlux, (size_t)m.F * 4 * fluxSlotsOf(c->N) * sizeof(double)));
}

// One fused Euler stage on four planes that may live in up to three states (rho | rhoU.x,rhoU.y | Ener) - the facade
// keeps rho, rhoU, Ener as separate fields like the reference - or in one 4-plane state.
struct PlaneRef { State* s; int plane; };

void eulerStagePlanes(hdg_context* c, const PlaneRef in[4], int inWhich, const PlaneRef aux[4], int auxWhich, int outWhich,
                      State& connState, double gamma, double dt, int fluxKind, double A, double B, int mode, int64_t elemBegin = 0,
                      int64_t elemEnd = -1, int64_t elemBegin2 = 0, int64_t elemEnd2 = 0, const int* octList = nullptr, int64_t nList = 0,
                      const PlaneRef* src = nullptr, const PlaneRef* out2 = nullptr, const PlaneRef* aux2 = nullptr, double A2 = 0.0, double B2 = 0.0)
{
    // src: the element (nodal) data are read from the CURRENT copies of these planes instead of in[] (whose ghost slots still supply
    // the boundary data); out2 / aux2 / A2 / B2: optional second result A2*aux2[current] + B2*(q + dt L) into the STAGE copies of out2
    if (fluxKind != HDG_FLUX_ROE && fluxKind != HDG_FLUX_LF)
        throw std::runtime_error("Euler stage: flux scheme must be Roe (godunovScheme{fluxScheme Roe;}) or LF (point-wise local Lax-Friedrichs, an extension)");
    refreshConn(c, connState);
    StageParams p{};
    p.geo = c->dGeo;
    p.conn = activeConn(connState);
    p.tables = c->dTables;
    p.nodeTab = c->dNodeTab;
    p.K = c->mesh.K;
    if (elemEnd < 0) elemEnd = c->mesh.K;
    if (elemBegin < 0 || elemBegin > elemEnd || elemEnd > c->mesh.K || (elemBegin & 7) || ((elemEnd & 7) && elemEnd != c->mesh.K))
        throw std::runtime_error("element range must be octet-aligned: begin % 8 == 0 and (end % 8 == 0 or end == K)");
    if (elemBegin2 < 0 || elemBegin2 > elemEnd2 || elemEnd2 > c->mesh.K || (elemBegin2 & 7) || ((elemEnd2 & 7) && elemEnd2 != c->mesh.K) ||
        (elemEnd2 > elemBegin2 && elemBegin2 < elemEnd))
        throw std::runtime_error("second element range must be octet-aligned and lie after the first");
    p.octBegin = elemBegin >> 3;
    p.octEnd = (elemEnd + 7) >> 3;
    p.octBegin2 = elemBegin2 >> 3;
    p.octEnd2 = (elemEnd2 + 7) >> 3;
    p.octList = octList;
    p.nList = nList;
    if (octList ? nList == 0 : (p.octEnd == p.octBegin && p.octEnd2 == p.octBegin2)) return;
    p.ghostBase = c->ghostBase;
    p.gamma = gamma;
    p.dt = dt;
    p.A = A;
    p.B = B;
    p.mode = mode;
    p.fluxKind = fluxKind;
    p.A2 = A2;
    p.B2 = B2;
    for (int f = 0; f < 4; ++f) {
        const size_t off = (size_t)in[f].plane * c->planeStride;
        p.qghost[f] = in[f].s->d[inWhich] + off;
        p.qin[f] = src ? src[f].s->d[0] + (size_t)src[f].plane * c->planeStride : p.qghost[f];
        p.qout[f] = in[f].s->d[outWhich] + off;
        p.qout2[f] = out2 ? out2[f].s->d[1] + (size_t)out2[f].plane * c->planeStride : nullptr;
        p.qaux2[f] = (out2 && aux2) ? aux2[f].s->d[0] + (size_t)aux2[f].plane * c->planeStride : p.qghost[f];
        p.qaux[f] = aux ? aux[f].s->d[auxWhich] + (size_t)aux[f].plane * c->planeStride : p.qin[f];
        p.res[f] = mode == 1 ? in[f].s->res + off : nullptr;
    }
    const int64_t nOct = octList ? nList : (p.octEnd - p.octBegin) + (p.octEnd2 - p.octBegin2);
    // a launch over most of the mesh runs as face-flux kernel (every dgFace once) + element kernel; thin launches (the rows next to
    // the processor patches of a decomposition, which must not wait for a pass over all faces) stay on the fused kernel
    if (useSplitStage(c) && 2 * nOct > (c->mesh.K + 7) / 8) {
        ensureSplit(c);
        p.flux = c->dFlux;
        p.faceOwner = c->dFaceOwner;
        p.elemFace = c->dElemFace;
        p.F = c->mesh.F;
        p.splitTables = c->dSplitTables;
        launchEulerSplit(c->N, p, true, c->smCount, c->stream);
        CUDA_OK(cudaGetLastError());
        c->launches += 2;
        return;
    }
    const int wpb = eulerWarpsPerBlock(c->N);
    const int grid = (int)std::min<int64_t>(c->eulerGrid, (nOct + wpb - 1) / wpb);      // one octet per warp at least
    launchEulerStage(c->N, p, grid, c->stream);
    CUDA_OK(cudaGetLastError());
    ++c->launches;
}

void eulerStage(hdg_context* c, State& s, double gamma, double dt, int fluxKind, int stageIndex, double A, double B, int mode,
                int64_t elemBegin = 0, int64_t elemEnd = -1, int64_t elemBegin2 = 0, int64_t elemEnd2 = 0)
{
    if (s.nPlanes != 4) throw std::runtime_error("hdg_euler_stage needs a 4-plane state (rho, rhoU.x, rhoU.y, Ener)");
    const PlaneRef pl[4] = {{&s, 0}, {&s, 1}, {&s, 2}, {&s, 3}};
    if (mode == 0) {
        // stage 0: current -> stage copy ; stage 1: stage copy (+ A * current) -> current, in place on q_n
        if (stageIndex == 0) eulerStagePlanes(c, pl, 0, pl, 0, 1, s, gamma, dt, fluxKind, A, B, 0, elemBegin, elemEnd, elemBegin2, elemEnd2);
        else                 eulerStagePlanes(c, pl, 1, pl, 0, 0, s, gamma, dt, fluxKind, A, B, 0, elemBegin, elemEnd, elemBegin2, elemEnd2);
    } else {
        eulerStagePlanes(c, pl, stageIndex & 1, nullptr, 0, (stageIndex + 1) & 1, s, gamma, dt, fluxKind, A, B, 1, elemBegin, elemEnd, elemBegin2, elemEnd2);
    }
}

void advectStage(hdg_context* c, State& T, State& U, double dt, int fluxKind, int stageIndex, double A, double B, int mode)
{
    if (T.nPlanes != 1 || U.nPlanes != 2) throw std::runtime_error("hdg_advect_stage needs a 1-plane T state and a 2-plane U state");
    if (fluxKind != HDG_FLUX_LF && fluxKind != HDG_FLUX_AVERAGE && fluxKind != HDG_FLUX_NONE)
        throw std::runtime_error("advection stage: flux scheme must be LF, average or none");
    refreshConn(c, T);
    refreshConn(c, U);
    AdvectParams p{};
    p.U = U.d[0];
    p.geo = c->dGeo;
    p.connT = T.conn;
    p.connU = U.conn;
    p.tables = c->dAdvTables;
    p.nodeTab = c->dNodeTab;
    p.K = c->mesh.K;
    p.planeStrideT = c->planeStride;
    p.planeStrideU = c->planeStride;
    if (T.patchKind == U.patchKind) {      // same boundary kinds => identical connectivity: the kernels compute the gather offsets once
        p.connU = T.conn;
        p.sameConn = 1;
    }
    p.anyReflect = std::find(U.patchKind.begin(), U.patchKind.end(), (int)HDG_BC_REFLECTIVE) != U.patchKind.end() ? 1 : 0;
    if (advectUsesTma(c->N)) {
        // the TMA kernel reads the velocity as (x,y) pairs: one 16-B gather per trace slot instead of two 8-B ones.  The pairs are a
        // derived copy of the two planes, rebuilt only when the state may have been written since (any non-read-only access)
        if (!U.zip) CUDA_OK(cudaMalloc(&U.zip, 2 * (size_t)c->planeStride * sizeof(double)));
        if (U.external || U.zipVersion != U.version || U.zipOf != U.d[0]) {
            launchZipPlanes(U.d[0], U.d[0] + c->planeStride, U.zip, c->planeStride, c->stream);
            ++c->launches;
            U.zipVersion = U.version;
            U.zipOf = U.d[0];
        }
        p.UZ = U.zip;
    }
    p.ghostBase = c->ghostBase;
    p.dt = dt;
    p.A = A;
    p.B = B;
    p.mode = mode;
    p.fluxKind = fluxKind;
    if (mode == 0) {
        if (stageIndex == 0) { p.Tin = T.d[0]; p.Taux = T.d[0]; p.Tout = T.d[1]; }
        else                 { p.Tin = T.d[1]; p.Taux = T.d[0]; p.Tout = T.d[0]; }
    } else {
        p.Tin = T.d[stageIndex & 1];
        p.Tout = T.d[(stageIndex + 1) & 1];
        p.res = T.res;
    }
    launchAdvectStage(c->N, p, c->advGrid, c->stream);
    CUDA_OK(cudaGetLastError());
    ++c->launches;
}

void ensureRes(hdg_context* c, State& s)
{
    if (s.res) return;
    const size_t bytes = (size_t)s.nPlanes * c->planeStride * sizeof(double);
    CUDA_OK(cudaMalloc(&s.res, bytes));
    CUDA_OK(cudaMemsetAsync(s.res, 0, bytes, c->stream));
}

// LSERK(5,4) coefficients (Carpenter & Kennedy), as declared in TUT/isentropicVortex/dgEulerFoam/createFields.H:119-131
const double kRk4a[5] = {0.0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                         -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
const double kRk4b[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0, 1720146321549.0 / 2090206949498.0,
                         3134564353537.0 / 4481467310338.0, 2277821191437.0 / 14882151754819.0};

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
// HOPEDG_TRACE=1: calls and host time per C-ABI entry point, printed when the process ends (what a solver on the facade spends where)
namespace {
struct ApiTrace {
    bool on = false;
    std::map<std::string, std::pair<long, double>> acc;
    ApiTrace() { const char* e = std::getenv("HOPEDG_TRACE"); on = e && std::atoi(e) != 0; }
    ~ApiTrace()
    {
        if (!on) return;
        std::fprintf(stderr, "HOPEDG_TRACE: %-34s %10s %12s\n", "entry point", "calls", "host ms");
        for (const auto& kv : acc) std::fprintf(stderr, "HOPEDG_TRACE: %-34s %10ld %12.3f\n", kv.first.c_str(), kv.second.first, kv.second.second * 1e3);
    }
};
ApiTrace g_trace;
struct TraceScope {
    const char* name;
    std::chrono::steady_clock::time_point t0;
    explicit TraceScope(const char* n) : name(n) { if (g_trace.on) t0 = std::chrono::steady_clock::now(); }
    ~TraceScope()
    {
        if (!g_trace.on) return;
        auto& a = g_trace.acc[name];
        ++a.first;
        a.second += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
};
}  // namespace

#define HDG_TRY(ctx) \
    if (!(ctx)) return 1; \
    TraceScope traceScope_(__func__); \
    try { \
        if (!(ctx)->hostOnly) cudaSetDevice((ctx)->device);
#define HDG_CATCH(ctx) \
    } catch (const std::exception& ex) { \
        (ctx)->err = ex.what(); \
        return 1; \
    } \
    return 0;

extern "C" {
namespace {
// hdg_euler_stage_fields_ex with exchange != 0 (defined next to the other processor-patch helpers)
void parFieldsStage(hdg_context* c, const PlaneRef in[4], const PlaneRef* aux, State& connState, double gamma, double dt, int fluxKind, double A, double B,
                    const PlaneRef* out2, const PlaneRef* aux2, double A2, double B2);
}
}

extern "C" {

const char* hdg_version(void) { return "hopedg-b200 0.1 (sm_100a, FP64 DMMA)"; }

static std::string g_createError;

int hdg_create(int device, hdg_context** out)
{
    if (!out) return 1;
    *out = nullptr;
    if (device == -1) {      // host-only context: operators + connectivity, no compute
        auto* c = new hdg_context();
        c->device = -1;
        c->hostOnly = true;
        *out = c;
        return 0;
    }
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        g_createError = std::string("no CUDA device available (") + cudaGetErrorString(e) + "); this library has no CPU fallback";
        return 2;
    }
    if (device < 0 || device >= n) { g_createError = "device index out of range"; return 3; }
    auto* c = new hdg_context();
    c->device = device;
    try {
        CUDA_OK(cudaSetDevice(device));
        cudaDeviceProp prop;
        CUDA_OK(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10) throw std::runtime_error(std::string("device ") + prop.name + " is not sm_100-class; kernels are built for sm_100a only");
        c->smCount = prop.multiProcessorCount;
        CUDA_OK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        {   // the halo stream's small kernels go ahead of pending stage-kernel blocks
            int lo = 0, hi = 0;
            CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CUDA_OK(cudaStreamCreateWithPriority(&c->haloStream, cudaStreamNonBlocking, hi));
        }
        CUDA_OK(cudaMalloc(&c->dPartial, 1024 * sizeof(double)));
    } catch (const std::exception& ex) {
        g_createError = ex.what();
        delete c;
        return 4;
    }
    *out = c;
    return 0;
}

void hdg_destroy(hdg_context* ctx) { delete ctx; }

const char* hdg_last_error(const hdg_context* ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int hdg_sync(hdg_context* ctx)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->haloStream));
    if (ctx->inStream) {
        CUDA_OK(cudaStreamSynchronize(ctx->inStream));
        CUDA_OK(cudaStreamSynchronize(ctx->outStream));
        for (auto& s : ctx->states)
            if (s) { s->readPending = s->downPending = false; s->downLo = s->downHi = nullptr; }
    }
    HDG_CATCH(ctx)
}

int hdg_set_order(hdg_context* ctx, int N)
{
    HDG_TRY(ctx)
    if (ctx->hasMesh) throw std::runtime_error("hdg_set_order must be called before the mesh is set");
    ctx->ref = buildRefElement(N);
    ctx->N = N;
    ctx->NpPad = npPadOf(N);
    ctx->NfpPad = nfpPadOf(N);
    std::vector<double> tab, adv, split;
    std::vector<int> nodeTab;
    buildTables(N, ctx->ref, tab, adv, nodeTab, split);
    ctx->nodeTabHost = nodeTab;
    if (ctx->hostOnly) { ctx->hasRef = true; return 0; }
    cudaFree(ctx->dTables); cudaFree(ctx->dAdvTables); cudaFree(ctx->dNodeTab); cudaFree(ctx->dSplitTables);
    CUDA_OK(cudaMalloc(&ctx->dSplitTables, split.size() * sizeof(double)));
    CUDA_OK(cudaMemcpy(ctx->dSplitTables, split.data(), split.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMalloc(&ctx->dTables, tab.size() * sizeof(double)));
    CUDA_OK(cudaMalloc(&ctx->dAdvTables, adv.size() * sizeof(double)));
    CUDA_OK(cudaMalloc(&ctx->dNodeTab, nodeTab.size() * sizeof(int)));
    CUDA_OK(cudaMemcpy(ctx->dTables, tab.data(), tab.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ctx->dAdvTables, adv.data(), adv.size() * sizeof(double), cudaMemcpyHostToDevice));
    CUDA_OK(cudaMemcpy(ctx->dNodeTab, nodeTab.data(), nodeTab.size() * sizeof(int), cudaMemcpyHostToDevice));
    int eb = 0, ab = 0;
    size_t es = 0, as = 0;
    stageOccupancy(N, &eb, &es, &ab, &as);
    if (eb < 1 || ab < 1) throw std::runtime_error("stage kernel does not fit on an SM at this order");
    ctx->eulerGrid = ctx->smCount * eb;     // persistent grid: resident blocks per SM x SM count
    ctx->advGrid = ctx->smCount * ab;
    ctx->hasRef = true;
    HDG_CATCH(ctx)
}

int hdg_get_sizes(const hdg_context* ctx, int32_t* Np, int32_t* Nfp, int32_t* Ng, int32_t* Nfg)
{
    if (!ctx || !ctx->hasRef) return 1;
    if (Np) *Np = ctx->ref.Np;
    if (Nfp) *Nfp = ctx->ref.Nfp;
    if (Ng) *Ng = ctx->ref.Ng;
    if (Nfg) *Nfg = ctx->ref.Nfg;
    return 0;
}

int64_t hdg_get_operator(const hdg_context* ctx, const char* what, double* out, int64_t cap)
{
    if (!ctx || !ctx->hasRef || !what) return -1;
    const RefElement& r = ctx->ref;
    const std::vector<double>* v = nullptr;
    const std::string w(what);
    if (w == "r") v = &r.r; else if (w == "s") v = &r.s; else if (w == "V") v = &r.V; else if (w == "invV") v = &r.invV;
    else if (w == "Dr") v = &r.Dr; else if (w == "Ds") v = &r.Ds; else if (w == "gr") v = &r.gr; else if (w == "gs") v = &r.gs;
    else if (w == "gw") v = &r.gw; else if (w == "Vg") v = &r.Vg; else if (w == "Dgr") v = &r.Dgr; else if (w == "Dgs") v = &r.Dgs;
    else if (w == "fx") v = &r.fx; else if (w == "fw") v = &r.fw; else if (w == "If") v = &r.If; else if (w == "Mref") v = &r.Mref;
    else if (w == "Pr") v = &r.Pr; else if (w == "Ps") v = &r.Ps; else if (w == "LIFT") v = &r.LIFT; else if (w == "Dwr") v = &r.Dwr;
    else if (w == "Dws") v = &r.Dws; else if (w == "LIFTn") v = &r.LIFTn; else if (w == "faceShift") v = &r.faceShift;
    if (!v) return -1;
    if (out && cap >= (int64_t)v->size()) std::memcpy(out, v->data(), v->size() * sizeof(double));
    return (int64_t)v->size();
}

int hdg_get_face_to_cell_index(const hdg_context* ctx, int32_t* out)
{
    if (!ctx || !ctx->hasRef || !out) return 1;
    for (size_t i = 0; i < ctx->ref.f2c.size(); ++i) out[i] = ctx->ref.f2c[i];
    return 0;
}

int hdg_get_node_table(const hdg_context* ctx, int32_t* out, int32_t cap)
{
    if (!ctx || !ctx->hasRef) return -1;
    const int n = (int)ctx->nodeTabHost.size();
    if (out && cap >= n) std::memcpy(out, ctx->nodeTabHost.data(), (size_t)n * sizeof(int32_t));
    return n;
}

int hdg_set_mesh_triangles(hdg_context* ctx, int64_t nPoints, const double* xy, int64_t K, const int32_t* tris,
                           const int32_t* pointEquiv, int32_t nPatches, const int32_t* patchStart, const int32_t* edgeCell,
                           const int32_t* edgePoints)
{
    HDG_TRY(ctx)
    if (!ctx->hasRef) throw std::runtime_error("hdg_set_order must be called before hdg_set_mesh_triangles");
    if (!xy || !tris) throw std::runtime_error("null mesh arrays");
    if (K > (int64_t)1 << 30) throw std::runtime_error("too many elements for label = int32");
    static const int32_t zero2[2] = {0, 0};
    ctx->hasMesh = false;
    ctx->procAddr = Mesh::LocalMesh();
    ctx->mesh.build(nPoints, xy, K, tris, pointEquiv, nPatches, nPatches ? patchStart : zero2, edgeCell, edgePoints, nullptr, nullptr);
    uploadMesh(ctx);
    HDG_CATCH(ctx)
}

int hdg_set_mesh_polymesh(hdg_context* ctx, const char* dir)
{
    HDG_TRY(ctx)
    if (!ctx->hasRef) throw std::runtime_error("hdg_set_order must be called before hdg_set_mesh_polymesh");
    if (!dir) throw std::runtime_error("null polyMesh directory");
    ctx->hasMesh = false;
    ctx->procAddr = Mesh::LocalMesh();
    ctx->mesh.readPolyMesh(dir);
    uploadMesh(ctx);
    HDG_CATCH(ctx)
}

int hdg_decompose_simple(const hdg_context* ctx, int32_t nx, int32_t ny, int32_t nz, double delta, int32_t* cellToProc)
{
    if (!ctx || !ctx->hasMesh || !cellToProc) return 1;
    try {
        const std::vector<int32_t> d = ctx->mesh.decomposeSimple(nx, ny, nz, delta);
        std::memcpy(cellToProc, d.data(), d.size() * sizeof(int32_t));
    } catch (const std::exception& ex) {
        const_cast<hdg_context*>(ctx)->err = ex.what();
        return 1;
    }
    return 0;
}

int hdg_decompose_graph(const hdg_context* ctx, int32_t nProcs, int32_t* cellToProc)
{
    if (!ctx || !ctx->hasMesh || !cellToProc) return 1;
    try {
        const std::vector<int32_t> d = ctx->mesh.decomposeGraph(nProcs);
        std::memcpy(cellToProc, d.data(), d.size() * sizeof(int32_t));
    } catch (const std::exception& ex) {
        const_cast<hdg_context*>(ctx)->err = ex.what();
        return 1;
    }
    return 0;
}

int hdg_decompose_from_dict(const hdg_context* ctx, const char* caseDir, int32_t* nProcs, int32_t* cellToProc)
{
    if (!ctx || !ctx->hasMesh || !caseDir || !nProcs || !cellToProc) return 1;
    try {
        const std::string root(caseDir);
        const Foam::dictionary d = Foam::dictionary::fromFile(root + "/system/decomposeParDict");
        const int n = (int)d.lookup("numberOfSubdomains").readScalar();
        const std::string method = d.lookup("method")[0];
        std::vector<int32_t> c2p;
        if (method == "simple") {
            const Foam::dictionary& c = d.subDict("simpleCoeffs");
            const Foam::ITstream& nn = c.lookup("n");                 // ( nx ny nz )
            if (nn.size() < 5) throw std::runtime_error("simpleCoeffs: n must read (nx ny nz)");
            const int nx = std::atoi(nn[1].c_str()), ny = std::atoi(nn[2].c_str()), nz = std::atoi(nn[3].c_str());
            if (nx * ny * nz != n)                                    // geomDecomp.C:44-51
                throw std::runtime_error("Wrong number of processor divisions in geomDecomp:\nNumber of domains    : " + std::to_string(n) +
                                         "\nWanted decomposition : (" + nn[1] + " " + nn[2] + " " + nn[3] + ")");
            c2p = ctx->mesh.decomposeSimple(nx, ny, nz, c.found("delta") ? c.lookup("delta").readScalar() : 0.001);
        } else if (method == "manual") {                              // manualCoeffs { dataFile "cellDecomposition"; } : labelList of K entries
            std::string file = d.subDict("manualCoeffs").lookup("dataFile")[0];
            if (file.size() >= 2 && file.front() == '"') file = file.substr(1, file.size() - 2);
            std::ifstream in(root + "/constant/" + file);
            if (!in) throw std::runtime_error("cannot open manual decomposition file constant/" + file);
            std::stringstream ss;
            ss << in.rdbuf();
            std::string t = ss.str();
            const size_t h = t.find("FoamFile");
            if (h != std::string::npos) t.erase(h, t.find('}', h) - h + 1);
            for (size_t a; (a = t.find("/*")) != std::string::npos;) t.erase(a, t.find("*/", a) + 2 - a);
            for (size_t a; (a = t.find("//")) != std::string::npos;) t.erase(a, t.find('\n', a) - a);
            const size_t open = t.find('(');
            if (open == std::string::npos) throw std::runtime_error("malformed labelList in constant/" + file);
            const char* c = t.c_str() + open + 1;
            while (true) {
                char* e;
                const long v = std::strtol(c, &e, 10);
                if (e == c) break;
                c2p.push_back((int32_t)v);
                c = e;
            }
            if ((int64_t)c2p.size() != ctx->mesh.K) throw std::runtime_error("manual decomposition: " + std::to_string(c2p.size()) + " labels for " + std::to_string(ctx->mesh.K) + " cells");
            for (int32_t v : c2p) if (v < 0 || v >= n) throw std::runtime_error("manual decomposition: processor label out of range");
        } else if (method == "scotch" || method == "metis" || method == "ptscotch") {
            // the reference hands the cell-cell graph to libscotch / libmetis (scotchDecomp.C, metisDecomp.C); neither library can be built
            // here, the native recursive-bisection graph partitioner takes the same graph (the cellToProc differs from scotch's)
            c2p = ctx->mesh.decomposeGraph(n);
        } else
            throw std::runtime_error("Unknown decompositionMethod " + method + "\n\nValid decompositionMethods here are : (manual metis scotch simple)");
        *nProcs = n;
        std::memcpy(cellToProc, c2p.data(), c2p.size() * sizeof(int32_t));
    } catch (const std::exception& ex) {
        const_cast<hdg_context*>(ctx)->err = ex.what();
        return 1;
    }
    return 0;
}

int hdg_mesh_decompose(const hdg_context* global, int32_t nProcs, const int32_t* cellToProc, int32_t rank, hdg_context* local)
{
    HDG_TRY(local)
    if (!global || !global->hasMesh || !cellToProc) throw std::runtime_error("hdg_mesh_decompose: the global context has no mesh");
    if (!local->hasRef) throw std::runtime_error("hdg_set_order must be called on the local context first");
    const std::vector<int32_t> c2p(cellToProc, cellToProc + global->mesh.K);
    Mesh::LocalMesh L = global->mesh.decompose(c2p, nProcs, rank);
    if (L.cellAddr.empty()) throw std::runtime_error("processor " + std::to_string(rank) + " owns no cells");
    local->hasMesh = false;
    local->mesh.build((int64_t)L.pointAddr.size(), L.xy.data(), (int64_t)L.cellAddr.size(), L.tris.data(),
                      L.pointEquiv.empty() ? nullptr : L.pointEquiv.data(), (int)L.names.size(),
                      L.patchStart.data(), L.edgeCell.data(), L.edgePts.data(), &L.names, &L.types);
    local->procAddr = std::move(L);
    uploadMesh(local);
    HDG_CATCH(local)
}

int hdg_mesh_proc_addressing(const hdg_context* ctx, int32_t* cellProcAddressing, int32_t* pointProcAddressing, int32_t* patchNbrProc,
                             int32_t* patchFaceGlobal)
{
    if (!ctx || !ctx->hasMesh) return 1;
    const Mesh::LocalMesh& L = ctx->procAddr;
    const Mesh& m = ctx->mesh;
    const bool dec = (int64_t)L.cellAddr.size() == m.K;
    if (cellProcAddressing) for (int64_t c = 0; c < m.K; ++c) cellProcAddressing[c] = dec ? L.cellAddr[c] : (int32_t)c;
    if (pointProcAddressing) for (int64_t p = 0; p < m.nPoints; ++p) pointProcAddressing[p] = dec ? L.pointAddr[p] : (int32_t)p;
    if (patchNbrProc) for (size_t p = 0; p < m.patches.size(); ++p) patchNbrProc[p] = dec ? L.patchNbrProc[p] : m.patches[p].nbrProc;
    if (patchFaceGlobal) {
        size_t o = 0;
        for (const Patch& P : m.patches) for (int32_t fid : P.faces) { patchFaceGlobal[o] = dec ? L.patchFaceGlobal[o] : fid; ++o; }
    }
    return 0;
}

int64_t hdg_mesh_num_points(const hdg_context* ctx) { return (ctx && ctx->hasMesh) ? ctx->mesh.nPoints : -1; }

int hdg_mesh_get_points(const hdg_context* ctx, double* xy)
{
    if (!ctx || !ctx->hasMesh || !xy) return 1;
    std::memcpy(xy, ctx->mesh.xy.data(), (size_t)ctx->mesh.nPoints * 2 * sizeof(double));
    return 0;
}

int hdg_mesh_counts(const hdg_context* ctx, int64_t* K, int64_t* F, int32_t* nPatches, int64_t* nGhostFaces)
{
    if (!ctx || !ctx->hasMesh) return 1;
    if (K) *K = ctx->mesh.K;
    if (F) *F = ctx->mesh.F;
    if (nPatches) *nPatches = (int32_t)ctx->mesh.patches.size();
    if (nGhostFaces) *nGhostFaces = ctx->mesh.nGhost;
    return 0;
}

int hdg_mesh_get_faces(const hdg_context* ctx, int32_t* faceOwner, int32_t* faceNbr, int32_t* faceLocO, int32_t* faceLocN,
                       int32_t* faceRot)
{
    if (!ctx || !ctx->hasMesh) return 1;
    const Mesh& m = ctx->mesh;
    const size_t b = (size_t)m.F * sizeof(int32_t);
    if (faceOwner) std::memcpy(faceOwner, m.faceOwner.data(), b);
    if (faceNbr) std::memcpy(faceNbr, m.faceNbr.data(), b);
    if (faceLocO) std::memcpy(faceLocO, m.faceLocO.data(), b);
    if (faceLocN) std::memcpy(faceLocN, m.faceLocN.data(), b);
    if (faceRot) std::memcpy(faceRot, m.faceRot.data(), b);
    return 0;
}

int hdg_mesh_get_cell_vertices(const hdg_context* ctx, int32_t* tris)
{
    if (!ctx || !ctx->hasMesh || !tris) return 1;
    std::memcpy(tris, ctx->mesh.tris.data(), (size_t)ctx->mesh.K * 3 * sizeof(int32_t));
    return 0;
}

int hdg_mesh_patch_info(const hdg_context* ctx, int32_t p, char* name, int32_t nameCap, char* type, int32_t typeCap, int32_t* nFaces)
{
    if (!ctx || !ctx->hasMesh || p < 0 || p >= (int32_t)ctx->mesh.patches.size()) return 1;
    const Patch& P = ctx->mesh.patches[p];
    if (name && nameCap > 0) { std::strncpy(name, P.name.c_str(), nameCap - 1); name[nameCap - 1] = 0; }
    if (type && typeCap > 0) { std::strncpy(type, P.type.c_str(), typeCap - 1); type[typeCap - 1] = 0; }
    if (nFaces) *nFaces = (int32_t)P.faces.size();
    return 0;
}

int hdg_mesh_patch_faces(const hdg_context* ctx, int32_t p, int32_t* dgFaceIndex)
{
    if (!ctx || !ctx->hasMesh || p < 0 || p >= (int32_t)ctx->mesh.patches.size() || !dgFaceIndex) return 1;
    const Patch& P = ctx->mesh.patches[p];
    std::memcpy(dgFaceIndex, P.faces.data(), P.faces.size() * sizeof(int32_t));
    return 0;
}

int hdg_mesh_node_coords(const hdg_context* ctx, double* out)
{
    if (!ctx || !ctx->hasMesh || !out) return 1;
    const Mesh& m = ctx->mesh;
    const RefElement& r = ctx->ref;
    for (int64_t k = 0; k < m.K; ++k) {
        const double* v0 = &m.xy[2 * (size_t)m.tris[3 * k]];
        const double* v1 = &m.xy[2 * (size_t)m.tris[3 * k + 1]];
        const double* v2 = &m.xy[2 * (size_t)m.tris[3 * k + 2]];
        for (int i = 0; i < r.Np; ++i)
            for (int d = 0; d < 2; ++d)       // triangleBaseFunction.C:303-313
                out[((size_t)k * r.Np + i) * 2 + d] = -(r.r[i] + r.s[i]) * 0.5 * v0[d] + (r.r[i] + 1) * 0.5 * v1[d] + (r.s[i] + 1) * 0.5 * v2[d];
    }
    for (const auto& kv : ctx->nodeShift)      // cells on curved patches (hdg_mesh_set_curved_patch)
        for (int i = 0; i < 2 * r.Np; ++i) out[(size_t)kv.first * r.Np * 2 + i] += kv.second[(size_t)i];
    return 0;
}

int hdg_mesh_patch_node_coords(const hdg_context* ctx, int32_t p, double* out)
{
    if (!ctx || !ctx->hasMesh || p < 0 || p >= (int32_t)ctx->mesh.patches.size() || !out) return 1;
    const Mesh& m = ctx->mesh;
    const RefElement& r = ctx->ref;
    size_t o = 0;
    for (int32_t fid : m.patches[p].faces) {
        const int64_t k = m.faceOwner[fid];
        const int lf = m.faceLocO[fid];
        const double* v0 = &m.xy[2 * (size_t)m.tris[3 * k]];
        const double* v1 = &m.xy[2 * (size_t)m.tris[3 * k + 1]];
        const double* v2 = &m.xy[2 * (size_t)m.tris[3 * k + 2]];
        for (int i = 0; i < r.Nfp; ++i) {
            const int n = r.f2cIdx(lf, 0, i);
            const auto sh = ctx->nodeShift.find(k);
            for (int d = 0; d < 2; ++d)
                out[o++] = -(r.r[n] + r.s[n]) * 0.5 * v0[d] + (r.r[n] + 1) * 0.5 * v1[d] + (r.s[n] + 1) * 0.5 * v2[d] +
                           (sh != ctx->nodeShift.end() ? sh->second[(size_t)2 * n + d] : 0.0);
        }
    }
    return 0;
}

/* Curved boundary (patch type `arc`): positions = where the Nfp nodes of every face of the patch lie on the curve (patch-dof order,
 * x y per node; arcDgPatch::positions moves the interior face nodes onto the parametric curve and keeps the end points).  As in the
 * reference (physicalElementData::updatePatchDofIndexMapping, physicalElementData.C:185-224) the displacement is blended into the
 * owner cell's dofLocation (triangleBaseFunction::addFaceShiftToCell).  NOTE what the reference does and does not do: dgMesh.C:110-113
 * runs initElements (metrics, mass matrices, cellD1dx, face normals - all from the straight-sided nodes, physicalCellElement.C:72-96)
 * BEFORE this displacement and never recomputes them, so a curved patch changes where fields and boundary values are SAMPLED
 * (dofLocation: initial conditions, setBoundaryValues, output), not the operators.  This entry point reproduces exactly that. */
int hdg_mesh_set_curved_patch(hdg_context* ctx, int32_t patch, const double* positions)
{
    HDG_TRY(ctx)
    if (!ctx->hasMesh || patch < 0 || patch >= (int32_t)ctx->mesh.patches.size() || !positions) throw std::runtime_error("hdg_mesh_set_curved_patch: bad arguments");
    const Mesh& m = ctx->mesh;
    const RefElement& r = ctx->ref;
    size_t o = 0;
    for (int32_t fid : m.patches[(size_t)patch].faces) {
        const int64_t k = m.faceOwner[fid];
        const int lf = m.faceLocO[fid];
        const double* v0 = &m.xy[2 * (size_t)m.tris[3 * k]];
        const double* v1 = &m.xy[2 * (size_t)m.tris[3 * k + 1]];
        const double* v2 = &m.xy[2 * (size_t)m.tris[3 * k + 2]];
        std::vector<double> d((size_t)2 * r.Nfp);      // displacement of the face nodes against the straight-sided (affine) positions
        for (int i = 0; i < r.Nfp; ++i) {
            const int n = r.f2cIdx(lf, 0, i);
            for (int c = 0; c < 2; ++c)
                d[(size_t)2 * i + c] = positions[o + 2 * i + c] - (-(r.r[n] + r.s[n]) * 0.5 * v0[c] + (r.r[n] + 1) * 0.5 * v1[c] + (r.s[n] + 1) * 0.5 * v2[c]);
        }
        o += (size_t)2 * r.Nfp;
        std::vector<double>& sh = ctx->nodeShift[k];
        if (sh.empty()) sh.assign((size_t)2 * r.Np, 0.0);
        for (int p = 0; p < r.Np; ++p)
            for (int i = 0; i < r.Nfp; ++i) {
                const double w = r.faceShift[((size_t)lf * r.Np + p) * r.Nfp + i];
                sh[(size_t)2 * p] += w * d[(size_t)2 * i];
                sh[(size_t)2 * p + 1] += w * d[(size_t)2 * i + 1];
            }
    }
    HDG_CATCH(ctx)
}

// ---- states ---------------------------------------------------------------------------------------------------
int hdg_state_create(hdg_context* ctx, int32_t nPlanes, int32_t* stateId)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (nPlanes < 1 || nPlanes > 16 || !stateId) throw std::runtime_error("bad nPlanes / null id");
    auto s = std::make_unique<State>();
    s->nPlanes = nPlanes;
    const size_t bytes = (size_t)nPlanes * ctx->planeStride * sizeof(double);
    for (int w = 0; w < 2; ++w) {
        CUDA_OK(cudaMalloc(&s->d[w], bytes));
        CUDA_OK(cudaMemset(s->d[w], 0, bytes));
    }
    s->patchKind.assign(ctx->mesh.patches.size(), HDG_BC_FIXED_VALUE);
    ctx->states.push_back(std::move(s));
    *stateId = (int32_t)ctx->states.size() - 1;
    HDG_CATCH(ctx)
}

int hdg_state_destroy(hdg_context* ctx, int32_t id)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (ctx->inStream) { CUDA_OK(cudaStreamSynchronize(ctx->inStream)); CUDA_OK(cudaStreamSynchronize(ctx->outStream)); }
    cudaFree(s.d[0]); cudaFree(s.d[1]); cudaFree(s.res); cudaFree(s.conn); cudaFree(s.connFrozen); cudaFree(s.zip);
    if (s.evUp) { cudaEventDestroy(s.evUp); cudaEventDestroy(s.evRead); cudaEventDestroy(s.evDown); }
    ctx->states[id].reset();
    HDG_CATCH(ctx)
}

int hdg_state_upload(hdg_context* ctx, int32_t id, int32_t plane0, int32_t nPlanes, const double* host, int32_t hostStride)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    if (!host || hostStride < nPlanes || plane0 < 0 || nPlanes < 1 || plane0 + nPlanes > s.nPlanes) throw std::runtime_error("hdg_state_upload: bad arguments");
    const Mesh& m = ctx->mesh;
    const size_t n = (size_t)m.K * ctx->ref.Np * hostStride;
    ctx->ensureStage(n);
    CUDA_OK(cudaMemcpyAsync(ctx->dStage, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    for (int c = 0; c < nPlanes; ++c) {
        launchAosToPlane(ctx->dStage + c, hostStride, s.d[0] + (size_t)(plane0 + c) * ctx->planeStride, m.K, ctx->ref.Np, ctx->NpPad, ctx->stream);
        ++ctx->launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(ctx->stream));     // the staging buffer and `host` are reusable on return
    HDG_CATCH(ctx)
}

int hdg_state_download(hdg_context* ctx, int32_t id, int32_t plane0, int32_t nPlanes, double* host, int32_t hostStride)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    if (!host || hostStride < nPlanes || plane0 < 0 || nPlanes < 1 || plane0 + nPlanes > s.nPlanes) throw std::runtime_error("hdg_state_download: bad arguments");
    const Mesh& m = ctx->mesh;
    const size_t n = (size_t)m.K * ctx->ref.Np * hostStride;
    ctx->ensureStage(n);
    if (hostStride > nPlanes) CUDA_OK(cudaMemsetAsync(ctx->dStage, 0, n * sizeof(double), ctx->stream));
    for (int c = 0; c < nPlanes; ++c) {
        launchPlaneToAos(s.d[0] + (size_t)(plane0 + c) * ctx->planeStride, ctx->dStage + c, hostStride, m.K, ctx->ref.Np, ctx->NpPad, ctx->stream);
        ++ctx->launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(host, ctx->dStage, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    HDG_CATCH(ctx)
}

namespace {
void ensureStateEvents(State& s)
{
    if (s.evUp) return;
    CUDA_OK(cudaEventCreateWithFlags(&s.evUp, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s.evRead, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreateWithFlags(&s.evDown, cudaEventDisableTiming));
}
}  // namespace

int hdg_state_upload_async(hdg_context* ctx, int32_t id, int32_t plane0, int32_t nPlanes, const double* host, int32_t hostStride)
{
    HDG_TRY(ctx)
    State& s = ctx->rawState(id);
    ++s.version;
    if (!host || hostStride < nPlanes || plane0 < 0 || nPlanes < 1 || plane0 + nPlanes > s.nPlanes) throw std::runtime_error("hdg_state_upload_async: bad arguments");
    ctx->ensureTransferStreams();
    ensureStateEvents(s);
    const Mesh& m = ctx->mesh;
    const size_t n = (size_t)m.K * ctx->ref.Np * hostStride;
    double* stage = ctx->ringAlloc(0, n);
    // the host buffer may still be the target of an asynchronous download of this state (only then does the copy wait for it: a new
    // request that arrives in its own buffer starts at once); the planes may still be read by it (evRead below)
    {
        const char* lo = reinterpret_cast<const char*>(host);
        const char* hi = lo + n * sizeof(double);
        if (s.downPending && lo < s.downHi && s.downLo < hi) CUDA_OK(cudaStreamWaitEvent(ctx->inStream, s.evDown, 0));
    }
    CUDA_OK(cudaMemcpyAsync(stage, host, n * sizeof(double), cudaMemcpyHostToDevice, ctx->inStream));
    if (s.readPending) CUDA_OK(cudaStreamWaitEvent(ctx->inStream, s.evRead, 0));
    if (s.usedSinceRead) {      // compute work enqueued so far may still use the planes (an asynchronous download already covers it)
        CUDA_OK(cudaEventRecord(ctx->evCompute, ctx->stream));
        CUDA_OK(cudaStreamWaitEvent(ctx->inStream, ctx->evCompute, 0));
        s.usedSinceRead = false;
    }
    for (int c = 0; c < nPlanes; ++c) {
        launchAosToPlane(stage + c, hostStride, s.d[0] + (size_t)(plane0 + c) * ctx->planeStride, m.K, ctx->ref.Np, ctx->NpPad, ctx->inStream);
        ++ctx->launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(s.evUp, ctx->inStream));
    s.upPending = true;
    HDG_CATCH(ctx)
}

int hdg_state_download_async(hdg_context* ctx, int32_t id, int32_t plane0, int32_t nPlanes, double* host, int32_t hostStride)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);      // orders the compute stream after a pending upload, so the event below covers it
    if (!host || hostStride < nPlanes || plane0 < 0 || nPlanes < 1 || plane0 + nPlanes > s.nPlanes) throw std::runtime_error("hdg_state_download_async: bad arguments");
    ctx->ensureTransferStreams();
    ensureStateEvents(s);
    const Mesh& m = ctx->mesh;
    const size_t n = (size_t)m.K * ctx->ref.Np * hostStride;
    double* stage = ctx->ringAlloc(1, n);
    CUDA_OK(cudaEventRecord(ctx->evCompute, ctx->stream));      // everything enqueued on the compute stream so far
    CUDA_OK(cudaStreamWaitEvent(ctx->outStream, ctx->evCompute, 0));
    if (hostStride > nPlanes) CUDA_OK(cudaMemsetAsync(stage, 0, n * sizeof(double), ctx->outStream));
    for (int c = 0; c < nPlanes; ++c) {
        launchPlaneToAos(s.d[0] + (size_t)(plane0 + c) * ctx->planeStride, stage + c, hostStride, m.K, ctx->ref.Np, ctx->NpPad, ctx->outStream);
        ++ctx->launches;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(s.evRead, ctx->outStream));
    s.readPending = true;
    s.usedSinceRead = false;
    CUDA_OK(cudaMemcpyAsync(host, stage, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->outStream));
    CUDA_OK(cudaEventRecord(s.evDown, ctx->outStream));
    {
        const char* lo = reinterpret_cast<const char*>(host);
        const char* hi = lo + n * sizeof(double);
        s.downLo = (s.downPending && s.downLo < lo) ? s.downLo : lo;
        s.downHi = (s.downPending && s.downHi > hi) ? s.downHi : hi;
    }
    s.downPending = true;
    HDG_CATCH(ctx)
}

int hdg_state_set_patch_kind(hdg_context* ctx, int32_t id, int32_t patch, int32_t kind)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    if (patch < 0 || patch >= (int32_t)s.patchKind.size()) throw std::runtime_error("patch index out of range");
    if (kind < HDG_BC_FIXED_VALUE || kind > HDG_BC_EMPTY) throw std::runtime_error("unknown patch field kind");
    if (s.patchKind[patch] != kind) { s.patchKind[patch] = kind; s.connDirty = true; }
    HDG_CATCH(ctx)
}

int hdg_state_set_patch_values(hdg_context* ctx, int32_t id, int32_t plane0, int32_t nPlanes, int32_t patch, const double* values,
                               int32_t hostStride)
{
    HDG_TRY(ctx)
    State& s = ctx->peekState(id);
    {   // element data and processor ghosts are untouched: exchanged ghosts stay current
        const bool fresh = s.haloVersion == s.version;
        ++s.version;
        if (fresh) s.haloVersion = s.version;
    }
    const Mesh& m = ctx->mesh;
    if (patch < 0 || patch >= (int32_t)m.patches.size()) throw std::runtime_error("patch index out of range");
    if (!values || hostStride < nPlanes || plane0 < 0 || nPlanes < 1 || plane0 + nPlanes > s.nPlanes) throw std::runtime_error("hdg_state_set_patch_values: bad arguments");
    const Patch& P = m.patches[patch];
    const int64_t nF = (int64_t)P.faces.size();
    if (nF == 0) return 0;
    const size_t n = (size_t)nF * ctx->ref.Nfp * hostStride;
    // asynchronous: pinned slot -> device slot -> ONE kernel for all planes and both copies (stage 2 reuses the t_n boundary data,
    // dgEulerFoam.C:73,99-113); `values` is free for the caller on return, nothing waits for the stream
    hdg_context::UpSlot& u = ctx->acquireUpSlot(n);
    std::memcpy(u.h, values, n * sizeof(double));
    CUDA_OK(cudaMemcpyAsync(u.d, u.h, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    const size_t goff = (size_t)plane0 * ctx->planeStride + ctx->ghostBase + P.ghostStart * ctx->NfpPad;
    launchPatchToGhostAll(u.d, hostStride, s.d[0] + goff, s.d[1] + goff, ctx->planeStride, nPlanes, nF, ctx->ref.Nfp, ctx->NfpPad, ctx->stream);
    ++ctx->launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaEventRecord(u.ev, ctx->stream));
    u.used = true;
    HDG_CATCH(ctx)
}

int hdg_state_copy(hdg_context* ctx, int32_t dst, int32_t src)
{
    HDG_TRY(ctx)
    State &d = ctx->state(dst), &s = ctx->state(src);
    if (d.nPlanes != s.nPlanes) throw std::runtime_error("hdg_state_copy: plane count mismatch");
    // field assignment copies the internal field and the boundary field (rho1 = rho, dgEulerFoam.C:70-72); the stage
    // copy receives the same data so that its ghost (boundary) slots are valid for the next stage
    const size_t bytes = (size_t)s.nPlanes * ctx->planeStride * sizeof(double);
    CUDA_OK(cudaMemcpyAsync(d.d[0], s.d[0], bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    // the stage copy only needs the ghost (boundary) region: its nodal part is overwritten by the next stage before anything reads it
    const size_t gb = (size_t)(ctx->planeStride - ctx->ghostBase) * sizeof(double);
    if (gb) CUDA_OK(cudaMemcpy2DAsync(d.d[1] + ctx->ghostBase, ctx->planeStride * sizeof(double), s.d[0] + ctx->ghostBase, ctx->planeStride * sizeof(double), gb,
                                      (size_t)s.nPlanes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d.patchKind != s.patchKind) { d.patchKind = s.patchKind; d.connDirty = true; }
    d.frozen = s.frozen;      // the frozen traces travel with the ghost region
    // ... and so do exchanged processor ghosts (s.version was bumped by the accessor above: compare against version - 1)
    if (s.haloVersion == s.version - 1 && s.haloFreshBuf == s.d[0]) { s.haloVersion = s.version; d.haloFreshBuf = d.d[0]; d.haloVersion = d.version; }
    HDG_CATCH(ctx)
}

int hdg_state_l1_diff(hdg_context* ctx, int32_t id, int32_t plane, const double* ref, int32_t hostStride, double* out)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    if (!ref || !out || plane < 0 || plane >= s.nPlanes || hostStride < 1) throw std::runtime_error("hdg_state_l1_diff: bad arguments");
    const Mesh& m = ctx->mesh;
    const size_t n = (size_t)m.K * ctx->ref.Np * hostStride;
    const size_t planeD = (size_t)ctx->Kpad * ctx->NpPad;
    ctx->ensureStage(n + planeD);
    double* scratch = ctx->dStage + n;
    CUDA_OK(cudaMemcpyAsync(ctx->dStage, ref, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    launchAosToPlane(ctx->dStage, hostStride, scratch, m.K, ctx->ref.Np, ctx->NpPad, ctx->stream);
    const int nb = 512;
    launchL1Diff(s.d[0] + (size_t)plane * ctx->planeStride, scratch, m.K, ctx->ref.Np, ctx->NpPad, ctx->dPartial, nb, ctx->stream);
    ctx->launches += 2;
    std::vector<double> part(nb);
    CUDA_OK(cudaMemcpyAsync(part.data(), ctx->dPartial, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (double v : part) sum += v;
    *out = sum;
    HDG_CATCH(ctx)
}

// ---- hot path ---------------------------------------------------------------------------------------------------
int hdg_euler_stage(hdg_context* ctx, int32_t id, double gamma, double dt, int32_t fluxKind, int32_t stageIndex, double a, double b)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (stageIndex != 0 && stageIndex != 1) throw std::runtime_error("stageIndex must be 0 or 1");
    eulerStage(ctx, ctx->state(id), gamma, dt, fluxKind, stageIndex, a, b, 0);
    HDG_CATCH(ctx)
}

int hdg_euler_stage_range(hdg_context* ctx, int32_t id, double gamma, double dt, int32_t fluxKind, int32_t stageIndex, double a, double b,
                          int64_t elemBegin, int64_t elemEnd, int64_t elemBegin2, int64_t elemEnd2)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (stageIndex != 0 && stageIndex != 1) throw std::runtime_error("stageIndex must be 0 or 1");
    eulerStage(ctx, ctx->state(id), gamma, dt, fluxKind, stageIndex, a, b, 0, elemBegin, elemEnd, elemBegin2, elemEnd2);
    HDG_CATCH(ctx)
}

int hdg_stream_wait(hdg_context* ctx, int32_t waiter, int32_t signaler)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    if ((waiter != 0 && waiter != 1) || (signaler != 0 && signaler != 1) || waiter == signaler) throw std::runtime_error("hdg_stream_wait: streams are 0 (compute) and 1 (halo)");
    cudaEvent_t ev;
    CUDA_OK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventRecord(ev, signaler == 0 ? ctx->stream : ctx->haloStream));
    CUDA_OK(cudaStreamWaitEvent(waiter == 0 ? ctx->stream : ctx->haloStream, ev, 0));
    CUDA_OK(cudaEventDestroy(ev));
    HDG_CATCH(ctx)
}

int hdg_euler_step_ssprk2(hdg_context* ctx, int32_t id, double gamma, double dt, int32_t fluxKind)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State& s = ctx->state(id);
    eulerStage(ctx, s, gamma, dt, fluxKind, 0, 0.0, 1.0, 0);
    eulerStage(ctx, s, gamma, dt, fluxKind, 1, 0.5, 0.5, 0);
    HDG_CATCH(ctx)
}

int hdg_euler_step_lserk45(hdg_context* ctx, int32_t id, double gamma, double dt, int32_t fluxKind)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State& s = ctx->state(id);
    ensureRes(ctx, s);
    for (int st = 0; st < 5; ++st) eulerStage(ctx, s, gamma, dt, fluxKind, st, kRk4a[st], kRk4b[st], 1);
    // 5 stages ping-pong current -> stage -> ... and end in the stage copy: swap so that `current` holds q^{n+1}
    std::swap(s.d[0], s.d[1]);
    HDG_CATCH(ctx)
}

int hdg_advect_stage(hdg_context* ctx, int32_t idT, int32_t idU, double dt, int32_t fluxKind, int32_t stageIndex, double a, double b)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (stageIndex != 0 && stageIndex != 1) throw std::runtime_error("stageIndex must be 0 or 1");
    advectStage(ctx, ctx->state(idT), ctx->peekState(idU), dt, fluxKind, stageIndex, a, b, 0);
    HDG_CATCH(ctx)
}

int hdg_advect_step_ssprk2(hdg_context* ctx, int32_t idT, int32_t idU, double dt, int32_t fluxKind)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State &T = ctx->state(idT), &U = ctx->peekState(idU);
    advectStage(ctx, T, U, dt, fluxKind, 0, 0.0, 1.0, 0);
    advectStage(ctx, T, U, dt, fluxKind, 1, 0.5, 0.5, 0);
    HDG_CATCH(ctx)
}

int hdg_advect_step_lserk45(hdg_context* ctx, int32_t idT, int32_t idU, double dt, int32_t fluxKind)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State &T = ctx->state(idT), &U = ctx->peekState(idU);
    ensureRes(ctx, T);
    for (int st = 0; st < 5; ++st) advectStage(ctx, T, U, dt, fluxKind, st, kRk4a[st], kRk4b[st], 1);
    std::swap(T.d[0], T.d[1]);      // 5 ping-pong stages end in the stage copy
    HDG_CATCH(ctx)
}

int hdg_euler_stage_fields(hdg_context* ctx, int32_t sRho, int32_t sRhoU, int32_t sEner, double gamma, double dt, int32_t fluxKind,
                           double a, double b, int32_t auxRho, int32_t auxRhoU, int32_t auxEner)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State &r = ctx->state(sRho), &u = ctx->state(sRhoU), &e = ctx->state(sEner);
    if (r.nPlanes != 1 || u.nPlanes != 2 || e.nPlanes != 1) throw std::runtime_error("hdg_euler_stage_fields: rho/Ener must be 1-plane and rhoU a 2-plane state");
    for (size_t p = 0; p < u.patchKind.size(); ++p) {
        const bool gu = u.patchKind[p] == HDG_BC_FIXED_VALUE || u.patchKind[p] == HDG_BC_PROCESSOR;
        const bool gr = r.patchKind[p] == HDG_BC_FIXED_VALUE || r.patchKind[p] == HDG_BC_PROCESSOR;
        const bool ge = e.patchKind[p] == HDG_BC_FIXED_VALUE || e.patchKind[p] == HDG_BC_PROCESSOR;
        if (gu != gr || gu != ge)
            throw std::runtime_error("patch " + ctx->mesh.patches[p].name + ": rho, rhoU and Ener must all be fixedValue/processor or all be "
                                     "zeroGradient/reflective on the fused Euler path");
    }
    if ((r.frozen || u.frozen || e.frozen) && !(r.frozen && u.frozen && e.frozen))
        throw std::runtime_error("hdg_euler_stage_fields: rho, rhoU and Ener must be frozen together (hdg_state_freeze_traces)");
    const PlaneRef in[4] = {{&r, 0}, {&u, 0}, {&u, 1}, {&e, 0}};
    if (a != 0.0) {
        State &ar = ctx->state(auxRho), &au = ctx->state(auxRhoU), &ae = ctx->state(auxEner);
        if (ar.nPlanes != 1 || au.nPlanes != 2 || ae.nPlanes != 1) throw std::runtime_error("hdg_euler_stage_fields: bad aux states");
        const PlaneRef aux[4] = {{&ar, 0}, {&au, 0}, {&au, 1}, {&ae, 0}};
        eulerStagePlanes(ctx, in, 0, aux, 0, 1, u, gamma, dt, fluxKind, a, b, 0);
    } else
        eulerStagePlanes(ctx, in, 0, nullptr, 0, 1, u, gamma, dt, fluxKind, 0.0, b, 0);
    r.frozen = u.frozen = e.frozen = false;      // the new fields are evaluated afresh (correctBoundaryConditions after the solve)
    HDG_CATCH(ctx)
}

int hdg_state_copy_ghosts(hdg_context* ctx, int32_t dst, int32_t src)
{
    HDG_TRY(ctx)
    State &d = ctx->peekState(dst), &s = ctx->peekState(src);
    if (d.nPlanes != s.nPlanes) throw std::runtime_error("hdg_state_copy_ghosts: plane count mismatch");
    const size_t gb = (size_t)(ctx->planeStride - ctx->ghostBase) * sizeof(double), pitch = ctx->planeStride * sizeof(double);
    if (gb)
        for (int w = 0; w < 2; ++w)
            CUDA_OK(cudaMemcpy2DAsync(d.d[w] + ctx->ghostBase, pitch, s.d[0] + ctx->ghostBase, pitch, gb, (size_t)s.nPlanes, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d.patchKind != s.patchKind) { d.patchKind = s.patchKind; d.connDirty = true; }
    d.frozen = s.frozen;
    HDG_CATCH(ctx)
}

int hdg_euler_stage_fields_ex(hdg_context* ctx, const hdg_euler_fields_stage* st)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (!st) throw std::runtime_error("hdg_euler_stage_fields_ex: null descriptor");
    auto triple = [&](const int32_t id[3], PlaneRef pl[4], const char* what) {
        State *s0 = &ctx->state(id[0]), *s1 = &ctx->state(id[1]), *s2 = &ctx->state(id[2]);
        if (s0->nPlanes != 1 || s1->nPlanes != 2 || s2->nPlanes != 1)
            throw std::runtime_error(std::string("hdg_euler_stage_fields_ex: ") + what + ": rho / Ener must be 1-plane and rhoU a 2-plane state");
        pl[0] = {s0, 0}; pl[1] = {s1, 0}; pl[2] = {s1, 1}; pl[3] = {s2, 0};
    };
    PlaneRef in[4], src[4], out2[4], aux[4], aux2[4];
    triple(st->s, in, "fields");
    State &r = *in[0].s, &u = *in[1].s, &e = *in[3].s;
    for (size_t p = 0; p < u.patchKind.size(); ++p) {
        const bool gu = u.patchKind[p] == HDG_BC_FIXED_VALUE || u.patchKind[p] == HDG_BC_PROCESSOR;
        const bool gr = r.patchKind[p] == HDG_BC_FIXED_VALUE || r.patchKind[p] == HDG_BC_PROCESSOR;
        const bool ge = e.patchKind[p] == HDG_BC_FIXED_VALUE || e.patchKind[p] == HDG_BC_PROCESSOR;
        if (gu != gr || gu != ge)
            throw std::runtime_error("patch " + ctx->mesh.patches[p].name + ": rho, rhoU and Ener must all be fixedValue/processor or all be "
                                     "zeroGradient/reflective on the fused Euler path");
    }
    if ((r.frozen || u.frozen || e.frozen) && !(r.frozen && u.frozen && e.frozen))
        throw std::runtime_error("hdg_euler_stage_fields_ex: rho, rhoU and Ener must be frozen together (hdg_state_freeze_traces)");
    const bool hasSrc = st->src[0] >= 0, hasOut2 = st->out2[0] >= 0, hasAux = st->a != 0.0, hasAux2 = hasOut2 && st->a2 != 0.0;
    if (hasSrc) triple(st->src, src, "src");
    if (hasOut2) triple(st->out2, out2, "out2");
    if (hasAux) triple(st->aux, aux, "aux");
    if (hasAux2) triple(st->aux2, aux2, "aux2");
    if (st->exchange && ctx->comm) {
        if (hasSrc) throw std::runtime_error("hdg_euler_stage_fields_ex: exchange cannot be combined with src[] (the traces sent are those of s[])");
        parFieldsStage(ctx, in, hasAux ? aux : nullptr, u, st->gamma, st->dt, st->fluxKind, hasAux ? st->a : 0.0, st->b, hasOut2 ? out2 : nullptr,
                       hasAux2 ? aux2 : nullptr, hasAux2 ? st->a2 : 0.0, st->b2);
    } else
        eulerStagePlanes(ctx, in, 0, hasAux ? aux : nullptr, 0, 1, u, st->gamma, st->dt, st->fluxKind, hasAux ? st->a : 0.0, st->b, 0, 0, -1, 0, 0, nullptr, 0,
                         hasSrc ? src : nullptr, hasOut2 ? out2 : nullptr, hasAux2 ? aux2 : nullptr, hasAux2 ? st->a2 : 0.0, st->b2);
    r.frozen = u.frozen = e.frozen = false;
    HDG_CATCH(ctx)
}

// ---- slope limiter ---------------------------------------------------------------------------------------------------
// column sums of the reference mass matrix (V V^T)^-1 = invV^T invV, halved (Trianglelimite.C:109-116; baseFunction.C:63-77)
static std::vector<double> limiterWeights(const RefElement& r)
{
    std::vector<double> mpp((size_t)r.Np, 0.0);
    for (int j = 0; j < r.Np; ++j) {
        double sum = 0;
        for (int i = 0; i < r.Np; ++i)
            for (int m = 0; m < r.Np; ++m) sum += r.invV[(size_t)m * r.Np + i] * r.invV[(size_t)m * r.Np + j];
        mpp[(size_t)j] = 0.5 * sum;
    }
    return mpp;
}

int hdg_mesh_conn_codes(const hdg_context* ctx, const int32_t* patchKind, int32_t nPatches, int32_t* out)
{
    if (!ctx || !ctx->hasMesh || !patchKind || !out || nPatches != (int32_t)ctx->mesh.patches.size()) return 1;
    try {
        static_assert(sizeof(int) == sizeof(int32_t), "patch kinds are passed as int");
        ctx->mesh.connCodes(reinterpret_cast<const int*>(patchKind), out);
    } catch (const std::exception&) { return 1; }
    return 0;
}

int hdg_mesh_boundary_slots(const hdg_context* ctx, int32_t* bslot, int32_t* ghostFirst)
{
    if (!ctx || !ctx->hasMesh || !bslot || (!ghostFirst && ctx->mesh.nGhost > 0)) return 1;
    ctx->mesh.boundarySlots(bslot, ghostFirst);
    return 0;
}

int hdg_limiter_weights(const hdg_context* ctx, double* mpp)
{
    if (!ctx || !ctx->hasRef || !mpp) return 1;
    const std::vector<double> w = limiterWeights(ctx->ref);
    std::memcpy(mpp, w.data(), w.size() * sizeof(double));
    return 0;
}

int hdg_state_freeze_traces(hdg_context* ctx, int32_t id)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State& s = ctx->state(id);
    const Mesh& m = ctx->mesh;
    for (size_t p = 0; p < m.patches.size(); ++p) {
        const int kind = s.patchKind[p];
        const int64_t nF = (int64_t)m.patches[p].faces.size();
        if (nF == 0 || (kind != HDG_BC_ZERO_GRADIENT && kind != HDG_BC_REFLECTIVE)) continue;
        ctx->ensureStage((size_t)nF * ctx->NfpPad * s.nPlanes);
        const HaloPatch& h = ctx->halo[p];
        launchHaloPack(s.d[0], ctx->planeStride, s.nPlanes, h.faceElem, h.faceLoc, ctx->dNodeTab, nF, ctx->ref.Nfp, ctx->NfpPad, ctx->NpPad,
                       ctx->dStage, ctx->stream, 0);
        launchHaloUnpack(ctx->dStage, s.d[0], ctx->planeStride, s.nPlanes, ctx->ghostBase + m.patches[p].ghostStart * ctx->NfpPad, nF,
                         ctx->NfpPad, ctx->stream);
        CUDA_OK(cudaGetLastError());
        ctx->launches += 2;
    }
    s.frozen = true;
    HDG_CATCH(ctx)
}

int hdg_state_thaw(hdg_context* ctx, int32_t id)
{
    HDG_TRY(ctx)
    ctx->state(id).frozen = false;
    HDG_CATCH(ctx)
}

int hdg_euler_limit(hdg_context* ctx, int32_t sRho, int32_t sRhoU, int32_t sEner, double gamma, double eps, double tol)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (ctx->commSize > 1) throw std::runtime_error("hdg_euler_limit: processor patches are not handled (the reference leaves them empty too, Trianglelimite.C:170-172)");
    State &r = ctx->state(sRho), &u = ctx->state(sRhoU), &e = ctx->state(sEner);
    if (r.nPlanes != 1 || u.nPlanes != 2 || e.nPlanes != 1) throw std::runtime_error("hdg_euler_limit: rho/Ener must be 1-plane and rhoU a 2-plane state");
    for (size_t p = 0; p < u.patchKind.size(); ++p) {
        const bool gu = u.patchKind[p] == HDG_BC_FIXED_VALUE, gr = r.patchKind[p] == HDG_BC_FIXED_VALUE, ge = e.patchKind[p] == HDG_BC_FIXED_VALUE;
        if (r.patchKind[p] == HDG_BC_PROCESSOR) throw std::runtime_error("hdg_euler_limit: processor patches are not handled");
        if (gu != gr || gu != ge) throw std::runtime_error("patch " + ctx->mesh.patches[p].name + ": rho, rhoU and Ener must all be fixedValue or none");
    }
    refreshConn(ctx, r);
    refreshConn(ctx, u);
    const Mesh& m = ctx->mesh;
    const RefElement& ref = ctx->ref;
    const int64_t K = m.K, tot = K + m.nGhost;
    const size_t nInts = (size_t)3 * K + (size_t)std::max<int64_t>(m.nGhost, 1);
    const size_t oVerts = 0, oR = oVerts + 6 * (size_t)K, oS = oR + ref.Np, oMpp = oS + ref.Np, oWork = (oMpp + ref.Np + 7) / 8 * 8;      // records start on a 64-B boundary
    const size_t nWork = 8 * (size_t)tot + 12 * (size_t)K + 8 * (size_t)tot;      // cell records, vertex records, cell-gradient records
    if (!ctx->dLimInts) {
        std::vector<int32_t> ints(nInts, 0);
        m.boundarySlots(ints.data(), ints.data() + 3 * K);
        std::vector<double> dbl(oWork, 0.0);
        for (int64_t k = 0; k < K; ++k)
            for (int v = 0; v < 3; ++v) {
                dbl[oVerts + 6 * (size_t)k + 2 * v] = m.xy[2 * (size_t)m.tris[3 * k + v]];
                dbl[oVerts + 6 * (size_t)k + 2 * v + 1] = m.xy[2 * (size_t)m.tris[3 * k + v] + 1];
            }
        const std::vector<double> mpp = limiterWeights(ref);
        for (int i = 0; i < ref.Np; ++i) { dbl[oR + i] = ref.r[i]; dbl[oS + i] = ref.s[i]; dbl[oMpp + i] = mpp[(size_t)i]; }
        {   // centroid weights of the affine node map (triangleBaseFunction.C:303-313), summed in node order as limCellAverages does
            double a = 0, b = 0, c = 0, w = 0;
            for (int i = 0; i < ref.Np; ++i) {
                const double wi = mpp[(size_t)i];
                a += -(ref.r[i] + ref.s[i]) * 0.5 * wi;
                b += (ref.r[i] + 1.0) * 0.5 * wi;
                c += (ref.s[i] + 1.0) * 0.5 * wi;
                w += wi;
            }
            ctx->limCabc[0] = a; ctx->limCabc[1] = b; ctx->limCabc[2] = c; ctx->limCabc[3] = w;
        }
        CUDA_OK(cudaMalloc(&ctx->dLimInts, nInts * sizeof(int32_t)));
        CUDA_OK(cudaMalloc(&ctx->dLimDoubles, (oWork + nWork) * sizeof(double)));
        CUDA_OK(cudaMemcpyAsync(ctx->dLimInts, ints.data(), nInts * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_OK(cudaMemcpyAsync(ctx->dLimDoubles, dbl.data(), oWork * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        CUDA_OK(cudaStreamSynchronize(ctx->stream));      // the host vectors go out of scope
    }
    LimiterView v{};
    v.K = K; v.nGhost = m.nGhost; v.ghostBase = ctx->ghostBase;
    v.Np = ref.Np; v.NpPad = ctx->NpPad; v.Nfp = ref.Nfp; v.NfpPad = ctx->NfpPad;
    double* planes[4] = {r.d[0], u.d[0], u.d[0] + ctx->planeStride, e.d[0]};
    for (int f = 0; f < 4; ++f) { v.q[f] = planes[f]; v.qout[f] = planes[f]; }
    v.connS = reinterpret_cast<const int*>(r.conn);
    v.connU = reinterpret_cast<const int*>(u.conn);
    v.bslot = ctx->dLimInts;
    v.ghostFirst = ctx->dLimInts + 3 * K;
    double* d = ctx->dLimDoubles;
    v.verts = d + oVerts; v.r = d + oR; v.s = d + oS; v.mpp = d + oMpp;
    v.nodeTab = ctx->dNodeTab;
    double* w = d + oWork;
    v.cell = w; w += 8 * tot;
    v.vtx = w; w += 12 * K;
    v.CV = w; w += 8 * tot;
    v.V = nullptr; v.A2 = nullptr;      // per-face arrays of the five-pass form (host harness only)
    v.gamma = gamma; v.eps = eps; v.tol = tol;
    {
        static const int streamPlanes = [] { const char* e = std::getenv("HDG_LIM_STREAM"); return e ? std::atoi(e) : 1; }();      // A/B aid
        v.streamPlanes = streamPlanes;
    }
    for (int i = 0; i < 4; ++i) v.cabc[i] = ctx->limCabc[i];
    ctx->launches += launchTriangleLimiter(v, ctx->stream);
    CUDA_OK(cudaGetLastError());
    HDG_CATCH(ctx)
}

int hdg_state_swap(hdg_context* ctx, int32_t id)
{
    HDG_TRY(ctx)
    State& s = ctx->peekState(id);
    const bool fresh = s.haloVersion == s.version;      // the exchanged ghosts travel with their buffer
    std::swap(s.d[0], s.d[1]);
    ++s.version;
    if (fresh) s.haloVersion = s.version;
    HDG_CATCH(ctx)
}

int hdg_state_axpby(hdg_context* ctx, int32_t dst, double a, int32_t x, double b, int32_t y)
{
    HDG_TRY(ctx)
    State &D = ctx->state(dst), &X = ctx->state(x), &Y = ctx->state(y);
    if (D.nPlanes != X.nPlanes || D.nPlanes != Y.nPlanes) throw std::runtime_error("hdg_state_axpby: plane count mismatch");
    launchAxpby(D.d[0], a, X.d[0], b, Y.d[0], (int64_t)D.nPlanes * ctx->planeStride, ctx->stream);
    D.frozen = X.frozen && Y.frozen;      // the ghost region is combined too: frozen only if both operands carry frozen traces
    {   // exchanged processor ghosts combine linearly with the field: current afterwards iff they were current in both operands
        // (the accessors above bumped each distinct state's version once per access)
        auto wasFresh = [&](State& S, int accesses) { return S.haloFreshBuf == S.d[0] && S.haloVersion + (uint64_t)accesses == S.version; };
        const int nx = 1 + (&X == &D ? 1 : 0) + (&X == &Y ? 1 : 0), ny = 1 + (&Y == &D ? 1 : 0) + (&X == &Y ? 1 : 0);
        const int nd = 1 + (&X == &D ? 1 : 0) + (&Y == &D ? 1 : 0);
        const bool fx = wasFresh(X, nx), fy = wasFresh(Y, ny);
        if (fx && &X != &D) X.haloVersion = X.version;
        if (fy && &Y != &D && &Y != &X) Y.haloVersion = Y.version;
        (void)nd;
        if (fx && fy) { D.haloFreshBuf = D.d[0]; D.haloVersion = D.version; }
    }
    CUDA_OK(cudaGetLastError());
    ++ctx->launches;
    HDG_CATCH(ctx)
}

// ---- communicator (one process per GPU) ------------------------------------------------------------------------------
static void ensureHaloBuffers(hdg_context* ctx, HaloPatch& h, int64_t need);
namespace {
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi& nccl()
{
    static NcclApi api;
    if (api.lib) return api;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
        api.lib = dlopen(name, RTLD_NOW | RTLD_LOCAL);
        if (api.lib) break;
    }
    if (!api.lib) throw std::runtime_error(std::string("cannot load libnccl.so.2: ") + dlerror());
    auto sym = [&](const char* n) {
        void* p = dlsym(api.lib, n);
        if (!p) throw std::runtime_error(std::string("libnccl lacks ") + n);
        return p;
    };
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
    api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
    api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
    api.Send = reinterpret_cast<decltype(api.Send)>(sym("ncclSend"));
    api.Recv = reinterpret_cast<decltype(api.Recv)>(sym("ncclRecv"));
    api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
    return api;
}
#define NCCL_OK(call)                                                                                        \
    do {                                                                                                     \
        ncclResult_t r_ = (call);                                                                            \
        if (r_ != ncclSuccess) throw std::runtime_error(std::string(#call) + " failed: " + nccl().GetErrorString(r_)); \
    } while (0)
}  // namespace

int hdg_comm_init(hdg_context* ctx, int32_t rank, int32_t worldSize, const char* idFile)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    if (ctx->comm) throw std::runtime_error("hdg_comm_init: the communicator exists already");
    if (worldSize < 1 || rank < 0 || rank >= worldSize || !idFile) throw std::runtime_error("hdg_comm_init: bad arguments");
    ncclUniqueId id;
    const std::string path(idFile), tmp = path + ".tmp";
    if (rank == 0) {      // rank 0 publishes the id through the (shared) file system: write, then rename
        NCCL_OK(nccl().GetUniqueId(&id));
        std::ofstream(tmp, std::ios::binary).write(reinterpret_cast<const char*>(&id), sizeof(id));
        if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("hdg_comm_init: cannot write " + path);
    } else {
        bool ok = false;
        for (int i = 0; i < 1200 && !ok; ++i) {      // up to 2 minutes
            std::ifstream in(path, std::ios::binary);
            if (in && in.read(reinterpret_cast<char*>(&id), sizeof(id))) ok = true;
            else std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
        if (!ok) throw std::runtime_error("hdg_comm_init: rank 0 did not publish " + path);
    }
    NCCL_OK(nccl().CommInitRank(&ctx->comm, worldSize, id, rank));
    if (rank == 0) std::remove(path.c_str());      // every rank has joined: the id file has served its purpose
    ctx->commDestroy = nccl().CommDestroy;
    ctx->commRank = rank;
    ctx->commSize = worldSize;
    CUDA_OK(cudaEventCreateWithFlags(&ctx->evHalo, cudaEventDisableTiming));
    CUDA_OK(cudaMalloc(&ctx->dReduce, 4096 * sizeof(double)));
    HDG_CATCH(ctx)
}

int hdg_comm_rank_size(const hdg_context* ctx, int32_t* rank, int32_t* size)
{
    if (!ctx) return 1;
    if (rank) *rank = ctx->commRank;
    if (size) *size = ctx->commSize;
    return 0;
}

// =============================================================================================================
// Overlapped processor-patch exchange inside the library (any decomposition)
//   replaces processorDgPatchField::initEvaluate/evaluate (processorDgPatchField.C:235-331): the reference exchanges before
//   every evaluation and hides nothing.  Here every stage is  [octets owning a processor face] -> pack -> transport -> unpack
//   on the halo stream, under the launch of [all other octets] on the compute stream.  Transport = grouped ncclSend/ncclRecv
//   (one process per GPU, hdg_comm_init) or peer copies between the contexts of ONE process (hdg_group_*).
// =============================================================================================================
namespace {

int nbrProcOf(const hdg_context* c, size_t p)
{
    const Mesh& m = c->mesh;
    const bool dec = (int64_t)c->procAddr.cellAddr.size() == m.K && c->procAddr.patchNbrProc.size() == m.patches.size();
    if (m.patches[p].nbrProc >= 0) return m.patches[p].nbrProc;
    return dec ? c->procAddr.patchNbrProc[p] : -1;
}

void buildParPlan(hdg_context* c)
{
    ParPlan& P = c->par;
    if (P.built) return;
    const Mesh& m = c->mesh;
    struct Ent { int nbr, tag, patch; };
    std::vector<Ent> ents;
    for (size_t p = 0; p < m.patches.size(); ++p) {
        const int q = nbrProcOf(c, p);
        if (q >= 0 && !m.patches[p].faces.empty()) ents.push_back({q, p < c->patchTag.size() ? c->patchTag[p] : 0, (int)p});
    }
    std::sort(ents.begin(), ents.end(), [](const Ent& x, const Ent& y) { return x.nbr != y.nbr ? x.nbr < y.nbr : (x.tag != y.tag ? x.tag < y.tag : x.patch < y.patch); });
    std::vector<int> fe, fl, fg;
    const int64_t nOct = c->Kpad / 8;
    std::vector<char> isB((size_t)nOct, 0);
    P.faceOff.assign(1, 0);
    for (const Ent& e : ents) {
        const Patch& pt = m.patches[(size_t)e.patch];
        for (size_t i = 0; i < pt.faces.size(); ++i) {
            const int32_t fid = pt.faces[i];
            fe.push_back(m.faceOwner[fid]);
            fl.push_back(m.faceLocO[fid]);
            fg.push_back((int)(pt.ghostStart + (int64_t)i));
            isB[(size_t)(m.faceOwner[fid] >> 3)] = 1;
        }
        P.patches.push_back(e.patch); P.nbr.push_back(e.nbr); P.tag.push_back(e.tag);
        P.faceOff.push_back((int64_t)fe.size());
    }
    P.nPF = (int64_t)fe.size();
    std::vector<int> ob, oi;
    for (int64_t o = 0; o < nOct; ++o) (isB[(size_t)o] ? ob : oi).push_back((int)o);
    P.nB = (int64_t)ob.size();
    P.nI = (int64_t)oi.size();
    auto up = [&](int*& d, const std::vector<int>& h) {
        CUDA_OK(cudaMalloc(&d, std::max<size_t>(h.size(), 1) * sizeof(int)));
        if (!h.empty()) CUDA_OK(cudaMemcpy(d, h.data(), h.size() * sizeof(int), cudaMemcpyHostToDevice));
    };
    up(P.dFaceElem, fe); up(P.dFaceLoc, fl); up(P.dFaceGhost, fg); up(P.dOctB, ob); up(P.dOctI, oi);
    const size_t nb = (size_t)std::max<int64_t>(P.nPF, 1) * 4 * c->NfpPad * sizeof(double);
    CUDA_OK(cudaMalloc(&P.send, nb));
    CUDA_OK(cudaMalloc(&P.recv, nb));
    CUDA_OK(cudaMemset(P.send, 0, nb));
    CUDA_OK(cudaMemset(P.recv, 0, nb));
    for (cudaEvent_t* e : {&P.evBoundary, &P.evHalo, &P.evPacked, &P.evCopied}) CUDA_OK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
    P.built = true;
}

// the four conserved planes of one copy, wherever they live (one 4-plane state, or rho | rhoU | Ener)
struct PlaneSet {
    PlaneRef pl[4];
    HaloPlanes ptrs(const hdg_context* c, int which) const
    {
        HaloPlanes h;
        for (int f = 0; f < 4; ++f) h.p[f] = pl[f].s->d[which] + (size_t)pl[f].plane * c->planeStride;
        return h;
    }
    bool fresh(int which) const
    {
        for (int f = 0; f < 4; ++f)
            if (pl[f].s->haloFreshBuf != pl[f].s->d[which] || pl[f].s->haloVersion != pl[f].s->version) return false;
        return true;
    }
    void markFresh(int which) const
    {
        for (int f = 0; f < 4; ++f) { pl[f].s->haloFreshBuf = pl[f].s->d[which]; pl[f].s->haloVersion = pl[f].s->version; }
    }
    void bump() const      // contents changed: each distinct state once
    {
        for (int f = 0; f < 4; ++f)
            if (f == 0 || pl[f].s != pl[f - 1].s) ++pl[f].s->version;
    }
};

// halo stream: pack the processor-face traces of copy `which` (after `after` on the compute stream when given)
void parPack(hdg_context* c, const PlaneSet& ps, int which, bool afterCompute)
{
    ParPlan& P = c->par;
    if (afterCompute) {
        CUDA_OK(cudaEventRecord(P.evBoundary, c->stream));
        CUDA_OK(cudaStreamWaitEvent(c->haloStream, P.evBoundary, 0));
    }
    launchHaloPackAll(ps.ptrs(c, which), 4, P.dFaceElem, P.dFaceLoc, c->dNodeTab, P.nPF, c->ref.Nfp, c->NfpPad, c->NpPad, P.send, c->haloStream);
    CUDA_OK(cudaGetLastError());
    ++c->launches;
}
void parUnpack(hdg_context* c, const PlaneSet& ps, int which)
{
    ParPlan& P = c->par;
    launchHaloUnpackAll(P.recv, ps.ptrs(c, which), 4, P.dFaceGhost, c->ghostBase, P.nPF, c->NfpPad, c->haloStream);
    CUDA_OK(cudaGetLastError());
    ++c->launches;
    CUDA_OK(cudaEventRecord(P.evHalo, c->haloStream));
    P.haloPending = true;
}
void parWaitHalo(hdg_context* c)      // compute stream: the ghosts of the last exchange have landed
{
    ParPlan& P = c->par;
    if (P.haloPending) { CUDA_OK(cudaStreamWaitEvent(c->stream, P.evHalo, 0)); P.haloPending = false; }
}
void parTransportNccl(hdg_context* c)
{
    ParPlan& P = c->par;
    if (P.patches.empty()) return;
    if (!c->comm) throw std::runtime_error("processor patches need a communicator: call hdg_comm_init first");
    const size_t per = (size_t)4 * c->NfpPad;
    NCCL_OK(nccl().GroupStart());
    for (size_t i = 0; i < P.patches.size(); ++i) {
        const size_t off = (size_t)P.faceOff[i] * per, n = (size_t)(P.faceOff[i + 1] - P.faceOff[i]) * per;
        NCCL_OK(nccl().Send(P.send + off, n, ncclDouble, P.nbr[i], c->comm, c->haloStream));
        NCCL_OK(nccl().Recv(P.recv + off, n, ncclDouble, P.nbr[i], c->comm, c->haloStream));
    }
    NCCL_OK(nccl().GroupEnd());
}
// contexts of one process (index = rank): every receive segment is a peer copy from the matching send segment of the neighbour
void parTransportGroup(hdg_context** ctxs, int n)
{
    for (int r = 0; r < n; ++r) {
        hdg_context* c = ctxs[r];
        ParPlan& P = c->par;
        CUDA_OK(cudaSetDevice(c->device));
        CUDA_OK(cudaEventRecord(P.evPacked, c->haloStream));
    }
    for (int r = 0; r < n; ++r) {
        hdg_context* c = ctxs[r];
        ParPlan& P = c->par;
        CUDA_OK(cudaSetDevice(c->device));
        const size_t per = (size_t)4 * c->NfpPad;
        for (size_t i = 0; i < P.patches.size(); ++i) {
            const int q = P.nbr[i];
            if (q < 0 || q >= n) throw std::runtime_error("neighbour processor " + std::to_string(q) + " is not in the group");
            ParPlan& Q = ctxs[q]->par;
            size_t j = 0;
            for (; j < Q.patches.size(); ++j)
                if (Q.nbr[j] == r && Q.tag[j] == P.tag[i]) break;
            if (j == Q.patches.size() || Q.faceOff[j + 1] - Q.faceOff[j] != P.faceOff[i + 1] - P.faceOff[i])
                throw std::runtime_error("processor patch " + c->mesh.patches[(size_t)P.patches[i]].name + " has no matching patch on processor " + std::to_string(q));
            CUDA_OK(cudaStreamWaitEvent(c->haloStream, Q.evPacked, 0));
            const size_t bytes = (size_t)(P.faceOff[i + 1] - P.faceOff[i]) * per * sizeof(double);
            CUDA_OK(cudaMemcpyPeerAsync(P.recv + (size_t)P.faceOff[i] * per, c->device, Q.send + (size_t)Q.faceOff[j] * per, ctxs[q]->device, bytes, c->haloStream));
        }
        CUDA_OK(cudaEventRecord(P.evCopied, c->haloStream));
        P.copiedPending = true;
    }
}
// a neighbour may still be copying out of this context's send buffer: the next pack waits for it
void parGroupGuardSend(hdg_context** ctxs, int n)
{
    for (int r = 0; r < n; ++r) {
        hdg_context* c = ctxs[r];
        CUDA_OK(cudaSetDevice(c->device));
        for (int q : c->par.nbr)
            if (q >= 0 && q < n && ctxs[q]->par.copiedPending) CUDA_OK(cudaStreamWaitEvent(c->haloStream, ctxs[q]->par.evCopied, 0));
    }
}

struct ParJob {
    hdg_context* c;
    PlaneSet q;            // the planes advanced
    const PlaneSet* aux;   // q_n of the SSP combination (nullptr: none)
    State* connState;
};

// one stage on every job: q[outWhich] = A*aux[auxWhich] + B*(q[inWhich] + dt*L(q[inWhich])), ghosts of q[outWhich] exchanged under the
// interior launch.  jobs.size() == 1 with NCCL transport, or the contexts of a group (index = rank).
void parStage(std::vector<ParJob>& jobs, bool group, int inWhich, int auxWhich, int outWhich, double gamma, double dt, int fluxKind, double A, double B)
{
    const int n = (int)jobs.size();
    std::vector<hdg_context*> ctxs;
    for (ParJob& j : jobs) ctxs.push_back(j.c);
    // ghosts of the input copy: fresh from the previous stage's exchange, else a blocking exchange now
    bool need = false;
    for (ParJob& j : jobs) { buildParPlan(j.c); need = need || !j.q.fresh(inWhich); }
    if (need) {
        if (group) parGroupGuardSend(ctxs.data(), n);
        for (ParJob& j : jobs) { cudaSetDevice(j.c->device); parPack(j.c, j.q, inWhich, true); }
        if (group) parTransportGroup(ctxs.data(), n); else parTransportNccl(jobs[0].c);
        for (ParJob& j : jobs) { cudaSetDevice(j.c->device); parUnpack(j.c, j.q, inWhich); }
    }
    for (ParJob& j : jobs) {
        hdg_context* c = j.c;
        cudaSetDevice(c->device);
        parWaitHalo(c);
        eulerStagePlanes(c, j.q.pl, inWhich, j.aux ? j.aux->pl : nullptr, auxWhich, outWhich, *j.connState, gamma, dt, fluxKind, A, B, 0, 0, -1, 0, 0,
                         c->par.dOctB, c->par.nB);
    }
    if (group) parGroupGuardSend(ctxs.data(), n);
    for (ParJob& j : jobs) { cudaSetDevice(j.c->device); parPack(j.c, j.q, outWhich, true); }
    if (group) parTransportGroup(ctxs.data(), n); else parTransportNccl(jobs[0].c);
    for (ParJob& j : jobs) {
        hdg_context* c = j.c;
        cudaSetDevice(c->device);
        parUnpack(c, j.q, outWhich);
        eulerStagePlanes(c, j.q.pl, inWhich, j.aux ? j.aux->pl : nullptr, auxWhich, outWhich, *j.connState, gamma, dt, fluxKind, A, B, 0, 0, -1, 0, 0,
                         c->par.dOctI, c->par.nI);
        j.q.bump();
        j.q.markFresh(outWhich);
    }
}

// One stage of three separate fields on a decomposed mesh, halo hidden behind compute (the mirror image of parStage, which needs the
// NEXT stage's input to look ahead): the traces of the CURRENT values travel on the halo stream while the octets without a processor
// face are advanced; the octets next to processor patches follow when the ghosts have landed.
void parFieldsStage(hdg_context* c, const PlaneRef in[4], const PlaneRef* aux, State& connState, double gamma, double dt, int fluxKind, double A, double B,
                    const PlaneRef* out2, const PlaneRef* aux2, double A2, double B2)
{
    buildParPlan(c);
    ParPlan& P = c->par;
    PlaneSet ps;
    for (int f = 0; f < 4; ++f) ps.pl[f] = in[f];
    if (P.nPF > 0) {
        parPack(c, ps, 0, true);
        parTransportNccl(c);
        parUnpack(c, ps, 0);
    }
    if (P.nI > 0)
        eulerStagePlanes(c, in, 0, aux, 0, 1, connState, gamma, dt, fluxKind, A, B, 0, 0, -1, 0, 0, P.dOctI, P.nI, nullptr, out2, aux2, A2, B2);
    parWaitHalo(c);
    if (P.nB > 0)
        eulerStagePlanes(c, in, 0, aux, 0, 1, connState, gamma, dt, fluxKind, A, B, 0, 0, -1, 0, 0, P.dOctB, P.nB, nullptr, out2, aux2, A2, B2);
    ps.markFresh(0);
}

PlaneSet planeSetOf(State& s)
{
    if (s.nPlanes != 4) throw std::runtime_error("the parallel Euler step needs a 4-plane state (rho, rhoU.x, rhoU.y, Ener)");
    PlaneSet ps;
    for (int f = 0; f < 4; ++f) ps.pl[f] = {&s, f};
    return ps;
}

}  // namespace

int hdg_mesh_set_patch_neighbour(hdg_context* ctx, int32_t patch, int32_t nbrRank, int32_t tag)
{
    HDG_TRY(ctx)
    if (!ctx->hasMesh || patch < 0 || patch >= (int32_t)ctx->mesh.patches.size()) throw std::runtime_error("hdg_mesh_set_patch_neighbour: bad patch");
    ctx->mesh.patches[(size_t)patch].nbrProc = nbrRank;
    ctx->patchTag.resize(ctx->mesh.patches.size(), 0);
    ctx->patchTag[(size_t)patch] = tag;
    if (!ctx->hostOnly) ctx->par.release();
    HDG_CATCH(ctx)
}

int hdg_par_counts(hdg_context* ctx, int64_t* nProcFaces, int64_t* nBoundaryOctets, int64_t* nInteriorOctets, int32_t* nNeighbours)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    buildParPlan(ctx);
    if (nProcFaces) *nProcFaces = ctx->par.nPF;
    if (nBoundaryOctets) *nBoundaryOctets = ctx->par.nB;
    if (nInteriorOctets) *nInteriorOctets = ctx->par.nI;
    if (nNeighbours) *nNeighbours = (int32_t)ctx->par.patches.size();
    HDG_CATCH(ctx)
}

/* the processor-patch halo of one state copy, blocking order (pack -> grouped send/recv -> unpack, then the compute stream continues):
 * processorDgPatchField::initEvaluate/evaluate (processorDgPatchField.C:235-331).  Marks the copy's ghosts as current. */
int hdg_halo_exchange(hdg_context* ctx, int32_t id, int32_t which)
{
    HDG_TRY(ctx)
    State& s = ctx->peekState(id);
    if (which != 0 && which != 1) throw std::runtime_error("hdg_halo_exchange: bad arguments");
    if (!ctx->comm) throw std::runtime_error("hdg_halo_exchange: call hdg_comm_init first");
    buildParPlan(ctx);
    ParPlan& P = ctx->par;
    if (P.patches.empty()) return 0;
    if (s.nPlanes > 4) throw std::runtime_error("hdg_halo_exchange: at most 4 planes per state");
    PlaneSet ps;
    for (int f = 0; f < 4; ++f) ps.pl[f] = {&s, std::min(f, s.nPlanes - 1)};      // planes beyond nPlanes repeat the last one (harmless duplicates)
    parPack(ctx, ps, which, true);
    parTransportNccl(ctx);
    parUnpack(ctx, ps, which);
    parWaitHalo(ctx);
    s.haloFreshBuf = s.d[which];
    s.haloVersion = s.version;
    HDG_CATCH(ctx)
}

/* One SSP-RK2 step (dgEulerFoam.C:67-117) of a 4-plane state on a processor mesh, the exchange of every stage's result overlapped with the
 * launch over the octets that own no processor face.  NCCL transport (hdg_comm_init); collective over the ranks that share patches. */
int hdg_euler_step_ssprk2_parallel(hdg_context* ctx, int32_t id, double gamma, double dt, int32_t fluxKind)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    State& s = ctx->peekState(id);
    std::vector<ParJob> jobs(1);
    jobs[0] = {ctx, planeSetOf(s), nullptr, &s};
    const PlaneSet aux = jobs[0].q;
    jobs[0].aux = &aux;
    parStage(jobs, false, 0, 0, 1, gamma, dt, fluxKind, 0.0, 1.0);
    parStage(jobs, false, 1, 0, 0, gamma, dt, fluxKind, 0.5, 0.5);
    HDG_CATCH(ctx)
}

/* The same step on n contexts of THIS process (index = processor number of the decomposition): one GPU each, or several on one GPU.
 * Transport = peer copies (cudaMemcpyPeerAsync) ordered by events; everything else - octet lists, pack / unpack kernels, ordering - is the
 * code path of hdg_euler_step_ssprk2_parallel.  A single-process multi-GPU driver, and the way the halo is tested on one GPU. */
int hdg_group_euler_step_ssprk2(hdg_context** ctxs, const int32_t* stateIds, int32_t n, double gamma, double dt, int32_t fluxKind)
{
    if (!ctxs || !stateIds || n < 1 || !ctxs[0]) return 1;
    try {
        std::vector<ParJob> jobs((size_t)n);
        std::vector<PlaneSet> aux((size_t)n);
        for (int r = 0; r < n; ++r) {
            if (!ctxs[r]) throw std::runtime_error("null context in the group");
            cudaSetDevice(ctxs[r]->device);
            ctxs[r]->requireMesh();
            State& s = ctxs[r]->peekState(stateIds[r]);
            aux[(size_t)r] = planeSetOf(s);
            jobs[(size_t)r] = {ctxs[r], aux[(size_t)r], &aux[(size_t)r], &s};
        }
        parStage(jobs, true, 0, 0, 1, gamma, dt, fluxKind, 0.0, 1.0);
        parStage(jobs, true, 1, 0, 0, gamma, dt, fluxKind, 0.5, 0.5);
    } catch (const std::exception& ex) {
        for (int r = 0; r < n; ++r)
            if (ctxs[r]) ctxs[r]->err = ex.what();
        return 1;
    }
    return 0;
}

/* host values summed over all ranks, in place (gSum of the reference's error print-outs); n <= 4096 */
int hdg_comm_allreduce_sum(hdg_context* ctx, double* hostValues, int32_t n)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    if (!hostValues || n < 1 || n > 4096) throw std::runtime_error("hdg_comm_allreduce_sum: bad arguments");
    if (!ctx->comm) return 0;      // one rank: nothing to add
    CUDA_OK(cudaMemcpyAsync(ctx->dReduce, hostValues, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_OK(nccl().AllReduce(ctx->dReduce, ctx->dReduce, (size_t)n, ncclDouble, ncclSum, ctx->comm, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(hostValues, ctx->dReduce, n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    HDG_CATCH(ctx)
}

/* one int64 per rank -> all ranks (sizes for the global dof ranges, dgMesh::updateLocalRange); out has commSize entries */
int hdg_comm_allgather_i64(hdg_context* ctx, int64_t value, int64_t* out)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    if (!out) throw std::runtime_error("hdg_comm_allgather_i64: bad arguments");
    if (!ctx->comm) { out[0] = value; return 0; }
    if (ctx->commSize > 2048) throw std::runtime_error("hdg_comm_allgather_i64: too many ranks");
    int64_t* d = reinterpret_cast<int64_t*>(ctx->dReduce);
    CUDA_OK(cudaMemcpyAsync(d + 2048, &value, sizeof(int64_t), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_OK(nccl().AllGather(d + 2048, d, 1, ncclInt64, ctx->comm, ctx->stream));
    CUDA_OK(cudaMemcpyAsync(out, d, ctx->commSize * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    HDG_CATCH(ctx)
}

// ---- halo ---------------------------------------------------------------------------------------------------------
int hdg_halo_counts(const hdg_context* ctx, int32_t patch, int64_t* nDoublesPerPlane)
{
    if (!ctx || !ctx->hasMesh || patch < 0 || patch >= (int32_t)ctx->mesh.patches.size() || !nDoublesPerPlane) return 1;
    *nDoublesPerPlane = (int64_t)ctx->mesh.patches[patch].faces.size() * ctx->NfpPad;
    return 0;
}

int hdg_halo_bind(hdg_context* ctx, int32_t patch, void* devSend, void* devRecv, int64_t capDoubles)
{
    HDG_TRY(ctx)
    ctx->requireMesh();
    if (patch < 0 || patch >= (int32_t)ctx->halo.size()) throw std::runtime_error("patch index out of range");
    HaloPatch& h = ctx->halo[patch];
    if (h.ownsBuffers) { cudaFree(h.send); cudaFree(h.recv); h.ownsBuffers = false; }
    h.send = (double*)devSend;
    h.recv = (double*)devRecv;
    h.capDoubles = capDoubles;
    HDG_CATCH(ctx)
}

static void ensureHaloBuffers(hdg_context* ctx, HaloPatch& h, int64_t need)
{
    if (h.capDoubles >= need && h.send && h.recv) return;
    if (h.send && !h.ownsBuffers) throw std::runtime_error("bound halo buffers are too small");
    if (h.ownsBuffers) { cudaFree(h.send); cudaFree(h.recv); }
    CUDA_OK(cudaMalloc(&h.send, need * sizeof(double)));
    CUDA_OK(cudaMalloc(&h.recv, need * sizeof(double)));
    h.capDoubles = need;
    h.ownsBuffers = true;
}

int hdg_halo_pack(hdg_context* ctx, int32_t id, int32_t which, int32_t patch, void** devSendBuf, int64_t* nDoubles)
{
    HDG_TRY(ctx)
    const bool pend = ctx->rawState(id).upPending || ctx->rawState(id).readPending;
    State& s = ctx->state(id);      // orders the COMPUTE stream after pending asynchronous transfers of this state ...
    if (pend) {                     // ... and the halo stream (where pack / unpack run) after the compute stream
        CUDA_OK(cudaEventRecord(ctx->evCompute, ctx->stream));
        CUDA_OK(cudaStreamWaitEvent(ctx->haloStream, ctx->evCompute, 0));
    }
    if (patch < 0 || patch >= (int32_t)ctx->halo.size() || (which != 0 && which != 1)) throw std::runtime_error("hdg_halo_pack: bad arguments");
    const int64_t nF = (int64_t)ctx->mesh.patches[patch].faces.size();
    const int64_t need = nF * ctx->NfpPad * s.nPlanes;
    HaloPatch& h = ctx->halo[patch];
    ensureHaloBuffers(ctx, h, need);
    launchHaloPack(s.d[which], ctx->planeStride, s.nPlanes, h.faceElem, h.faceLoc, ctx->dNodeTab, nF, ctx->ref.Nfp, ctx->NfpPad, ctx->NpPad,
                   h.send, ctx->haloStream);
    CUDA_OK(cudaGetLastError());
    ++ctx->launches;
    if (devSendBuf) *devSendBuf = h.send;
    if (nDoubles) *nDoubles = need;
    HDG_CATCH(ctx)
}

int hdg_halo_recv_buffer(hdg_context* ctx, int32_t id, int32_t patch, void** devRecvBuf, int64_t* nDoubles)
{
    HDG_TRY(ctx)
    State& s = ctx->state(id);
    if (patch < 0 || patch >= (int32_t)ctx->halo.size()) throw std::runtime_error("hdg_halo_recv_buffer: bad arguments");
    const int64_t need = (int64_t)ctx->mesh.patches[patch].faces.size() * ctx->NfpPad * s.nPlanes;
    HaloPatch& h = ctx->halo[patch];
    ensureHaloBuffers(ctx, h, need);
    if (devRecvBuf) *devRecvBuf = h.recv;
    if (nDoubles) *nDoubles = need;
    HDG_CATCH(ctx)
}

int hdg_halo_unpack(hdg_context* ctx, int32_t id, int32_t which, int32_t patch)
{
    HDG_TRY(ctx)
    const bool pend = ctx->rawState(id).upPending || ctx->rawState(id).readPending;
    State& s = ctx->state(id);
    if (pend) {
        CUDA_OK(cudaEventRecord(ctx->evCompute, ctx->stream));
        CUDA_OK(cudaStreamWaitEvent(ctx->haloStream, ctx->evCompute, 0));
    }
    if (patch < 0 || patch >= (int32_t)ctx->halo.size() || (which != 0 && which != 1)) throw std::runtime_error("hdg_halo_unpack: bad arguments");
    const Patch& P = ctx->mesh.patches[patch];
    HaloPatch& h = ctx->halo[patch];
    if (!h.recv) throw std::runtime_error("hdg_halo_unpack: no receive buffer");
    launchHaloUnpack(h.recv, s.d[which], ctx->planeStride, s.nPlanes, ctx->ghostBase + P.ghostStart * ctx->NfpPad, (int64_t)P.faces.size(),
                     ctx->NfpPad, ctx->haloStream);
    CUDA_OK(cudaGetLastError());
    ++ctx->launches;
    HDG_CATCH(ctx)
}

void* hdg_stream(hdg_context* ctx, int32_t which) { return ctx ? (void*)(which == 0 ? ctx->stream : ctx->haloStream) : nullptr; }

int64_t hdg_launch_count(const hdg_context* ctx) { return ctx ? ctx->launches : 0; }

int hdg_euler_stage_kernels(hdg_context* ctx, char* out, int32_t cap)
{
    if (!ctx || !ctx->hasRef || !out || cap < 1) return 0;
    const std::string n = std::to_string(ctx->N);
    const bool split = useSplitStage(ctx);
    const std::string fk = ctx->N <= 2 ? "eulerFacePairFluxKernel<" : "eulerFaceFluxKernel<";      // N = 1, 2: two faces per DMMA row
    const std::string s = split ? fk + n + ">+eulerElemKernel<" + n + ">" : "eulerStageKernel<" + n + ">";
    std::snprintf(out, (size_t)cap, "%s", s.c_str());
    return split ? 2 : 1;
}

int hdg_measure_fp64_peak(hdg_context* ctx, double seconds, double* tflops)
{
    HDG_TRY(ctx)
    ctx->requireDevice();
    if (!tflops || !(seconds > 0) || seconds > 30) throw std::runtime_error("hdg_measure_fp64_peak: bad arguments");
    const int blocks = ctx->smCount * 4;
    ctx->ensureStage((size_t)blocks * 512);
    cudaEvent_t e0, e1;
    CUDA_OK(cudaEventCreate(&e0));
    CUDA_OK(cudaEventCreate(&e1));
    const int iters = 4096;
    launchFp64Peak(ctx->dStage, blocks, iters, ctx->stream);      // warm-up
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    double flops = 0, ms = 0;
    const auto t0 = std::chrono::steady_clock::now();
    do {      // back-to-back launches for `seconds`: a sustained figure at the clocks the stage kernel sees
        float m;
        CUDA_OK(cudaEventRecord(e0, ctx->stream));
        double f = 0;
        for (int i = 0; i < 8; ++i) f += launchFp64Peak(ctx->dStage, blocks, iters, ctx->stream);
        CUDA_OK(cudaEventRecord(e1, ctx->stream));
        CUDA_OK(cudaEventSynchronize(e1));
        CUDA_OK(cudaEventElapsedTime(&m, e0, e1));
        flops += f;
        ms += m;
        ctx->launches += 8;
    } while (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < seconds);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = flops / (ms * 1e-3) * 1e-12;
    HDG_CATCH(ctx)
}

void* hdg_state_device_ptr(hdg_context* ctx, int32_t id, int32_t which)
{
    if (!ctx || id < 0 || id >= (int)ctx->states.size() || !ctx->states[id] || (which != 0 && which != 1)) return nullptr;
    ctx->states[id]->external = true;
    return ctx->states[id]->d[which];
}

int hdg_layout(const hdg_context* ctx, int64_t* Kpad, int32_t* NpPad, int32_t* NfpPad, int64_t* planeStride, int64_t* ghostBase,
               int32_t* eulerGrid, int32_t* advectGrid)
{
    if (!ctx || !ctx->hasMesh) return 1;
    if (Kpad) *Kpad = ctx->Kpad;
    if (NpPad) *NpPad = ctx->NpPad;
    if (NfpPad) *NfpPad = ctx->NfpPad;
    if (planeStride) *planeStride = ctx->planeStride;
    if (ghostBase) *ghostBase = ctx->ghostBase;
    if (eulerGrid) *eulerGrid = ctx->eulerGrid;
    if (advectGrid) *advectGrid = ctx->advGrid;
    return 0;
}

}  // extern "C"

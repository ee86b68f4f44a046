// Device helpers shared by the stage kernels (dg_kernels.cu, dg_euler_split.cu): DMMA wrapper, TMA staging of the operator
// tables, MUFU-seeded reciprocal / rsqrt, and the point-wise Euler physics (volume flux, Roe flux).
#pragma once
#include <cuda_runtime.h>

#include "dg_kernels.cuh"

namespace hdg {

// D(8x8) += A(8x4) * B(4x8);  lane = 4*g + t :  A[g][t],  B[t][g],  D[g][2t], D[g][2t+1]
__device__ __forceinline__ void prefetchL1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetchL2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b)
{
#ifdef HDG_DMMA_VOLATILE
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
#else
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
#endif
                 : "+d"(d[0]), "+d"(d[1])
                 : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------------------
// Operator-table staging: one TMA bulk copy (cp.async.bulk -> SASS UBLKCP) per <= 32 KB chunk, completion on an mbarrier.
// The tables are shared by every element the persistent block will ever process, so this runs once per block.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void stageTables(double* dstSmem, const double* srcGlobal, int nDoubles, unsigned long long* mbar)
{
    const unsigned bar = (unsigned)__cvta_generic_to_shared(mbar);
    const unsigned bytes = (unsigned)nDoubles * 8u;            // multiple of 16: every table is a whole number of 32-lane rows
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        for (unsigned off = 0; off < bytes; off += 32768u) {
            const unsigned n = min(32768u, bytes - off);
            const unsigned dst = (unsigned)__cvta_generic_to_shared(reinterpret_cast<char*>(dstSmem) + off);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                         "l"(reinterpret_cast<const char*>(srcGlobal) + off), "r"(n), "r"(bar)
                         : "memory");
        }
    }
    unsigned done = 0;
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(done) : "r"(bar) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// Point-wise physics
// ---------------------------------------------------------------------------------------------------------

// 1/x and 1/sqrt(x) from the MUFU seed (about 20 mantissa bits) + one cubic step / two Newton steps: full double precision to ~1 ulp for
// normal, finite, positive-magnitude arguments, without the IEEE slow-path subroutine of `1.0/x` / `sqrt` (the
// densities, sound speeds etc. divided by here are O(1) physical quantities; tolerance to the oracle is 1e-12).
__device__ __forceinline__ double fastRcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#ifndef HDG_RCP_NEWTON
    // one cubic step: r (1 + e + e^2), e = 1 - x r: the seed's 2^-23 becomes 2^-69 in three dependent operations (two Newton steps
    // need four); the last fma rounds once
    const double e = fma(-x, r, 1.0);
    const double t = fma(e, e, e);
    return fma(r, t, r);
#else
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
#endif
}
__device__ __forceinline__ double fastRsqrt(double x)
{
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double h = 0.5 * x;
    double e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    e = fma(-h * y, y, 0.5);
    y = fma(y, e, y);
    return y;
}

// Contravariant Euler fluxes at one cubature point:  Gr = rx*Fx + ry*Fy,  Gs = sx*Fx + sy*Fy  with
//   U = rhoU/rho, p = (gamma-1)(E - rho|U|^2/2)                    (dgEulerFoam.C:81-82)
//   rho : U rho ; rhoU : U rhoU + p I ; E : U E + U p              (dgEulerFoam.C:86-90 volume terms)
__device__ __forceinline__ void eulerVolumeFlux(const double q[4], double rx, double ry, double sx, double sy, double gm1,
                                                double Gr[4], double Gs[4])
{
    const double ir = fastRcp(q[0]);
    const double u = q[1] * ir, v = q[2] * ir;
    const double p = gm1 * (q[3] - 0.5 * (q[0] * (u * u + v * v)));
    const double Ur = rx * u + ry * v, Us = sx * u + sy * v;
    const double Ep = q[3] + p;
    Gr[0] = q[0] * Ur;
    Gs[0] = q[0] * Us;
    Gr[1] = q[1] * Ur + rx * p;
    Gs[1] = q[1] * Us + sx * p;
    Gr[2] = q[2] * Ur + ry * p;
    Gs[2] = q[2] * Us + sy * p;
    Gr[3] = Ep * Ur;
    Gs[3] = Ep * Us;
}

// Roe flux F*.n, M = owner side, P = neighbour side, n = owner's outward normal (RoeFlux.C:131-177)
__device__ __forceinline__ void roeFlux(const double qM[4], const double qP[4], double nx, double ny, double gm1, double fl[4])
{
    const double QM2 = nx * qM[1] + ny * qM[2], QP2 = nx * qP[1] + ny * qP[2];
    const double QM3 = nx * qM[2] - ny * qM[1], QP3 = nx * qP[2] - ny * qP[1];
    const double rhoM = qM[0], rhoP = qP[0], EM = qM[3], EP = qP[3];
    const double isM = fastRsqrt(rhoM), isP = fastRsqrt(rhoP);     // 1/sqrt(rho): gives sqrt(rho) and 1/rho
    const double irM = isM * isM, irP = isP * isP;
    const double uM = QM2 * irM, uP = QP2 * irP, vM = QM3 * irM, vP = QP3 * irP;
    const double pM = gm1 * (EM - 0.5 * (QM2 * uM + QM3 * vM));
    const double pP = gm1 * (EP - 0.5 * (QP2 * uP + QP3 * vP));
    const double HM = (EM + pM) * irM, HP = (EP + pP) * irP;
    double fR = (QM2 + QP2) * 0.5;
    double fU = (QM2 * uM + pM + QP2 * uP + pP) * 0.5;
    double fV = (QM3 * uM + QP3 * uP) * 0.5;
    double fE = (uM * (EM + pM) + uP * (EP + pP)) * 0.5;
    const double rMs = rhoM * isM, rPs = rhoP * isP;
    const double rhob = rMs * rPs;
    const double is = fastRcp(rMs + rPs);
    const double u = (rMs * uM + rPs * uP) * is;
    const double v = (rMs * vM + rPs * vP) * is;
    const double H = (rMs * HM + rPs * HP) * is;
    const double c2 = gm1 * (H - 0.5 * (u * u + v * v));
    const double ac2 = fabs(c2);                                   // c = sqrt(fabs(c2) + e), e = 0 (RoeFlux.C:162-163)
    const double ic = fastRsqrt(ac2);
    const double c = ac2 * ic;
    const double ic2 = copysign(ic * ic, c2);
    const double du = uP - uM, dp = pP - pM;
    const double dw1 = (-0.5 * rhob * du * ic + 0.5 * dp * ic2) * fabs(u - c);
    const double dw2 = ((rhoP - rhoM) - dp * ic2) * fabs(u);
    const double dw3 = (rhob * (vP - vM)) * fabs(u);
    const double dw4 = (0.5 * rhob * du * ic + 0.5 * dp * ic2) * fabs(u + c);
    fR -= (dw1 + dw2 + dw4) * 0.5;
    fU -= (dw1 * (u - c) + dw2 * u + dw4 * (u + c)) * 0.5;
    fV -= (dw1 * v + dw2 * v + dw3 + dw4 * v) * 0.5;
    fE -= (dw1 * (H - u * c) + dw2 * (u * u + v * v) * 0.5 + dw3 * v + dw4 * (H + u * c)) * 0.5;
    fl[0] = fR;
    fl[1] = nx * fU - ny * fV;
    fl[2] = ny * fU + nx * fV;
    fl[3] = fE;
}

// Local Lax-Friedrichs (Rusanov) flux F*.n for the Euler system, point-wise: the central flux minus half the largest wave speed
// |u.n| + c of the two states times the jump.  The reference defines NO Lax-Friedrichs flux for the Euler system (its godunovScheme
// knows Roe only, godunovFlux/fluxSchemes/scheme/); this is the counterpart of its scalar LFFlux (simpleFlux/schemes/LFFlux) that
// BASELINE's north_star asks for, offered through HDG_FLUX_LF on the Euler entry points.  Antisymmetric under (M <-> P, n -> -n).
__device__ __forceinline__ void rusanovFlux(const double qM[4], const double qP[4], double nx, double ny, double gm1, double fl[4])
{
    const double irM = fastRcp(qM[0]), irP = fastRcp(qP[0]);
    const double unM = (nx * qM[1] + ny * qM[2]) * irM, unP = (nx * qP[1] + ny * qP[2]) * irP;
    const double pM = gm1 * (qM[3] - 0.5 * (qM[1] * qM[1] + qM[2] * qM[2]) * irM);
    const double pP = gm1 * (qP[3] - 0.5 * (qP[1] * qP[1] + qP[2] * qP[2]) * irP);
    const double c2M = fabs((gm1 + 1.0) * pM * irM), c2P = fabs((gm1 + 1.0) * pP * irP);
    const double cM = c2M * fastRsqrt(c2M), cP = c2P * fastRsqrt(c2P);
    const double lM = fabs(unM) + cM, lP = fabs(unP) + cP;
    const double lam = lM > lP ? lM : lP;
    fl[0] = 0.5 * ((qM[0] * unM + qP[0] * unP) - lam * (qP[0] - qM[0]));
    fl[1] = 0.5 * ((qM[1] * unM + pM * nx + qP[1] * unP + pP * nx) - lam * (qP[1] - qM[1]));
    fl[2] = 0.5 * ((qM[2] * unM + pM * ny + qP[2] * unP + pP * ny) - lam * (qP[2] - qM[2]));
    fl[3] = 0.5 * (((qM[3] + pM) * unM + (qP[3] + pP) * unP) - lam * (qP[3] - qM[3]));
}
// the flux scheme of a stage: 0 Roe, 1 Lax-Friedrichs (HDG_FLUX_* of include/hopedg.h)
__device__ __forceinline__ void eulerFaceFluxPoint(int fluxKind, const double qM[4], const double qP[4], double nx, double ny, double gm1, double fl[4])
{
    if (fluxKind == 1) rusanovFlux(qM, qP, nx, ny, gm1, fl);
    else roeFlux(qM, qP, nx, ny, gm1, fl);
}

}  // namespace hdg

/* ORACLE / CPU BASELINE (test infrastructure, NOT product code).
 *
 * C port of the reference's per-stage loop structure for dgEulerFoam ("reference-faithful mode"):
 * AoS element-contiguous fields, separate Gauss-field interpolation pass (dgGaussField.C:188-269), point-wise
 * gther_U / gther_p (dgEulerFoam.C:81-82), face-flux arrays from a Roe pass (RoeFlux.C:46-191), then THREE separate
 * equation passes per stage that stream the stored per-element cellD1dx (Ng x Np x 2) and dense mass matrix
 * (defaultConvectionScheme.C:48-129, defaultGrad.C:87-166, dgLduMatrix.C:316-321, dgMatrix.C:383-406) and a block
 * solve with the pre-factored element mass matrices (dgMesh.C:129-172: PETSc ILU(0) == exact LU on dense blocks).
 * All geometry/operators are supplied by the numpy oracle (oracle/dg_oracle.py); this file only times the loops.
 * Threads: static element/face ranges (OpenMP) stand in for the reference's MPI ranks.
 * Compiled with gcc -O3 (no -ffast-math), the reference's c++Opt (wmake/rules/linux64Gcc/c++Opt:1-2).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int K, F, Np, Ng, Nfp, Nfg;
    const double *Vg, *If;            /* Ng x Np, Nfg x Nfp */
    const double *D1x, *D1y;          /* K x Ng x Np */
    const double *WJ;                 /* K x Ng */
    const double *M;                  /* K x Np x Np */
    const double *Lchol;              /* K x Np x Np lower Cholesky factor of M (pre-factored once) */
    const double *fnx, *fny, *fWJ;    /* F x Nfg */
    const int32_t *mapO, *mapN;       /* F x Nfp global dof ids (owner / rotated neighbour) */
    const int32_t *cellFace;          /* K x 3 dgFace ids */
    const int32_t *faceOwner, *faceLocO, *faceLocN; /* F */
    const int32_t *f2c;               /* 3 x 2 x Nfp */
    const int32_t *faceRot;           /* F */
} RefCase;

static void roe(double nx, double ny, double rhoM, double ruM, double rvM, double EM, double rhoP, double ruP, double rvP, double EP,
                double g, double *fR, double *fUx, double *fUy, double *fE)
{
    double QM2 = nx * ruM + ny * rvM, QP2 = nx * ruP + ny * rvP, QM3 = nx * rvM - ny * ruM, QP3 = nx * rvP - ny * ruP;
    double uM = QM2 / rhoM, uP = QP2 / rhoP, vM = QM3 / rhoM, vP = QP3 / rhoP;
    double pM = (g - 1) * (EM - 0.5 * (QM2 * uM + QM3 * vM)), pP = (g - 1) * (EP - 0.5 * (QP2 * uP + QP3 * vP));
    double HM = (EM + pM) / rhoM, HP = (EP + pP) / rhoP;
    double r = (QM2 + QP2) / 2, fu = (QM2 * uM + pM + QP2 * uP + pP) / 2, fv = (QM3 * uM + QP3 * uP) / 2;
    double fe = (uM * (EM + pM) + uP * (EP + pP)) / 2;
    double rMs = sqrt(rhoM), rPs = sqrt(rhoP), rhob = rMs * rPs;
    double u = (rMs * uM + rPs * uP) / (rMs + rPs), v = (rMs * vM + rPs * vP) / (rMs + rPs), H = (rMs * HM + rPs * HP) / (rMs + rPs);
    double c2 = (g - 1) * (H - 0.5 * (u * u + v * v)), c = sqrt(fabs(c2) + 0.0);
    double dw1 = (-0.5 * rhob * (uP - uM) / c + 0.5 * (pP - pM) / c2) * fabs(u - c);
    double dw2 = ((rhoP - rhoM) - (pP - pM) / c2) * fabs(u);
    double dw3 = (rhob * (vP - vM)) * fabs(u);
    double dw4 = (0.5 * rhob * (uP - uM) / c + 0.5 * (pP - pM) / c2) * fabs(u + c);
    r -= (dw1 + dw2 + dw4) / 2;
    fu -= (dw1 * (u - c) + dw2 * u + dw4 * (u + c)) / 2;
    fv -= (dw1 * v + dw2 * v + dw3 + dw4 * v) / 2;
    fe -= (dw1 * (H - u * c) + dw2 * (u * u + v * v) / 2 + dw3 * v + dw4 * (H + u * c)) / 2;
    *fR = r; *fUx = nx * fu - ny * fv; *fUy = ny * fu + nx * fv; *fE = fe;
}

/* one equation pass: q_new = solve(M, dt*(b + M*q_old/dt)), b = volume - surface (+ extra volume source field) */
static void equation_pass(const RefCase *c, const double *q_old, int stride, const double *gq /*K*Ng cell values of q*/,
                          const double *Ux, const double *Uy, const double *extraX, const double *extraY /* K*Ng or NULL */,
                          const double *flux /*F*Nfg*/, double dt, double *q_new)
{
    const int Np = c->Np, Ng = c->Ng, Nfp = c->Nfp, Nfg = c->Nfg;
#pragma omp parallel
    {
        double *b = (double *)malloc(sizeof(double) * (Np * 3 + Ng * 2 + Nfg));
        double *src = b + Np, *y = src + Np, *tx = y + Np, *ty = tx + Ng, *tf = ty + Ng;
#pragma omp for schedule(static)
        for (int k = 0; k < c->K; ++k) {
            const double *D1x = c->D1x + (size_t)k * Ng * Np, *D1y = c->D1y + (size_t)k * Ng * Np, *WJ = c->WJ + (size_t)k * Ng;
            for (int j = 0; j < Np; ++j) b[j] = 0.0;
            for (int g = 0; g < Ng; ++g) {
                const size_t i = (size_t)k * Ng + g;
                tx[g] = Ux[i] * gq[i] * WJ[g];
                ty[g] = Uy[i] * gq[i] * WJ[g];
                if (extraX) { tx[g] += extraX[i] * WJ[g]; ty[g] += extraY[i] * WJ[g]; }
            }
            for (int g = 0; g < Ng; ++g)
                for (int j = 0; j < Np; ++j) b[j] += D1x[g * Np + j] * tx[g] + D1y[g * Np + j] * ty[g];
            for (int lf = 0; lf < 3; ++lf) {
                const int f = c->cellFace[3 * k + lf];
                const int owner = (c->faceOwner[f] == k && c->faceLocO[f] == lf);
                const int32_t *map = c->f2c + ((owner ? lf * 2 : lf * 2 + (c->faceRot[f] == 1)) * Nfp);
                for (int i = 0; i < Nfg; ++i) tf[i] = c->fWJ[(size_t)f * Nfg + i] * flux[(size_t)f * Nfg + i];
                for (int i = 0; i < Nfg; ++i)
                    for (int j = 0; j < Nfp; ++j) {
                        if (owner) b[map[j]] -= tf[i] * c->If[i * Nfp + j];
                        else       b[map[j]] += tf[i] * c->If[i * Nfp + j];
                    }
            }
            /* source = q_old/dt ; b += M*source ; b *= dt */
            const double *M = c->M + (size_t)k * Np * Np;
            for (int j = 0; j < Np; ++j) src[j] = q_old[((size_t)k * Np + j) * stride] / dt;
            for (int i = 0; i < Np; ++i) {
                double s = 0.0;
                for (int j = 0; j < Np; ++j) s += M[i * Np + j] * src[j];
                b[i] = (b[i] + s) * dt;
            }
            /* block solve with the pre-factored Cholesky factor L L^T */
            const double *L = c->Lchol + (size_t)k * Np * Np;
            for (int i = 0; i < Np; ++i) {
                double s = b[i];
                for (int j = 0; j < i; ++j) s -= L[i * Np + j] * y[j];
                y[i] = s / L[i * Np + i];
            }
            for (int i = Np - 1; i >= 0; --i) {
                double s = y[i];
                for (int j = i + 1; j < Np; ++j) s -= L[j * Np + i] * b[j];
                b[i] = s / L[i * Np + i];
            }
            for (int j = 0; j < Np; ++j) q_new[((size_t)k * Np + j) * stride] = b[j];
        }
        free(b);
    }
}

static void gauss_field(const RefCase *c, const double *q, int stride, double *cell, double *own, double *nbr)
{
    const int Np = c->Np, Ng = c->Ng, Nfp = c->Nfp, Nfg = c->Nfg;
#pragma omp parallel for schedule(static)
    for (int k = 0; k < c->K; ++k)
        for (int g = 0; g < Ng; ++g) {
            double s = 0.0;
            for (int j = 0; j < Np; ++j) s += c->Vg[g * Np + j] * q[((size_t)k * Np + j) * stride];
            cell[(size_t)k * Ng + g] = s;
        }
#pragma omp parallel for schedule(static)
    for (int f = 0; f < c->F; ++f)
        for (int i = 0; i < Nfg; ++i) {
            double so = 0.0, sn = 0.0;
            for (int j = 0; j < Nfp; ++j) {
                so += c->If[i * Nfp + j] * q[(size_t)c->mapO[(size_t)f * Nfp + j] * stride];
                sn += c->If[i * Nfp + j] * q[(size_t)c->mapN[(size_t)f * Nfp + j] * stride];
            }
            own[(size_t)f * Nfg + i] = so;
            nbr[(size_t)f * Nfg + i] = sn;
        }
}

/* one forward-Euler sub-step of the reference solver on AoS fields rho[K*Np], rhoU[K*Np*2], E[K*Np]  (periodic mesh) */
static void euler_stage(const RefCase *c, const double *rho, const double *rhoU, const double *E, double gamma, double dt,
                        double *rho1, double *rhoU1, double *E1, double *work)
{
    const size_t nC = (size_t)c->K * c->Ng, nF = (size_t)c->F * c->Nfg;
    double *rc = work, *uxc = rc + nC, *uyc = uxc + nC, *ec = uyc + nC, *Ux = ec + nC, *Uy = Ux + nC, *p = Uy + nC, *pUx = p + nC, *pUy = pUx + nC;
    double *ro = pUy + nC, *rn = ro + nF, *uxo = rn + nF, *uxn = uxo + nF, *uyo = uxn + nF, *uyn = uyo + nF, *eo = uyn + nF, *en = eo + nF;
    double *fR = en + nF, *fUx = fR + nF, *fUy = fUx + nF, *fE = fUy + nF;
    gauss_field(c, rho, 1, rc, ro, rn);
    gauss_field(c, rhoU, 2, uxc, uxo, uxn);
    gauss_field(c, rhoU + 1, 2, uyc, uyo, uyn);
    gauss_field(c, E, 1, ec, eo, en);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nC; ++i) {
        Ux[i] = uxc[i] / rc[i];
        Uy[i] = uyc[i] / rc[i];
        p[i] = (gamma - 1.0) * (ec[i] - 0.5 * (rc[i] * (Ux[i] * Ux[i] + Uy[i] * Uy[i])));
    }
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nF; ++i)
        roe(c->fnx[i], c->fny[i], ro[i], uxo[i], uyo[i], eo[i], rn[i], uxn[i], uyn[i], en[i], gamma, fR + i, fUx + i, fUy + i, fE + i);
    /* rho: ddt + div(U,rho,fluxRho) */
    equation_pass(c, rho, 1, rc, Ux, Uy, NULL, NULL, fR, dt, rho1);
    /* rhoU: + div(U,rhoU,fluxRhoU) + grad(p): x-momentum extra = (p,0), y-momentum extra = (0,p) */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nC; ++i) { pUx[i] = p[i]; pUy[i] = 0.0; }
    equation_pass(c, rhoU, 2, uxc, Ux, Uy, pUx, pUy, fUx, dt, rhoU1);
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nC; ++i) { pUx[i] = 0.0; pUy[i] = p[i]; }
    equation_pass(c, rhoU + 1, 2, uyc, Ux, Uy, pUx, pUy, fUy, dt, rhoU1 + 1);
    /* E: + div(U,E,fluxEner) + div(U,p) */
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < nC; ++i) { pUx[i] = Ux[i] * p[i]; pUy[i] = Uy[i] * p[i]; }
    equation_pass(c, E, 1, ec, Ux, Uy, pUx, pUy, fE, dt, E1);
}

size_t refcpu_work_doubles(int K, int F, int Ng, int Nfg) { return (size_t)K * Ng * 9 + (size_t)F * Nfg * 12; }

/* nsteps SSP-RK2 steps (dgEulerFoam.C:67-117) in place on rho/rhoU/E; returns 0 */
int refcpu_euler_steps(const RefCase *c, double *rho, double *rhoU, double *E, double gamma, double dt, int nsteps, int threads,
                       double *work, double *tmp /* 2 x (K*Np*4) */)
{
#ifdef _OPENMP
    if (threads > 0) omp_set_num_threads(threads);
#endif
    const size_t n = (size_t)c->K * c->Np;
    double *r1 = tmp, *u1 = r1 + n, *e1 = u1 + 2 * n, *r2 = e1 + n, *u2 = r2 + n, *e2 = u2 + 2 * n;
    for (int s = 0; s < nsteps; ++s) {
        euler_stage(c, rho, rhoU, E, gamma, dt, r1, u1, e1, work);
        euler_stage(c, r1, u1, e1, gamma, dt, r2, u2, e2, work);
#pragma omp parallel for schedule(static)
        for (size_t i = 0; i < n; ++i) {
            rho[i] = 0.5 * rho[i] + 0.5 * r2[i];
            rhoU[2 * i] = 0.5 * rhoU[2 * i] + 0.5 * u2[2 * i];
            rhoU[2 * i + 1] = 0.5 * rhoU[2 * i + 1] + 0.5 * u2[2 * i + 1];
            E[i] = 0.5 * E[i] + 0.5 * e2[i];
        }
    }
    return 0;
}

int refcpu_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

// See ref_element.hpp for the reference citations.
#include "ref_element.hpp"

#include <algorithm>
#include <cmath>
#include <stdexcept>
#include <string>

#include "cubature_tri.inc"

namespace hdg {

// ------------------------------------------------------------------------------------------------
// dense helpers
// ------------------------------------------------------------------------------------------------
Mat matmul(const Mat& A, int m, int k, const Mat& B, int n)
{
    Mat C((size_t)m * n, 0.0);
    for (int i = 0; i < m; ++i)
        for (int l = 0; l < k; ++l) {
            const double a = A[(size_t)i * k + l];
            for (int j = 0; j < n; ++j) C[(size_t)i * n + j] += a * B[(size_t)l * n + j];
        }
    return C;
}

Mat transpose(const Mat& A, int m, int n)
{
    Mat T((size_t)m * n);
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < n; ++j) T[(size_t)j * m + i] = A[(size_t)i * n + j];
    return T;
}

// Gauss-Jordan with partial pivoting (the reference uses PETSc dense LU + MatMatSolve, Legendre.C:540-618;
// the inverse is unique, agreement is O(eps*cond)).
Mat inverse(const Mat& Ain, int n)
{
    Mat A(Ain), I((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) I[(size_t)i * n + i] = 1.0;
    for (int c = 0; c < n; ++c) {
        int piv = c;
        for (int r = c + 1; r < n; ++r)
            if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
        if (A[(size_t)piv * n + c] == 0.0) throw std::runtime_error("hdg::inverse: singular matrix");
        if (piv != c)
            for (int j = 0; j < n; ++j) {
                std::swap(A[(size_t)piv * n + j], A[(size_t)c * n + j]);
                std::swap(I[(size_t)piv * n + j], I[(size_t)c * n + j]);
            }
        const double d = 1.0 / A[(size_t)c * n + c];
        for (int j = 0; j < n; ++j) { A[(size_t)c * n + j] *= d; I[(size_t)c * n + j] *= d; }
        for (int r = 0; r < n; ++r) {
            if (r == c) continue;
            const double f = A[(size_t)r * n + c];
            if (f == 0.0) continue;
            for (int j = 0; j < n; ++j) {
                A[(size_t)r * n + j] -= f * A[(size_t)c * n + j];
                I[(size_t)r * n + j] -= f * I[(size_t)c * n + j];
            }
        }
    }
    return I;
}

// cyclic Jacobi eigen-solver for a small symmetric matrix; eigenvectors in the columns of Q
static void symEig(Mat A, int n, std::vector<double>& lam, Mat& Q)
{
    Q.assign((size_t)n * n, 0.0);
    for (int i = 0; i < n; ++i) Q[(size_t)i * n + i] = 1.0;
    for (int sweep = 0; sweep < 100; ++sweep) {
        double off = 0.0;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) off += A[(size_t)p * n + q] * A[(size_t)p * n + q];
        if (off < 1e-60) break;
        for (int p = 0; p < n; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[(size_t)p * n + q];
                if (apq == 0.0) continue;
                const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
                const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
                for (int k = 0; k < n; ++k) {
                    const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
                    A[(size_t)k * n + p] = c * akp - s * akq;
                    A[(size_t)k * n + q] = s * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {
                    const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
                    A[(size_t)p * n + k] = c * apk - s * aqk;
                    A[(size_t)q * n + k] = s * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double qkp = Q[(size_t)k * n + p], qkq = Q[(size_t)k * n + q];
                    Q[(size_t)k * n + p] = c * qkp - s * qkq;
                    Q[(size_t)k * n + q] = s * qkp + c * qkq;
                }
            }
    }
    lam.resize(n);
    for (int i = 0; i < n; ++i) lam[i] = A[(size_t)i * n + i];
}

// ------------------------------------------------------------------------------------------------
// Jacobi polynomials and Gauss rules (Legendre.C)
// ------------------------------------------------------------------------------------------------
static double factorialGamma(double xd)   // Legendre::gamma(int x) = (x-1)!   (Legendre.C:243-247)
{
    const int x = (int)xd;
    double g = 1.0;
    for (int i = 2; i < x; ++i) g *= i;
    return g;
}

std::vector<double> jacobiP(const std::vector<double>& x, double a, double b, int N)
{
    const size_t n = x.size();
    const double gamma0 = std::pow(2.0, a + b + 1) / (a + b + 1) * factorialGamma(a + 1) * factorialGamma(b + 1) / factorialGamma(a + b + 1);
    std::vector<double> p0(n, 1.0 / std::sqrt(gamma0));
    if (N == 0) return p0;
    const double gamma1 = (a + 1) * (b + 1) / (a + b + 3) * gamma0;
    std::vector<double> p1(n);
    for (size_t i = 0; i < n; ++i) p1[i] = ((a + b + 2) / 2 * x[i] + (a - b) / 2) / std::sqrt(gamma1);
    if (N == 1) return p1;
    double aold = 2.0 / (2 + a + b) * std::sqrt((a + 1) * (b + 1) / (a + b + 3));
    std::vector<double> p2(n);
    for (int i = 1; i < N; ++i) {
        const double h1 = 2.0 * i + a + b;
        const double anew = 2.0 / (h1 + 2) * std::sqrt((i + 1) * (i + 1 + a + b) * (i + 1 + a) * (i + 1 + b) / (h1 + 1) / (h1 + 3));
        const double bnew = -(a * a - b * b) / h1 / (h1 + 2);
        for (size_t k = 0; k < n; ++k) p2[k] = 1.0 / anew * (-aold * p0[k] + (x[k] - bnew) * p1[k]);
        p0.swap(p1);
        p1.swap(p2);
        aold = anew;
    }
    return p1;
}

static std::vector<double> gradJacobiP(const std::vector<double>& x, double a, double b, int N)
{
    if (N == 0) return std::vector<double>(x.size(), 0.0);
    std::vector<double> d = jacobiP(x, a + 1, b + 1, N - 1);
    const double f = std::sqrt(N * (N + a + b + 1));
    for (double& v : d) v *= f;
    return d;
}

void jacobiGQ(double a, double b, int N, std::vector<double>& x, std::vector<double>& w)
{
    x.assign(N + 1, 0.0);
    w.assign(N + 1, 0.0);
    if (N == 0) { x[0] = -(a - b) / (a + b + 2); w[0] = 2; return; }
    const int n = N + 1;
    Mat A((size_t)n * n, 0.0);   // zero diagonal: the reference only ever uses alpha == beta (Legendre.C:77-87)
    for (int i = 0; i < N; ++i) {
        const int j = i + 1;
        const double v = 2.0 / (2 * i + a + b + 2) * std::sqrt(j * (j + a + b) * (j + a) * (j + b) / (2 * i + a + b + 1) / (2 * i + a + b + 3));
        A[(size_t)i * n + j] = A[(size_t)j * n + i] = v;
    }
    std::vector<double> lam;
    Mat Q;
    symEig(A, n, lam, Q);
    std::vector<int> ord(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int p, int q) { return lam[p] < lam[q]; });
    const double scale = std::pow(2.0, a + b + 1) / (a + b + 1) * factorialGamma(a + 1) * factorialGamma(b + 1) / factorialGamma(a + b + 1);
    for (int i = 0; i < n; ++i) {
        x[i] = lam[ord[i]];
        const double v0 = Q[(size_t)0 * n + ord[i]];
        w[i] = v0 * v0 * scale;
    }
    for (int i = 0; i <= (N - 1) / 2; ++i) {              // weight symmetrisation (Legendre.C:154-157)
        const double t = 0.5 * (w[i] + w[N - i]);
        w[i] = w[N - i] = t;
    }
}

// n-point Gauss-Jacobi rule for (1-x)^a (1+x)^b by Golub-Welsch with the full recurrence (diagonal included).  Not a reference function
// (its JacobiGQ is only valid for a == b): used for the collapsed cubature of the orders beyond the reference's table.
void gaussJacobi(double a, double b, int n, std::vector<double>& x, std::vector<double>& w)
{
    Mat A((size_t)n * n, 0.0);
    for (int k = 0; k < n; ++k) {
        const double d = (2 * k + a + b) * (2 * k + a + b + 2);
        A[(size_t)k * n + k] = d != 0 ? (b * b - a * a) / d : (b - a) / (a + b + 2);
        if (k + 1 < n) {
            const double h = 2 * k + a + b;
            const double v = 2.0 / (h + 2) * std::sqrt((k + 1) * (k + 1 + a + b) * (k + 1 + a) * (k + 1 + b) / (h + 1) / (h + 3));
            A[(size_t)k * n + k + 1] = A[(size_t)(k + 1) * n + k] = v;
        }
    }
    std::vector<double> lam;
    Mat Q;
    symEig(A, n, lam, Q);
    std::vector<int> ord(n);
    for (int i = 0; i < n; ++i) ord[i] = i;
    std::sort(ord.begin(), ord.end(), [&](int p, int q) { return lam[p] < lam[q]; });
    const double mu0 = std::pow(2.0, a + b + 1) * std::tgamma(a + 1) * std::tgamma(b + 1) / std::tgamma(a + b + 2);
    x.resize(n);
    w.resize(n);
    for (int i = 0; i < n; ++i) {
        x[i] = lam[ord[i]];
        const double v0 = Q[(size_t)0 * n + ord[i]];
        w[i] = v0 * v0 * mu0;
    }
}

// cubature of the reference triangle exact to degree >= order: conical product of m-point Gauss-Legendre (a) and Gauss-Jacobi(1,0) (b),
// m = order/2 + 1; r = (1+a)(1-b)/2 - 1, s = b, weight w_a w_b / 2; b outer, a inner.  For volIntOrder > 28 (N = 9, 10) only.
static void collapsedCubature(int order, std::vector<double>& r, std::vector<double>& s, std::vector<double>& w)
{
    const int m = order / 2 + 1;
    std::vector<double> xa, wa, xb, wb;
    gaussJacobi(0, 0, m, xa, wa);
    gaussJacobi(1, 0, m, xb, wb);
    r.clear(); s.clear(); w.clear();
    for (int j = 0; j < m; ++j)
        for (int i = 0; i < m; ++i) {
            r.push_back((1 + xa[i]) * (1 - xb[j]) / 2 - 1);
            s.push_back(xb[j]);
            w.push_back(wa[i] * wb[j] / 2);
        }
}

std::vector<double> jacobiGL(double a, double b, int N)
{
    std::vector<double> x(N + 1, 0.0);
    x[0] = -1.0;
    x[N] = 1.0;
    if (N < 2) return x;
    std::vector<double> xi, wi;
    jacobiGQ(a + 1, b + 1, N - 2, xi, wi);
    for (int i = 1; i < N; ++i) x[i] = xi[i - 1];
    return x;
}

static Mat vandermonde1D(int N, const std::vector<double>& r)
{
    const int n = (int)r.size();
    Mat V((size_t)n * (N + 1));
    for (int j = 0; j <= N; ++j) {
        const std::vector<double> p = jacobiP(r, 0, 0, j);
        for (int i = 0; i < n; ++i) V[(size_t)i * (N + 1) + j] = p[i];
    }
    return V;
}

static void rsToAb(const std::vector<double>& r, const std::vector<double>& s, std::vector<double>& a, std::vector<double>& b)
{
    a.resize(r.size());
    b = s;
    for (size_t i = 0; i < r.size(); ++i) a[i] = (s[i] != 1.0) ? 2 * (1 + r[i]) / (1 - s[i]) - 1 : -1.0;
}

static Mat vandermonde2D(int N, const std::vector<double>& r, const std::vector<double>& s)
{
    const int n = (int)r.size(), Np = (N + 1) * (N + 2) / 2;
    std::vector<double> a, b;
    rsToAb(r, s, a, b);
    Mat V((size_t)n * Np);
    int sk = 0;
    for (int i = 0; i <= N; ++i)
        for (int j = 0; j <= N - i; ++j, ++sk) {
            const std::vector<double> h1 = jacobiP(a, 0, 0, i), h2 = jacobiP(b, 2 * i + 1, 0, j);
            for (int k = 0; k < n; ++k) V[(size_t)k * Np + sk] = std::sqrt(2.0) * h1[k] * h2[k] * std::pow(1 - b[k], i);
        }
    return V;
}

static void gradVandermonde2D(int N, const std::vector<double>& r, const std::vector<double>& s, Mat& Vr, Mat& Vs)
{
    const int n = (int)r.size(), Np = (N + 1) * (N + 2) / 2;
    std::vector<double> a, b;
    rsToAb(r, s, a, b);
    Vr.assign((size_t)n * Np, 0.0);
    Vs.assign((size_t)n * Np, 0.0);
    int sk = 0;
    for (int i = 0; i <= N; ++i)
        for (int j = 0; j <= N - i; ++j, ++sk) {
            const std::vector<double> fa = jacobiP(a, 0, 0, i), gb = jacobiP(b, 2 * i + 1, 0, j);
            const std::vector<double> dfa = gradJacobiP(a, 0, 0, i), dgb = gradJacobiP(b, 2 * i + 1, 0, j);
            const double c = std::pow(2.0, i + 0.5);
            for (int k = 0; k < n; ++k) {
                const double hb = 0.5 * (1 - b[k]);
                double vr, vs;
                if (i > 0) {
                    vr = dfa[k] * gb[k] * c * std::pow(hb, i - 1);
                    const double tmp = dgb[k] * std::pow(hb, i) - 0.5 * i * gb[k] * std::pow(hb, i - 1);
                    vs = (dfa[k] * gb[k] * 0.5 * (1 + a[k]) * std::pow(hb, i - 1) + fa[k] * tmp) * c;
                } else {
                    vr = dfa[k] * gb[k] * c;
                    vs = (dfa[k] * gb[k] * 0.5 * (1 + a[k]) + fa[k] * dgb[k]) * c;
                }
                Vr[(size_t)k * Np + sk] = vr;
                Vs[(size_t)k * Np + sk] = vs;
            }
        }
}

// ------------------------------------------------------------------------------------------------
// Warp & Blend nodes (triangleBaseFunction.C:122-233)
// ------------------------------------------------------------------------------------------------
static std::vector<double> warpFactor(int N, const std::vector<double>& rout)
{
    const std::vector<double> lgl = jacobiGL(0, 0, N);
    std::vector<double> req(N + 1);
    for (int i = 0; i <= N; ++i) req[i] = 2.0 / N * i - 1;
    const Mat Veq = vandermonde1D(N, req);
    const Mat invVeqT = inverse(transpose(Veq, N + 1, N + 1), N + 1);
    const int n = (int)rout.size();
    Mat P((size_t)(N + 1) * n);
    for (int i = 0; i <= N; ++i) {
        const std::vector<double> p = jacobiP(rout, 0, 0, i);
        for (int k = 0; k < n; ++k) P[(size_t)i * n + k] = p[k];
    }
    const Mat L = matmul(invVeqT, N + 1, N + 1, P, n);
    std::vector<double> warp(n, 0.0);
    for (int k = 0; k < n; ++k)
        for (int i = 0; i <= N; ++i) warp[k] += L[(size_t)i * n + k] * (lgl[i] - req[i]);
    for (int k = 0; k < n; ++k) {
        const double zerof = (std::fabs(rout[k]) < 1.0 - 1.0e-10) ? 1.0 : 0.0;
        const double sf = 1.0 - (zerof * rout[k]) * (zerof * rout[k]);
        warp[k] = warp[k] / sf + warp[k] * (zerof - 1.0);
    }
    return warp;
}

static void warpBlendNodes(int N, std::vector<double>& r, std::vector<double>& s)
{
    static const double alpopt[15] = {0.0000, 0.0000, 1.4152, 0.1001, 0.2751, 0.9800, 1.0999, 1.2832,
                                      1.3648, 1.4773, 1.4959, 1.5743, 1.5770, 1.6223, 1.6258};
    const int Np = (N + 1) * (N + 2) / 2;
    const double alpha = (N < 16) ? alpopt[N - 1] : 5.0 / 3.0;
    std::vector<double> L1(Np), L2(Np), L3(Np), x(Np), y(Np);
    int sk = 0;
    for (int n = 1; n <= N + 1; ++n)
        for (int m = 1; m <= N + 2 - n; ++m, ++sk) {
            L1[sk] = (double)(n - 1) / N;
            L3[sk] = (double)(m - 1) / N;
        }
    std::vector<double> d1(Np), d2(Np), d3(Np);
    for (int i = 0; i < Np; ++i) {
        L2[i] = 1.0 - L1[i] - L3[i];
        x[i] = L3[i] - L2[i];
        y[i] = (2 * L1[i] - L2[i] - L3[i]) / std::sqrt(3.0);
        d1[i] = L3[i] - L2[i];
        d2[i] = L1[i] - L3[i];
        d3[i] = L2[i] - L1[i];
    }
    std::vector<double> w1 = warpFactor(N, d1), w2 = warpFactor(N, d2), w3 = warpFactor(N, d3);
    const double pi = 3.14159265358979323846;
    const double c1 = std::cos(2 * pi / 3), c2 = std::cos(4 * pi / 3), s1 = std::sin(2 * pi / 3), s2 = std::sin(4 * pi / 3);
    r.resize(Np);
    s.resize(Np);
    for (int i = 0; i < Np; ++i) {
        const double a1 = 4 * L2[i] * L3[i] * w1[i] * (1 + (alpha * L1[i]) * (alpha * L1[i]));
        const double a2 = 4 * L1[i] * L3[i] * w2[i] * (1 + (alpha * L2[i]) * (alpha * L2[i]));
        const double a3 = 4 * L1[i] * L2[i] * w3[i] * (1 + (alpha * L3[i]) * (alpha * L3[i]));
        const double X = x[i] + 1 * a1 + c1 * a2 + c2 * a3;
        const double Y = y[i] + 0 * a1 + s1 * a2 + s2 * a3;
        const double l1 = (std::sqrt(3.0) * Y + 1.0) / 3.0;
        const double l2 = (-3.0 * X - std::sqrt(3.0) * Y + 2.0) / 6.0;
        const double l3 = (3.0 * X - std::sqrt(3.0) * Y + 2.0) / 6.0;
        r[i] = l3 - l2 - l1;
        s[i] = l1 - l2 - l3;
    }
}

// ------------------------------------------------------------------------------------------------
RefElement buildRefElement(int N)
{
    if (N < 1) throw std::runtime_error("baseOrder must be >= 1");
    const int volOrder = 3 * (N + 1), faceOrder = 2 * (N + 1);   // gaussIntegration.C:66-68
    if (volOrder > 33)      // the reference stops at 28 (N = 8, gaussTriangleIntegration.C:59-64); N = 9, 10 use the collapsed rule below
        throw std::runtime_error("volIntOrder_ = " + std::to_string(volOrder) + " is not implemented");
    RefElement e;
    e.N = N;
    e.Np = (N + 1) * (N + 2) / 2;
    e.Nfp = N + 1;
    const int Np = e.Np, Nfp = e.Nfp;
    warpBlendNodes(N, e.r, e.s);
    e.V = vandermonde2D(N, e.r, e.s);
    e.invV = inverse(e.V, Np);
    Mat Vr, Vs;
    gradVandermonde2D(N, e.r, e.s, Vr, Vs);
    e.Dr = matmul(Vr, Np, Np, e.invV, Np);
    e.Ds = matmul(Vs, Np, Np, e.invV, Np);

    // faceToCellIndex_ (triangleBaseFunction.C:75-94)
    e.f2c.assign(3 * 2 * Nfp, 0);
    auto F = [&](int f, int rot, int i) -> int& { return e.f2c[(f * 2 + rot) * Nfp + i]; };
    for (int i = 0; i < Nfp; ++i) F(0, 0, i) = i;
    F(1, 0, 0) = N;
    F(2, 0, 0) = Np - 1;
    for (int i = 1; i < Nfp; ++i) {
        F(1, 0, i) = F(1, 0, i - 1) + (Nfp - i);
        F(2, 0, i) = F(2, 0, i - 1) - i - 1;
    }
    for (int f = 0; f < 3; ++f)
        for (int i = 0; i < Nfp; ++i) F(f, 1, i) = F(f, 0, N - i);

    // cell cubature: dataTable(volOrder) (gaussTriangleIntegrationDataTable.C)
    if (volOrder <= 28) {
        const int o0 = kCubOffset[volOrder - 1], o1 = kCubOffset[volOrder];
        e.gr.assign(kCubR + o0, kCubR + o1);
        e.gs.assign(kCubS + o0, kCubS + o1);
        e.gw.assign(kCubW + o0, kCubW + o1);
    } else      // beyond the reference's table (BASELINE configs[3]: N = 9, 10): own collapsed Gauss-Jacobi rule, parity unpinned
        collapsedCubature(volOrder, e.gr, e.gs, e.gw);
    e.Ng = (int)e.gr.size();
    const int Ng = e.Ng;
    e.Vg = matmul(vandermonde2D(N, e.gr, e.gs), Ng, Np, e.invV, Np);
    Mat Vgr, Vgs;
    gradVandermonde2D(N, e.gr, e.gs, Vgr, Vgs);
    e.Dgr = matmul(Vgr, Ng, Np, e.invV, Np);
    e.Dgs = matmul(Vgs, Ng, Np, e.invV, Np);

    // face Gauss rule and trace interpolation (gaussTriangleIntegration.C:80-95, lineBaseFunction.C:52-63)
    jacobiGQ(0, 0, faceOrder / 2, e.fx, e.fw);
    e.Nfg = (int)e.fx.size();
    const int Nfg = e.Nfg;
    const std::vector<double> lgl = jacobiGL(0, 0, N);
    const Mat invV1 = inverse(vandermonde1D(N, lgl), Nfp);
    e.If = matmul(vandermonde1D(N, e.fx), Nfg, Nfp, invV1, Nfp);

    // quadrature mass matrix of the reference element (physicalCellElement.C:101-110 with J = 1)
    e.Mref.assign((size_t)Np * Np, 0.0);
    for (int g = 0; g < Ng; ++g)
        for (int i = 0; i < Np; ++i)
            for (int j = 0; j < Np; ++j) e.Mref[(size_t)i * Np + j] += e.Vg[(size_t)g * Np + i] * e.gw[g] * e.Vg[(size_t)g * Np + j];
    const Mat Minv = inverse(e.Mref, Np);

    // Pr = Mref^-1 Dgr^T diag(w), Ps likewise (Np x Ng)
    Mat DrTw((size_t)Np * Ng), DsTw((size_t)Np * Ng);
    for (int j = 0; j < Np; ++j)
        for (int g = 0; g < Ng; ++g) {
            DrTw[(size_t)j * Ng + g] = e.Dgr[(size_t)g * Np + j] * e.gw[g];
            DsTw[(size_t)j * Ng + g] = e.Dgs[(size_t)g * Np + j] * e.gw[g];
        }
    e.Pr = matmul(Minv, Np, Np, DrTw, Ng);
    e.Ps = matmul(Minv, Np, Np, DsTw, Ng);

    // LIFT_f = Mref^-1 E_f If^T diag(fw)  (Np x Nfg per face)
    e.LIFT.assign((size_t)Np * 3 * Nfg, 0.0);
    for (int f = 0; f < 3; ++f) {
        Mat B((size_t)Np * Nfg, 0.0);
        for (int i = 0; i < Nfp; ++i) {
            const int node = F(f, 0, i);
            for (int g = 0; g < Nfg; ++g) B[(size_t)node * Nfg + g] = e.If[(size_t)g * Nfp + i] * e.fw[g];
        }
        const Mat L = matmul(Minv, Np, Np, B, Nfg);
        for (int j = 0; j < Np; ++j)
            for (int g = 0; g < Nfg; ++g) e.LIFT[(size_t)j * 3 * Nfg + f * Nfg + g] = L[(size_t)j * Nfg + g];
    }

    // curved boundary faces: the displacement d_i of the Nfp nodes of face f (in the face's own traversal order) moves every node p of
    // the cell by blend_f(p) * sum_i [V1D(vr_p) invV1D]_{p,i} d_i, vr = r for face 0, s for faces 1 and 2; for face 2 the displacement is
    // taken in reversed order; blend = (1+r)/(1-s) for face 1, -(r+s)/(1-vr) for faces 0 and 2, 1 where 1 - vr < 1e-7
    // (triangleBaseFunction.C:421-466, the Gordon-Hall blending of Hesthaven & Warburton's MakeCylinder2D)
    e.faceShift.assign((size_t)3 * Np * Nfp, 0.0);
    for (int f = 0; f < 3; ++f) {
        const std::vector<double>& vr = f == 0 ? e.r : e.s;
        const Mat W = matmul(vandermonde1D(N, vr), Np, Nfp, invV1, Nfp);      // Np x Nfp
        for (int p = 0; p < Np; ++p) {
            // the reference skips the blend FACTOR there (:451-452) and still adds the unblended value: the far end of the face's parameter
            const double blend = std::fabs(1.0 - vr[p]) < 1e-7 ? 1.0 : (f == 1 ? (e.r[p] + 1) / (1 - vr[p]) : -(e.r[p] + e.s[p]) / (1 - vr[p]));
            for (int i = 0; i < Nfp; ++i)
                e.faceShift[((size_t)f * Np + p) * Nfp + i] = blend * W[(size_t)p * Nfp + (f == 2 ? Nfp - 1 - i : i)];
        }
    }

    // nodal collapses used by the scalar-advection kernel
    e.Dwr = matmul(e.Pr, Np, Ng, e.Vg, Np);
    e.Dws = matmul(e.Ps, Np, Ng, e.Vg, Np);
    e.LIFTn.assign((size_t)Np * 3 * Nfp, 0.0);
    for (int f = 0; f < 3; ++f)
        for (int j = 0; j < Np; ++j)
            for (int i = 0; i < Nfp; ++i) {
                double acc = 0.0;
                for (int g = 0; g < Nfg; ++g) acc += e.LIFT[(size_t)j * 3 * Nfg + f * Nfg + g] * e.If[(size_t)g * Nfp + i];
                e.LIFTn[(size_t)j * 3 * Nfp + f * Nfp + i] = acc;
            }
    return e;
}

}  // namespace hdg

// Reference (standard) triangle element of order N: nodes, Vandermonde matrices, cubature,
// interpolation / projection / lift operators.  Host-side, double precision, no dependencies.
//
// Reference behaviour restated (paths relative to HopeFOAM-0.1/src/DG/):
//   element/polynomials/Legendre/Legendre.C:39-299,387-450                  Jacobi polynomials, GQ/GL nodes, Vandermonde
//   element/baseFunctions/straightBaseFunctions/triangleBaseFunction/triangleBaseFunction.C:46-246  nodes, face maps
//   element/gaussIntegration/gaussIntegration/gaussIntegration.C:62-69      quadrature orders 3(N+1) / 2(N+1)
//   element/gaussIntegration/gaussTriangleIntegration/gaussTriangleIntegration.C:45-97  Vg, Dg, If
// The reference applies per-element quadrature matrices (cellD1dx, massMatrix); for affine triangles those
// collapse to the reference-element operators Pr, Ps, LIFT built here (derivation in DESIGN.md §3).
#pragma once
#include <vector>

namespace hdg {

using Mat = std::vector<double>;   // row-major dense matrix

struct RefElement {
    int N = 0, Np = 0, Nfp = 0, Ng = 0, Nfg = 0;
    std::vector<double> r, s;            // Warp&Blend nodes (Np)
    Mat V, invV;                         // Np x Np
    Mat Dr, Ds;                          // nodal differentiation (Np x Np)
    std::vector<double> gr, gs, gw;      // cell cubature (Ng)
    Mat Vg, Dgr, Dgs;                    // Ng x Np
    std::vector<double> fx, fw;          // face Gauss rule (Nfg)
    Mat If;                              // Nfg x Nfp, nodal face trace -> face Gauss points
    std::vector<int> f2c;                // faceToCellIndex_[3][2][Nfp]
    Mat Mref;                            // Vg^T diag(gw) Vg   (Np x Np)
    Mat Pr, Ps;                          // Mref^-1 Dg{r,s}^T diag(gw)   (Np x Ng)
    Mat LIFT;                            // Np x (3*Nfg): Mref^-1 E_f If^T diag(fw), face-major columns
    Mat Dwr, Dws;                        // Pr*Vg, Ps*Vg (Np x Np) weak nodal derivative (advection collapse)
    Mat LIFTn;                           // Np x (3*Nfp): LIFT_f * If (nodal-flux lift, advection collapse)
    Mat faceShift;                       // 3 x Np x Nfp: node displacement of the cell caused by the displacement of the nodes of face f
                                         // (curved `arc` patches, triangleBaseFunction::addFaceShiftToCell, triangleBaseFunction.C:421-466)

    int f2cIdx(int face, int rot, int i) const { return f2c[(face * 2 + rot) * Nfp + i]; }
};

// throws std::runtime_error for unsupported orders (N < 1 or N > 10; N = 9, 10 use an own collapsed cubature, the reference's table ends at N = 8)
RefElement buildRefElement(int N);

// small dense helpers (row-major)
Mat matmul(const Mat& A, int m, int k, const Mat& B, int n);
Mat transpose(const Mat& A, int m, int n);
Mat inverse(const Mat& A, int n);

// 1-D pieces, exposed for tests
void jacobiGQ(double alpha, double beta, int N, std::vector<double>& x, std::vector<double>& w);
void gaussJacobi(double alpha, double beta, int n, std::vector<double>& x, std::vector<double>& w);
std::vector<double> jacobiGL(double alpha, double beta, int N);
std::vector<double> jacobiP(const std::vector<double>& x, double alpha, double beta, int N);

}  // namespace hdg

// hopeScalarTransportFoam - explicit DG scalar advection  dT/dt + div(U T) = 0  on the HopeFOAM operator interface:
//     dg::solveEquation(dgm::ddt(T) + dgc::div(U, T))        with   divSchemes { div(U,T) default LF; }
// (the composition of SURVEY.md §3.3: EquationConvectionScheme Type3, HopeFOAM-0.1/src/DG/DG/dgc/dgcDiv.C:88-107 +
//  DG/simpleFlux/schemes/LFFlux/LFFlux.C:105-211), SSP-RK2 like dgEulerFoam.  Case: 0/{T,U}; U is a nodal dgVectorField.
// BASELINE configs[0]: Gaussian pulse, fixedValue patches hold the translated exact pulse (refreshed each step at t_n).
#include "dgCFD.H"

using namespace Foam;

static scalar pulse(const vector& x, scalar t, const vector& U0)
{
    const scalar dx = x.x() + 0.3 - U0.x() * t, dy = x.y() + 0.3 - U0.y() * t;
    return std::exp(-(dx * dx + dy * dy) / (2 * 0.1 * 0.1));
}

int main(int argc, char* argv[])
{
    argList args(argc, argv);
    Time runTime(args);
    dgMesh mesh(runTime);

    dgScalarField T("T", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE);
    dgVectorField U("U", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE);
    const vector U0 = U.internalField()[0];          // uniform advection velocity of the test case

    const List<vector> px = mesh.dofLocation();
    {
        Field<scalar>& t = T.primitiveFieldRef();
        forAll(px, i) t[i] = pulse(px[i], 0.0, U0);
    }
    dgScalarField T1("T1", T);

    while (runTime.run()) {
        runTime++;
        T1 = T;
        const scalar tn = runTime.value() - runTime.deltaTValue();
        for (label p = 0; p < mesh.nPatches(); ++p) {
            if (!T1.boundaryField()[p].fixesValue() || mesh.patchNDof(p) == 0) continue;
            const List<vector> bx = mesh.patchDofLocation(p);
            forAll(bx, i) T1.boundaryField()[p][i] = pulse(bx[i], tn, U0);
        }
        dg::solveEquation(dgm::ddt(T1) + dgc::div(U, T1));
        dg::solveEquation(dgm::ddt(T1) + dgc::div(U, T1));
        T = 0.5 * T + 0.5 * T1;
        T.correctBoundaryConditions();
        runTime.write();
    }
    {
        const Field<scalar>& t = T.internalField();
        scalar err = 0;
        forAll(px, i) err += mag(t[i] - pulse(px[i], runTime.value(), U0));
        Pstream::sumReduce(&err, 1);      // gSum over the ranks of a parallel run
        Info << std::setprecision(16) << "TError: " << err / mesh.localRange().second() << endl;
    }
    runTime.writeNow();
    return 0;
}

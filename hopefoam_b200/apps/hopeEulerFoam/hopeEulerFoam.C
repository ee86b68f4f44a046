// hopeEulerFoam - compressible Euler solver on the B200-native DG stage, written against the HopeFOAM operator interface
// (dgm::ddt, dgc::div, dgc::grad, dg::godunovScheme, dg::solveEquation on dgScalarField/dgVectorField) provided by dgCFD.H.
//
// It performs the algorithm of HopeFOAM-0.1/tutorials/DG/2D/isentropicVortex/dgEulerFoam (dgEulerFoam.C:64-131: SSP-RK2 made of
// two forward-Euler solves per conserved field, exact-solution fixedValue boundary refreshed once per step at t_n) on an
// unmodified HopeFOAM case directory: system/{controlDict,dgSchemes,dgSolution}, constant/{polyMesh,transportProperties},
// 0/{rho,rhoU,Ener}.  Usage:  hopeEulerFoam -case <caseDir>
#include "dgCFD.H"

using namespace Foam;

namespace
{

// isentropic vortex, centre (5,0) advected with (1,0), beta = 5 (setNonUniformInlet.H:3-27, setBoundaryValues.H:17-46)
struct Vortex
{
    scalar gamma, beta;
    void eval(const vector& x, scalar t, scalar& rho, vector& rhoU, scalar& Ener) const
    {
        const scalar pi = constant::mathematical::pi;
        const scalar dx = x.x() - 5.0 - t, dy = x.y();
        const scalar r = dx * dx + dy * dy;
        rho = std::pow(1.0 - (gamma - 1.0) * (beta * beta) * std::exp(2.0 * (1.0 - r)) / (16.0 * gamma * pi * pi), 1.0 / (gamma - 1.0));
        rhoU = vector(1 - beta * std::exp(1 - r) * dy / (2.0 * pi), beta * std::exp(1 - r) * dx / (2.0 * pi), 0.0) * rho;
        Ener = std::pow(rho, gamma) / (gamma - 1.0) + 0.5 * magSqr(rhoU) / rho;
    }
};

void setInitialFields(const dgMesh& mesh, const Vortex& vx, dgScalarField& rho, dgVectorField& rhoU, dgScalarField& Ener)
{
    const List<vector> px = mesh.dofLocation();
    Field<scalar>& r = rho.primitiveFieldRef();
    Field<vector>& m = rhoU.primitiveFieldRef();
    Field<scalar>& e = Ener.primitiveFieldRef();
    forAll(px, i) vx.eval(px[i], 0.0, r[i], m[i], e[i]);
    // boundary field := interior trace (patchInternalField), then the fixedValue patches keep it
    for (label p = 0; p < mesh.nPatches(); ++p) {
        rho.patchInternalField(p, rho.boundaryFieldRef()[p]);
        rhoU.patchInternalField(p, rhoU.boundaryFieldRef()[p]);
        Ener.patchInternalField(p, Ener.boundaryFieldRef()[p]);
        rho.boundaryFieldRef()[p].markDirty();
        rhoU.boundaryFieldRef()[p].markDirty();
        Ener.boundaryFieldRef()[p].markDirty();
    }
}

// exact solution at time t on every patch that fixes its value
void setBoundaryValues(const dgScalarField& rho, const dgVectorField& rhoU, const dgScalarField& Ener, const Vortex& vx, const dimensionedScalar& time)
{
    const dgMesh& mesh = rho.mesh();
    for (label p = 0; p < mesh.nPatches(); ++p) {
        if (!rho.boundaryField()[p].fixesValue() || mesh.patchNDof(p) == 0) continue;
        const List<vector> px = mesh.patchDofLocation(p);
        forAll(px, i) vx.eval(px[i], time.value(), rho.boundaryField()[p][i], rhoU.boundaryField()[p][i], Ener.boundaryField()[p][i]);
    }
}

}  // namespace

int main(int argc, char* argv[])
{
    argList args(argc, argv);
    Time runTime(args);
    dgMesh mesh(runTime);

    const IOdictionary transportProperties(IOobject("transportProperties", runTime.constant(), mesh, IOobject::MUST_READ_IF_MODIFIED, IOobject::NO_WRITE));
    const dimensionedScalar gamma = transportProperties.lookup("gamma");

    Info << "Reading fields rho, rhoU, Ener\n" << endl;
    dgScalarField rho("rho", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE);
    dgVectorField rhoU("rhoU", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE);
    dgScalarField Ener("Ener", runTime.timeName(), mesh, IOobject::MUST_READ, IOobject::AUTO_WRITE);

    dgGaussVectorField gther_U("gther_U", rhoU);
    dgGaussScalarField gther_p("gther_p", rho);

    Info << "Create Riemann solver\n" << endl;
    dg::godunovScheme Godunov(mesh, gther_U, gther_p, gamma.value(), rho.gaussField(), rhoU.gaussField(), Ener.gaussField());

    const Vortex vortex{gamma.value(), 5.0};
    setInitialFields(mesh, vortex, rho, rhoU, Ener);

    dgScalarField rho1("rho1", rho);
    dgVectorField rhoU1("rhoU1", rhoU);
    dgScalarField Ener1("Ener1", Ener);

    const dimensionedScalar one("one", gamma.dimensions(), 1.0);

    while (runTime.run()) {
        runTime++;

        // ---- SSP-RK2, stage 1 -------------------------------------------------------------------------------
        rho1 = rho;
        rhoU1 = rhoU;
        Ener1 = Ener;
        setBoundaryValues(rho1, rhoU1, Ener1, vortex, runTime - runTime.deltaT());
        rho1.storeOldTime();
        rhoU1.storeOldTime();
        Ener1.storeOldTime();
        rho1.updateGaussField();
        rhoU1.updateGaussField();
        Ener1.updateGaussField();

        gther_U = rhoU1.gaussField() / rho1.gaussField();
        gther_p = (gamma - one) * (Ener1.gaussField() - 0.5 * (rho1.gaussField() * magSqr(gther_U)));

        Godunov.update(gther_U, gther_p, gamma.value(), rho1.gaussField(), rhoU1.gaussField(), Ener1.gaussField());

        dg::solveEquation(dgm::ddt(rho1) + dgc::div(gther_U, rho1, Godunov.fluxRho()));
        dg::solveEquation(dgm::ddt(rhoU1) + dgc::div(gther_U, rhoU1, Godunov.fluxRhoU()) + dgc::grad(gther_p));
        dg::solveEquation(dgm::ddt(Ener1) + dgc::div(gther_U, Ener1, Godunov.fluxEner()) + dgc::div(gther_U, gther_p));

        rho1.storeOldTime();
        rhoU1.storeOldTime();
        Ener1.storeOldTime();

        // ---- stage 2 (boundary data stay at t_n, as in the reference) ------------------------------------------
        rho1.updateGaussField();
        rhoU1.updateGaussField();
        Ener1.updateGaussField();

        gther_U = rhoU1.gaussField() / rho1.gaussField();
        gther_p = (gamma - one) * (Ener1.gaussField() - 0.5 * rho1.gaussField() * magSqr(gther_U));

        Godunov.update(gther_U, gther_p, gamma.value(), rho1.gaussField(), rhoU1.gaussField(), Ener1.gaussField());

        dg::solveEquation(dgm::ddt(rho1) + dgc::div(gther_U, rho1, Godunov.fluxRho()));
        dg::solveEquation(dgm::ddt(rhoU1) + dgc::div(gther_U, rhoU1, Godunov.fluxRhoU()) + dgc::grad(gther_p));
        dg::solveEquation(dgm::ddt(Ener1) + dgc::div(gther_U, Ener1, Godunov.fluxEner()) + dgc::div(gther_U, gther_p));

        rho = 0.5 * rho + 0.5 * rho1;
        rhoU = 0.5 * rhoU + 0.5 * rhoU1;
        Ener = 0.5 * Ener + 0.5 * Ener1;

        rho.correctBoundaryConditions();
        rhoU.correctBoundaryConditions();
        Ener.correctBoundaryConditions();

        runTime.write();

        Info << "Time = " << runTime.timeName() << nl << endl;
        Info << "ExecutionTime = " << runTime.elapsedCpuTime() << " s" << "  ClockTime = " << runTime.elapsedClockTime() << " s" << nl << endl;
    }

    Info << runTime.value() << endl;

    // error against the exact solution at the final time (eulerError.H:32-38): sum |q - q_exact| / nDof
    {
        const List<vector> px = mesh.dofLocation();
        const Field<scalar>& r = rho.internalField();
        const Field<vector>& m = rhoU.internalField();
        scalar eR = 0, eU = 0;
        forAll(px, i) {
            scalar rx, ex;
            vector mx;
            vortex.eval(px[i], runTime.value(), rx, mx, ex);
            eR += mag(rx - r[i]);
            eU += mag(mx - m[i]);
        }
        Pstream::sumReduce(&eR, 1);      // gSum over the ranks of a parallel run
        Pstream::sumReduce(&eU, 1);
        Info << std::setprecision(16);
        Info << "rhoError: " << eR / mesh.localRange().second() << endl;
        Info << "rhoUError: " << eU / mesh.localRange().second() << endl;
    }
    runTime.writeNow();
    return 0;
}

"""Test utility: write a HopeFOAM-style case directory (system/, constant/, 0/) around a generated triangle mesh."""
from pathlib import Path

import numpy as np

from tests.polymesh_writer import write_polymesh

HDR = """FoamFile
{{
    version     2.0;
    format      ascii;
    class       {cls};
    object      {obj};
}}
// * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * * //
"""


def write_euler_case(case, mg, N, dt, end_time, patches=None, bc_type="fixedValue", write_interval=1000000, bc_types=None):
    """bc_types: optional {patchName: {fieldName: type}} overriding bc_type (e.g. a slip wall: rho/Ener zeroGradient, rhoU reflective)."""
    case = Path(case)
    (case / "system").mkdir(parents=True, exist_ok=True)
    (case / "constant").mkdir(exist_ok=True)
    (case / "0").mkdir(exist_ok=True)
    patches = patches or [("boundary", "patch", mg["patch_edges"][0])]
    write_polymesh(case / "constant" / "polyMesh", mg["xy"], mg["tris"], patches)
    (case / "system" / "controlDict").write_text(HDR.format(cls="dictionary", obj="controlDict") + f"""
application     dgEulerFoam;
startFrom       startTime;
startTime       0;
stopAt          endTime;
endTime         {end_time};
deltaT          {dt};
writeControl    timeStep;
writeInterval   {write_interval};
purgeWrite      0;
writeFormat     ascii;
writePrecision  16;
writeCompression off;
timeFormat      general;
timePrecision   6;
runTimeModifiable true;
""")
    (case / "system" / "dgSolution").write_text(HDR.format(cls="dictionary", obj="dgSolution") + f"""
DG
{{
    meshDimension 2;
    baseOrder {N};   // polynomial order
}}
solvers
{{
    "rho(1|2|3)"
    {{
        tolerance   1e-10;
        relTol      0;
        kspSolver   preonly;
        kspPC       ilu;
    }}
}}
""")
    (case / "system" / "dgSchemes").write_text(HDR.format(cls="dictionary", obj="dgSchemes") + """
ddtSchemes
{
    default         Euler;
}
gradSchemes
{
    default         none;
    grad(gther_p)   default none;
}
divSchemes
{
    default         none;
    div(gther_U,rho1)   default none;
    div(gther_U,rhoU1)  default none;
    div(gther_U,Ener1)  default none;
    div(gther_U,gther_p) default none;
    div(U,T)        default LF;
    div(U,T1)       default LF;
}
laplacianSchemes
{
    default         none;
}
godunovScheme
{
    fluxScheme      Roe;
    limiteScheme    Triangle;
}
""")
    (case / "constant" / "transportProperties").write_text(HDR.format(cls="dictionary", obj="transportProperties") + """
gamma gamma [0 0 0 0 0 0 0] 1.4;
""")
    def field(name, cls, dims, uni):
        body = HDR.format(cls=cls, obj=name) + f"\ndimensions      {dims};\n\ninternalField   uniform {uni};\n\nboundaryField\n{{\n"
        for pname, ptype, _ in patches:
            bt = (bc_types or {}).get(pname, {}).get(name, bc_type)
            body += f"    {pname}\n    {{\n        type            {bt};\n"
            if bt == "fixedValue":
                body += f"        value           uniform {uni};\n"
            body += "    }\n"
        body += "    frontAndBackPlanes\n    {\n        type            empty;\n    }\n}\n"
        (case / "0" / name).write_text(body)
    field("T", "dgScalarField", "[0 0 0 0 0 0 0]", "0")
    field("U", "dgVectorField", "[0 1 -1 0 0 0 0]", "(1 0.5 0)")
    field("p", "dgScalarField", "[1 -1 -2 0 0 0 0]", "1")           # read (and otherwise unused) by the tutorial solver's createFields.H
    field("rho", "dgScalarField", "[1 -3 0 0 0 0 0]", "1")
    field("rhoU", "dgVectorField", "[1 -2 -1 0 0 0 0]", "(1 0 0)")
    field("Ener", "dgScalarField", "[1 -1 -2 0 0 0 0]", "3")
    return case


def read_field(path, n_cmpt):
    """internalField nonuniform List<...> N ( ... ) of a written time-directory field."""
    import re
    t = Path(path).read_text()
    i = t.index("internalField")
    j = t.index("(", i)
    n = int(re.search(r"(\d+)\s*$", t[i:j]).group(1))
    end = t.index("\n)\n", j)
    vals = np.array(t[j + 1:end].replace("(", " ").replace(")", " ").split(), dtype=float)
    return vals.reshape(n, n_cmpt) if n_cmpt > 1 else vals

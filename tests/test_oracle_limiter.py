"""oracle.triangle_limit - the restatement of the reference's `Triangle` slope limiter (Trianglelimite.C:61-864).  The reference
publishes no numbers for a limited run, so the restatement is parity-unpinned; these tests hold it to the properties the algorithm
guarantees (constant states are fixed points, cell averages are kept, a linear density is reproduced away from the boundary, the
density floor of :823-827)."""
import numpy as np
import pytest

from hopefoam_b200 import meshgen
from oracle import dg_oracle as o
from tests import helpers as H


def _case(n, N, kind):
    mg = meshgen.jittered_square(n)
    return o.Case(H.oracle_mesh(mg), N, bc_kinds=[kind])


def _bvals(case, rho, U, E):
    bR, bU, bE = [case.patch_internal(rho, 0)], [case.patch_internal(U, 0)], [case.patch_internal(E, 0)]
    case.evaluate_bc(rho, bR)
    case.evaluate_bc(U, bU, is_vector=True)
    case.evaluate_bc(E, bE)
    return bR, bU, bE


def _weights(case):
    return np.linalg.inv(case.ref.V @ case.ref.V.T).sum(0) / 2


@pytest.mark.parametrize("kind", [o.BC_FIXED, o.BC_ZEROGRAD])
def test_constant_state_is_a_fixed_point(kind):
    case = _case(5, 4, kind)
    x = case.geo.x[..., 0]
    rho, U, E = np.full_like(x, 1.3), np.stack([np.full_like(x, 0.4), np.full_like(x, -0.2)], -1), np.full_like(x, 2.5)
    r, u, e = o.triangle_limit(case, rho, U, E, *_bvals(case, rho, U, E))
    assert np.abs(r - rho).max() < 1e-14 and np.abs(u - U).max() < 1e-14 and np.abs(e - E).max() < 1e-14


def test_linear_density_reproduced_and_averages_kept():
    case = _case(10, 3, o.BC_ZEROGRAD)
    m = case.mesh
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho = 1 + 0.01 * x + 0.02 * y
    uu, vv, p = 0.3 + 0.002 * x - 0.001 * y, 0.1 + 0.001 * y, 1.0 + 0.003 * x
    U = np.stack([rho * uu, rho * vv], -1)
    E = p / 0.4 + 0.5 * rho * (uu ** 2 + vv ** 2)
    r, u, e = o.triangle_limit(case, rho, U, E, *_bvals(case, rho, U, E))
    w = _weights(case)
    assert abs(w.sum() - 1) < 1e-13
    assert np.abs(r @ w - rho @ w).max() < 1e-14 and np.abs(e @ w - E @ w).max() < 1e-13          # cell means of rho and E survive
    layer = np.full(m.K, 99)
    layer[m.face_owner[m.patches[0]["faces"]]] = 0
    for it in (1, 2):
        for f in np.nonzero(m.face_nbr >= 0)[0]:
            a, b = m.face_owner[f], m.face_nbr[f]
            if layer[a] == it - 1 and layer[b] > it:
                layer[b] = it
            if layer[b] == it - 1 and layer[a] > it:
                layer[a] = it
    inner = layer >= 2                                   # the ghost-cell gradients reach two layers of cells
    assert inner.sum() > 50
    assert np.abs(r - rho)[inner].max() < 1e-14          # all neighbour gradients agree: weights 1/3 each, the plane comes back
    assert np.abs(u - U)[inner].max() < 1e-4 and np.abs(e - E)[inner].max() < 1e-4      # linearised products: second-order remainder


def test_density_floor():
    """A cell whose limited slope would push a node below tol gets its density slope halved until it does not (:823-827)."""
    case = _case(4, 2, o.BC_ZEROGRAD)
    x, y = case.geo.x[..., 0], case.geo.x[..., 1]
    rho = 0.012 + 0.04 * np.maximum(x - 5.0, 0.0)        # a plateau just above tol and a steep ramp -> large downwind slopes
    rho[case.mesh.K // 2] *= 0.9
    U, E = np.zeros(x.shape + (2,)), np.full_like(x, 2.5)
    r, u, e = o.triangle_limit(case, rho, U, E, *_bvals(case, rho, U, E))
    assert r.min() >= 1e-2 - 1e-15
    assert np.isfinite(r).all() and np.isfinite(e).all()


def test_reflective_ghost_removes_normal_momentum():
    """Uniform flow along a slip wall is a fixed point; the ghost cell of a reflective face carries the owner's average with the
    normal momentum removed once (:176-205)."""
    mg = meshgen.jittered_square(5)
    e = mg["patch_edges"][0]
    om = o.build_connectivity(mg["xy"], mg["tris"], [[(int(c), (int(a), int(b))) for c, a, b in e]], [{"name": "wall", "type": "wall"}],
                              point_equiv=mg["point_equiv"])
    case = o.Case(om, 3, bc_kinds=[o.BC_REFLECTIVE])
    x = case.geo.x[..., 0]
    rho, U, E = np.full_like(x, 1.0), np.zeros(x.shape + (2,)), np.full_like(x, 2.5)
    r, u, en = o.triangle_limit(case, rho, U, E, *_bvals(case, rho, U, E))
    assert np.abs(r - rho).max() < 1e-14 and np.abs(u).max() < 1e-14 and np.abs(en - E).max() < 1e-14

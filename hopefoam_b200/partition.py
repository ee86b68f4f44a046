"""Strip partitions of the periodic benchmark square for the weak-scaling runs (BASELINE.json configs[2]) and the
per-stage halo exchange between them.

Partition = the reference's `simple` geometric decomposition with n = (1 P 1) (system/decomposeParDict method simple,
HopeFOAM-0.1/tutorials/DG/2D/isentropicVortex/system/decomposeParDict:24-33) applied to a mesh that is P strips tall: rank r
owns strip r.  Cut faces become two processor patches per rank (0 = bottom, towards rank r-1; 1 = top, towards rank r+1,
periodic wrap), whose faces are listed in ascending column order on both sides - the analogue of the reference's
"inter-processor faces in ascending global face id" rule (applications/utilities/DG/dgDecomposePar/domainDecompositionMesh.C:215-240).

The exchange itself is processorDgPatchField::initEvaluate/evaluate (src/DG/fields/dgPatchFields/constraint/processor/
processorDgPatchField.C:235-331) restated for GPUs: the library packs the owner-side nodal traces (already reversed per
face) of all planes into ONE buffer per neighbour, NCCL send/recv moves it over NVLink, the library unpacks into the
ghost slots.  One message per neighbour per stage instead of the reference's 6 rounds.
"""
from __future__ import annotations

import numpy as np

from . import capi, meshgen


def strip_partition(n: int, world: int, rank: int, seed=20240501):
    """Local mesh of rank `rank`: n x n x 2 jittered triangles on [0,10] x [-5+10r, 5+10r]; x-periodic locally; when
    world > 1 the bottom and top rows of edges are processor patches, else y-periodic as well."""
    if world == 1:
        mg = meshgen.jittered_square(n, periodic=True, seed=seed)
        mg["y_shift"] = 0.0
        mg["peers"] = []
        mg["n"] = n
        return mg
    y0 = -5.0 + 10.0 * rank
    mg = meshgen.jittered_square(n, y0=y0, y1=y0 + 10.0, periodic=False, seed=seed + rank)
    eq = np.arange((n + 1) * (n + 1), dtype=np.int32).reshape(n + 1, n + 1)
    eq[:, n] = eq[:, 0]                                   # glue left/right columns
    mg["point_equiv"] = eq.reshape(-1)
    edges = mg["patch_edges"][0]                          # order: bottom (n), right (n), top (n), left (n)
    bottom, top = edges[0:n], edges[2 * n:3 * n]          # both listed in ascending column order by the generator
    mg["patch_edges"] = [np.ascontiguousarray(bottom), np.ascontiguousarray(top)]
    mg["y_shift"] = 10.0 * rank
    mg["peers"] = [(rank - 1) % world, (rank + 1) % world]
    mg["n"] = n
    return mg


def exchange_messages(dist, send, recv, peers):
    """Post the two sends / two receives of one stage and wait.  send/recv = [bottom, top] tensors, peers = (down, up).
    Message order per peer matters when down == up (world == 2): my bottom trace is the peer's TOP ghost, so receives
    are posted top-first.  Works with any torch.distributed backend (NCCL on GPUs, gloo in the CPU tests)."""
    down, up = peers
    ops = [dist.P2POp(dist.isend, send[0], down), dist.P2POp(dist.isend, send[1], up),
           dist.P2POp(dist.irecv, recv[1], up), dist.P2POp(dist.irecv, recv[0], down)]
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class HaloExchanger:
    """Per-stage exchange of the two processor patches of a strip through torch.distributed (NCCL).

    Pack / NCCL / unpack run on the library's HALO stream; `step_ssprk2` splits every stage into a launch over the two rows of
    partition-boundary elements (which need the halo) and a launch over the interior, ordered so that the exchange for the
    NEXT stage (which only needs the boundary rows of this stage) overlaps this stage's interior launch."""

    def __init__(self, ctx: capi.Context, sid: int, part, dist, torch, n_planes=4, share=None):
        """`share`: another exchanger of the same context whose message buffers are re-used (the buffers are bound to the context's
        processor patches; all halo work of a context runs in order on its halo stream, so several states can share them)."""
        self.ctx, self.sid, self.dist, self.torch = ctx, sid, dist, torch
        self.peers = part["peers"]
        self.halo_stream = torch.cuda.ExternalStream(ctx.stream(1))
        self.send, self.recv = ([], []) if share is None else (share.send, share.recv)
        for p in (0, 1):
            ctx.set_patch_kind(sid, p, capi.BC_PROCESSOR)
            if share is not None:
                continue
            cnt = ctx.halo_count(p) * n_planes
            s = torch.zeros(cnt, dtype=torch.float64, device="cuda")
            r = torch.zeros(cnt, dtype=torch.float64, device="cuda")
            ctx.halo_bind(p, s.data_ptr(), r.data_ptr(), cnt)
            self.send.append(s)
            self.recv.append(r)
        # boundary rows of the strip: the first and the last row of quads = elements [0, 2n) and [K-2n, K), widened to octets
        K, n2 = ctx.K, part.get("row_elems", 2 * part["n"])
        self.lo_end = min(K, (n2 + 7) // 8 * 8)
        self.hi_begin = max(self.lo_end, (K - n2) // 8 * 8)
        self.K = K
        self._primed = False

    def _exchange_on_halo_stream(self, which: int):
        ctx = self.ctx
        with self.torch.cuda.stream(self.halo_stream):
            ctx.halo_pack(self.sid, which, 0)
            ctx.halo_pack(self.sid, which, 1)
            exchange_messages(self.dist, self.send, self.recv, self.peers)
            ctx.halo_unpack(self.sid, which, 0)
            ctx.halo_unpack(self.sid, which, 1)

    def exchange(self, which: int):
        """Blocking-order exchange (no overlap): halo of copy `which` (0 current, 1 stage) before the stage that reads it."""
        self.ctx.stream_wait(1, 0)
        self._exchange_on_halo_stream(which)
        self.ctx.stream_wait(0, 1)

    def step_ssprk2(self, gamma: float, dt: float):
        """One SSP-RK2 step with the exchange of stage s+1 overlapped with the interior launch of stage s."""
        ctx, sid = self.ctx, self.sid
        if not self._primed:                       # halo of the very first stage
            self.exchange(0)
            self._primed = True
        for stage, (a, b) in enumerate(((0.0, 1.0), (0.5, 0.5))):
            ctx.stream_wait(0, 1)                  # ghosts of this stage's input copy have landed
            ctx.euler_stage_range(sid, gamma, dt, stage, a, b, 0, self.lo_end, self.hi_begin, self.K)    # both boundary rows, one launch
            ctx.stream_wait(1, 0)                  # boundary rows of this stage are final -> exchange for the next stage ...
            self._exchange_on_halo_stream(1 - stage)      # stage 0 wrote the stage copy (1); stage 1 wrote the current copy (0)
            ctx.euler_stage_range(sid, gamma, dt, stage, a, b, self.lo_end, self.hi_begin)    # ... overlaps the interior


class GeneralHalo:
    """Halo exchange for an arbitrary decomposition built by `Context.set_mesh_from_decomposition` (dgDecomposePar rules): one
    processor patch per neighbour rank, both sides list the cut faces in the same (ascending global face id) order, so one packed
    message per neighbour per stage suffices.  Serial with the stage launch (no interior/boundary split: the boundary cells of
    a general partition are not contiguous)."""

    def __init__(self, ctx: capi.Context, sid: int, dist, torch, n_planes=4):
        self.ctx, self.sid, self.dist, self.torch = ctx, sid, dist, torch
        self.halo_stream = torch.cuda.ExternalStream(ctx.stream(1))
        nbr = ctx.proc_addressing()["patch_nbr_proc"]
        self.patches = [(p, int(q)) for p, q in enumerate(nbr) if q >= 0]
        self.buf = {}
        for p, q in self.patches:
            ctx.set_patch_kind(sid, p, capi.BC_PROCESSOR)
            cnt = ctx.halo_count(p) * n_planes
            s = torch.zeros(cnt, dtype=torch.float64, device="cuda")
            r = torch.zeros(cnt, dtype=torch.float64, device="cuda")
            ctx.halo_bind(p, s.data_ptr(), r.data_ptr(), cnt)
            self.buf[p] = (s, r)

    def exchange(self, which: int):
        ctx, dist = self.ctx, self.dist
        ctx.stream_wait(1, 0)
        with self.torch.cuda.stream(self.halo_stream):
            for p, _ in self.patches:
                ctx.halo_pack(self.sid, which, p)
            ops = []
            for p, q in self.patches:
                ops.append(dist.P2POp(dist.isend, self.buf[p][0], q))
                ops.append(dist.P2POp(dist.irecv, self.buf[p][1], q))
            if ops:
                for req in dist.batch_isend_irecv(ops):
                    req.wait()
            for p, _ in self.patches:
                ctx.halo_unpack(self.sid, which, p)
        ctx.stream_wait(0, 1)

    def step_ssprk2(self, gamma: float, dt: float):
        self.exchange(0)
        self.ctx.euler_stage(self.sid, gamma, dt, 0, 0.0, 1.0)
        self.exchange(1)
        self.ctx.euler_stage(self.sid, gamma, dt, 1, 0.5, 0.5)


def sector_partition(n_r: int, n_theta: int, world: int, rank: int, r0=0.5, r1=20.0):
    """Angular-sector partition of the cylinder O-grid (BASELINE configs[4]): rank r owns theta in [2 pi r/P, 2 pi (r+1)/P) with n_theta
    rings.  Patches: 0 = cut towards rank r-1 (processor), 1 = cut towards rank r+1 (processor), 2 = cylinder wall, 3 = far field.
    With world == 1 the annulus is closed on itself and only the wall / far-field patches remain (ids 0, 1)."""
    if world == 1:
        mg = meshgen.ogrid_sector(n_r, n_theta, 0.0, 2 * np.pi, r0, r1, closed=True)
        mg["patch_edges"] = [mg["sides"]["left"], mg["sides"]["right"]]
        mg["peers"], mg["n"], mg["row_elems"] = [], n_r, 2 * n_r
        mg["wall_patch"], mg["farfield_patch"] = 0, 1
        return mg
    t0, t1 = 2 * np.pi * rank / world, 2 * np.pi * (rank + 1) / world
    mg = meshgen.ogrid_sector(n_r, n_theta, t0, t1, r0, r1)
    mg["patch_edges"] = [mg["sides"]["bottom"], mg["sides"]["top"], mg["sides"]["left"], mg["sides"]["right"]]
    mg["peers"] = [(rank - 1) % world, (rank + 1) % world]
    mg["n"], mg["row_elems"] = n_r, 2 * n_r
    mg["wall_patch"], mg["farfield_patch"] = 2, 3
    return mg


def global_mesh(n: int, world: int, seed=20240501):
    """The undecomposed mesh the strips of `strip_partition` tile: strips concatenated in rank order (cells and points),
    doubly periodic through the canonical point map.  Used to check multi-GPU runs against one GPU on the same mesh."""
    xs, ts, eqs = [], [], []
    P = (n + 1) * (n + 1)
    for r in range(world):
        y0 = -5.0 + 10.0 * r
        mg = meshgen.jittered_square(n, y0=y0, y1=y0 + 10.0, periodic=False, seed=seed + r)
        xs.append(mg["xy"])
        ts.append(mg["tris"] + r * P)
        eq = (np.arange(P, dtype=np.int32) + r * P).reshape(n + 1, n + 1)
        eq[:, n] = eq[:, 0]
        eqs.append(eq)
    for r in range(world):                       # top row of strip r == bottom row of strip r+1 (periodic wrap)
        eqs[r][n, :] = eqs[(r + 1) % world][0, :]
    if world == 1:
        eqs[0][n, :] = eqs[0][0, :]
    # resolve chains (corner points are glued twice)
    eq = np.concatenate([e.reshape(-1) for e in eqs])
    for _ in range(3):
        eq = eq[eq]
    return {"xy": np.concatenate(xs), "tris": np.concatenate(ts).astype(np.int32), "point_equiv": eq.astype(np.int32),
            "patch_edges": []}

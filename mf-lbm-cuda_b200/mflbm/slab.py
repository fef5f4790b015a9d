"""x-slab domain decomposition of one lattice across the GPUs of a box: one process per GPU, NCCL over NVLink.

The reference is single-GPU (README.md:119 lists multi-GPU as future work); this is new work.

Partition: contiguous x ranges, full y and z.  x is never periodic and both x ends are walls
(/root/reference/src/IO_multiphase.cpp:210-213), so the ranks form an open chain.

What crosses a face and when (AA pattern, see DESIGN.md "x-slab exchange"):

  even step  collide (local) -> PDF halo kind 0: my first/last REAL column, the five populations per component
             that leave through that face (ex = -1 on the left face, ex = +1 on the right) -> neighbour's GHOST
             column, so that the next odd pull finds them
  odd step   collide (pull from x-e, push to x+e) -> PDF halo kind 1: what I pushed into my GHOST columns
             -> the neighbour's REAL boundary column (the owner of those nodes)
  both       -> boundary kernels (inlet/outlet/periodic/porous plate; they read the freshly landed columns and
             write phi ghost planes) -> phi halo kind 2: four real columns per side -> neighbour's four phi ghost
             columns (the colour-gradient chain depends on phi within Chebyshev radius 4) -> gradient chain, each
             stage evaluated redundantly on the ghost columns it needs.

No collective is needed on the data path; the monitor sums are combined with all_reduce.

`SlabStepper` is backend-agnostic: it drives any object with the small `SlabBackend` interface below.  The product
backend is `CudaSlab` (libmflbm.so through the C ABI, halo buffers exposed as zero-copy torch tensors); the CPU
tests drive the same stepper over gloo with a numpy backend to check the protocol.
"""
from __future__ import annotations

import ctypes as C
import json
import os
import sys
import time
from dataclasses import dataclass

import numpy as np

KIND_PDF_EVEN, KIND_PDF_ODD, KIND_PHI = 0, 1, 2
LEFT, RIGHT = 0, 1


@dataclass
class SlabRange:
    rank: int
    world: int
    x0: int          # global index (1-based) of the first real column
    nx_local: int

    @property
    def x1(self) -> int:
        return self.x0 + self.nx_local - 1

    @property
    def has_left(self) -> bool:
        return self.rank > 0

    @property
    def has_right(self) -> bool:
        return self.rank < self.world - 1


def partition(nx_global: int, world: int, rank: int) -> SlabRange:
    """contiguous, balanced x ranges: the first nx % world ranks get one extra column"""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(nx_global, world)
    if base < 4 and world > 1:
        raise ValueError(f"{nx_global} columns over {world} slabs: a slab must be at least 4 columns wide (phi halo)")
    nx_local = base + (1 if rank < rem else 0)
    x0 = 1 + rank * base + min(rank, rem)
    return SlabRange(rank, world, x0, nx_local)


def balanced_cuts(col_cost, world: int, side_cost: float = 0.0, min_width: int = 4) -> list[int]:
    """x cuts [0, c1, ..., nx] that equalise  sum(col_cost over the slab's columns) + side_cost * (neighbours of the slab).

    col_cost[i]: work of global column i+1 (its fluid nodes: the collide kernels run one thread per fluid node); side_cost: what
    one neighbour costs a slab per step (halo pack / unpack kernels, arrival waits, the chain's redundant ghost columns), in the
    same unit.  End slabs have one neighbour, interior slabs two, so equal widths leave the end ranks waiting; in a random pack
    the fluid count per slab also varies by a few per cent.  Deterministic: every rank computes the same cuts."""
    col_cost = np.asarray(col_cost, dtype=np.float64)
    nx = int(col_cost.size)
    if world < 1 or nx < world * min_width:
        raise ValueError(f"{nx} columns over {world} slabs: a slab must be at least {min_width} columns wide (phi halo)")
    cs = np.cumsum(col_cost)
    target = (cs[-1] + (2 * world - 2) * side_cost) / world
    cuts = [0]
    for r in range(world - 1):
        want = (r + 1) * target - side_cost * (2 * (r + 1) - 1)      # column cost left of the cut: ranks 0..r carry 2(r+1)-1 sides
        c = int(np.searchsorted(cs, want))                           # cs[c] >= want > cs[c-1]: cutting after column c or c+1
        if c < nx and (c == 0 or abs(cs[c] - want) < abs(cs[c - 1] - want)):
            c += 1
        c = max(c, cuts[-1] + min_width)
        c = min(c, nx - (world - 1 - r) * min_width)
        cuts.append(c)
    cuts.append(nx)
    return cuts


def partition_balanced(col_cost, world: int, rank: int, side_cost: float = 0.0) -> SlabRange:
    """the slab of `rank` under balanced_cuts (same SlabRange as partition())"""
    if not (0 <= rank < world):
        raise ValueError("bad rank/world")
    cuts = balanced_cuts(col_cost, world, side_cost)
    return SlabRange(rank, world, cuts[rank] + 1, cuts[rank + 1] - cuts[rank])


class SlabStepper:
    """Per-step sequencing of one slab: compute phases interleaved with the two halo exchanges.

    backend interface:
        step_phase(ntime, phase)            phase 0 collide, 1 boundary kernels, 2 gradient chain
        halo_pack(kind) / halo_unpack(kind)
        halo_tensors(kind, side) -> (send, recv) torch tensors (device of the backend)
    """

    def __init__(self, backend, rng: SlabRange, group=None):
        import torch.distributed as dist
        self.b, self.rng, self.dist, self.group = backend, rng, dist, group
        self.exchanges = 0
        self.last_ntime = None

    def exchange(self, kind: int) -> None:
        dist, r = self.dist, self.rng
        if r.world == 1:
            return
        if getattr(self.b, "p2p", False):
            # one kernel per neighbour writes the message into its receive buffer over NVLink and publishes a sequence
            # number there; the unpack kernels spin on my own flags.  No NCCL call, no host synchronisation.
            self.b.halo_push(kind)
            self.b.halo_unpack_wait(kind)
            self.exchanges += 1
            return
        self.b.halo_pack(kind)
        ops = []
        # a message sent through my right face lands in my right neighbour's "left" receive buffer and vice versa
        if r.has_left:
            s, v = self.b.halo_tensors(kind, LEFT)
            ops += [dist.P2POp(dist.isend, s, r.rank - 1, self.group), dist.P2POp(dist.irecv, v, r.rank - 1, self.group)]
        if r.has_right:
            s, v = self.b.halo_tensors(kind, RIGHT)
            ops += [dist.P2POp(dist.isend, s, r.rank + 1, self.group), dist.P2POp(dist.irecv, v, r.rank + 1, self.group)]
        for w in dist.batch_isend_irecv(ops):
            w.wait()
        self.b.halo_unpack(kind)
        self.exchanges += 1

    def step(self, ntime: int) -> None:
        self.b.step_phase(ntime, 0)
        self.exchange(KIND_PDF_ODD if ntime % 2 else KIND_PDF_EVEN)
        self.b.step_phase(ntime, 1)
        self.exchange(KIND_PHI)
        self.b.step_phase(ntime, 2)
        self.last_ntime = ntime

    def run(self, ntime_first: int, nsteps: int) -> None:
        if getattr(self.b, "p2p", False) and nsteps > 0:
            # the library sequences collide / push / wait-unpack / boundaries / push / wait-unpack / chain itself and replays
            # step pairs, messages included, from a CUDA graph
            self.b.solver.run(ntime_first, nsteps)
            self.exchanges += 2 * nsteps
            self.last_ntime = ntime_first + nsteps - 1
            return
        for n in range(nsteps):
            self.step(ntime_first + n)

    def settle(self) -> None:
        """Make every rank's OWN columns complete before they are gathered (checkpoint, field output).

        After an even step the "before odd" boundary kernels (inlet, outlet, porous plate) of my neighbour write the
        populations its next pull needs into ITS ghost copy of my boundary column (planes k = 0, nz+1, Z_porous_plate);
        nothing on my side ever reads those entries, so my copy is stale.  Sending the ghost columns back to their owner
        (the kind-1 message) makes the owner's copy equal to the single-domain array.  After an odd step the regular
        kind-1 exchange has already done this."""
        if self.last_ntime is not None and self.last_ntime % 2 == 0:
            self.exchange(KIND_PDF_ODD)
            self.exchanges -= 1


class _DeviceArray:
    """zero-copy view of a device buffer owned by libmflbm.so, for torch.as_tensor (CUDA array interface v2)"""

    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


class CudaSlab:
    """SlabBackend over libmflbm.so: one mflbm.Solver created with an mflbm_slab, on the current torch CUDA stream."""

    def __init__(self, params, prec: str, rng: SlabRange, device: int, stream=None):
        import torch
        import mflbm
        self.torch = torch
        self.rng = rng
        self.stream = stream if stream is not None else torch.cuda.Stream(device=device)
        slab = mflbm.Slab(rng.x0, rng.nx_local, int(rng.has_left), int(rng.has_right)) if rng.world > 1 else None
        self.solver = mflbm.Solver(params, prec, slab=slab, device=device, stream=self.stream.cuda_stream)
        self.prec = prec
        self._t = {}
        if rng.world > 1:
            typestr = "<f8" if prec == "f64" else "<f4"
            with torch.cuda.device(device):
                for kind in (0, 1, 2):
                    for side in (LEFT, RIGHT):
                        if (side == LEFT and not rng.has_left) or (side == RIGHT and not rng.has_right):
                            continue
                        s, r, n = self.solver.halo_buffers(kind, side)
                        self._t[(kind, side)] = (torch.as_tensor(_DeviceArray(s, n, typestr), device=f"cuda:{device}"),
                                                 torch.as_tensor(_DeviceArray(r, n, typestr), device=f"cuda:{device}"))

    def step_phase(self, ntime, phase):
        self.solver.step_phase(ntime, phase)

    def connect_p2p(self, dist, group=None) -> None:
        """Exchange CUDA IPC handles with the neighbour ranks (once) so that halo messages go through peer memory.
        Every rank exports the one allocation that holds its receive buffers and arrival flags; the offsets inside it are
        the same on every rank that has the same y/z extents, but they are sent along anyway.  Collective: every rank
        takes part in every step, and all of them fall back to NCCL send/recv together if any import fails."""
        import mflbm
        import torch
        r = self.rng
        self.p2p = False
        if r.world == 1:
            return
        mine, why = None, ""
        try:
            base, nbytes = self.solver.halo_p2p_region()
            mine = {"handle": mflbm.ipc_export(base), "offsets": {}}
            for kind in (0, 1, 2):
                for side in (LEFT, RIGHT):
                    recv, flag = self.solver.halo_p2p_local(kind, side)
                    mine["offsets"][(kind, side)] = (recv - base, flag - base)
        except Exception as e:
            mine, why = None, str(e)
        everyone = [None] * r.world
        dist.all_gather_object(everyone, mine, group=group)
        ok = 1 if mine is not None else 0
        self._peer_bases = {}
        for side, nb in ((LEFT, r.rank - 1), (RIGHT, r.rank + 1)):
            if not ok or nb < 0 or nb >= r.world:
                continue
            try:
                if everyone[nb] is None:
                    raise RuntimeError(f"rank {nb} exported nothing")
                pbase = mflbm.ipc_import(everyone[nb]["handle"])
                self._peer_bases[side] = pbase
                opposite = RIGHT if side == LEFT else LEFT     # my left neighbour receives my message in ITS right-side buffer
                for kind in (0, 1, 2):
                    orecv, oflag = everyone[nb]["offsets"][(kind, opposite)]
                    self.solver.halo_p2p_connect(kind, side, pbase + orecv, pbase + oflag)
            except Exception as e:
                ok, why = 0, str(e)
        flag = torch.tensor([ok], dtype=torch.int32, device=f"cuda:{torch.cuda.current_device()}")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
        self.p2p = bool(int(flag[0]))
        if not self.p2p and why:
            print(f"[rank {r.rank}] peer-memory halo transport unavailable ({why}); using NCCL send/recv", flush=True)

    def save_checkpoint(self, path, stepper: "SlabStepper", nx_global: int, ntime_next: int, dist) -> None:
        """collective: settle the slabs, download every rank's state and write the reference-format checkpoint file"""
        stepper.settle()
        st = self.solver.download_state(fields=("pdf", "phi"), convective=self.solver.params.outlet_BC == 1)
        write_checkpoint_slabs(path, self.rng, nx_global, st, ntime_next, float(self.solver.params.force_z), float(self.solver.params.rho_in), dist)

    def load_checkpoint(self, path, nx_global: int, W_in=None, rho_in_new: bool = False) -> int:
        """Restart this slab from a checkpoint file written by any decomposition; returns the step index to continue with.
        As in the reference's restart (/root/reference/src/Init_multiphase.cpp:519-520) and in mflbm_run, force_z and rho_in come
        from the file (rho_in_new = True keeps the control file's rho_in, `rho_in_new` key).  W_in: this slab's columns of the
        inlet velocity profile (needed unless init_state has already uploaded it)."""
        P = self.solver.params
        c = read_checkpoint_slab(path, self.rng, nx_global, int(P.ny), int(P.nz), self.solver.rt, P.outlet_BC == 1)
        P.force_z = c["force_z"]
        if not rho_in_new:
            P.rho_in = c["rho_in"]
        self.solver.set_params(P)
        self.solver.upload_state(pdf=c["pdf"], phi=c["phi"], W_in=W_in, f_convec=c.get("f_convec"), g_convec=c.get("g_convec"), phi_convec=c.get("phi_convec"))
        self.solver.color_gradient()
        return c["ntime_next"]

    def close(self) -> None:
        """destroy the solver and close the CUDA IPC mappings of the neighbours' buffers"""
        import mflbm
        self.solver.close()
        for pbase in getattr(self, "_peer_bases", {}).values():
            try:
                mflbm.ipc_release(pbase)
            except Exception:
                pass
        self._peer_bases = {}

    def halo_push(self, kind):
        self.solver.halo_push(kind)

    def halo_unpack_wait(self, kind):
        self.solver.halo_unpack_wait(kind)

    def halo_pack(self, kind):
        self.solver.halo_pack(kind)

    def halo_unpack(self, kind):
        self.solver.halo_unpack(kind)

    def halo_tensors(self, kind, side):
        return self._t[(kind, side)]


def reduce_monitor(m: dict, rng: SlabRange, params, dist, device=None) -> dict:
    """combine per-slab monitor sums into the global figures of src/Monitor.cpp:111-171 (SUM / MAX all_reduce)"""
    import torch
    keys = ["vol1_sum", "vol2_sum", "mass1_sum", "mass2_sum", "vol1_full", "vol2_full", "mass1_full", "mass2_full",
            "fl1_avg", "fl2_avg", "fl1_avg_whole", "fl2_avg_whole"]
    keys += [k for k in ("pre_w_sum", "pre_nw_sum", "n_w", "n_nw", "outlet_phase1_count") if k in m]   # capillary-pressure / breakthrough monitors (src/Monitor.cpp:376-459)
    sums = torch.tensor([float(m[k]) for k in keys] + list(m["kinetic_energy"]), dtype=torch.float64, device=device)
    mx = torch.tensor([m["umax"], float(m["nan_detected"])], dtype=torch.float64, device=device)
    if rng.world > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    out = dict(zip(keys, sums[:len(keys)].tolist()))
    out["kinetic_energy"] = sums[len(keys):].tolist()
    if "profiles" in m:   # the 7 per-slice sums behind results/out1.output/profile/* (src/Monitor.cpp:34-80): additive over slabs
        names = sorted(m["profiles"])
        prof = torch.tensor(np.stack([np.asarray(m["profiles"][k], dtype=np.float64) for k in names]), dtype=torch.float64, device=device)
        if rng.world > 1:
            dist.all_reduce(prof, op=dist.ReduceOp.SUM)
        out["profiles"] = {k: prof[i].cpu().numpy() for i, k in enumerate(names)}
    out["umax"], out["nan_detected"] = float(mx[0]), int(mx[1])
    out["saturation"] = out["vol1_sum"] / (out["vol1_sum"] + out["vol2_sum"])
    out["saturation_full_domain"] = out["vol1_full"] / (out["vol1_full"] + out["vol2_full"])
    out["ca"] = ((out["fl1_avg"] + out["fl2_avg"]) / float(params.A_xy)) * float(params.la_nu1) / float(params.lbm_gamma)
    return out


# ----------------------------------------------------------------------------------------------------------
# checkpoint of a decomposed lattice in the reference's single-file layout (results/out2.checkpoint/id0000,
# /root/reference/src/IO_multiphase.cpp:252-305, read back by src/Init_multiphase.cpp:501-538):
#     int32 ntime + 1 | T force_z | T rho_in | pdf[2][19][nz+2][ny+2][nx+2] | phi[nz+8][ny+8][nx+8]
#     | convective outlet only: f_convec[19][ny+2][nx+2] | g_convec[19][ny+2][nx+2] | phi_convec[ny+2][nx+2]
# Every rank writes (reads) the x columns it owns straight into (from) the one file through a memory map: no funnel
# through rank 0, and a file written by N slabs restarts on M slabs or in the single-GPU programs.
# ----------------------------------------------------------------------------------------------------------
def checkpoint_layout(nx: int, ny: int, nz: int, dtype, convective: bool) -> tuple[dict, int]:
    """{field: (byte offset, shape, x ghost width)} of the global file and its total size"""
    s = np.dtype(dtype).itemsize
    off = 4 + 2 * s
    out = {}
    for name, shape, g in (("pdf", (2, 19, nz + 2, ny + 2, nx + 2), 1), ("phi", (nz + 8, ny + 8, nx + 8), 4)) + (
            (("f_convec", (19, ny + 2, nx + 2), 1), ("g_convec", (19, ny + 2, nx + 2), 1), ("phi_convec", (ny + 2, nx + 2), 1)) if convective else ()):
        out[name] = (off, shape, g)
        off += int(np.prod(shape)) * s
    return out, off


def _owned_columns(rng: SlabRange, g: int) -> tuple[slice, slice]:
    """(local, global) index slices along x of the columns `rng` owns in an array with g ghost columns: its real columns,
    plus the ghost columns of the lattice itself on a side without a neighbour"""
    lo = 1 if rng.has_left else 1 - g            # local x, inclusive
    hi = rng.nx_local if rng.has_right else rng.nx_local + g
    return slice(lo + g - 1, hi + g), slice(rng.x0 + lo + g - 2, rng.x0 + hi + g - 1)


def write_checkpoint_slabs(path, rng: SlabRange, nx_global: int, local: dict, ntime_next: int, force_z: float, rho_in: float, dist=None) -> None:
    """local: this slab's arrays in the reference layout of an (nx_local, ny, nz) lattice - what Solver.download_state returns
    after SlabStepper.settle(): "pdf" [2,19,nz+2,ny+2,nxl+2], "phi" [nz+8,ny+8,nxl+8], with a convective outlet also "f_convec",
    "g_convec" [19,ny+2,nxl+2] and "phi_convec" [ny+2,nxl+2] (flat arrays are reshaped).  Collective over `dist` when world > 1."""
    pdf = np.asarray(local["pdf"])
    dtype = pdf.dtype
    nxl = rng.nx_local
    nz, ny = pdf.shape[-3] - 2, pdf.shape[-2] - 2
    convective = local.get("f_convec") is not None
    layout, total = checkpoint_layout(nx_global, ny, nz, dtype, convective)
    multi = rng.world > 1 and dist is not None
    if rng.rank == 0:
        with open(path, "wb") as f:
            f.write(np.int32(ntime_next).tobytes())
            f.write(np.asarray([force_z, rho_in], dtype=dtype).tobytes())
            f.truncate(total)
    if multi:
        dist.barrier()
    mm = np.memmap(path, dtype=np.uint8, mode="r+")
    if mm.size != total:
        raise IOError(f"{path}: {mm.size} bytes, expected {total}")
    for name, (off, gshape, g) in layout.items():
        lshape = gshape[:-1] + (nxl + 2 * g,)
        a = np.asarray(local[name], dtype=dtype).reshape(lshape)
        dst = np.ndarray(gshape, dtype=dtype, buffer=mm, offset=off)
        ls, gs = _owned_columns(rng, g)
        dst[..., gs] = a[..., ls]
    mm.flush()
    del mm
    if multi:
        dist.barrier()


def read_checkpoint_slab(path, rng: SlabRange, nx_global: int, ny: int, nz: int, dtype, convective: bool) -> dict:
    """This slab's part of a checkpoint file, ghost columns included (1 for the PDFs and the convective buffers, 4 for phi: the
    neighbour's real columns, i.e. what the halo exchanges would have delivered), in the layout Solver.upload_state takes.
    Returns the arrays plus "ntime_next", "force_z", "rho_in".  Restart: upload_state(pdf, phi, convective buffers), then
    color_gradient() rebuilds cn_* / c_norm from phi as the reference does (src/main.cpp:112)."""
    layout, total = checkpoint_layout(nx_global, ny, nz, dtype, convective)
    mm = np.memmap(path, dtype=np.uint8, mode="r")
    if mm.size != total:
        raise IOError(f"{path}: {mm.size} bytes, expected {total} for a {nx_global} x {ny} x {nz} lattice")
    s = np.dtype(dtype).itemsize
    out = {"ntime_next": int(np.frombuffer(mm, np.int32, 1, 0)[0])}
    out["force_z"], out["rho_in"] = (float(v) for v in np.frombuffer(mm, dtype, 2, 4))
    for name, (off, gshape, g) in layout.items():
        src = np.ndarray(gshape, dtype=dtype, buffer=mm, offset=off)
        out[name] = np.ascontiguousarray(src[..., rng.x0 - 1:rng.x0 - 1 + rng.nx_local + 2 * g])   # global x0-g .. x1+g
    del mm
    return out


# ----------------------------------------------------------------------------------------------------------
# phase-field output of a decomposed lattice: the reference's legacy-VTK "small_" file (VTK_legacy_writer_3D type 2,
# /root/reference/src/IO_multiphase.cpp:500-716: STRUCTURED_POINTS, one big-endian float per interior node, phi zeroed in
# solids), every rank writing its own x columns of the one file.
# ----------------------------------------------------------------------------------------------------------
def vtk_header(nx: int, ny: int, nz: int, name: str = "phi", typ: str = "float") -> bytes:
    return (f"# vtk DataFile Version 3.0\nvtk output\nBINARY\nDATASET STRUCTURED_POINTS\nDIMENSIONS {nx} {ny} {nz}\n"
            f"ORIGIN 1 1 1\nSPACING 1 1 1\nPOINT_DATA {nx * ny * nz}\nSCALARS {name} {typ}\nLOOKUP_TABLE default\n").encode()


def write_vtk_phase_slabs(path, rng: SlabRange, nx_global: int, phi_local: np.ndarray, solid_local: np.ndarray, dist=None) -> None:
    """phi_local: this slab's phi in the 4-ghost reference layout [nz+8, ny+8, nxl+8] (Solver.download_state); solid_local: its
    interior wall flags [nz, ny, nxl] (non-zero = solid).  Collective over `dist` when world > 1."""
    nz, ny, nxl = solid_local.shape
    if phi_local.shape != (nz + 8, ny + 8, nxl + 8) or nxl != rng.nx_local:
        raise ValueError("phi_local / solid_local do not match the slab")
    head = vtk_header(nx_global, ny, nz)
    total = len(head) + 4 * nx_global * ny * nz
    multi = rng.world > 1 and dist is not None
    if rng.rank == 0:
        with open(path, "wb") as f:
            f.write(head)
            f.truncate(total)
    if multi:
        dist.barrier()
    mm = np.memmap(path, dtype=np.uint8, mode="r+")
    if mm.size != total:
        raise IOError(f"{path}: {mm.size} bytes, expected {total}")
    field = np.ndarray((nz, ny, nx_global), dtype=">f4", buffer=mm, offset=len(head))
    v = np.where(solid_local != 0, 0.0, phi_local[4:-4, 4:-4, 4:-4]).astype(np.float32)
    field[:, :, rng.x0 - 1:rng.x0 - 1 + nxl] = v
    mm.flush()
    del mm
    if multi:
        dist.barrier()

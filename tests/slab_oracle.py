"""Test infrastructure: a SlabBackend (see mflbm/slab.py) whose compute is the C oracle restricted to the slab's
x range, with numpy halo buffers.  Each rank holds a full-domain oracle context but only ever computes its own
columns; everything it does not own or receive through a halo exchange is poisoned with NaN, so a protocol error
(a column read before it was exchanged) shows up as NaN in the result."""
from __future__ import annotations

import numpy as np
import torch

EXP = (1, 7, 9, 11, 13)    # ex = +1
EXM = (2, 8, 10, 12, 14)   # ex = -1


class OracleSlab:
    def __init__(self, oracle, rng):
        self.o, self.r = oracle, rng
        o, r = oracle, rng
        self.pdf = o.arr("pdf")          # [2,19,nz+2,ny+2,nx+2]  index x+0 (ghost 1)
        self.phi = o.arr("phi")          # [nz+8,ny+8,nx+8]       index x+3
        self.ctl = o.control
        nan = np.nan
        # poison what this rank must never rely on
        gx = lambda x, g: x + g - 1      # global column x (1-based) -> array index with g ghosts
        self.pdf[..., :max(0, gx(r.x0 - 1, 1))] = nan
        self.pdf[..., gx(r.x1 + 1, 1) + 1:] = nan
        if r.has_left:
            self.phi[..., :gx(r.x0 - 4, 4)] = nan
        if r.has_right:
            self.phi[..., gx(r.x1 + 4, 4) + 1:] = nan
        for k in ("cn_x", "cn_y", "cn_z", "c_norm"):
            a = o.arr(k)
            if r.has_left:
                a[..., :gx(r.x0, 2)] = nan
            if r.has_right:
                a[..., gx(r.x1, 2) + 1:] = nan
        c = o.arr("curv")
        if r.has_left:
            c[..., :gx(r.x0, 1)] = nan
        if r.has_right:
            c[..., gx(r.x1, 1) + 1:] = nan
        npdf = 10 * (o.nz + 2) * (o.ny + 2)
        nphi = 4 * (o.nz + 8) * (o.ny + 8)
        dt = torch.float64 if o.prec == "f64" else torch.float32
        self.buf = {(k, s): (torch.zeros(nphi if k == 2 else npdf, dtype=dt), torch.zeros(nphi if k == 2 else npdf, dtype=dt))
                    for k in (0, 1, 2) for s in (0, 1)}

    # ---- compute phases (call order of src/main_iteration_GPU.cu:1890-2055, x restricted) ------------------
    def step_phase(self, ntime, phase):
        o, r, L, c = self.o, self.r, self.o.lib, self.ctl
        odd = int(ntime % 2 != 0)
        x0, x1 = r.x0, r.x1
        lo, hi = x0 - int(r.has_left), x1 + int(r.has_right)
        if phase == 0:
            L.orc_collide(o.h, odd, x0, x1)
        elif phase == 1:
            if c["kper"]:
                L.orc_periodic_pdf(o.h, 2, odd, lo, hi); L.orc_periodic_phi(o.h, 2, lo, hi)
            if c["jper"]:
                L.orc_periodic_pdf(o.h, 1, odd, lo, hi); L.orc_periodic_phi(o.h, 1, lo, hi)
            if c["jper"] and c["kper"]:
                L.orc_periodic_pdf_edges(o.h, odd, lo, hi); L.orc_periodic_phi(o.h, 3, lo, hi)
            if c["kper"] == 0 and c["domain_wall_status_z_min"] == 0 and c["domain_wall_status_z_max"] == 0:
                if c["inlet_BC"] == 1: L.orc_inlet_velocity(o.h, odd, x0, x1)
                elif c["inlet_BC"] == 2: L.orc_inlet_pressure(o.h, odd, x0, x1)
                if c["outlet_BC"] == 1: L.orc_outlet_convective(o.h, odd, x0, x1)
                elif c["outlet_BC"] == 2: L.orc_outlet_pressure(o.h, odd, x0, x1)
            if c["porous_plate_cmd"] != 0:
                L.orc_porous_plate(o.h, odd, x0, x1)
                if not odd:
                    # pass-through copies of the non-blocked component (:1770-1780) also on the ghost columns: they change
                    # real-plane values after the PDF halo was sent (same rule as k_porous_plate's extended x range)
                    zp, cmd, ny = c["Z_porous_plate"], c["porous_plate_cmd"], o.ny
                    if 1 <= zp <= o.nz and cmd in (1, 2):
                        gp = 1 if cmd == 1 else 0
                        for col in ([x0 - 1] if r.has_left else []) + ([x1 + 1] if r.has_right else []):
                            for q in (5, 11, 12, 15, 16):
                                oq = (0, 2, 1, 4, 3, 6, 5, 10, 9, 8, 7, 14, 13, 12, 11, 18, 17, 16, 15)[q]
                                self.pdf[gp, oq, zp, 1:ny + 1, col] = self.pdf[gp, oq, zp + 1, 1:ny + 1, col]
                                self.pdf[gp, q, zp, 1:ny + 1, col] = self.pdf[gp, q, zp - 1, 1:ny + 1, col]
        else:
            L.orc_extrapolate_phi_to_solid(o.h, x0 - 3, x1 + 3)
            L.orc_normal_directions(o.h, x0 - 2, x1 + 2)
            L.orc_alter_color_gradient(o.h, x0 - 2, x1 + 2)
            L.orc_extrapolate_normal_to_solid(o.h, x0 - 1, x1 + 1)
            L.orc_csf_curvature(o.h, x0, x1)

    # ---- halo buffers: same content as k_halo_pdf / k_halo_phi (mf-lbm-cuda_b200/csrc/kernels_aux.cuh) --------
    def halo_tensors(self, kind, side):
        return self.buf[(kind, side)]

    def _pdf_col(self, col, slots):
        """view list of the ten [nz+2, ny+2] planes of global column `col`"""
        return [self.pdf[g, q, :, :, col] for g in (0, 1) for q in slots]

    def _copy_pdf(self, buf, col, slots, pack):
        n = (self.o.nz + 2) * (self.o.ny + 2)
        for m, plane in enumerate(self._pdf_col(col, slots)):
            if pack:
                buf[m * n:(m + 1) * n] = torch.from_numpy(np.ascontiguousarray(plane).reshape(-1))
            else:
                plane[...] = buf[m * n:(m + 1) * n].numpy().reshape(plane.shape)

    def _copy_phi(self, buf, col0, pack):
        n = (self.o.nz + 8) * (self.o.ny + 8)
        for m in range(4):
            plane = self.phi[:, :, col0 + m + 3]
            if pack:
                buf[m * n:(m + 1) * n] = torch.from_numpy(np.ascontiguousarray(plane).reshape(-1))
            else:
                plane[...] = buf[m * n:(m + 1) * n].numpy().reshape(plane.shape)

    def halo_pack(self, kind):
        r = self.r
        if kind == 0:
            if r.has_left: self._copy_pdf(self.buf[(0, 0)][0], r.x0, EXM, True)
            if r.has_right: self._copy_pdf(self.buf[(0, 1)][0], r.x1, EXP, True)
        elif kind == 1:
            if r.has_left: self._copy_pdf(self.buf[(1, 0)][0], r.x0 - 1, EXP, True)
            if r.has_right: self._copy_pdf(self.buf[(1, 1)][0], r.x1 + 1, EXM, True)
        else:
            if r.has_left: self._copy_phi(self.buf[(2, 0)][0], r.x0, True)
            if r.has_right: self._copy_phi(self.buf[(2, 1)][0], r.x1 - 3, True)

    def halo_unpack(self, kind):
        r = self.r
        if kind == 0:
            if r.has_left: self._copy_pdf(self.buf[(0, 0)][1], r.x0 - 1, EXP, False)
            if r.has_right: self._copy_pdf(self.buf[(0, 1)][1], r.x1 + 1, EXM, False)
        elif kind == 1:
            if r.has_left: self._copy_pdf(self.buf[(1, 0)][1], r.x0, EXM, False)
            if r.has_right: self._copy_pdf(self.buf[(1, 1)][1], r.x1, EXP, False)
        else:
            if r.has_left: self._copy_phi(self.buf[(2, 0)][1], r.x0 - 4, False)
            if r.has_right: self._copy_phi(self.buf[(2, 1)][1], r.x1 + 1, False)

    # ---- owned part of the state, for gathering -----------------------------------------------------------
    def owned(self):
        o, r = self.o, self.r
        lo1 = 0 if not r.has_left else r.x0
        hi1 = o.nx + 1 if not r.has_right else r.x1
        sl = lambda g: slice(lo1 + g - 1 if r.has_left else 0, hi1 + g if r.has_right else None)
        return dict(pdf=self.pdf[..., sl(1)].copy(), phi=self.phi[..., sl(4)].copy(), cn_x=o.arr("cn_x")[..., sl(2)].copy(),
                    cn_y=o.arr("cn_y")[..., sl(2)].copy(), cn_z=o.arr("cn_z")[..., sl(2)].copy(), c_norm=o.arr("c_norm")[..., sl(2)].copy(),
                    curv=o.arr("curv")[..., sl(1)].copy())

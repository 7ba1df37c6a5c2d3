"""Device-side state of the mechanics hot path and thin wrappers over the C ABI.

``Engine`` owns every device buffer (as torch CUDA tensors: PyTorch is used for allocation,
stream handling and host<->device copies only) and fills the ``sic_problem_t`` the CUDA library
borrows raw pointers from.  No arithmetic of the hot path happens here.

Layouts (see include/safeincave_cuda.h): per-cell fields are SoA ``(rows, cell_stride)`` with
Voigt order [xx,yy,zz,xy,xz,yz]; nodal vectors are ``(n_nodes, 3)`` contiguous (dof = 3*node+c).
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np
import torch

from . import _lib as L


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


def tet_geometry(coords: torch.Tensor, cells: torch.Tensor):
    """Constant P1 shape-function gradients (N,4,3) and volumes (N,) of tetrahedra.
    vol = |det|/6 as in Grid.py:115-137; grad(lambda_1) = (e2 x e3)/det etc."""
    x0, x1, x2, x3 = (coords[cells[:, a]] for a in range(4))
    e1, e2, e3 = x1 - x0, x2 - x0, x3 - x0
    c23 = torch.linalg.cross(e2, e3)
    c31 = torch.linalg.cross(e3, e1)
    c12 = torch.linalg.cross(e1, e2)
    det = (e1 * c23).sum(dim=1)
    g1, g2, g3 = c23 / det[:, None], c31 / det[:, None], c12 / det[:, None]
    g0 = -(g1 + g2 + g3)
    grad = torch.stack([g0, g1, g2, g3], dim=1)
    return grad, det.abs() / 6.0


@dataclass
class ElemSpec:
    kind: int          # L.ELEM_*
    param_off: int     # offset inside a material-table row


class ElemState:
    """Device state of one non-elastic element (MaterialProps.py:567-574, 1089-1092)."""

    def __init__(self, kind, ns, device):
        z = lambda rows: torch.zeros((rows, ns), dtype=torch.float64, device=device)
        self.kind = kind
        self.eps_old, self.rate_old, self.rate, self.eps_k = z(6), z(6), z(6), z(6)
        # internal-state block (sic_elem_t.desai): Desai 21 rows, Munson-Dawson 20, Mohr-Coulomb / Matsuoka-Nakai 1
        self.desai = z(L.ISV_ROWS[kind]) if kind in L.ISV_ROWS else None
        if kind == L.ELEM_DESAI:
            self.desai[L.DS_H].fill_(1.0)
        elif kind == L.ELEM_MUNSON_DAWSON:            # MaterialProps.py:2052-2063
            self.desai[L.MD_H].fill_(1.0)
            self.desai[L.MD_F].fill_(1.0)
            self.desai[L.MD_ETS].fill_(1.0)

    SNAP = ("rate", "rate_old", "eps_old", "eps_k")
    # rows of the internal-state block that save_internal_state keeps (MomentumEquation.py:456-477)
    SNAP_ROWS = {L.ELEM_DESAI: [L.DS_ALPHA, L.DS_QSI, L.DS_QSI_OLD, L.DS_FVP],
                 L.ELEM_MUNSON_DAWSON: [L.MD_ZETA, L.MD_ZETA_OLD]}

    def snapshot(self):
        """save_internal_state, MomentumEquation.py:456-477 (alpha, qsi, qsi_old, Fvp for Desai; zeta, zeta_old
        for Munson-Dawson)."""
        snap = {k: getattr(self, k).clone() for k in self.SNAP}
        rows = self.SNAP_ROWS.get(self.kind)
        if rows:
            snap["desai"] = self.desai[rows].clone()
        return snap

    def restore(self, snap):
        for k in self.SNAP:
            getattr(self, k).copy_(snap[k])
        rows = self.SNAP_ROWS.get(self.kind)
        if rows:
            self.desai[rows] = snap["desai"]


class Engine:
    def __init__(self, coords, cells, device="cuda", operator_only=False, geometry_only=False):
        self.lib = L.load()
        if not torch.cuda.is_available():
            raise L.SicError("safeincave_b200 needs a CUDA device (no CPU fallback)")
        self._setup(coords, cells, device, operator_only, geometry_only)

    def _setup(self, coords, cells, device, operator_only=False, geometry_only=False):
        """Allocate and fill every buffer of sic_problem_t (torch index plumbing, once per mesh).
        operator_only: a coarse multigrid level -- mesh, scatter plan and C_T, no constitutive state.
        geometry_only: connectivity, gradients and volumes only (the heat equation's view of the mesh)."""
        self.device = torch.device(device)
        coords = torch.as_tensor(coords, dtype=torch.float64).to(self.device).contiguous()
        cells = torch.as_tensor(cells).to(self.device, dtype=torch.int64).contiguous()
        self.N, self.M = int(cells.shape[0]), int(coords.shape[0])
        self.ns = max(128, (self.N + 127) // 128 * 128)   # whole 128-cell tiles for the TMA-staged operator
        self.coords = coords
        grad, vol = tet_geometry(coords, cells)
        # orientation does not matter for P1 gradients; degenerate cells are a mesh error
        if self.N and not bool(torch.isfinite(grad).all()):
            raise ValueError("degenerate tetrahedron (zero volume) in mesh")
        ns, dev = self.ns, self.device
        self.conn = torch.zeros((4, ns), dtype=torch.int32, device=dev)
        self.conn[:, :self.N] = cells.t().to(torch.int32)
        self.grad = torch.zeros((12, ns), dtype=torch.float64, device=dev)
        self.grad[:, :self.N] = grad.reshape(self.N, 12).t()
        self.vol = torch.zeros(ns, dtype=torch.float64, device=dev)
        self.vol[:self.N] = vol
        self.launches = 0
        self.halo, self._halo_keep = None, None
        if geometry_only:
            self.operator_only = True
            return
        # the same geometry, tiled for the operator kernel (sic_geom_tile_t: grad[12][128], vol[128], conn[4][128])
        nt = ns // 128
        self.geom_tiles = torch.zeros((nt, 1920), dtype=torch.float64, device=dev)
        self.geom_tiles[:, :1536] = self.grad.reshape(12, nt, 128).permute(1, 0, 2).reshape(nt, 1536)
        self.geom_tiles[:, 1536:1664] = self.vol.reshape(nt, 128)
        self.geom_tiles[:, 1664:].view(torch.int32).copy_(self.conn.reshape(4, nt, 128).permute(1, 0, 2).reshape(nt, 512))
        self._build_scatter_plan()
        self.operator_only = bool(operator_only)
        z = lambda rows: None if operator_only else torch.zeros((rows, ns), dtype=torch.float64, device=dev)
        self.sig, self.sig_k, self.eps, self.eps_prev = z(6), z(6), z(6), z(6)
        self.eps_rhs = z(6)
        self.CT = torch.zeros((ns // 128, 36, 128), dtype=torch.float64, device=dev)    # tiled, SIC_CT_INDEX
        z1 = lambda: None if operator_only else torch.zeros(ns, dtype=torch.float64, device=dev)
        self.T, self.T0 = z1(), z1()                                  # MomentumEquation.py:116-117
        self.n_singular = torch.zeros(1, dtype=torch.int32, device=dev)
        self.n_clamped = torch.zeros(1, dtype=torch.int32, device=dev)
        self.err_out = torch.zeros(2, dtype=torch.float64, device=dev)
        self.err_scratch = torch.zeros(2 * max(1, self.lib.sic_post_blocks(self.N)), dtype=torch.float64, device=dev)
        self.mat_id = torch.zeros(ns, dtype=torch.int32, device=dev)
        self.mat_table = torch.zeros((1, 8), dtype=torch.float64, device=dev)
        self.spring_off, self.thermo_off, self.n_thermo = 0, 6, 0
        self.elems: list[ElemState] = []
        self.elem_specs: list[ElemSpec] = []
        self._ksp_work = {}
        self._prob = None
        self.launches = 0   # kernels launched through this engine (bench.py's gpu_launches)
        self.time_operator = False          # bracket one operator launch per Krylov batch with CUDA events
        self.profile = False                # bracket the constitutive kernels with CUDA events (bench.py)
        self._prof_events = {}
        self.halo = None                    # L.SicHalo when the mesh is partitioned over several GPUs
        self._halo_keep = None
        self.op_ms, self.op_samples, self.op_launches = 0.0, 0, 0

    def _build_scatter_plan(self):
        """Per tile of 128 cells: unique nodes (tile-interior first) and, per unique node, its (cell, slot)
        references inside the tile (sic_problem_t.tile_ptr / tile_nint / tile_nodes / ent_ptr / ent)."""
        dev, ns, M = self.device, self.ns, max(self.M, 1)
        nt = ns // 128
        conn = self.conn.long()                                   # (4, ns); padded cells reference node 0
        cell = torch.arange(ns, device=dev)
        tile = (cell // 128).repeat(4)                            # ref order: slot-major
        local = ((cell % 128) * 4).repeat(4) + torch.arange(4, device=dev).repeat_interleave(ns)
        node = conn.reshape(-1)
        total_refs = torch.bincount(node, minlength=M)            # references of each node in the whole mesh
        pair = tile * M + node
        upair, inv, cnt = torch.unique(pair, return_inverse=True, return_counts=True)
        utile, unode = upair // M, upair % M
        interior = cnt == total_refs[unode]
        # order the unique (tile,node) pairs: by tile, interior first, then node id
        key = (utile * 2 + (~interior).long()) * M + unode
        order = torch.argsort(key)
        rank = torch.empty_like(order)
        rank[order] = torch.arange(order.numel(), device=dev)
        new_inv = rank[inv]                                       # unique index of every reference, in the new order
        ref_order = torch.argsort(new_inv, stable=True)
        self.ent = local[ref_order].to(torch.int16).contiguous()  # values < 512 fit; read back as uint16
        cnt_o = cnt[order]
        self.ent_ptr = torch.zeros(order.numel() + 1, dtype=torch.int32, device=dev)
        self.ent_ptr[1:] = torch.cumsum(cnt_o, 0).to(torch.int32)
        self.tile_nodes = unode[order].to(torch.int32).contiguous()
        per_tile = torch.bincount(utile, minlength=nt)
        self.tile_ptr = torch.zeros(nt + 1, dtype=torch.int32, device=dev)
        self.tile_ptr[1:] = torch.cumsum(per_tile, 0).to(torch.int32)
        self.tile_nint = torch.bincount(utile[interior], minlength=nt).to(torch.int32).contiguous()
        # position of every cell node in its tile's list of unique nodes ([4][ns], < 512): the compressed multigrid
        # operator gathers x once per unique node of a tile and indexes the shared-memory copy with these
        self.lidx = (new_inv - self.tile_ptr.long()[tile]).to(torch.int16).reshape(4, ns).contiguous()

    # ------------------------------------------------------------------ material
    def set_material(self, table, mat_id, spring_off, thermo_off, n_thermo, elem_specs, keep_state=False, keep=None):
        table = torch.as_tensor(np.ascontiguousarray(table), dtype=torch.float64)
        self.mat_table = table.to(self.device).contiguous()
        mid = torch.as_tensor(mat_id).to(self.device, dtype=torch.int32)
        self.mat_id.zero_()
        self.mat_id[:self.N] = mid
        self.spring_off, self.thermo_off, self.n_thermo = int(spring_off), int(thermo_off), int(n_thermo)
        if len(elem_specs) > L.SIC_MAX_ELEMS:
            raise L.SicError(f"at most {L.SIC_MAX_ELEMS} non-elastic elements are supported")
        if keep is not None:
            # second set_material of a staged run (Simulators.py:1213-1326, nobian/Simulation/Run.py:1503-1506): the
            # elements that were already attached keep their device state, new ones start from zero
            self.elems = [self.elems[k] if (k is not None and k >= 0) else ElemState(s.kind, self.ns, self.device)
                          for k, s in zip(keep, elem_specs)]
        elif not keep_state or len(elem_specs) != len(self.elems):
            self.elems = [ElemState(s.kind, self.ns, self.device) for s in elem_specs]
        self.elem_specs = list(elem_specs)
        self._prob = None

    def problem(self) -> L.SicProblem:
        if self._prob is not None:
            return self._prob
        P = L.SicProblem()
        P.abi_version = L.SIC_ABI_VERSION
        P.n_cells, P.cell_stride, P.n_nodes = self.N, self.ns, self.M
        P.conn, P.grad, P.vol = _ptr(self.conn), _ptr(self.grad), _ptr(self.vol)
        P.geom_tiles = _ptr(self.geom_tiles)
        P.tile_ptr, P.tile_nint, P.tile_nodes = _ptr(self.tile_ptr), _ptr(self.tile_nint), _ptr(self.tile_nodes)
        P.ent_ptr, P.ent = _ptr(self.ent_ptr), _ptr(self.ent)
        P.mat_id, P.mat_table = _ptr(self.mat_id), _ptr(self.mat_table)
        P.n_rows, P.row_len = int(self.mat_table.shape[0]), int(self.mat_table.shape[1])
        P.spring_off, P.n_thermo, P.thermo_off = self.spring_off, self.n_thermo, self.thermo_off
        P.n_elems = len(self.elems)
        for i, (st, sp) in enumerate(zip(self.elems, self.elem_specs)):
            e = P.elems[i]
            e.kind, e.param_off = sp.kind, sp.param_off
            e.eps_old, e.rate_old, e.rate, e.eps_k = _ptr(st.eps_old), _ptr(st.rate_old), _ptr(st.rate), _ptr(st.eps_k)
            e.desai = _ptr(st.desai)
        P.T, P.T0 = _ptr(self.T), _ptr(self.T0)
        P.sig, P.sig_k, P.eps, P.eps_prev = _ptr(self.sig), _ptr(self.sig_k), _ptr(self.eps), _ptr(self.eps_prev)
        P.CT, P.eps_rhs = _ptr(self.CT), _ptr(self.eps_rhs)
        P.n_singular = _ptr(self.n_singular)
        self._prob = P
        return P

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _pp(self):
        return ctypes.byref(self.problem())

    # ------------------------------------------------------------------ measurement
    def _tic(self, name):
        if not self.profile:
            return None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        self._prof_events.setdefault(name, []).append((e0, e1))
        return e1

    @staticmethod
    def _toc(e1):
        if e1 is not None:
            e1.record()

    def profile_summary(self, reset=True):
        """{kernel: (launches, total ms)} from the recorded CUDA events (synchronises)."""
        torch.cuda.synchronize(self.device)
        out = {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in self._prof_events.items()}
        if reset:
            self._prof_events = {}
        return out

    # ------------------------------------------------------------------ part (1)
    def tangent(self, dt, theta):
        t = self._tic("tangent")
        L.check(self.lib.sic_tangent(self._pp(), float(dt), float(theta), self._stream()), "sic_tangent")
        self._toc(t)
        self.launches += 1

    def elastic_tangent(self):
        L.check(self.lib.sic_elastic_tangent(self._pp(), self._stream()), "sic_elastic_tangent")
        self.launches += 1

    def post(self, u, dt, theta, kelvin_phi2, flags):
        up = _ptr(u) if u is not None else ctypes.c_void_p(0)
        t = self._tic("post")
        L.check(self.lib.sic_post(self._pp(), up, float(dt), float(theta), float(kelvin_phi2), int(flags),
                                  _ptr(self.err_out), _ptr(self.err_scratch), self._stream()), "sic_post")
        self._toc(t)
        self.launches += 2 if (flags & L.POST_ERROR) else 1

    def commit(self, dt, theta):
        t = self._tic("commit")
        L.check(self.lib.sic_commit(self._pp(), float(dt), float(theta), self._stream()), "sic_commit")
        self._toc(t)
        self.launches += 1

    def commit_rates(self):
        L.check(self.lib.sic_commit_rates(self._pp(), self._stream()), "sic_commit_rates")
        self.launches += 1

    def desai_initial_hardening(self, elem, Fvp_0):
        self.n_clamped.zero_()
        L.check(self.lib.sic_desai_initial_hardening(self._pp(), int(elem), float(Fvp_0), _ptr(self.n_clamped),
                                                     self._stream()), "sic_desai_initial_hardening")
        self.launches += 1
        return int(self.n_clamped.item())

    # ------------------------------------------------------------------ part (2)
    def apply(self, x, y, fixed):
        L.check(self.lib.sic_apply(self._pp(), _ptr(x), _ptr(y), _ptr(fixed), self._stream()), "sic_apply")
        self.launches += 2

    def _ph(self):
        return ctypes.byref(self.halo) if self.halo is not None else None

    def residual0(self, b_ext, x0, r, fixed):
        L.check(self.lib.sic_residual0(self._pp(), _ptr(b_ext), _ptr(x0), _ptr(r), _ptr(fixed), self._ph(),
                                       self._stream()), "sic_residual0")
        self.launches += 2

    def block_jacobi(self, dinv, fixed):
        t = self._tic("block_jacobi")
        L.check(self.lib.sic_block_jacobi(self._pp(), _ptr(dinv), _ptr(fixed), self._ph(), self._stream()),
                "sic_block_jacobi")
        self._toc(t)
        self.launches += 2

    # ------------------------------------------------------------------ several GPUs
    def set_partition(self, part, comm, p2p=None):
        """Attach the halo plan of safeincave_b200.partition.Partition and an NCCL communicator handle."""
        if part.n_ranks <= 1:
            self.halo = None
            return
        if len(part.peers) > L.SIC_MAX_PEERS:
            raise L.SicError("too many neighbouring ranks")
        dev = self.device
        idx = torch.cat([s.to(dev) for s in part.shared]).to(torch.int32).contiguous()
        owner_w = part.owner_w.to(dev, dtype=torch.float64).contiguous()
        total = int(idx.numel())
        send = torch.zeros(9 * max(total, 1), dtype=torch.float64, device=dev)
        recv = torch.zeros(9 * max(total, 1), dtype=torch.float64, device=dev)
        h = L.SicHalo()
        h.n_ranks, h.rank, h.n_peers, h.n_shared_total = part.n_ranks, part.rank, len(part.peers), total
        off = 0
        for i, (peer, sh) in enumerate(zip(part.peers, part.shared)):
            h.peer[i] = int(peer)
            h.peer_off[i] = off
            off += int(sh.numel())
        h.peer_off[len(part.peers)] = off
        h.idx, h.owner_w, h.send_buf, h.recv_buf = _ptr(idx), _ptr(owner_w), _ptr(send), _ptr(recv)
        h.comm = comm
        h.p2p = p2p
        tile_order = None
        if p2p is not None and hasattr(self, "tile_ptr") and total > 0:
            # launch order of the operator's tiles for the fused operator + exchange kernel: tiles that touch an
            # interface node first (their sums are what the neighbours wait for), interior tiles after them
            iface = torch.zeros(max(self.M, 1), dtype=torch.bool, device=dev)
            iface[idx.long()] = True
            touches = iface[self.conn.long()].any(dim=0)                       # (ns,) cells with an interface node
            touches[self.N:] = False
            tile_touch = touches.reshape(-1, 128).any(dim=1)
            tile_order = torch.cat([torch.nonzero(tile_touch).flatten(), torch.nonzero(~tile_touch).flatten()]).to(torch.int32).contiguous()
            h.tile_order, h.n_iface_tiles = _ptr(tile_order), int(tile_touch.sum().item())
        self.halo = h
        self._halo_keep = (idx, owner_w, send, recv, tile_order)
        self.owner_w = owner_w

    def halo_sum(self, vec, ncomp):
        if self.halo is None:
            return
        L.check(self.lib.sic_exchange(self._ph(), _ptr(vec), int(ncomp), None, 0, self._stream()), "sic_exchange")
        self.launches += 2

    def neumann(self, tri, area_n, bc_of_tri, bc_par, b):
        n_tri = int(tri.shape[1])
        n_bc = int(bc_par.shape[0])
        L.check(self.lib.sic_neumann(n_tri, _ptr(tri), _ptr(area_n), _ptr(bc_of_tri), _ptr(self.coords), n_bc,
                                     _ptr(bc_par), _ptr(b), self._stream()), "sic_neumann")
        self.launches += 1

    # ------------------------------------------------------------------ part (3)
    def ksp_solve(self, method, b_ext, x, fixed, dinv, rtol=1e-10, atol=0.0, max_it=10000, check_every=25,
                  guess_nonzero=False):
        key = method
        need = int(self.lib.sic_ksp_workspace_doubles(self.M, method))
        w = self._ksp_work.get(key)
        if w is None or w.numel() < need:
            w = torch.zeros(need, dtype=torch.float64, device=self.device)
            self._ksp_work = {key: w}     # one workspace at a time
        ksp = L.SicKsp()
        ksp.method, ksp.max_it, ksp.rtol, ksp.atol = int(method), int(max_it), float(rtol), float(atol)
        ksp.check_every, ksp.use_graph = int(check_every), 0
        ksp.time_operator = 1 if self.time_operator else 0
        ksp.guess_nonzero = 1 if guess_nonzero else 0
        L.check(self.lib.sic_ksp_solve(self._pp(), ctypes.byref(ksp), _ptr(b_ext), _ptr(x), _ptr(fixed), _ptr(dinv),
                                       _ptr(w), self._ph(), self._stream()), "sic_ksp_solve")
        per_it = {L.KSP_CG: 4, L.KSP_CGCG: 3}.get(method, 7)
        self.launches += 3 + per_it * int(ksp.iterations)
        self.op_ms += float(ksp.op_ms)
        self.op_samples += int(ksp.op_samples)
        self.op_launches += (2 if method == L.KSP_BICGSTAB else 1) * int(ksp.iterations)
        return ksp

    def guess_extrapolate(self, iterates, x, want_coef=False):
        """x <- prediction of the next Newton iterate from ``iterates`` = [u0 (latest), u1, u2(, u3)]
        (sic_guess_extrapolate).  Returns (a, b, terms, misfit1, misfit2) if want_coef (synchronises) else None."""
        n = len(iterates)
        need = int(self.lib.sic_guess_workspace_doubles(self.M))
        w = getattr(self, "_guess_work", None)
        if w is None or w.numel() < need:
            w = self._guess_work = torch.zeros(need, dtype=torch.float64, device=self.device)
        coef = (ctypes.c_double * 5)() if want_coef else None
        L.check(self.lib.sic_guess_extrapolate(self.M, n, _ptr(iterates[0]), _ptr(iterates[1]), _ptr(iterates[2]),
                                               _ptr(iterates[3]) if n > 3 else None, _ptr(x), self._ph(), _ptr(w),
                                               coef, self._stream()), "sic_guess_extrapolate")
        self.launches += 2 + (2 if self.halo is not None else 0)
        return (coef[0], coef[1], int(coef[2]), coef[3], coef[4]) if want_coef else None

    def fp64_peak(self):
        out = ctypes.c_double(0.0)
        L.check(self.lib.sic_fp64_peak(ctypes.byref(out), self._stream()), "sic_fp64_peak")
        return out.value

    # ------------------------------------------------------------------ host <-> device helpers
    def get_CT(self):
        """(N,6,6) host numpy copy of the tiled tangent."""
        return self.CT.permute(0, 2, 1).reshape(self.ns, 36)[:self.N].reshape(self.N, 6, 6).cpu().numpy()

    def put_CT(self, ct):
        """Upload an (N,6,6) tangent into the tiled device layout."""
        t = torch.zeros((self.ns, 36), dtype=torch.float64, device=self.device)
        t[:self.N] = torch.as_tensor(np.ascontiguousarray(ct), dtype=torch.float64).reshape(self.N, 36).to(self.device)
        self.CT.copy_(t.reshape(self.ns // 128, 128, 36).permute(0, 2, 1))

    def put6(self, dst, v6):
        """(N,6) host/any array -> SoA (6, ns) device tensor."""
        t = torch.as_tensor(np.ascontiguousarray(v6), dtype=torch.float64).to(self.device)
        dst[:, :self.N] = t.t()

    def get6(self, src):
        return src[:, :self.N].t().contiguous().cpu().numpy()

    def put1(self, dst, v):
        dst[:self.N] = torch.as_tensor(np.ascontiguousarray(v), dtype=torch.float64).to(self.device)

    def get1(self, src):
        return src[:self.N].cpu().numpy()

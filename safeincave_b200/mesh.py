"""Tetrahedral mesh ingest and services for the hot path (host side, numpy/torch plumbing).

* ``read_msh``     Gmsh MSH 2.2 and 4.1 ASCII reader written from the file-format description
                   (the reference goes through dolfinx.io.gmshio + meshio, Grid.py:275-313; neither
                   exists here).  Keeps tetrahedra, tagged triangles and the physical-name table.
* ``TetMesh``      coordinates, cells, cell tags, ORIENTED boundary triangles (right-hand normal
                   points out of the domain) and their tags.
* ``red_refine``   Bey's regular refinement (each tet -> 8, each boundary triangle -> 4, tags and
                   orientation inherited) used to build the 10M-80M cell synthetic cavern meshes of
                   BASELINE config 5 without gmsh.
* ``morton_order`` space-filling-curve renumbering of cells and nodes (gather locality on the GPU
                   and contiguous chunks for the multi-GPU partition).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np
import torch


@dataclass
class TetMesh:
    coords: np.ndarray                 # (M,3) float64
    cells: np.ndarray                  # (N,4) int64
    cell_tags: np.ndarray              # (N,) int64 physical id of the volume
    tris: np.ndarray                   # (F,3) int64 boundary triangles, outward oriented
    tri_tags: np.ndarray               # (F,) int64 physical id of the surface
    names: dict = field(default_factory=lambda: {1: {}, 2: {}, 3: {}})   # dim -> {name: tag}

    @property
    def n_cells(self):
        return int(self.cells.shape[0])

    @property
    def n_nodes(self):
        return int(self.coords.shape[0])

    def save_npz(self, path):
        names = np.array([f"{d}|{n}|{t}" for d, m in self.names.items() for n, t in m.items()])
        np.savez_compressed(path, coords=self.coords, cells=self.cells.astype(np.int32),
                            cell_tags=self.cell_tags.astype(np.int32), tris=self.tris.astype(np.int32),
                            tri_tags=self.tri_tags.astype(np.int32), names=names)

    @staticmethod
    def load_npz(path):
        z = np.load(path, allow_pickle=False)
        names = {1: {}, 2: {}, 3: {}}
        for s in z["names"]:
            d, n, t = str(s).split("|")
            names[int(d)][n] = int(t)
        return TetMesh(z["coords"].astype(np.float64), z["cells"].astype(np.int64), z["cell_tags"].astype(np.int64),
                       z["tris"].astype(np.int64), z["tri_tags"].astype(np.int64), names)


# ----------------------------------------------------------------------------------------------
# MSH reader
# ----------------------------------------------------------------------------------------------
def _sections(path):
    with open(path, "r") as f:
        lines = f.read().split("\n")
    sec, i = {}, 0
    while i < len(lines):
        ln = lines[i].strip()
        if ln.startswith("$") and not ln.startswith("$End"):
            name, j = ln[1:], i + 1
            while j < len(lines) and lines[j].strip() != "$End" + name:
                j += 1
            sec[name] = lines[i + 1:j]
            i = j
        i += 1
    return sec


def _physical_names(sec):
    names = {0: {}, 1: {}, 2: {}, 3: {}}
    for ln in sec.get("PhysicalNames", [])[1:]:
        parts = ln.split(None, 2)
        if len(parts) == 3:
            names[int(parts[0])][parts[2].strip().strip('"')] = int(parts[1])
    return {d: names[d] for d in (1, 2, 3)}


def _read_v2(sec):
    body = sec["Nodes"]
    n = int(body[0])
    arr = np.array([ln.split() for ln in body[1:1 + n]], dtype=np.float64)
    node_ids, xyz = arr[:, 0].astype(np.int64), arr[:, 1:4]
    tets, tet_tags, tris, tri_tags = [], [], [], []
    body = sec["Elements"]
    for ln in body[1:1 + int(body[0])]:
        p = ln.split()
        etype, ntags = int(p[1]), int(p[2])
        phys = int(p[3]) if ntags > 0 else 0
        nodes = p[3 + ntags:]
        if etype == 4:
            tets.append(nodes[:4]); tet_tags.append(phys)
        elif etype == 2:
            tris.append(nodes[:3]); tri_tags.append(phys)
    return node_ids, xyz, np.array(tets, dtype=np.int64).reshape(-1, 4), np.array(tet_tags, dtype=np.int64), \
        np.array(tris, dtype=np.int64).reshape(-1, 3), np.array(tri_tags, dtype=np.int64)


def _read_v4(sec):
    ent = sec["Entities"]
    npnt, ncur, nsur, nvol = (int(v) for v in ent[0].split())
    phys = {2: {}, 3: {}}
    k = 1 + npnt + ncur
    for dim, cnt in ((2, nsur), (3, nvol)):
        for ln in ent[k:k + cnt]:
            p = ln.split()
            nphys = int(p[7])
            phys[dim][int(p[0])] = int(p[8]) if nphys > 0 else 0
        k += cnt
    body = sec["Nodes"]
    nblocks, nnodes = (int(v) for v in body[0].split()[:2])
    node_ids = np.empty(nnodes, dtype=np.int64)
    xyz = np.empty((nnodes, 3))
    k, filled = 1, 0
    for _ in range(nblocks):
        nb = int(body[k].split()[3])
        k += 1
        node_ids[filled:filled + nb] = [int(v) for v in body[k:k + nb]]
        k += nb
        xyz[filled:filled + nb] = [[float(v) for v in ln.split()[:3]] for ln in body[k:k + nb]]
        k += nb
        filled += nb
    body = sec["Elements"]
    nblocks = int(body[0].split()[0])
    tets, tet_tags, tris, tri_tags = [], [], [], []
    k = 1
    for _ in range(nblocks):
        edim, etag, etype, nb = (int(v) for v in body[k].split())
        k += 1
        if etype == 4:
            a = np.array([ln.split()[1:5] for ln in body[k:k + nb]], dtype=np.int64)
            tets.append(a); tet_tags.append(np.full(nb, phys[3].get(etag, 0)))
        elif etype == 2:
            a = np.array([ln.split()[1:4] for ln in body[k:k + nb]], dtype=np.int64)
            tris.append(a); tri_tags.append(np.full(nb, phys[2].get(etag, 0)))
        k += nb
    cat = lambda lst, w: np.concatenate(lst) if lst else np.zeros((0, w), dtype=np.int64)
    cat1 = lambda lst: np.concatenate(lst).astype(np.int64) if lst else np.zeros(0, dtype=np.int64)
    return node_ids, xyz, cat(tets, 4), cat1(tet_tags), cat(tris, 3), cat1(tri_tags)


def read_msh(path) -> TetMesh:
    """Read a Gmsh .msh (format 2.2 or 4.1, ASCII).  Nodes not used by a tetrahedron are dropped
    and the rest renumbered 0..M-1 in file order; untagged triangles are dropped."""
    sec = _sections(path)
    version = float(sec["MeshFormat"][0].split()[0])
    if int(sec["MeshFormat"][0].split()[1]) != 0:
        raise ValueError(f"{path}: binary MSH files are not supported")
    node_ids, xyz, tets, tet_tags, tris, tri_tags = (_read_v4 if version >= 4 else _read_v2)(sec)
    names = _physical_names(sec)
    lut = np.full(int(node_ids.max()) + 1, -1, dtype=np.int64)
    used = np.unique(tets)
    lut[used] = np.arange(used.size)
    order = np.argsort(node_ids)
    pos = order[np.searchsorted(node_ids[order], used)]
    coords = xyz[pos]
    keep = tri_tags > 0
    tris, tri_tags = tris[keep], tri_tags[keep]
    mesh = TetMesh(coords, lut[tets], tet_tags, lut[tris], tri_tags, names)
    if (mesh.tris < 0).any():
        raise ValueError(f"{path}: a tagged triangle uses a node that no tetrahedron uses")
    orient_boundary(mesh)
    return mesh


def orient_boundary(mesh: TetMesh):
    """Order the nodes of every boundary triangle so that its right-hand normal points out of the
    adjacent tetrahedron (the outward FacetNormal of MomentumEquation.py:252-253)."""
    if mesh.tris.shape[0] == 0:
        return
    c = mesh.cells
    faces = np.concatenate([c[:, [1, 2, 3]], c[:, [0, 2, 3]], c[:, [0, 1, 3]], c[:, [0, 1, 2]]])
    opp = np.concatenate([c[:, 0], c[:, 1], c[:, 2], c[:, 3]])
    fkey = np.sort(faces, axis=1)
    tkey = np.sort(mesh.tris, axis=1)
    M = mesh.n_nodes
    pack = lambda k: (k[:, 0] * M + k[:, 1]) * M + k[:, 2] if M < 2_000_000 else None
    if pack(fkey) is None:
        raise ValueError("orient_boundary: mesh too large for the packed face key; refine a smaller mesh instead")
    fk, tk = pack(fkey), pack(tkey)
    order = np.argsort(fk, kind="stable")
    fks = fk[order]
    lo = np.searchsorted(fks, tk, side="left")
    hi = np.searchsorted(fks, tk, side="right")
    if (hi - lo < 1).any():
        raise ValueError("a tagged triangle is not a face of any tetrahedron")
    interior = hi - lo > 1
    owner = order[lo]
    x = mesh.coords
    a, b, d = (x[mesh.tris[:, k]] for k in range(3))
    nrm = np.cross(b - a, d - a)
    inward = ((x[opp[owner]] - a) * nrm).sum(axis=1) > 0
    flip = inward & ~interior
    mesh.tris[flip] = mesh.tris[flip][:, [0, 2, 1]]
    mesh.tri_interior = interior


# ----------------------------------------------------------------------------------------------
# refinement and ordering (torch: runs on the GPU for the 10M+ cell meshes)
# ----------------------------------------------------------------------------------------------
def red_refine(mesh: TetMesh, device="cpu", return_edges=False):
    """One level of Bey's red refinement.  Conforming; children of a parent are contiguous (8p..8p+7), the
    coarse nodes keep their ids and the midpoint of the k-th edge (sorted by (lo, hi)) gets id M + k.
    return_edges: also return the (E,2) end nodes of those edges (the multigrid transfer tables need them)."""
    dev = torch.device(device)
    cells = torch.as_tensor(mesh.cells, device=dev)
    tris = torch.as_tensor(mesh.tris, device=dev)
    coords = torch.as_tensor(mesh.coords, device=dev)
    M = mesh.n_nodes
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    ekeys = []
    for a, b in pairs:
        lo = torch.minimum(cells[:, a], cells[:, b])
        hi = torch.maximum(cells[:, a], cells[:, b])
        ekeys.append(lo * M + hi)
    ekeys = torch.stack(ekeys, dim=1)                     # (N,6)
    uniq, inv = torch.unique(ekeys.reshape(-1), return_inverse=True)
    mid = (M + inv).reshape(-1, 6)                        # midpoint node id of each cell edge
    lo, hi = uniq // M, uniq % M
    new_coords = torch.cat([coords, 0.5 * (coords[lo] + coords[hi])])
    v0, v1, v2, v3 = (cells[:, k] for k in range(4))
    m01, m02, m03, m12, m13, m23 = (mid[:, k] for k in range(6))
    kids = [(v0, m01, m02, m03), (m01, v1, m12, m13), (m02, m12, v2, m23), (m03, m13, m23, v3),
            (m01, m02, m03, m13), (m01, m02, m12, m13), (m02, m03, m13, m23), (m02, m12, m13, m23)]
    new_cells = torch.stack([torch.stack(k, dim=1) for k in kids], dim=1).reshape(-1, 4)
    new_ctags = torch.as_tensor(mesh.cell_tags, device=dev).repeat_interleave(8)
    if tris.shape[0]:
        def emid(a, b):
            key = torch.minimum(a, b) * M + torch.maximum(a, b)
            return M + torch.searchsorted(uniq, key)
        a, b, c = tris[:, 0], tris[:, 1], tris[:, 2]
        ab, bc, ac = emid(a, b), emid(b, c), emid(a, c)
        tk = [(a, ab, ac), (ab, b, bc), (ac, bc, c), (ab, bc, ac)]
        new_tris = torch.stack([torch.stack(k, dim=1) for k in tk], dim=1).reshape(-1, 3)
        new_ttags = torch.as_tensor(mesh.tri_tags, device=dev).repeat_interleave(4)
    else:
        new_tris, new_ttags = tris, torch.as_tensor(mesh.tri_tags, device=dev)
    fine = TetMesh(new_coords.cpu().numpy(), new_cells.cpu().numpy(), new_ctags.cpu().numpy(),
                   new_tris.cpu().numpy(), new_ttags.cpu().numpy(), mesh.names)
    if return_edges:
        return fine, torch.stack([lo, hi], dim=1).cpu().numpy()
    return fine


def _spread3(v):
    v = v & 0x1FFFFF
    v = (v | (v << 32)) & 0x1F00000000FFFF
    v = (v | (v << 16)) & 0x1F0000FF0000FF
    v = (v | (v << 8)) & 0x100F00F00F00F00F
    v = (v | (v << 4)) & 0x10C30C30C30C30C3
    v = (v | (v << 2)) & 0x1249249249249249
    return v


def morton_keys(points: torch.Tensor) -> torch.Tensor:
    lo = points.min(dim=0).values
    span = (points.max(dim=0).values - lo).clamp_min(1e-300)
    q = ((points - lo) / span * (2 ** 21 - 1)).to(torch.int64)
    return _spread3(q[:, 0]) | (_spread3(q[:, 1]) << 1) | (_spread3(q[:, 2]) << 2)


def morton_order(mesh: TetMesh, device="cpu", keep_cell_order=False) -> TetMesh:
    """Renumber nodes and cells along a Morton curve (nodes by position, cells by centroid).
    keep_cell_order: renumber the nodes only (a red-refined mesh lists the eight children of every parent cell
    together, in the parents' order: the NESTED order a multi-GPU multigrid hierarchy is partitioned in)."""
    dev = torch.device(device)
    coords = torch.as_tensor(mesh.coords, device=dev)
    cells = torch.as_tensor(mesh.cells, device=dev)
    nperm = torch.argsort(morton_keys(coords))
    ninv = torch.empty_like(nperm)
    ninv[nperm] = torch.arange(nperm.numel(), device=dev)
    cells = ninv[cells]
    coords = coords[nperm]
    if keep_cell_order:
        cperm = torch.arange(cells.shape[0], device=dev)
    else:
        cperm = torch.argsort(morton_keys(coords[cells].mean(dim=1)))
    tris = ninv[torch.as_tensor(mesh.tris, device=dev)] if mesh.tris.shape[0] else torch.as_tensor(mesh.tris)
    out = TetMesh(coords.cpu().numpy(), cells[cperm].cpu().numpy(), mesh.cell_tags[cperm.cpu().numpy()],
                  tris.cpu().numpy(), mesh.tri_tags.copy(), mesh.names)
    out.cell_perm = cperm.cpu().numpy()     # new cell i was old cell cell_perm[i]
    out.node_perm = nperm.cpu().numpy()
    return out


def tri_area_normals(mesh: TetMesh) -> np.ndarray:
    """(F,3) outward normal times area of the oriented boundary triangles."""
    x = mesh.coords
    a, b, c = (x[mesh.tris[:, k]] for k in range(3))
    return 0.5 * np.cross(b - a, c - a)

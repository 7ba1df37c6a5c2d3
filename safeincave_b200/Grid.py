"""Grid handler with the interface of the reference's ``GridHandlerGMSH`` (safeincave/Grid.py:27-579).

Reads Gmsh MSH 2.2 / 4.1 ASCII with this package's own reader (the reference delegates to
dolfinx.io.gmshio and meshio, Grid.py:275-313) and exposes the attributes user scripts touch:
``mesh.geometry.x``, ``mesh.topology.connectivity(3,0).array``, ``mesh.comm``, ``n_elems``,
``n_nodes``, ``region_indices``, ``region_names``, ``get_boundary_names`` / ``get_boundary_tag(s)``,
``Lx/Ly/Lz``, ``volumes``, ``get_parameter``, ``A_csr`` / ``B_csr`` / ``smoother`` (lazy).

Cell and node numbering is this package's own (T11: the reference's is whatever DOLFINx chose; user
tensors are indexed through ``region_indices``, so any self-consistent numbering is compatible).
"""
from __future__ import annotations

import os

import numpy as np
import torch as to

from .mesh import TetMesh, morton_order, read_msh


class _Comm:
    """Stand-in for mpi4py's communicator: rank / size and an allreduce over torch.distributed."""

    @property
    def rank(self):
        return to.distributed.get_rank() if to.distributed.is_available() and to.distributed.is_initialized() else 0

    @property
    def size(self):
        return to.distributed.get_world_size() if to.distributed.is_available() and to.distributed.is_initialized() else 1

    def Get_rank(self):
        return self.rank

    def Get_size(self):
        return self.size

    def allreduce(self, value, op=None):
        return value

    def Barrier(self):
        pass


class _Conn:
    def __init__(self, arr):
        self.array = arr


class _Topology:
    dim = 3

    def __init__(self, cells):
        self._cells = cells

    def connectivity(self, d0, d1):
        if (d0, d1) != (3, 0):
            raise NotImplementedError("only the cell->vertex connectivity is available")
        return _Conn(self._cells.reshape(-1))


class _Geometry:
    dim = 3

    def __init__(self, x):
        self.x = x


class _MeshTags:
    """Stand-in for dolfinx.mesh.MeshTags (what get_boundaries / get_subdomains return, Grid.py:392-412): entity
    indices and their gmsh physical tags, in THIS package's entity numbering."""

    def __init__(self, dim, indices, values):
        self.dim = dim
        self.indices = np.asarray(indices, dtype=np.int32)
        self.values = np.asarray(values, dtype=np.int32)

    def find(self, tag):
        return self.indices[self.values == tag]


class _Mesh:
    def __init__(self, coords, cells):
        self.geometry = _Geometry(coords)
        self.topology = _Topology(cells)
        self.comm = _Comm()


class GridHandlerGMSH:
    def __init__(self, geometry_name=None, grid_folder=None, tetmesh: TetMesh | None = None, reorder=True):
        self.grid_folder, self.geometry_name = grid_folder, geometry_name
        if tetmesh is None:
            tetmesh = read_msh(os.path.join(grid_folder, f"{geometry_name}.msh"))
        if reorder and tetmesh.n_cells > 0 and not hasattr(tetmesh, "cell_perm"):
            tetmesh = morton_order(tetmesh)
        self.tetmesh = tetmesh
        self.comm = _Comm()
        self.rank = self.comm.rank
        self.mesh = _Mesh(tetmesh.coords, tetmesh.cells)
        self.domain_dim, self.boundary_dim = 3, 2
        self.n_elems, self.n_nodes = tetmesh.n_cells, tetmesh.n_nodes
        self.dolfin_tags = {d: dict(tetmesh.names.get(d, {})) for d in (1, 2, 3)}
        self.tags = self.dolfin_tags
        x = tetmesh.coords
        self.Lx, self.Ly, self.Lz = (float(x[:, k].max() - x[:, k].min()) for k in range(3))
        # Grid.py:496-536
        self.region_names = self.get_subdomain_names()
        self.n_regions = len(self.region_names)
        self.tags_dict = {tag: name for name, tag in self.dolfin_tags[3].items()}
        self.region_indices = {name: np.nonzero(tetmesh.cell_tags == tag)[0].tolist()
                               for name, tag in self.dolfin_tags[3].items()}
        self.subdomain_tags = {name: [] for name in self.region_names}
        # Grid.py:359-368: boundary name -> facet (triangle) indices
        self.boundary_tags = {name: np.nonzero(tetmesh.tri_tags == tag)[0].tolist()
                              for name, tag in self.dolfin_tags[2].items()}
        self._smoother = None
        self._volumes = None

    @classmethod
    def from_mesh(cls, tetmesh: TetMesh, reorder=True):
        return cls(tetmesh=tetmesh, reorder=reorder)

    @classmethod
    def from_hierarchy(cls, hierarchy):
        """Grid on the finest level of a ``multigrid.Hierarchy``; the coarser levels stay attached
        (``grid.hierarchy``) for the multigrid preconditioner (PC type ``mg``)."""
        grid = cls(tetmesh=hierarchy.finest, reorder=False)
        grid.hierarchy = hierarchy
        return grid

    def refine(self, levels: int, nested=True, device="cpu"):
        """A NEW grid on this grid's mesh regularly (red / Bey) refined ``levels`` times, with this mesh as the coarsest
        level of the attached hierarchy: PC type ``mg`` (the stand-in for the reference's ``gamg``,
        nobian/Simulation/run_interlayer.py:2114-2116) then works for ANY mesh loaded from a .msh file, not only for the
        bench's synthetic ones.  Region / boundary tags are inherited by the children; per-cell material tensors must be
        built for the refined grid (``refined.get_parameter`` / ``region_indices`` follow the new numbering)."""
        from .multigrid import refine_hierarchy
        if levels < 1:
            raise ValueError("refine: levels must be >= 1")
        grid = type(self).from_hierarchy(refine_hierarchy(self.tetmesh, int(levels), device=device, nested=nested))
        grid.grid_folder, grid.geometry_name = self.grid_folder, self.geometry_name
        return grid

    # --- tag queries (Grid.py:392-494)
    def get_boundaries(self):
        return _MeshTags(2, np.arange(self.tetmesh.tris.shape[0]), self.tetmesh.tri_tags)

    def get_subdomains(self):
        return _MeshTags(3, np.arange(self.tetmesh.n_cells), self.tetmesh.cell_tags)

    def get_boundary_names(self):
        return list(self.dolfin_tags[2].keys())

    def get_boundary_tags(self, name):
        return None if name is None else self.boundary_tags[name]

    def get_boundary_tag(self, name):
        return None if name is None else self.dolfin_tags[2][name]

    def get_subdomain_names(self):
        return list(self.dolfin_tags[3].keys())

    def get_subdomain_tag(self, name):
        return self.dolfin_tags[3][name]

    # --- Grid.py:538-579 (NB: produces float32 tensors for scalars / region lists, T1)
    def get_parameter(self, param):
        if type(param) == int or type(param) == float:
            return to.tensor([param for _ in range(self.n_elems)])
        elif len(param) == self.n_regions:
            out = to.zeros(self.n_elems)
            for i, region in enumerate(self.region_indices.keys()):
                out[self.region_indices[region]] = param[i]
            return out
        elif len(param) == self.n_elems:
            return param if type(param) == to.Tensor else to.tensor(param)
        raise Exception("Size of parameter list does not match neither # of elements nor # of regions.")

    # --- Grid.py:139-242 (built lazily: output-only)
    @property
    def volumes(self):
        if self._volumes is None:
            x = self.tetmesh.coords[self.tetmesh.cells]
            e = x[:, 1:] - x[:, :1]
            self._volumes = np.abs(np.einsum("ni,ni->n", e[:, 0], np.cross(e[:, 1], e[:, 2]))) / 6.0
        return self._volumes

    def build_smoother(self):
        import scipy.sparse as sp
        N, M, cells, vol = self.n_elems, self.n_nodes, self.tetmesh.cells, self.volumes
        rows, cols = cells.ravel(), np.repeat(np.arange(N), 4)
        A = sp.csr_matrix((np.repeat(vol, 4), (rows, cols)), shape=(M, N))
        self.A_csr = sp.diags(1.0 / np.asarray(A.sum(axis=1)).ravel()) @ A
        self.B_csr = sp.csr_matrix((np.full(4 * N, 0.25), (cols, rows)), shape=(N, M))
        self.smoother = self.B_csr @ self.A_csr
        self._smoother = True

    def __getattr__(self, name):
        if name == "hierarchy":
            return None
        if name in ("A_csr", "B_csr", "smoother"):
            self.build_smoother()
            return self.__dict__[name]
        raise AttributeError(name)

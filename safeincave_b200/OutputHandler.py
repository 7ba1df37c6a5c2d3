"""``SaveFields`` with the interface of the reference's safeincave/OutputHandler.py:27-169 (``set_output_folder``,
``add_output_field``, ``initialize``, ``save_fields(t)``, ``save_mesh``).

The reference writes one XDMF/HDF5 time series per field through dolfinx.io.XDMFFile; h5py / dolfinx are not
part of this stack, so the same layout is written as XDMF 3 with the heavy data in raw little-endian binary
files next to it (``Format="Binary"``; ParaView reads it, and so does ``numpy.fromfile``):

    <output_folder>/<field>/<field>.xdmf          temporal collection, one grid per saved time
    <output_folder>/<field>/mesh_geometry.bin     (M,3) float64     mesh_topology.bin  (N,4) int32
    <output_folder>/<field>/<field>_<k>.bin       values of the k-th saved time, float64

Fields are looked up on the equation by name, like the reference does (``getattr(eq, field_name)``): nodal vectors
(``u``), nodal scalars (``p_nodes``, ``q_nodes``, ``T``), cell scalars (``p_elems``, ``q_elems``) and cell tensors
(``sig``, ``eps_tot``: 9 components, row-major 3x3).  Output happens outside the timed hot path."""
from __future__ import annotations

import os
import shutil

import numpy as np


class SaveFields:
    def __init__(self, eq):
        self.eq = eq
        self.fields_data = []
        self.output_fields = []

    def set_output_folder(self, output_folder: str) -> None:
        self.output_folder = output_folder

    def add_output_field(self, field_name: str, label_name: str) -> None:
        self.fields_data.append({"field_name": field_name, "label_name": label_name})

    # ------------------------------------------------------------------ helpers
    def _values(self, field_name):
        """(array, center, attribute type) of a field of the equation."""
        eq = self.eq
        tm = eq.grid.tetmesh
        N, M = tm.n_cells, tm.n_nodes
        f = getattr(eq, field_name)
        if hasattr(f, "to_tensor"):                       # CellField: symmetric tensor per cell
            a = f.to_tensor().numpy().reshape(N, 9)
        elif hasattr(f, "x"):                             # dolfinx-like Function view
            a = np.asarray(f.x.array, dtype=np.float64)
        elif hasattr(f, "detach"):                        # torch tensor (p_elems, q_nodes, ...)
            a = f.detach().cpu().numpy().astype(np.float64)
        else:
            a = np.asarray(f, dtype=np.float64)
        a = a.reshape(-1)
        for rows, center in ((M, "Node"), (N, "Cell")):
            if rows and a.size % rows == 0 and a.size // rows in (1, 3, 9):
                comp = a.size // rows
                return a.reshape(rows, comp), center, {1: "Scalar", 3: "Vector", 9: "Tensor"}[comp]
        raise ValueError(f"field '{field_name}' has {a.size} values: neither nodal nor cell data of this mesh")

    def initialize(self) -> None:
        """OutputHandler.py:116-131: one file per field, mesh written once."""
        tm = self.eq.grid.tetmesh
        self.output_fields = []
        for fd in self.fields_data:
            folder = os.path.join(self.output_folder, fd["field_name"])
            os.makedirs(folder, exist_ok=True)
            np.ascontiguousarray(tm.coords, dtype="<f8").tofile(os.path.join(folder, "mesh_geometry.bin"))
            np.ascontiguousarray(tm.cells, dtype="<i4").tofile(os.path.join(folder, "mesh_topology.bin"))
            self.output_fields.append({"folder": folder, "name": fd["field_name"], "label": fd["label_name"], "steps": []})
            self._write_xdmf(self.output_fields[-1])

    def save_fields(self, t: float) -> None:
        """OutputHandler.py:133-151."""
        if self.eq.grid.mesh.comm.rank != 0 and self.eq.grid.mesh.comm.size > 1:
            return
        for out in self.output_fields:
            a, center, kind = self._values(out["name"])
            k = len(out["steps"])
            fname = f"{out['name']}_{k:06d}.bin"
            np.ascontiguousarray(a, dtype="<f8").tofile(os.path.join(out["folder"], fname))
            out["steps"].append((float(t), fname, a.shape, center, kind))
            self._write_xdmf(out)

    def _write_xdmf(self, out):
        tm = self.eq.grid.tetmesh
        N, M = tm.n_cells, tm.n_nodes
        item = lambda dims, dtype, prec, path: (f'<DataItem Dimensions="{dims}" NumberType="{dtype}" Precision="{prec}" '
                                                f'Format="Binary" Endian="Little">{path}</DataItem>')
        lines = ['<?xml version="1.0"?>', '<Xdmf Version="3.0">', ' <Domain>',
                 '  <Grid Name="TimeSeries" GridType="Collection" CollectionType="Temporal">']
        for t, fname, shape, center, kind in out["steps"]:
            lines += [f'   <Grid Name="mesh" GridType="Uniform">', f'    <Time Value="{t!r}"/>',
                      f'    <Topology TopologyType="Tetrahedron" NumberOfElements="{N}">',
                      '     ' + item(f"{N} 4", "Int", 4, "mesh_topology.bin"), '    </Topology>',
                      '    <Geometry GeometryType="XYZ">', '     ' + item(f"{M} 3", "Float", 8, "mesh_geometry.bin"),
                      '    </Geometry>',
                      f'    <Attribute Name="{out["label"]}" AttributeType="{kind}" Center="{center}">',
                      '     ' + item(f"{shape[0]} {shape[1]}", "Float", 8, fname), '    </Attribute>', '   </Grid>']
        lines += ['  </Grid>', ' </Domain>', '</Xdmf>']
        with open(os.path.join(out["folder"], f"{out['name']}.xdmf"), "w") as f:
            f.write("\n".join(lines) + "\n")

    def save_mesh(self) -> None:
        """OutputHandler.py:153-169: copy the .msh next to the results (when the grid came from a file)."""
        grid = self.eq.grid
        dest = os.path.join(self.output_folder, "mesh")
        os.makedirs(dest, exist_ok=True)
        if getattr(grid, "grid_folder", None) and getattr(grid, "geometry_name", None):
            src = os.path.join(grid.grid_folder, f"{grid.geometry_name}.msh")
            if os.path.isfile(src):
                shutil.copy(src, dest)
                return
        grid.tetmesh.save_npz(os.path.join(dest, "mesh.npz"))

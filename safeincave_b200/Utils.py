"""Helpers with the names of the reference's safeincave/Utils.py that user scripts import.

Units (Utils.py:34-40), JSON io (:42-81), field samplers (:285-342, vectorised here) and
``dotdot_torch`` (:251-283).  The UFL helpers (epsilon, dotdot_ufl, tensor2voigt, voigt2tensor,
project) have no counterpart: what those forms compute is done by the CUDA kernels.
"""
import json

import numpy as np
import torch as to

GPa = 1e9
MPa = 1e6
kPa = 1e3
minute = 60
hour = 60 * minute
day = 24 * hour
year = 365 * day


def read_json(file_name):
    with open(file_name, "r") as f:
        return json.load(f)


def save_json(data, file_name):
    with open(file_name, "w") as f:
        json.dump(data, f, indent=4)


def numpy2torch(numpy_array):
    return to.tensor(np.asarray(numpy_array), dtype=to.float64)


def dotdot_torch(C_voigt, eps_tensor):
    """sigma = C : eps in tensorial Voigt form (host utility, same result layout as Utils.py:251-283)."""
    idx = ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))
    ev = to.stack([eps_tensor[:, i, j] for i, j in idx], dim=1).to(to.float64)
    sv = to.bmm(C_voigt.to(to.float64), ev.unsqueeze(2)).squeeze(2)
    out = to.zeros_like(eps_tensor, dtype=to.float64)
    for k, (i, j) in enumerate(idx):
        out[:, i, j] = sv[:, k]
        out[:, j, i] = sv[:, k]
    return out


def create_field_nodes(grid, fun):
    """fun(x, y, z) sampled at the mesh nodes (Utils.py:285-311)."""
    x = np.asarray(grid.mesh.geometry.x)
    try:
        v = fun(x[:, 0], x[:, 1], x[:, 2])
        v = np.broadcast_to(np.asarray(v, dtype=np.float64), (x.shape[0],))
    except Exception:
        v = np.array([fun(*p) for p in x], dtype=np.float64)
    return to.tensor(np.array(v), dtype=to.float64)


def create_field_elems(grid, fun):
    """fun(x, y, z) sampled at the cell centroids (Utils.py:313-342)."""
    x = np.asarray(grid.mesh.geometry.x)
    conn = np.asarray(grid.mesh.topology.connectivity(3, 0).array).reshape(grid.n_elems, 4)
    c = x[conn].sum(axis=1) / 4
    try:
        v = fun(c[:, 0], c[:, 1], c[:, 2])
        v = np.broadcast_to(np.asarray(v, dtype=np.float64), (c.shape[0],))
    except Exception:
        v = np.array([fun(*p) for p in c], dtype=np.float64)
    return to.tensor(np.array(v), dtype=to.float64)

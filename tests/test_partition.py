"""The N>1 path on CPU: world_size-2 and -3 gloo jobs check that the cell partition + halo plan
reproduce the single-domain operator, loads, constrained-dof sets and dot products exactly the way
the CUDA library uses them (operator partial sums -> halo sum; owner-weighted dots -> allreduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import constitutive as oc
from oracle import fem
from safeincave_b200.mesh import TetMesh, morton_order, red_refine
from safeincave_b200.partition import build_partition, halo_sum_reference

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _mesh():
    tm = TetMesh.load_npz(os.path.join(GOLD, "mesh_cube_coarse.npz"))
    return morton_order(red_refine(tm))


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        tm = _mesh()
        N, M = tm.n_cells, tm.n_nodes
        rng = np.random.default_rng(7)
        CT = oc.iso_matrix(102e9 * (1 + 0.3 * rng.random(N)), 0.3 * np.ones(N)) + 1e9 * rng.standard_normal((N, 6, 6))
        x = rng.standard_normal(3 * M)
        y_ref = fem.assemble_K(tm.coords, tm.cells, CT) @ x
        part = build_partition(tm.cells, M, rank, world)
        loc = part.local_mesh(tm)
        ln = part.local_nodes.numpy()
        c0, c1 = part.cell_range
        # operator: local partial sums + halo sum == global K x on the local nodes
        K_loc = fem.assemble_K(loc.coords, loc.cells, CT[c0:c1])
        x_loc = x.reshape(M, 3)[ln].reshape(-1)
        y_loc = torch.tensor((K_loc @ x_loc).reshape(-1, 3))
        halo_sum_reference(part, y_loc)
        err_op = np.abs(y_loc.numpy() - y_ref.reshape(M, 3)[ln]).max() / np.abs(y_ref).max()
        # owner-weighted dot == global dot
        w = part.owner_w.numpy()
        d = torch.tensor([(w[:, None] * x.reshape(M, 3)[ln] * y_ref.reshape(M, 3)[ln]).sum()])
        dist.all_reduce(d)
        err_dot = abs(d.item() - x @ y_ref) / abs(x @ y_ref)
        # cell-wise energy needs no weights: sum_e x_e^T K_e x_e over the rank's cells
        e = torch.tensor([x_loc @ (K_loc @ x_loc)])
        dist.all_reduce(e)
        err_energy = abs(e.item() - x @ y_ref) / abs(x @ y_ref)
        # Neumann load: triangles are assigned to exactly one rank, halo sum completes interface nodes
        bc = dict(tag=tm.names[2]["TOP"], direction=2, density=1000.0, ref_pos=1.0, gravity=-9.81,
                  values=[5e6, 5e6], time_values=[0, 1])
        b_ref = fem.neumann_load(tm.coords, tm.tris, tm.tri_tags, [bc], 0.0).reshape(M, 3)
        b_loc = torch.tensor(fem.neumann_load(loc.coords, loc.tris, loc.tri_tags, [bc], 0.0).reshape(-1, 3))
        halo_sum_reference(part, b_loc)
        err_neu = np.abs(b_loc.numpy() - b_ref[ln]).max() / np.abs(b_ref).max()
        n_tris = torch.tensor([loc.tris.shape[0]])
        dist.all_reduce(n_tris)
        # Dirichlet node sets come from the global triangles
        tag = tm.names[2]["WEST"]
        glob = np.unique(tm.tris[tm.tri_tags == tag])
        want = np.nonzero(np.isin(ln, glob))[0]
        ok_dir = np.array_equal(np.sort(loc.boundary_nodes[tag]), want)
        # every cell on exactly one rank; owners partition the nodes
        n_owned = torch.tensor([w.sum()])
        dist.all_reduce(n_owned)
        out.put((rank, err_op, err_dot, err_energy, err_neu, int(n_tris.item()) == tm.tris.shape[0], ok_dir,
                 int(n_owned.item()) == M, len(part.peers)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_partition_halo_plan_gloo(world):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = [out.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, err_op, err_dot, err_energy, err_neu, tris_ok, dir_ok, owned_ok, n_peers in res:
        assert err_op < 1e-13 and err_dot < 1e-13 and err_energy < 1e-13 and err_neu < 1e-13
        assert tris_ok and dir_ok and owned_ok
        assert n_peers >= 1


def test_partition_single_rank_is_identity():
    tm = _mesh()
    part = build_partition(tm.cells, tm.n_nodes, 0, 1)
    assert part.peers == [] and bool((part.owner_w == 1).all())
    assert torch.equal(part.cells_local, torch.as_tensor(tm.cells))
    loc = part.local_mesh(tm)
    assert loc.tris.shape == tm.tris.shape

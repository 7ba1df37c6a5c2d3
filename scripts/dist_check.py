"""torchrun --nproc-per-node N scripts/dist_check.py [--levels L] [--ksp cg] [--pc jacobi|mg] [--min-cells-per-rank K]
Runs the cavern case partitioned over N GPUs and (on every rank, on its own GPU) the same case on the whole mesh;
compares displacement / stress / creep strain on the rank's local nodes and cells.

--pc mg: multigrid CG with the bench's solver settings (extrapolated guess, lagged setup) on a NESTED hierarchy: every
level with at least --min-cells-per-rank cells per rank is partitioned (csrc/mg.cu), the ones below are replicated.
--heat: Simulator_TM instead of Simulator_M (HeatDiffusion on the partitioned grid, csrc/heat.cu; thermal strain and
thermally activated creep in the momentum equation): temperatures compared as well."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import safeincave_b200 as sf  # noqa: E402
from safeincave_b200 import cases, distributed  # noqa: E402
from safeincave_b200.mesh import TetMesh  # noqa: E402
from safeincave_b200.multigrid import refine_hierarchy  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--levels", type=int, default=1)
ap.add_argument("--ksp", default="cg")
ap.add_argument("--pc", default="jacobi", choices=["jacobi", "mg"])
ap.add_argument("--min-cells-per-rank", type=int, default=200_000)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--rtol", type=float, default=1e-12)
ap.add_argument("--tol", type=float, default=1e-8)
ap.add_argument("--heat", action="store_true")
ap.add_argument("--fused-exchange", action="store_true", help="V-cycle operator + halo exchange as one launch (k_mg_ebe_pc_x)")
a = ap.parse_args()

ctx = distributed.init()
dev = ctx.device
from safeincave_b200 import _lib  # noqa: E402
_lib.load().sic_mg_set_fused_exchange(1 if a.fused_exchange else 0)
h = refine_hierarchy(TetMesh.load_npz(os.path.join(ROOT, "tests/golden/mesh_cavern_regular.npz")), a.levels, device=dev,
                     nested=True)
tm = h.finest
gg = sf.GridHandlerGMSH.from_hierarchy(h)
case = cases.cavern_case(gg, n_steps=a.steps, ksp_type=a.ksp, rtol=a.rtol)


def settings(eq):
    eq.solver.single_reduction = os.environ.get("SIC_CGCG", "0") == "1"
    if a.pc == "mg":
        eq.solver.getPC().setType("mg")
        eq.solver.initial_guess_nonzero = True
        eq.solver.guess_extrapolation = True
        eq.solver.mg_setup_first = 2


heat = heat1 = None
if a.heat:
    from tests import heat_checks as H
    hc = H.heat_case(gg, "cavern_regular")
    case["dt"], case["t_final_run"], case["thermo_alpha"] = 20 * H.DAY, a.steps * 20 * H.DAY, 44e-6


def thermo(eq, sim_m, grid, part):
    """Simulator_TM around the momentum equation of `sim_m` (examples/thermomechanics/2_cavern/main.py's set-up)."""
    hcl = dict(hc, T0=hc["T0"][part.local_nodes.cpu().numpy()] if part is not None else hc["T0"])
    ht = H.build_heat(sf, grid, hcl)
    eq.mat.add_to_thermoelastic(sf.Thermoelastic(case["thermo_alpha"] * torch.ones(grid.n_elems, dtype=torch.float64)))
    eq.set_material(eq.mat)
    return ht, sf.Simulator_TM(eq, ht, sim_m.t_control, [], compute_elastic_response=True, verbose=False)


grid, part = distributed.partition_grid(ctx, tm, hierarchy=h if a.pc == "mg" else None,
                                        min_cells_per_rank=a.min_cells_per_rank)
eq, sim = cases.build(case, grid, device=dev, part=part, ctx=ctx)
settings(eq)
if a.heat:
    heat, sim = thermo(eq, sim, grid, part)
sim.verbose = False
hist = sim.run()
lc = eq.mg.lc if eq.mg is not None else None
level_cells = [e.N for e in eq.mg.engines] if eq.mg is not None else None
# single-domain run of the same case on this rank's GPU
eq1, sim1 = cases.build(case, gg, device=dev)
settings(eq1)
if a.heat:
    heat1, sim1 = thermo(eq1, sim1, gg, None)
sim1.verbose = False
hist1 = sim1.run()
ln = part.local_nodes.to(dev)
c0, c1 = part.cell_range
rel = lambda x, y: float((x - y).abs().max() / y.abs().max())
e_u = rel(eq.X, eq1.X[ln])
e_s = rel(eq.engine.sig[:, :eq.engine.N], eq1.engine.sig[:, c0:c1])
e_c = rel(eq.engine.elems[0].eps_old[:, :eq.engine.N], eq1.engine.elems[0].eps_old[:, c0:c1])
its = [r["iterations"] for r in hist], [r["iterations"] for r in hist1]
print(f"rank {ctx.rank}/{ctx.world}: cells {eq.engine.N} nodes {eq.engine.M} peers {part.peers} interface {part.n_interface} "
      f"| pc {a.pc} distributed from level {lc} level cells {level_cells} fused operator+exchange launches "
      f"{int(eq.engine.lib.sic_mg_fused_exchange_launches())} "
      f"| u err {e_u:.2e} sig err {e_s:.2e} eps_cr err {e_c:.2e} | newton {its} ksp {[r['ksp_iterations'] for r in hist]} "
      f"vs {[r['ksp_iterations'] for r in hist1]}", flush=True)
assert e_u < a.tol and e_s < a.tol and e_c < a.tol and its[0] == its[1]
if a.heat:
    T0 = torch.as_tensor(hc["T0"], device=dev)
    e_T = rel(heat.T_dev[:heat.n_nodes] - T0[ln], heat1.T_dev[ln] - T0[ln])
    print(f"rank {ctx.rank}: temperature change err {e_T:.2e}, max |dT| {float((heat1.T_dev - T0).abs().max()):.2f} K, "
          f"heat iterations {[k[0] for k in heat.ksp_log]} vs {[k[0] for k in heat1.ksp_log]}", flush=True)
    assert e_T < a.tol
ctx.barrier()
if ctx.rank == 0:
    print("DIST CHECK OK")

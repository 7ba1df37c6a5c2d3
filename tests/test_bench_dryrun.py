"""bench.py's GPU arm executed end to end on the HOST (tests/hostemu engine, stubbed CUDA events / pinned memory) on a
small stand-in workload: the JSON line is assembled by code that otherwise only ever runs on the B200 box at round end,
so a typo there would cost the round's headline number.  Checks the line's contract (keys, units, roofline and e2e
objects), both preconditioner branches and the validity gate; the numbers themselves mean nothing here."""
import json
import os
import sys
import types

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def synchronize(self):
        pass

    def elapsed_time(self, other):
        return 1.0


@pytest.fixture()
def dry(monkeypatch):
    import bench
    import safeincave_b200 as sf
    from safeincave_b200 import cases, distributed
    from safeincave_b200.engine import Engine
    from safeincave_b200.mesh import TetMesh
    from tests.hostemu import EmuEngine

    class DryEngine(EmuEngine):            # keep the product's event-based profiling hooks (stub events)
        _tic = Engine._tic

        def fp64_peak(self):               # the DFMA micro-benchmark would take minutes on the emulator
            return 1.0e12

    monkeypatch.setattr(sf.LinearMomentum, "engine_cls", DryEngine)
    monkeypatch.setattr(distributed, "init", lambda device=None: distributed.DistContext(0, 1, torch.device("cpu")))
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a, **k: None)
    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)
    # stand-in workload: the 48-cell cube with a body force (non-uniform stress, 5-6 Newton iterations per step)
    load = TetMesh.load_npz
    monkeypatch.setattr(TetMesh, "load_npz", staticmethod(lambda path: load(os.path.join(GOLD, "mesh_cube_coarse.npz"))))

    def case(grid, n_steps=None, ksp_type="cg", rtol=1e-10, **kw):
        c = cases.triaxial_case(grid, n_steps=n_steps, ksp_override=ksp_type)
        c["g"] = [0.0, 0.0, -5e3]
        c["ksp"]["rtol"] = rtol
        return c
    monkeypatch.setattr(cases, "cavern_case", case)
    return bench


def run(bench, capsys, *flags, cpu_baseline=False):
    argv = ["bench.py", "--steps", "2", "--warmup", "1"] + ([] if cpu_baseline else ["--no-cpu-baseline"]) + list(flags)
    old = sys.argv
    sys.argv = argv
    try:
        bench.run_b200(bench.parse())
    finally:
        sys.argv = old
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    return json.loads(lines[0])


def check_contract(line, pc):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "clocks", "gpu_launches", "roofline", "e2e"):
        assert key in line, key
    assert line["metric"] == "cell_updates_per_s" and line["unit"] == "cell-updates/s" and line["dtype"] == "f64"
    assert line["n_gpus"] == 1 and line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True
    assert line["vs_baseline"] is None and line["data"] == "synthetic" and "workload" in line["config"]
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["gpu_launches"] > 0
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and r["peak"] > 0 and r["achieved"] > 0
    assert r["frac"] == pytest.approx(r["achieved"] / r["peak"])
    e = line["e2e"]
    assert e["value"] > 0 and e["unit"] == line["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0
    assert line["config"]["newton_iterations"] >= 8 and line["config"]["krylov_iterations"] > 0
    assert "extrapolated" in line["config"]["warm_start"]
    assert ("multigrid" in line["config"]["preconditioner"]) == (pc == "mg")


def test_bench_line_multigrid(dry, capsys):
    line = run(dry, capsys, "--pc", "mg", "--levels", "2")
    check_contract(line, "mg")
    mg = line["constitutive"]["mg"]
    assert mg["setup_lag"] == 2 and 0 < mg["setups"] < mg["solves"] and len(mg["lambda_max"]) == 3


def test_bench_line_cpu_arm_and_parity_check(dry, capsys):
    """The CPU leg of the GPU line (oracle/cpu_step.py on the same mesh family) and the parity check of the timed
    solver settings against its fields."""
    line = run(dry, capsys, "--pc", "mg", "--levels", "2", "--cpu-levels", "1", "--cpu-threads", "2", cpu_baseline=True)
    check_contract(line, "mg")
    cb, pc = line["cpu_baseline"], line["parity_check"]
    assert cb["kind"] == "port" and cb["cores"] == 2 and cb["value"] > 0 and cb["n_cells"] == 384 and "PCG" in cb["sample"]
    assert pc["u_max_rel"] < 1e-8 and pc["sigma_max_rel"] < 1e-8
    assert pc["newton_iterations_gpu"] == pc["newton_iterations_cpu"] and "mg_lag=2" in pc["solver_settings"]


def test_reference_arm_line(dry, capsys, monkeypatch):
    monkeypatch.setattr(sys, "argv", ["bench.py", "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-threads", "2"])
    dry.run_reference(dry.parse())
    line = json.loads([ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")][-1])
    assert line["impl"] == "reference" and line["metric"] == "cell_updates_per_s" and line["value"] > 0
    assert line["cpu_baseline"]["cores"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the line names the GPU arm's workload (default --levels 4) and says which member of the family the sample was
    assert line["config"]["n_cells"] == 14346 * 8 ** 4 and "x8^4" in line["config"]["workload"]
    assert line["config"]["sample_n_cells"] == 384 and line["cpu_baseline_unrefined_lu"]["cores"] == 1


def test_bench_line_block_jacobi(dry, capsys):
    line = run(dry, capsys, "--pc", "jacobi", "--levels", "1", "--warm-start", "1")
    assert "previous Newton iterate" in line["config"]["warm_start"]
    line["config"]["warm_start"] = "extrapolated"      # the rest of the contract is the same
    check_contract(line, "jacobi")


def test_bench_refuses_a_run_that_did_not_converge(dry, capsys):
    with pytest.raises(RuntimeError, match="timed steps invalid"):
        run(dry, capsys, "--pc", "jacobi", "--levels", "1", "--max-it", "3")


def test_smoke_entry_point_on_the_emulator(monkeypatch, capsys):
    """__graft_entry__.smoke() (two steps of config 1 vs the oracle, then the same on a 2-level multigrid) through the
    host-emulated engine: the entry point's own Python must not be what fails on the GPU box."""
    import __graft_entry__ as g
    import safeincave_b200 as sf
    from tests.hostemu import EmuEngine
    monkeypatch.setattr(sf.LinearMomentum, "engine_cls", EmuEngine)
    g.smoke()
    out = capsys.readouterr().out
    assert "smoke: u err" in out and "smoke (multigrid CG, 2 levels)" in out


def test_bench_line_two_emulated_ranks(dry, capsys, monkeypatch):
    """The N > 1 branch of run_b200 (collective in-process multigrid probe, Morton-chunk partition, finest level
    distributed / coarse levels replicated, max-over-ranks timing, rank 0 prints) on two emulated ranks."""
    import threading

    from safeincave_b200 import distributed
    from tests.hostemu.ranks import run_ranks
    local = threading.local()
    monkeypatch.setattr(distributed, "init", lambda device=None: local.ctx)
    monkeypatch.setattr(sys, "argv", ["bench.py", "--gpus", "2", "--steps", "1", "--warmup", "1", "--levels", "2",
                                      "--no-cpu-baseline"])
    # the probe's stand-in mesh is the cube as well (TetMesh.load_npz is redirected by the fixture)

    def body(ctx):
        local.ctx = ctx
        dry.run_b200(dry.parse())
        return ctx.rank

    assert sorted(run_ranks(2, body)) == [0, 1]
    lines = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                   # rank 0 alone prints
    line = json.loads(lines[0])
    assert line["n_gpus"] == 2 and line["scaling"] == "strong" and "cpu_baseline" not in line
    assert line["config"]["cells_per_gpu"] * 2 == line["config"]["n_cells"]
    assert "MG_PROBE_OK" in line["config"]["pc_choice"] and "multigrid" in line["config"]["preconditioner"]
    assert "partitioned by the cells. ancestors" in line["config"]["partition"].replace("'", ".")


def test_single_gpu_multigrid_probe(dry, capsys):
    """bench.probe_mg (what `--pc auto` runs in a child process on one GPU) with the timed run's solver settings."""
    assert dry.probe_mg(levels=2, device="cpu", warm_start=2, mg_lag=2) == 0
    assert "MG_PROBE_OK" in capsys.readouterr().out

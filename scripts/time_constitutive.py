"""Times the constitutive kernels of the bench workload in isolation (CUDA events, warm state):
    python scripts/time_constitutive.py [--levels 3] [--tag name]
Prints one line: average ms of k_tangent / k_post / k_commit and their algorithmic GB/s.  Used with
scripts/build_variants.py to A/B compile-time variants of constitutive.cu on the GPU box."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--levels", type=int, default=3)
    ap.add_argument("--tag", default="")
    ap.add_argument("--reps", type=int, default=5)
    a = ap.parse_args()
    import torch
    import safeincave_b200 as sf
    from safeincave_b200 import cases
    from safeincave_b200.mesh import TetMesh
    from safeincave_b200.multigrid import refine_hierarchy
    dev = torch.device("cuda:0")
    h = refine_hierarchy(TetMesh.load_npz(os.path.join(ROOT, "tests", "golden", "mesh_cavern_regular.npz")), a.levels, device=dev,
                         nested=True)
    grid = sf.GridHandlerGMSH.from_hierarchy(h)
    case = cases.cavern_case(grid, n_steps=2, ksp_type="cg", rtol=1e-10)
    eq, sim = cases.build(case, grid, device=dev)
    eq.solver.getPC().setType("mg")
    eq.solver.setGuessExtrapolation(True)
    sim.verbose = False
    sim.initialize()
    sim.step()
    eng = eq.engine
    dt = case["dt"]
    N, M, S = eng.N, eng.M, len(eng.elems)

    def timed(fn):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
        fn()
        torch.cuda.synchronize()
        for e0, e1 in ev:
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        return sum(e0.elapsed_time(e1) for e0, e1 in ev) / a.reps
    t_tan = timed(lambda: eng.tangent(dt, eq.theta))
    t_post = timed(lambda: eq.newton_post(dt))
    t_commit = timed(lambda: eng.commit(dt, eq.theta))
    tan_bytes = N * 8 * (6 + 2 + 42 + 24 * S) + 4 * N
    post_bytes = N * (16 + 8 * (12 + 36 + 6 + 6 + 12 + 6 * S)) + 24 * M
    chk = float(eng.CT.double().abs().sum())
    print(f"CONSTITUTIVE {a.tag:28s} cells={N} tangent {t_tan:.3f} ms ({tan_bytes / t_tan / 1e6:.0f} GB/s)  post {t_post:.3f} ms "
          f"({post_bytes / t_post / 1e6:.0f} GB/s)  commit {t_commit:.3f} ms  checksum(C_T)={chk:.17e}", flush=True)


if __name__ == "__main__":
    main()

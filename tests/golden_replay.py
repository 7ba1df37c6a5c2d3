"""Replay a ``tests/golden/constitutive_*.npz`` sequence (made by ``oracle/gen_golden.py``
from the unmodified reference) through any implementation that offers the
``OracleMaterial`` interface, and report the worst mismatch per recorded quantity.

Shared by the CPU oracle tests and the GPU parity tests.
"""
import os

import numpy as np

from oracle import constitutive as oc

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, f"constitutive_{name}.npz"), allow_pickle=False))


def build_oracle_material(g):
    """Instantiate the numpy oracle from the parameters stored in a golden file."""
    spec = [str(s) for s in g["spec"]]
    N = g["sig0"].shape[0]
    mat = oc.OracleMaterial(N)
    P = lambda kind, k: g[f"param/{kind}/{k}"]
    mat.add_spring(P("spring", "E"), P("spring", "nu"))
    for kind in spec:
        if kind == "kelvin":
            mat.add(oc.Kelvin(P(kind, "eta"), P(kind, "E"), P(kind, "nu")))
        elif kind in ("dislocation", "dislocation_n45"):
            mat.add(oc.Dislocation(P(kind, "A"), P(kind, "Q"), P(kind, "n")))
        elif kind == "pressure_solution":
            mat.add(oc.PressureSolution(P(kind, "A"), P(kind, "d"), P(kind, "Q")))
        elif kind == "desai":
            kw = {k: P(kind, k) for k in oc.DesaiParams.names}
            mat.add(oc.Desai(P(kind, "alpha_0"), **kw))
        elif kind == "thermo":
            mat.add_thermoelastic(P(kind, "alpha"))
        elif kind == "munson_dawson":
            mat.add(oc.MunsonDawson(**{k: P(kind, k) for k in oc.MunsonDawsonParams.names}))
        elif kind in ("mohr_coulomb", "matsuoka_nakai"):
            cls = oc.MohrCoulomb if kind == "mohr_coulomb" else oc.MatsuokaNakai
            mat.add(cls(*[P(kind, k) for k in INTERLAYER_NAMES]))
    return mat


INTERLAYER_NAMES = ("mu_1", "N_1", "cohesion", "friction_angle", "dilation_angle", "sigma_t")


def cell_rel_err(a, b):
    """max over cells of (max|a-b| / max|b|) with the maxima taken per cell."""
    a = np.asarray(a, dtype=np.float64).reshape(a.shape[0], -1)
    b = np.asarray(b, dtype=np.float64).reshape(b.shape[0], -1)
    scale = np.max(np.abs(b), axis=1)
    diff = np.max(np.abs(a - b), axis=1)
    ok = scale > 0
    out = 0.0
    if ok.any():
        out = float(np.max(diff[ok] / scale[ok]))
    return out


STATE_FIELDS = ("rate", "rate_old", "eps_old", "eps_k")
ISV_FIELDS = ("alpha", "alpha_0", "Fvp", "qsi", "qsi_old", "r", "h", "P", "zeta", "zeta_old", "F")
GOLD_NAME = {"rate": "eps_ne_rate", "rate_old": "eps_ne_rate_old", "eps_old": "eps_ne_old",
             "eps_k": "eps_ne_k"}

# absolute floors below which a quantity is numerical noise in BOTH implementations
# (e.g. the Desai flow rate when Fvp is a round-off-sized number right after
# compute_initial_hardening, or the residue r = alpha - alpha_0*(1+O(eps)))
ATOL = {"rate": 1e-22, "rate_old": 1e-22, "eps_old": 1e-20, "eps_k": 1e-20, "r": 1e-15,
        "B": 1e-20, "B_elem": 1e-20, "P": 1e-22, "Fvp": 1.0, "qsi": 1e-18, "eps_rhs": 1e-18,
        "zeta": 1e-20, "zeta_old": 1e-20}


def err(a, b, atol=0.0):
    """max|a-b| / max(max|b|, atol): scale-relative error of one recorded array."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    bad = np.isnan(a) != np.isnan(b)
    if bad.any():
        return float("inf")
    ok = ~np.isnan(b)
    if not ok.any():
        return 0.0
    scale = max(float(np.max(np.abs(b[ok]))), atol)
    if scale == 0.0:
        return float(np.max(np.abs(a[ok])))
    return float(np.max(np.abs(a[ok] - b[ok])) / scale)


def gold(g, key):
    v = g[key]
    return oc.to_voigt(v) if (v.ndim == 3 and v.shape[1:] == (3, 3)) else v


def inject(mat, g, tag, with_tangent=False):
    """Overwrite the implementation's per-element state with the reference's record."""
    for i, e in enumerate(mat.elems):
        pre = f"{tag}/e{i}"
        names = STATE_FIELDS + ("alpha", "alpha_0", "Fvp", "qsi", "qsi_old", "zeta", "zeta_old", "F")
        if with_tangent:
            names = names + ("r", "h", "P")
        for f in names:
            key = f"{pre}/{GOLD_NAME.get(f, f)}"
            if key in g and hasattr(e, f):
                setattr(e, f, gold(g, key).copy())
        if with_tangent and hasattr(e, "h_small"):
            # the reference flags |h| < 1e-6 (Desai; 1e-12 MunsonDawson) and resets h to 1, P to 0
            # (MaterialProps.py:1473-1498, 2275-2311)
            P = gold(g, f"{pre}/P")
            e.h_small = (gold(g, f"{pre}/h") == 1.0) & (np.abs(P).max(axis=1) == 0.0)


def compare(mat, g, tag, phase, errs, names):
    for i, e in enumerate(mat.elems):
        pre = f"{tag}/e{i}"
        for f in names:
            key = f"{pre}/{GOLD_NAME.get(f, f)}"
            if key in g and hasattr(e, f):
                errs.setdefault(f"{phase}:{f}[{e.kind}]", []).append(
                    err(getattr(e, f), gold(g, key), ATOL.get(f, 0.0)))


def replay(g, mat, isolate=True):
    """Drive ``mat`` through the recorded sequence and return {quantity: worst error}.

    isolate=True: before every phase the implementation's state is overwritten with the
    reference's record of the previous phase, so each phase is checked on its own and the
    finite-difference noise of one tangent does not leak into later comparisons.
    isolate=False: free-running (errors accumulate the way they would in a simulation).
    """
    errs = {}
    dt, theta = float(g["dt"]), float(g["theta"])
    T, T0 = g["T"], g["T0"]
    sig = oc.to_voigt(g["sig0"])
    noise = oc.to_voigt(g["noise"])
    if bool(g["desai_init"]):
        for e in mat.elems:
            if e.kind == "desai":
                e.initial_hardening(sig, 0.0)
    mat.eval_rates(sig, 0.0 * theta, T)
    mat.commit_rates()
    compare(mat, g, "init", "init", errs, STATE_FIELDS + ("alpha", "alpha_0", "Fvp", "F"))
    prev = "init"
    for step in range(int(g["n_steps"])):
        for it in range(int(g["n_iters"])):
            tag = f"s{step}i{it}"
            if isolate:
                sig_k = gold(g, f"{tag}/sig_k")
                inject(mat, g, prev)
            else:
                sig_k = sig.copy()
            CT, eps_rhs = mat.tangent_phase(sig_k, T, T0, dt, theta)
            if getattr(mat, "G", None) is not None:    # the CUDA path never materialises G, B
                errs.setdefault("tan:G", []).append(err(mat.G, g[f"{tag}/G"]))
                errs.setdefault("tan:B", []).append(err(mat.B, gold(g, f"{tag}/B"), ATOL["B"]))
            errs.setdefault("tan:CT", []).append(err(CT, g[f"{tag}/CT"]))
            errs.setdefault("tan:eps_rhs", []).append(err(eps_rhs, gold(g, f"{tag}/eps_rhs"), ATOL["eps_rhs"]))
            compare(mat, g, tag + "/tan", "tan", errs, ("eps_k", "qsi", "r", "h", "P"))
            for i, e in enumerate(mat.elems):
                if getattr(e, "G", None) is not None:
                    errs.setdefault(f"tan:G_elem[{e.kind}]", []).append(err(e.G, g[f"{tag}/tan/e{i}/G"]))
                    errs.setdefault(f"tan:B_elem[{e.kind}]", []).append(
                        err(e.B, gold(g, f"{tag}/tan/e{i}/B"), ATOL["B_elem"]))
            if isolate:
                eps_tot = gold(g, f"{tag}/eps_tot")
                inject(mat, g, tag + "/tan", with_tangent=True)
                mat.CT = g[f"{tag}/CT"].copy()
                mat.eps_rhs = gold(g, f"{tag}/eps_rhs").copy()
            else:
                scale = 1 + float(g["load_step"]) / (it + 1) * (1 + noise)
                eps_tot = oc.ddot(mat.C_inv, sig_k * scale) + eps_rhs
            sig = mat.post_phase(eps_tot, sig_k, T, dt, theta)
            errs.setdefault("post:sig", []).append(err(sig, gold(g, f"{tag}/sig")))
            compare(mat, g, tag + "/post", "post", errs, ("rate", "alpha", "Fvp", "zeta", "F"))
            prev = tag + "/post"
            if isolate:
                sig = gold(g, f"{tag}/sig")
        if isolate:
            inject(mat, g, prev)
        mat.commit(sig, sig_k, dt, theta)
        compare(mat, g, f"s{step}/commit", "commit", errs, ("rate_old", "eps_old", "qsi_old", "zeta_old"))
        prev = f"s{step}/commit"
    return {k: max(v) for k, v in errs.items()}


def record(g, mat):
    """Run ``mat`` free-running over the golden's load sequence (same procedure as
    oracle/gen_golden.py) and return a record with the golden files' key layout, so that
    ``replay(record(g, oracle), cuda, isolate=True)`` checks the CUDA path against the ORACLE phase
    by phase on bit-identical inputs."""
    rec = {k: v for k, v in g.items() if "/" not in k or k.startswith("param/")}
    dt, theta = float(g["dt"]), float(g["theta"])
    T, T0 = g["T"], g["T0"]
    sig = oc.to_voigt(g["sig0"])
    noise = oc.to_voigt(g["noise"])
    C_inv = np.asarray(mat.C_inv)

    def snap(tag):
        for i, e in enumerate(mat.elems):
            for f in STATE_FIELDS + ISV_FIELDS:
                if hasattr(e, f):
                    rec[f"{tag}/e{i}/{GOLD_NAME.get(f, f)}"] = np.array(getattr(e, f), dtype=np.float64)
            if getattr(e, "G", None) is not None:
                rec[f"{tag}/e{i}/G"] = np.array(e.G)
                rec[f"{tag}/e{i}/B"] = np.array(e.B)

    if bool(g["desai_init"]):
        for e in mat.elems:
            if e.kind == "desai":
                e.initial_hardening(sig, 0.0)
    mat.eval_rates(sig, 0.0 * theta, T)
    mat.commit_rates()
    snap("init")
    for step in range(int(g["n_steps"])):
        for it in range(int(g["n_iters"])):
            tag = f"s{step}i{it}"
            sig_k = sig.copy()
            CT, eps_rhs = mat.tangent_phase(sig_k, T, T0, dt, theta)
            rec[f"{tag}/sig_k"] = sig_k
            rec[f"{tag}/CT"] = np.array(CT)
            rec[f"{tag}/eps_rhs"] = np.array(eps_rhs)
            if getattr(mat, "G", None) is not None:
                rec[f"{tag}/G"] = np.array(mat.G)
                rec[f"{tag}/B"] = np.array(mat.B)
            snap(tag + "/tan")
            scale = 1 + float(g["load_step"]) / (it + 1) * (1 + noise)
            eps_tot = oc.ddot(C_inv, sig_k * scale) + rec[f"{tag}/eps_rhs"]
            rec[f"{tag}/eps_tot"] = eps_tot
            sig = np.array(mat.post_phase(eps_tot, sig_k, T, dt, theta))
            rec[f"{tag}/sig"] = sig.copy()
            snap(tag + "/post")
        mat.commit(sig, sig_k, dt, theta)
        snap(f"s{step}/commit")
    return rec

"""Adapter that lets tests/golden_replay.py drive the CUDA constitutive kernels (through the
C ABI) with the same interface as oracle.constitutive.OracleMaterial."""
import numpy as np
import torch

import safeincave_b200 as sf
from safeincave_b200 import _lib as L
from safeincave_b200.engine import Engine

KIND_NAME = {L.ELEM_KELVIN: "kelvin", L.ELEM_DISLOCATION: "dislocation",
             L.ELEM_PRESSURE_SOL: "pressure_solution", L.ELEM_DESAI: "desai", L.ELEM_MUNSON_DAWSON: "munson_dawson",
             L.ELEM_MOHR_COULOMB: "mohr_coulomb", L.ELEM_MATSUOKA_NAKAI: "matsuoka_nakai"}
# name -> row of the element's internal-state block, per element kind
ROWS = {
    L.ELEM_DESAI: {"alpha": L.DS_ALPHA, "alpha_0": L.DS_ALPHA0, "Fvp": L.DS_FVP, "qsi": L.DS_QSI,
                   "qsi_old": L.DS_QSI_OLD, "r": L.DS_R, "h": L.DS_H, "P": L.DS_P, "h_small": L.DS_HSMALL},
    L.ELEM_MUNSON_DAWSON: {"zeta": L.MD_ZETA, "zeta_old": L.MD_ZETA_OLD, "F": L.MD_F, "r": L.MD_R, "h": L.MD_H,
                           "P": L.MD_P, "h_small": L.MD_HSMALL},
    L.ELEM_MOHR_COULOMB: {"Fvp": L.VP_FVP}, L.ELEM_MATSUOKA_NAKAI: {"Fvp": L.VP_FVP},
}
DS_ROW = {"alpha": L.DS_ALPHA, "alpha_0": L.DS_ALPHA0, "Fvp": L.DS_FVP, "qsi": L.DS_QSI,
          "qsi_old": L.DS_QSI_OLD, "r": L.DS_R, "h": L.DS_H}


def disjoint_tets(n):
    """n unit tetrahedra that share no node: a mesh for kernels that only need cells."""
    base = np.array([[0., 0., 0.], [1., 0., 0.], [0., 1., 0.], [0., 0., 1.]])
    coords = (base[None, :, :] + 2.0 * np.arange(n)[:, None, None] * np.array([1., 0., 0.])).reshape(-1, 3)
    cells = np.arange(4 * n).reshape(n, 4)
    return coords, cells


class GpuElemView:
    G = None
    B = None

    def __init__(self, eng, i):
        object.__setattr__(self, "_eng", eng)
        object.__setattr__(self, "_st", eng.elems[i])
        object.__setattr__(self, "kind", KIND_NAME[eng.elems[i].kind])

    def __getattr__(self, name):
        eng, st = self._eng, self._st
        if name in ("rate", "rate_old", "eps_old", "eps_k"):
            return eng.get6(getattr(st, name))
        rows = ROWS.get(st.kind, {})
        if name in rows:
            r = rows[name]
            if name == "P":
                return eng.get6(st.desai[r:r + 6])
            v = eng.get1(st.desai[r])
            return v != 0 if name == "h_small" else v
        raise AttributeError(name)

    def __setattr__(self, name, value):
        eng, st = self._eng, self._st
        rows = ROWS.get(st.kind, {})
        if name in ("rate", "rate_old", "eps_old", "eps_k"):
            eng.put6(getattr(st, name), value)
        elif name in rows:
            r = rows[name]
            if name == "P":
                eng.put6(st.desai[r:r + 6], value)
            else:
                eng.put1(st.desai[r], np.asarray(value, dtype=np.float64))
        else:
            raise AttributeError(name)

    def initial_hardening(self, sig, Fvp_0):
        eng = self._eng
        eng.put6(eng.sig, sig)
        idx = eng.elems.index(self._st)
        eng.desai_initial_hardening(idx, Fvp_0)


class GpuMaterial:
    """Built from the parameters stored in a golden file."""
    G = None
    B = None

    engine_cls = Engine          # tests/hostemu swaps in EmuEngine to run the same kernels on the host

    def __init__(self, g, dtype=torch.float64):
        t = lambda a: torch.tensor(np.asarray(a), dtype=dtype)
        spec = [str(s) for s in g["spec"]]
        N = g["sig0"].shape[0]
        P = lambda kind, k: t(g[f"param/{kind}/{k}"])
        mat = sf.Material(N)
        mat.add_to_elastic(sf.Spring(P("spring", "E"), P("spring", "nu")))
        for kind in spec:
            if kind == "kelvin":
                mat.add_to_non_elastic(sf.Viscoelastic(P(kind, "eta"), P(kind, "E"), P(kind, "nu")))
            elif kind in ("dislocation", "dislocation_n45"):
                mat.add_to_non_elastic(sf.DislocationCreep(P(kind, "A"), P(kind, "Q"), P(kind, "n")))
            elif kind == "pressure_solution":
                mat.add_to_non_elastic(sf.PressureSolutionCreep(P(kind, "A"), P(kind, "d"), P(kind, "Q")))
            elif kind == "desai":
                names = sf.ViscoplasticDesai.param_names
                mat.add_to_non_elastic(sf.ViscoplasticDesai(*[P(kind, k) for k in names], P(kind, "alpha_0")))
            elif kind == "thermo":
                mat.add_to_thermoelastic(sf.Thermoelastic(P(kind, "alpha")))
            elif kind == "munson_dawson":
                mat.add_to_non_elastic(sf.MunsonDawsonCreep(*[P(kind, k) for k in sf.MunsonDawsonCreep.param_names]))
            elif kind in ("mohr_coulomb", "matsuoka_nakai"):
                cls = sf.MohrCoulombViscoplastic if kind == "mohr_coulomb" else sf.MatsuokaNakaiViscoplastic
                mat.add_to_non_elastic(cls(*[P(kind, k) for k in cls.param_names]))
        coords, cells = disjoint_tets(N)
        self.eng = self.engine_cls(coords, cells)
        mat.bind(self.eng)
        self.mat = mat
        self.elems = [GpuElemView(self.eng, i) for i in range(len(self.eng.elems))]
        self.kelvin_phi2 = -1.0
        self.C_inv = mat.C_inv.numpy()

    # --- OracleMaterial interface
    def tangent_phase(self, sig_k, T, T0, dt, theta):
        e = self.eng
        e.put6(e.sig_k, sig_k)
        e.put1(e.T, T)
        e.put1(e.T0, T0)
        e.tangent(dt, theta)
        self.kelvin_phi2 = dt * (1 - theta)
        return self.CT, self.eps_rhs

    @property
    def CT(self):
        e = self.eng
        return e.get_CT()

    @CT.setter
    def CT(self, v):
        e = self.eng
        e.put_CT(np.asarray(v))

    @property
    def eps_rhs(self):
        return self.eng.get6(self.eng.eps_rhs)

    @eps_rhs.setter
    def eps_rhs(self, v):
        self.eng.put6(self.eng.eps_rhs, v)

    def post_phase(self, eps, sig_k, T, dt, theta):
        e = self.eng
        e.put6(e.eps, eps)
        e.put6(e.sig_k, sig_k)
        e.put1(e.T, T)
        e.post(None, dt, theta, self.kelvin_phi2, L.POST_STRESS | L.POST_INCREMENT | L.POST_RATES)
        return e.get6(e.sig)

    def eval_rates(self, sig, phi1, T):
        e = self.eng
        e.put6(e.sig, sig)
        e.put1(e.T, T)
        e.post(None, phi1, 1.0, self.kelvin_phi2, L.POST_RATES)

    def commit_rates(self):
        self.eng.commit_rates()

    def commit(self, sig, sig_k, dt, theta):
        e = self.eng
        e.put6(e.sig, sig)
        e.put6(e.sig_k, sig_k)
        e.commit(dt, theta)

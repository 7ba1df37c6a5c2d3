"""Host-side mirror of the reference's constitutive API (safeincave/MaterialProps.py).

Same class names, constructor signatures and attribute names as the reference
(``Material``:22-331, ``Thermoelastic``:333-382, ``Spring``:385-539, ``NonElasticElement``:543-789,
``Viscoelastic``:795-885, ``DislocationCreep``:890-961, ``PressureSolutionCreep``:964-1034,
``ViscoplasticDesai``:1037-1562) so that user scripts run unchanged -- but these objects hold no
arithmetic: they describe the material to the CUDA engine (a de-duplicated parameter table plus a
per-cell row index) and expose the device state as lazily pulled host tensors.

Parameters are canonicalised to float64 on entry; quantities the reference derives at SETUP time
(C, C^-1, Kelvin C1) are derived here with the same torch expressions in the user's dtype, so
float32 user tensors give the same stiffness entries as the reference.
"""
from __future__ import annotations

import numpy as np
import torch as to

from . import _lib as L
from .engine import ElemSpec, Engine

VOIGT = ((0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2))


def voigt_to_tensor(v6: to.Tensor) -> to.Tensor:
    """(N,6) -> symmetric (N,3,3) (Utils.py:276-282)."""
    n = v6.shape[0]
    t = to.zeros((n, 3, 3), dtype=v6.dtype, device=v6.device)
    for k, (i, j) in enumerate(VOIGT):
        t[:, i, j] = v6[:, k]
        t[:, j, i] = v6[:, k]
    return t


def tensor_to_voigt(t33: to.Tensor) -> to.Tensor:
    """(N,3,3) -> (N,6), upper entries (Utils.py:269-274)."""
    return to.stack([t33[:, i, j] for i, j in VOIGT], dim=1)


def iso_to_66(c11, c12, c44):
    n = c11.shape[0]
    C = to.zeros((n, 6, 6), dtype=to.float64)
    for i in range(3):
        C[:, i, i] = c11
        C[:, i + 3, i + 3] = c44
        for j in range(3):
            if i != j:
                C[:, i, j] = c12
    return C


def _iso_entries(E, nu):
    """c11, c12, c44 exactly as MaterialProps.py:478-486 / 826-833 evaluates them (user dtype)."""
    a0 = E / ((1 + nu) * (1 - 2 * nu))
    return (a0 * (1 - nu)).double(), (a0 * nu).double(), (a0 * (1 - 2 * nu)).double()


class _ParamOwner:
    """Anything that contributes per-cell parameter columns to the material table."""
    param_names: tuple = ()

    def _columns(self):
        return [getattr(self, "_p_" + n) for n in self.param_names]

    def __setattr__(self, name, value):
        """Re-assigning a parameter AFTER the material was attached (e.g. ``mc.mu_1 = mc._real_mu_1`` to switch an
        interlayer model on once the stress field has equilibrated, nobian/Simulation/run_interlayer.py:1664-1669)
        rebuilds the device-side parameter table, keeping the state."""
        object.__setattr__(self, name, value)
        if name in self.param_names and "_p_" + name in self.__dict__:
            v = to.as_tensor(value)
            if v.ndim == 1 and v.shape[0] == self.__dict__["_p_" + name].shape[0]:
                object.__setattr__(self, "_p_" + name, v.detach().cpu())
                mat = self.__dict__.get("_material")
                if mat is not None and mat._engine is not None:
                    mat.rebind()

    def _set_params(self, **kw):
        n = None
        for k, v in kw.items():
            v = to.as_tensor(v)
            if v.ndim != 1:
                raise ValueError(f"{type(self).__name__}: parameter {k} must be a 1-D per-cell tensor")
            n = v.shape[0] if n is None else n
            if v.shape[0] != n:
                raise ValueError(f"{type(self).__name__}: parameter {k} has {v.shape[0]} entries, expected {n}")
            object.__setattr__(self, "_p_" + k, v.detach().cpu())
            object.__setattr__(self, k, v)
        return n


class Thermoelastic(_ParamOwner):
    """eps_th = alpha (T - T0) I  (MaterialProps.py:333-382)."""
    param_names = ("alpha",)

    def __init__(self, alpha, name="thermoelastic"):
        self.n_elems = self._set_params(alpha=alpha)
        self.name = name

    def derived(self, rows):
        return [rows["alpha"].double()]


class Spring(_ParamOwner):
    """Linear isotropic spring (MaterialProps.py:385-539)."""
    param_names = ("E", "nu")

    def __init__(self, E, nu, name="spring"):
        self.n_elems = self._set_params(E=E, nu=nu)
        self.name = name

    def initialize(self):
        """Builds C, C_inv (N,6,6) and K like the reference (:422-438).  Only called for API
        compatibility / small meshes; the engine uses the per-row entries from derived()."""
        c11, c12, c44 = _iso_entries(self._p_E, self._p_nu)
        self.C = iso_to_66(c11, c12, c44)
        self.C_inv = to.linalg.inv(self.C)
        self.K = self._p_E / (3 * (1 - 2 * self._p_nu))

    def derived(self, rows):
        c11, c12, c44 = _iso_entries(rows["E"], rows["nu"])
        Cinv = to.linalg.inv(iso_to_66(c11, c12, c44))        # as :503 (LAPACK inverse of C)
        return [c11, c12, c44, Cinv[:, 0, 0], Cinv[:, 0, 1], Cinv[:, 3, 3]]


class NonElasticElement(_ParamOwner):
    """Base of the non-elastic elements (MaterialProps.py:543-789).  State lives on the device;
    the attributes below pull it to the host on access, as (N,3,3) / (N,) float64 tensors."""
    kind = 0
    n_row_params = 0

    def __init__(self, n_elems):
        self.n_elems = n_elems
        self._engine: Engine | None = None
        self._index = -1

    # --- binding to the device engine
    def _bind(self, engine, index):
        self._engine, self._index = engine, index

    def _state(self):
        if self._engine is None:
            raise RuntimeError(f"element '{self.name}' is not attached to a momentum equation yet "
                               "(call LinearMomentum.set_material first)")
        return self._engine.elems[self._index]

    def _pull6(self, name):
        st = self._state()
        n = self._engine.N
        return voigt_to_tensor(getattr(st, name)[:, :n].t()).cpu()

    def _push6(self, name, value):
        st = self._state()
        v = tensor_to_voigt(to.as_tensor(value, dtype=to.float64)).to(self._engine.device)
        getattr(st, name)[:, :self._engine.N] = v.t()

    eps_ne_rate = property(lambda s: s._pull6("rate"), lambda s, v: s._push6("rate", v))
    eps_ne_rate_old = property(lambda s: s._pull6("rate_old"), lambda s, v: s._push6("rate_old", v))
    eps_ne_old = property(lambda s: s._pull6("eps_old"), lambda s, v: s._push6("eps_old", v))
    eps_ne_k = property(lambda s: s._pull6("eps_k"), lambda s, v: s._push6("eps_k", v))

    def derived(self, rows):
        return [rows[n].double() for n in self.param_names]


class Viscoelastic(NonElasticElement):
    """Kelvin-Voigt element (MaterialProps.py:795-885)."""
    kind = L.ELEM_KELVIN
    param_names = ("eta", "E", "nu")
    n_row_params = 4

    def __init__(self, eta, E, nu, name="kelvin_voigt"):
        super().__init__(self._set_params(eta=eta, E=E, nu=nu))
        self.name = name

    def derived(self, rows):
        c11, c12, c44 = _iso_entries(rows["E"], rows["nu"])
        return [rows["eta"].double(), c11, c12, c44]


class DislocationCreep(NonElasticElement):
    """Power-law dislocation creep (MaterialProps.py:890-961), R = 8.32."""
    kind = L.ELEM_DISLOCATION
    param_names = ("A", "Q", "n")
    n_row_params = 3

    def __init__(self, A, Q, n, name="creep"):
        super().__init__(self._set_params(A=A, Q=Q, n=n))
        self.R = 8.32
        self.name = name


class PressureSolutionCreep(NonElasticElement):
    """Pressure-solution creep (MaterialProps.py:964-1034), R = 8.32."""
    kind = L.ELEM_PRESSURE_SOL
    param_names = ("A", "d", "Q")
    n_row_params = 3

    def __init__(self, A, d, Q, name="creep"):
        super().__init__(self._set_params(A=A, d=d, Q=Q))
        self.R = 8.32
        self.name = name


class ViscoplasticDesai(NonElasticElement):
    """Desai viscoplasticity with hardening variable alpha (MaterialProps.py:1037-1562)."""
    kind = L.ELEM_DESAI
    param_names = ("mu_1", "N_1", "a_1", "eta", "n", "beta_1", "beta", "m", "gamma", "sigma_t")
    n_row_params = 10

    def __init__(self, mu_1, N_1, a_1, eta, n, beta_1, beta, m, gamma, sigma_t, alpha_0, name="desai"):
        super().__init__(self._set_params(mu_1=mu_1, N_1=N_1, a_1=a_1, eta=eta, n=n, beta_1=beta_1, beta=beta,
                                          m=m, gamma=gamma, sigma_t=sigma_t))
        self.name = name
        self.F_0 = 1.0
        self._alpha_0_init = to.as_tensor(alpha_0).detach().double().cpu().clone()

    def _bind(self, engine, index):
        super()._bind(engine, index)
        ds = engine.elems[index].desai
        a0 = self._alpha_0_init.to(engine.device)
        ds[L.DS_ALPHA0, :engine.N] = a0      # :1086
        ds[L.DS_ALPHA, :engine.N] = a0       # :1089
        ds[L.DS_ALPHA0, engine.N:] = 1.0
        ds[L.DS_ALPHA, engine.N:] = 1.0
        pending = self.__dict__.pop("_pending_hardening", None)
        if pending is not None:               # compute_initial_hardening was called before the element was attached
            keep = engine.sig.clone()
            self.compute_initial_hardening(pending[0], pending[1])
            engine.sig.copy_(keep)

    def _row(self, r):
        return self._state().desai[r, :self._engine.N].cpu()

    def _set_row(self, r, v):
        self._state().desai[r, :self._engine.N] = to.as_tensor(v, dtype=to.float64).to(self._engine.device)

    alpha = property(lambda s: s._row(L.DS_ALPHA), lambda s, v: s._set_row(L.DS_ALPHA, v))
    alpha_0 = property(lambda s: s._row(L.DS_ALPHA0), lambda s, v: s._set_row(L.DS_ALPHA0, v))
    Fvp = property(lambda s: s._row(L.DS_FVP), lambda s, v: s._set_row(L.DS_FVP, v))
    qsi = property(lambda s: s._row(L.DS_QSI), lambda s, v: s._set_row(L.DS_QSI, v))
    qsi_old = property(lambda s: s._row(L.DS_QSI_OLD), lambda s, v: s._set_row(L.DS_QSI_OLD, v))
    r = property(lambda s: s._row(L.DS_R))
    h = property(lambda s: s._row(L.DS_H))

    @property
    def P(self):
        st, n = self._state(), self._engine.N
        return voigt_to_tensor(st.desai[L.DS_P:L.DS_P + 6, :n].t()).cpu()

    def compute_initial_hardening(self, stress, Fvp_0=0.0):
        """MaterialProps.py:1248-1288 on the device.  ``stress``: (N,3,3) tensor (host or device)."""
        eng = self._engine
        if eng is None:
            # the reference's scripts call this BEFORE add_to_non_elastic / set_material (Simulators.py:1271-1276,
            # nobian/Simulation/Run.py:1500-1504): keep the stress and run the kernel when the element is attached
            if stress is None:
                raise RuntimeError("compute_initial_hardening(None) needs the material attached (set_material) first")
            s = stress.to_tensor() if hasattr(stress, "to_tensor") else to.as_tensor(stress)
            if s.ndim == 3:
                s = tensor_to_voigt(s.double())
            self._pending_hardening = (s.detach().double().cpu().clone(), float(Fvp_0))
            return
        if hasattr(stress, "to_tensor"):       # a CellField handle (LinearMomentum.sig)
            stress = None if stress.buf.data_ptr() == eng.sig.data_ptr() else stress.to_tensor()
        if stress is not None:
            s = to.as_tensor(stress)
            if s.ndim == 3:
                s = tensor_to_voigt(s.double())
            eng.sig[:, :eng.N] = s.to(eng.device, dtype=to.float64).t()
        n_clamped = eng.desai_initial_hardening(self._index, Fvp_0)
        if n_clamped > 0:
            import sys
            print(f"[DESAI INIT] Clamped alpha_0 for {n_clamped}/{self.n_elems} elements", file=sys.stderr)
        self.n_disabled = n_clamped


class _IsvRows:
    """Row accessors of the element's internal-state block on the device."""

    def _row(self, r):
        return self._state().desai[r, :self._engine.N].cpu()

    def _set_row(self, r, v):
        self._state().desai[r, :self._engine.N] = to.as_tensor(v, dtype=to.float64).to(self._engine.device)

    def _rows6(self, r0):
        st, n = self._state(), self._engine.N
        return voigt_to_tensor(st.desai[r0:r0 + 6, :n].t()).cpu()


class MunsonDawsonCreep(NonElasticElement, _IsvRows):
    """Munson-Dawson transient + steady-state creep with the internal variable zeta
    (MaterialProps.py:1971-2346), R = 8.32.  Parameters are promoted to float64 as the reference does (:2035-2049)."""
    kind = L.ELEM_MUNSON_DAWSON
    param_names = ("A", "Q", "n", "K0", "c", "m", "alpha_w", "beta_w", "delta", "mu")
    n_row_params = 10

    def __init__(self, A, Q, n, K0, c, m, alpha_w, beta_w, delta, mu, name="creep_munson_dawson"):
        super().__init__(self._set_params(A=A, Q=Q, n=n, K0=K0, c=c, m=m, alpha_w=alpha_w, beta_w=beta_w,
                                          delta=delta, mu=mu))
        self.R = 8.32
        self.name = name

    zeta = property(lambda s: s._row(L.MD_ZETA), lambda s, v: s._set_row(L.MD_ZETA, v))
    zeta_old = property(lambda s: s._row(L.MD_ZETA_OLD), lambda s, v: s._set_row(L.MD_ZETA_OLD, v))
    F = property(lambda s: s._row(L.MD_F))
    _eps_t_star = property(lambda s: s._row(L.MD_ETS))
    r = property(lambda s: s._row(L.MD_R))
    h = property(lambda s: s._row(L.MD_H))
    P = property(lambda s: s._rows6(L.MD_P))


def _dp_alpha(angle):
    """2 sin(a) / (sqrt(3) (3 - sin(a))) with the reference's expression and dtype (MaterialProps.py:1644, 1648)."""
    sn = to.sin(angle)
    return 2.0 * sn / (np.sqrt(3.0) * (3.0 - sn))


class MohrCoulombViscoplastic(NonElasticElement, _IsvRows):
    """Mohr-Coulomb (Drucker-Prager fit in triaxial compression) with tension cut-off and Perzyna overstress
    (MaterialProps.py:1565-1746).  alpha_F, k_F, alpha_Q are derived on the host with the reference's torch
    expressions (:1640-1648), the kernels receive them per material row."""
    kind = L.ELEM_MOHR_COULOMB
    param_names = ("mu_1", "N_1", "cohesion", "friction_angle", "dilation_angle", "sigma_t")
    n_row_params = 6

    def __init__(self, mu_1, N_1, cohesion, friction_angle, dilation_angle, sigma_t, name="mohr_coulomb"):
        super().__init__(self._set_params(mu_1=mu_1, N_1=N_1, cohesion=cohesion, friction_angle=friction_angle,
                                          dilation_angle=dilation_angle, sigma_t=sigma_t))
        self.name = name
        self.F_0 = 1.0

    def derived(self, rows):
        phi = rows["friction_angle"]
        sin_phi, cos_phi = to.sin(phi), to.cos(phi)
        alpha_F = 2.0 * sin_phi / (np.sqrt(3.0) * (3.0 - sin_phi))
        k_F = 6.0 * rows["cohesion"] * cos_phi / (np.sqrt(3.0) * (3.0 - sin_phi))
        return [rows["mu_1"].double(), rows["N_1"].double(), alpha_F.double(), k_F.double(),
                _dp_alpha(rows["dilation_angle"]).double(), rows["sigma_t"].double()]

    Fvp = property(lambda s: s._row(L.VP_FVP))


class MatsuokaNakaiViscoplastic(NonElasticElement, _IsvRows):
    """Matsuoka-Nakai criterion in the NFC (n = 1) form of Panteghini & Lagioia on the principal stresses, with
    Houlsby cohesive shift, tension cut-off, Drucker-Prager flow and Perzyna overstress (MaterialProps.py:1749-1968).
    The principal stresses come from a fixed six-sweep cyclic Jacobi iteration on the device (the reference calls
    torch.linalg.eigvalsh, :1882): agreement 1e-15 relative."""
    kind = L.ELEM_MATSUOKA_NAKAI
    param_names = ("mu_1", "N_1", "cohesion", "friction_angle", "dilation_angle", "sigma_t")
    n_row_params = 6

    def __init__(self, mu_1, N_1, cohesion, friction_angle, dilation_angle, sigma_t, name="matsuoka_nakai"):
        super().__init__(self._set_params(mu_1=mu_1, N_1=N_1, cohesion=cohesion, friction_angle=friction_angle,
                                          dilation_angle=dilation_angle, sigma_t=sigma_t))
        self.name = name
        self.F_0 = 1.0

    def derived(self, rows):
        phi = rows["friction_angle"]
        sin_phi, cos_phi = to.sin(phi), to.cos(phi)
        k_nfc = np.sqrt(2.0) * sin_phi                                              # :1816
        small = sin_phi.abs() < 1e-10
        safe = to.where(small, to.ones_like(sin_phi), sin_phi)
        shift = to.where(small, to.zeros_like(sin_phi), rows["cohesion"] * cos_phi / safe)   # :1820-1826
        return [rows["mu_1"].double(), rows["N_1"].double(), k_nfc.double(), shift.double(),
                _dp_alpha(rows["dilation_angle"]).double(), rows["sigma_t"].double()]

    Fvp = property(lambda s: s._row(L.VP_FVP))


class Material:
    """Aggregate of elastic, thermoelastic and non-elastic elements (MaterialProps.py:22-331)."""

    def __init__(self, n_elems: int):
        self.n_elems = n_elems
        self.elems_ne: list[NonElasticElement] = []
        self.elems_th: list[Thermoelastic] = []
        self.elems_e: list[Spring] = []
        self._engine: Engine | None = None
        self._layout = None

    def set_density(self, density):
        self.density = density

    def set_specific_heat_capacity(self, cp):
        self.cp = cp

    def set_thermal_conductivity(self, k):
        self.k = k

    def set_thermal_expansion(self, alpha_th):
        self.alpha_th = alpha_th

    def add_to_elastic(self, elem: Spring):
        if not isinstance(elem, Spring):
            raise TypeError("add_to_elastic expects a Spring")
        self.elems_e.append(elem)
        self.K = elem._p_E / (3 * (1 - 2 * elem._p_nu))      # :146-148
        self.E = elem.E
        self.ShearMod = 3 * self.K * elem._p_E / (9 * self.K - elem._p_E)

    def add_to_non_elastic(self, elem: NonElasticElement):
        if not isinstance(elem, NonElasticElement) or elem.kind == 0:
            raise TypeError(
                f"{type(elem).__name__} has no CUDA implementation in safeincave_b200 (supported: Viscoelastic, "
                "DislocationCreep, PressureSolutionCreep, ViscoplasticDesai, MunsonDawsonCreep, MohrCoulombViscoplastic, "
                "MatsuokaNakaiViscoplastic); there is no CPU fallback")
        if len(self.elems_ne) >= L.SIC_MAX_ELEMS:
            raise ValueError(f"at most {L.SIC_MAX_ELEMS} non-elastic elements")
        self.elems_ne.append(elem)
        elem._material = self
        if self._engine is not None:      # the material is shared by reference with the equation in the reference
            self.bind(self._engine)       # (Simulators.py:1276 adds Desai without a second set_material)

    def rebind(self):
        """Rebuild the parameter table on the device after a parameter tensor was re-assigned; state is kept."""
        engine = self._engine
        table, ids, layout = self.build_table(device=engine.device)
        engine.set_material(table.numpy(), ids, layout["spring_off"], layout["thermo_off"], layout["n_thermo"],
                            layout["specs"], keep_state=True)
        self._layout, self._table, self._ids = layout, table, ids.cpu()

    def add_to_thermoelastic(self, elem: Thermoelastic):
        if len(self.elems_th) >= L.SIC_MAX_THERMO:
            raise ValueError(f"at most {L.SIC_MAX_THERMO} thermoelastic elements")
        self.elems_th.append(elem)

    # ------------------------------------------------------------------ table
    def build_table(self, device="cpu"):
        """De-duplicate the per-cell parameter tuples into rows and derive the per-row constants.
        Returns (table (n_rows,row_len) float64, mat_id (N,) int64, layout dict)."""
        if not self.elems_e:
            raise RuntimeError("Material needs at least one Spring (add_to_elastic)")
        owners = list(self.elems_e) + list(self.elems_th) + list(self.elems_ne)
        for o in owners:
            if o.n_elems != self.n_elems:
                raise ValueError(f"element '{o.name}' has {o.n_elems} cells, material has {self.n_elems}")
        # combined id of the unique parameter tuple of every cell
        ids = to.zeros(self.n_elems, dtype=to.int64, device=device)
        first = None
        for o in owners:
            for col in o._columns():
                u, inv = to.unique(col.to(device).double(), return_inverse=True)
                if u.numel() > 1:
                    _, ids = to.unique(ids * u.numel() + inv, return_inverse=True)
        n_rows = int(ids.max().item()) + 1 if self.n_elems else 1
        # representative cell of every row
        first = to.full((n_rows,), self.n_elems, dtype=to.int64, device=device)
        first.scatter_reduce_(0, ids, to.arange(self.n_elems, device=device), reduce="amin")
        first = first.cpu()
        cols, layout = [], {}
        spring = [to.zeros(n_rows, dtype=to.float64) for _ in range(6)]
        for s in self.elems_e:          # C += elem.C and C_inv += elem.C_inv independently (:141-142)
            rows = {n: getattr(s, "_p_" + n)[first] for n in s.param_names}
            for k, v in enumerate(s.derived(rows)):
                spring[k] = spring[k] + v
        layout["spring_off"] = 0
        cols += spring
        layout["thermo_off"] = len(cols)
        for t in self.elems_th:
            rows = {n: getattr(t, "_p_" + n)[first] for n in t.param_names}
            cols += t.derived(rows)
        specs = []
        for e in self.elems_ne:
            rows = {n: getattr(e, "_p_" + n)[first] for n in e.param_names}
            specs.append(ElemSpec(e.kind, len(cols)))
            d = e.derived(rows)
            assert len(d) == e.n_row_params
            cols += d
        table = to.stack([c.double() for c in cols], dim=1).contiguous()
        layout["specs"] = specs
        layout["n_thermo"] = len(self.elems_th)
        return table, ids, layout

    def bind(self, engine: Engine):
        """Attach to the device engine.  Called again by a later ``set_material`` of the same engine -- the reference's
        staged runs add ViscoplasticDesai after the equilibrium stage and call ``mom_eq.set_material(mat)`` a second
        time (Simulators.py:1213-1326, nobian/Simulation/Run.py:1503-1506) -- elements that are already attached keep
        their state (in the reference it lives in the element objects), new ones start from zero."""
        table, ids, layout = self.build_table(device=engine.device)
        keep = [e._index if (e._engine is engine and 0 <= e._index < len(engine.elems)
                             and engine.elems[e._index].kind == e.kind) else None for e in self.elems_ne]
        engine.set_material(table.numpy(), ids, layout["spring_off"], layout["thermo_off"], layout["n_thermo"],
                            layout["specs"], keep=keep)
        self._engine, self._layout, self._table, self._ids = engine, layout, table, ids.cpu()
        for i, (e, k) in enumerate(zip(self.elems_ne, keep)):
            if k is None:
                e._bind(engine, i)
            else:
                e._index = i

    # ------------------------------------------------------------------ reference attributes (lazy)
    def _rows66(self, k0):
        t = self._table[self._ids]
        return iso_to_66(t[:, k0], t[:, k0 + 1], t[:, k0 + 2])

    @property
    def C(self):
        return self._rows66(0)

    @property
    def C_inv(self):
        return self._rows66(3)

    @property
    def CT(self):
        eng = self._engine
        return to.as_tensor(eng.get_CT())

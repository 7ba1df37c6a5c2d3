"""Facade with the petsc4py.PETSc.KSP methods the reference's scripts touch:
``KSP().create(comm)``, ``setType``, ``getPC().setType``, ``setTolerances``, ``getType``,
``getPC().getType``, ``getTolerances``, ``setOperators``, ``getIterationNumber``,
``getResidualNorm``, ``getConvergedReason``.  It only records the configuration; the solve itself
is ``sic_ksp_solve`` (block-Jacobi CG / BiCGStab on the matrix-free operator):

    cg                       -> SIC_KSP_CG
    bicg, bcgs, gmres, ...   -> SIC_KSP_BICGSTAB   (non-symmetric tangents)
    preonly (+ lu)           -> SIC_KSP_BICGSTAB with rtol 1e-13
    PC mg (or gamg on a grid that carries its refinement hierarchy, Grid.from_hierarchy)
                             -> CG preconditioned by a geometric-multigrid V-cycle (csrc/mg.cu)
    any other PC type        -> nodal 3x3 block Jacobi

T10: the examples' max_it = 100 silently truncates PETSc solves on large meshes; here the solve runs
to the requested rtol (``min_max_it``) unless ``respect_max_it`` is set.
"""
from . import _lib as L


class PC:
    def __init__(self):
        self._type = "bjacobi"

    def setType(self, t):
        self._type = str(t)

    def getType(self):
        return self._type


class KSP:
    def __init__(self):
        self._type = "cg"
        self._pc = PC()
        self.rtol, self.atol, self.divtol, self.max_it = 1e-5, 1e-50, 1e4, 10000
        self.respect_max_it = False
        self.min_max_it = 200000
        self.check_every = 25
        self.initial_guess_nonzero = False
        # with initial_guess_nonzero: predict the next Newton iterate from the last three differences of the time
        # step's iterates (sic_guess_extrapolate) instead of starting from the previous one; the solve still runs to rtol
        self.guess_extrapolation = False
        self.single_reduction = False      # cg -> Chronopoulos-Gear CG; symmetric (elastic) tangents only, see header
        self.mg_max_it, self.mg_check_every = 500, 4
        # PC mg: rebuild the preconditioner (Galerkin coarse tangents, block-Jacobi blocks, lambda_max) only on the
        # first `mg_setup_first` Newton iterations of a time step and afterwards only while the Newton error is above
        # `mg_setup_error`; in between the previous tangent's preconditioner is kept (the Krylov operator itself is
        # always the current tangent, so only the convergence RATE can change).  0: rebuild for every tangent.
        self.mg_setup_first, self.mg_setup_error = 0, 1e-3
        self._its, self._rnorm, self._reason = 0, 0.0, 0
        self.total_iterations = 0

    def create(self, comm=None):
        self.comm = comm
        return self

    def setType(self, t):
        self._type = str(t)

    def getType(self):
        return self._type

    def getPC(self):
        return self._pc

    def setTolerances(self, rtol=None, atol=None, divtol=None, max_it=None):
        if rtol is not None:
            self.rtol = float(rtol)
        if atol is not None:
            self.atol = float(atol)
        if divtol is not None:
            self.divtol = float(divtol)
        if max_it is not None:
            self.max_it = int(max_it)

    def getTolerances(self):
        return self.rtol, self.atol, self.divtol, self.max_it

    def setInitialGuessNonzero(self, flag):
        self.initial_guess_nonzero = bool(flag)

    def setGuessExtrapolation(self, flag):
        """Not in petsc4py: see ``guess_extrapolation`` above (implies a nonzero initial guess)."""
        self.guess_extrapolation = bool(flag)
        if flag:
            self.initial_guess_nonzero = True

    def setOperators(self, *a, **k):
        pass

    def setFromOptions(self):
        pass

    def getIterationNumber(self):
        return self._its

    def getResidualNorm(self):
        return self._rnorm

    def getConvergedReason(self):
        return self._reason

    # ---- used by LinearMomentum
    def method(self):
        t = self._type.lower()
        if t == "cg":
            return L.KSP_CGCG if self.single_reduction else L.KSP_CG
        return L.KSP_BICGSTAB

    def uses_multigrid(self, grid):
        pc = self._pc.getType().lower()
        has = getattr(grid, "hierarchy", None) is not None and grid.hierarchy.n_levels > 1
        if pc == "mg" and not has:
            raise ValueError("PC type 'mg' needs a grid built with GridHandlerGMSH.from_hierarchy(refine_hierarchy(...))")
        return has and pc in ("mg", "gamg")

    def effective(self):
        rtol = 1e-13 if self._type.lower() == "preonly" else self.rtol
        max_it = self.max_it if self.respect_max_it else max(self.max_it, self.min_max_it)
        return rtol, max(self.atol, 0.0) if self.atol > 1e-40 else 0.0, max_it

    def record(self, ksp):
        self._its, self._rnorm, self._reason = int(ksp.iterations), float(ksp.rnorm), int(ksp.reason)
        self.total_iterations += int(ksp.iterations)


class _PETScNamespace:
    """``from safeincave_b200.Solver import PETSc`` stands in for ``from petsc4py import PETSc``."""
    KSP = KSP


PETSc = _PETScNamespace()

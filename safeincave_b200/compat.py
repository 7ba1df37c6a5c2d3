"""Stand-ins that let the reference's example scripts run against this package WITHOUT editing them
(SURVEY 7 step 1 / 8b "other imports user scripts make").

    import safeincave_b200.compat as compat; compat.install()      # before the script's own imports
    import safeincave as sf                  -> safeincave_b200
    from petsc4py import PETSc               -> the KSP facade of Solver.py
    import dolfinx as do                     -> do.fem.Function(space) with a writable ``.x.array`` and ``.name``
    from mpi4py import MPI                   -> MPI.COMM_WORLD.rank / .size / Barrier / allreduce, MPI.SUM

Real installations of those packages are never shadowed: a module is only registered when it cannot be imported.
The function-space tokens ``DG0_1``, ``DG0_3x3`` (and ``C``) that the examples' ``LinearMomentum`` subclasses touch
in ``initialize()`` / ``run_after_solve()`` (examples/mechanics/1_triaxial/main.py:9-24) are provided by
``LinearMomentum`` itself (see ``Space`` / ``Function`` below)."""
from __future__ import annotations

import importlib
import sys
import types

import numpy as np


class Space:
    """What a user hook needs from a dolfinx function space: how many values a Function on it holds."""

    def __init__(self, n_entities: int, block: int, name: str = ""):
        self.n_entities, self.block, self.name = int(n_entities), int(block), name

    @property
    def size(self):
        return self.n_entities * self.block


class _X:
    def __init__(self, n):
        self.array = np.zeros(n, dtype=np.float64)

    def scatter_forward(self):
        pass


class Function:
    """``dolfinx.fem.Function`` as the examples use it: ``f.x.array[:] = ...`` and ``f.name``."""

    def __init__(self, space: Space, name: str = "f"):
        if not isinstance(space, Space):
            raise TypeError("Function expects one of the momentum equation's spaces (DG0_1, DG0_3x3, DG0_6x6, V)")
        self.function_space = space
        self.x = _X(space.size)
        self.name = name


def _missing(name):
    try:
        importlib.import_module(name)
        return False
    except Exception:
        return True


def install(force=False):
    """Register the stand-ins in ``sys.modules`` (idempotent).  Returns the list of names that were registered."""
    import safeincave_b200 as sf
    from . import Grid, HeatBC, HeatEquation, MaterialProps, MomentumBC, MomentumEquation, OutputHandler, ScreenOutput, \
        Simulators, Solver, TimeHandler, Utils
    done = []
    if force or "safeincave" not in sys.modules and _missing("safeincave"):
        sys.modules["safeincave"] = sf
        for m in (Grid, HeatBC, HeatEquation, MaterialProps, MomentumBC, MomentumEquation, OutputHandler, ScreenOutput,
                  Simulators, Solver, TimeHandler, Utils):
            sys.modules["safeincave." + m.__name__.rsplit(".", 1)[1]] = m
        done.append("safeincave")
    if force or _missing("petsc4py"):
        p = types.ModuleType("petsc4py")
        p.PETSc = Solver.PETSc
        sys.modules["petsc4py"] = p
        done.append("petsc4py")
    if force or _missing("dolfinx"):
        d = types.ModuleType("dolfinx")
        d.fem = types.ModuleType("dolfinx.fem")
        d.fem.Function = Function
        d.default_scalar_type = float
        sys.modules["dolfinx"], sys.modules["dolfinx.fem"] = d, d.fem
        done.append("dolfinx")
    if force or _missing("mpi4py"):
        m = types.ModuleType("mpi4py")
        m.MPI = types.ModuleType("mpi4py.MPI")
        m.MPI.COMM_WORLD = Grid._Comm()
        m.MPI.SUM = "sum"
        sys.modules["mpi4py"], sys.modules["mpi4py.MPI"] = m, m.MPI
        done.append("mpi4py")
    return done

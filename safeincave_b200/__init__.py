"""safeincave_b200 — B200-native implementation of SafeInCave's per-time-step mechanics hot path.

Public names mirror the reference's flat re-export (safeincave/__init__.py:16-61) for the parts
this package implements.
"""
__version__ = "0.1.0"

from .MaterialProps import (Material, NonElasticElement, Spring, Thermoelastic, Viscoelastic,  # noqa: F401
                            DislocationCreep, PressureSolutionCreep, ViscoplasticDesai, MunsonDawsonCreep,
                            MohrCoulombViscoplastic, MatsuokaNakaiViscoplastic)
from .Grid import GridHandlerGMSH  # noqa: F401
from .MomentumEquation import LinearMomentumBase, LinearMomentum, CellField  # noqa: F401
from .HeatEquation import HeatDiffusion  # noqa: F401
from .Simulators import Simulator_M, Simulator_TM, Simulator_T  # noqa: F401
from .TimeHandler import TimeControllerBase, TimeController, TimeControllerParabolic  # noqa: F401
from .Solver import KSP, PETSc  # noqa: F401
from .OutputHandler import SaveFields  # noqa: F401
from .ScreenOutput import ScreenPrinter  # noqa: F401
from . import HeatBC, MomentumBC, Utils  # noqa: F401

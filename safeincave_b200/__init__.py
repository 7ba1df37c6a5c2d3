"""safeincave_b200 — B200-native implementation of SafeInCave's per-time-step mechanics hot path.

Public names mirror the reference's flat re-export (safeincave/__init__.py:16-61) for the parts
this package implements.
"""
__version__ = "0.1.0"

from .MaterialProps import (Material, NonElasticElement, Spring, Thermoelastic, Viscoelastic,  # noqa: F401
                            DislocationCreep, PressureSolutionCreep, ViscoplasticDesai)

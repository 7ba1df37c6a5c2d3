"""Import the UNMODIFIED reference constitutive module in this container.

TEST INFRASTRUCTURE ONLY.  Used by ``oracle/gen_golden.py`` (run in the build
container, where ``/root/reference`` is mounted) to produce the golden vectors
committed under ``tests/golden/``.  Nothing at run time (tests -m gpu, smoke,
bench) may import this: ``/root/reference`` does not exist on the GPU box.

``import safeincave`` itself fails here (``safeincave/__init__.py:14-15`` pulls
in dolfinx / mpi4py).  ``safeincave/MaterialProps.py`` only needs
``dotdot_torch`` and ``MPa`` from ``safeincave/Utils.py`` (MaterialProps.py:20),
so we register a fake ``safeincave`` package whose ``Utils`` is built by
exec-ing the text span Utils.py:251-283 (``dotdot_torch``) plus the unit
constants (Utils.py:34-40), then load MaterialProps.py by path.  No reference
source is copied into this repository.
"""
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SAFEINCAVE_REFERENCE", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "safeincave", "MaterialProps.py"))


def load_reference_material_props():
    """Return the reference ``safeincave.MaterialProps`` module object."""
    if "safeincave.MaterialProps" in sys.modules and getattr(
            sys.modules["safeincave"], "__refstub__", False):
        return sys.modules["safeincave.MaterialProps"]
    pkg_dir = os.path.join(REF_ROOT, "safeincave")
    with open(os.path.join(pkg_dir, "Utils.py")) as f:
        lines = f.readlines()
    # Utils.py:251-283 is dotdot_torch (1-based, inclusive)
    span = "".join(lines[250:283])
    utils = types.ModuleType("safeincave.Utils")
    import torch as to
    utils.__dict__["to"] = to
    exec(compile(span, os.path.join(pkg_dir, "Utils.py"), "exec"), utils.__dict__)
    utils.GPa, utils.MPa, utils.kPa = 1e9, 1e6, 1e3
    utils.minute = 60
    utils.hour = 60 * utils.minute
    utils.day = 24 * utils.hour
    utils.year = 365 * utils.day
    pkg = types.ModuleType("safeincave")
    pkg.__path__ = [pkg_dir]
    pkg.__refstub__ = True
    pkg.Utils = utils
    sys.modules["safeincave"] = pkg
    sys.modules["safeincave.Utils"] = utils
    spec = importlib.util.spec_from_file_location(
        "safeincave.MaterialProps", os.path.join(pkg_dir, "MaterialProps.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["safeincave.MaterialProps"] = mod
    spec.loader.exec_module(mod)
    pkg.MaterialProps = mod
    return mod

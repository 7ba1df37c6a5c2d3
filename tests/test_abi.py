"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol that
include/safeincave_cuda.h declares (no compute calls without a GPU), and the ctypes mirrors of the
structs have the layout the header prescribes."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "safeincave_cuda.h")


@pytest.fixture(scope="module")
def lib():
    from safeincave_b200 import _lib, build
    if not os.path.isfile(_lib.LIB_PATH):
        build.build()
    return _lib.load()


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sic_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(lib):
    from safeincave_b200 import _lib
    names = declared_functions()
    assert len(names) >= 15
    assert set(names) == set(_lib.EXPORTS)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    assert lib.sic_abi_version() == _lib.SIC_ABI_VERSION


def test_struct_layouts_match_the_header(tmp_path):
    """Compile a tiny C program against the header and compare sizeof/offsetof with ctypes."""
    from safeincave_b200 import _lib
    c = tmp_path / "layout.c"
    c.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "safeincave_cuda.h"\n'
                 'int main(){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(sic_problem_t), sizeof(sic_elem_t), sizeof(sic_ksp_t),'
                 ' offsetof(sic_problem_t, elems), offsetof(sic_problem_t, T), offsetof(sic_problem_t, n_singular),'
                 ' offsetof(sic_ksp_t, op_ms), sizeof(sic_mg_level_t), offsetof(sic_mg_level_t, lambda_max),'
                 ' offsetof(sic_mg_level_t, halo), sizeof(sic_mg_opts_t), offsetof(sic_mg_opts_t, power_its_warm), sizeof(sic_heat_t), offsetof(sic_heat_t, fixed)); return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(c), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [ctypes.sizeof(_lib.SicProblem), ctypes.sizeof(_lib.SicElem), ctypes.sizeof(_lib.SicKsp),
            _lib.SicProblem.elems.offset, _lib.SicProblem.T.offset, _lib.SicProblem.n_singular.offset,
            _lib.SicKsp.op_ms.offset, ctypes.sizeof(_lib.SicMgLevel), _lib.SicMgLevel.lambda_max.offset,
            _lib.SicMgLevel.halo.offset, ctypes.sizeof(_lib.SicMgOpts), _lib.SicMgOpts.power_its_warm.offset, ctypes.sizeof(_lib.SicHeat),
            _lib.SicHeat.fixed.offset]
    assert got == want


def test_null_arguments_fail_loudly(lib):
    assert lib.sic_apply(None, None, None, None, None) != 0
    assert b"null" in lib.sic_last_error()


def test_unknown_element_is_rejected_not_emulated():
    import torch
    import safeincave_b200 as sf

    class MyCreep(sf.NonElasticElement):
        def __init__(self):
            super().__init__(4)
            self.name = "mine"

    mat = sf.Material(4)
    with pytest.raises(TypeError, match="no CPU fallback"):
        mat.add_to_non_elastic(MyCreep())


def test_material_table_deduplicates_rows():
    import torch
    import safeincave_b200 as sf
    n = 1000
    E = 102e9 * torch.ones(n)
    E[500:] = 180e9
    A = 1.9e-20 * torch.ones(n)
    A[250:750] = 0.0
    mat = sf.Material(n)
    mat.add_to_elastic(sf.Spring(E, 0.3 * torch.ones(n)))
    mat.add_to_non_elastic(sf.DislocationCreep(A, 51600 * torch.ones(n), 3 * torch.ones(n)))
    table, ids, layout = mat.build_table()
    assert table.shape[0] == 4 and ids.shape[0] == n
    # each cell's row reproduces its parameters; C follows torch's float32 promotion (T1)
    a0 = E / ((1 + 0.3 * torch.ones(n)) * (1 - 2 * 0.3 * torch.ones(n)))
    assert torch.equal(table[ids, 0], (a0 * (1 - 0.3 * torch.ones(n))).double())
    assert torch.equal(table[ids, layout["specs"][0].param_off], A.double())

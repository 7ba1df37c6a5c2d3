"""ctypes binding of libsafeincave_cuda.so (see include/safeincave_cuda.h).

There is NO fallback: if the CUDA library is missing, or an entry point returns an error, an
exception is raised.  The CPU oracle under ``oracle/`` is test infrastructure and is never
imported from here.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_int, c_int32, c_int64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsafeincave_cuda.so")       # the one library of the product: no override, no stand-in

SIC_ABI_VERSION = 13
SIC_MAX_ELEMS = 8
SIC_MAX_THERMO = 4
SIC_MAX_PEERS = 16

ELEM_KELVIN, ELEM_DISLOCATION, ELEM_PRESSURE_SOL, ELEM_DESAI = 1, 2, 3, 4
DS_ALPHA, DS_ALPHA0, DS_QSI, DS_QSI_OLD, DS_FVP, DS_R, DS_H, DS_HSMALL, DS_P, DS_ALPHA_K, DS_Q = (
    0, 1, 2, 3, 4, 5, 6, 7, 8, 14, 15)
DESAI_ROWS = 21
ELEM_MUNSON_DAWSON, ELEM_MOHR_COULOMB, ELEM_MATSUOKA_NAKAI = 5, 6, 7
MD_ZETA, MD_ZETA_OLD, MD_F, MD_ETS, MD_R, MD_H, MD_HSMALL, MD_P, MD_ZETA_K, MD_Q, MD_ROWS = 0, 1, 2, 3, 4, 5, 6, 7, 13, 14, 20
VP_FVP, VP_ROWS = 0, 1
ISV_ROWS = {ELEM_DESAI: DESAI_ROWS, ELEM_MUNSON_DAWSON: MD_ROWS, ELEM_MOHR_COULOMB: VP_ROWS, ELEM_MATSUOKA_NAKAI: VP_ROWS}
POST_STRAIN, POST_STRESS, POST_INCREMENT, POST_RATES, POST_ERROR = 1, 2, 4, 8, 16
KSP_CG, KSP_BICGSTAB, KSP_CGCG = 1, 2, 3

# every symbol include/safeincave_cuda.h declares
EXPORTS = (
    "sic_last_error", "sic_abi_version", "sic_device_info", "sic_tangent", "sic_elastic_tangent",
    "sic_post", "sic_post_blocks", "sic_commit", "sic_commit_rates", "sic_desai_initial_hardening",
    "sic_apply", "sic_residual0", "sic_block_jacobi", "sic_neumann", "sic_ksp_workspace_doubles",
    "sic_ksp_solve", "sic_guess_workspace_doubles", "sic_guess_extrapolate", "sic_fp64_peak", "sic_comm_unique_id", "sic_comm_init", "sic_comm_destroy", "sic_halo_sum",
    "sic_allreduce_sum", "sic_p2p_create", "sic_p2p_connect", "sic_p2p_destroy", "sic_p2p_error", "sic_exchange",
    "sic_mg_workspace_doubles", "sic_mg_setup", "sic_mg_solve", "sic_mg_vcycle", "sic_mg_fused_coarse_launches",
    "sic_mg_graph_captures", "sic_mg_set_fused_exchange", "sic_mg_fused_exchange_launches",
    "sic_heat_workspace_doubles", "sic_heat_step", "sic_heat_cell_mean", "sic_node_volumes", "sic_pq_fields",
)


class SicElem(ctypes.Structure):
    _fields_ = [("kind", c_int32), ("param_off", c_int32), ("eps_old", c_void_p), ("rate_old", c_void_p),
                ("rate", c_void_p), ("eps_k", c_void_p), ("desai", c_void_p)]


class SicProblem(ctypes.Structure):
    _fields_ = [
        ("abi_version", c_int32), ("n_cells", c_int32), ("cell_stride", c_int32), ("n_nodes", c_int32),
        ("conn", c_void_p), ("grad", c_void_p), ("vol", c_void_p), ("geom_tiles", c_void_p),
        ("tile_ptr", c_void_p), ("tile_nint", c_void_p), ("tile_nodes", c_void_p), ("ent_ptr", c_void_p), ("ent", c_void_p),
        ("mat_id", c_void_p), ("mat_table", c_void_p), ("n_rows", c_int32), ("row_len", c_int32),
        ("spring_off", c_int32), ("n_thermo", c_int32), ("thermo_off", c_int32), ("n_elems", c_int32),
        ("elems", SicElem * SIC_MAX_ELEMS),
        ("T", c_void_p), ("T0", c_void_p),
        ("sig", c_void_p), ("sig_k", c_void_p), ("eps", c_void_p), ("eps_prev", c_void_p),
        ("CT", c_void_p), ("eps_rhs", c_void_p), ("n_singular", c_void_p),
    ]


class SicKsp(ctypes.Structure):
    _fields_ = [("method", c_int32), ("max_it", c_int32), ("rtol", c_double), ("atol", c_double),
                ("check_every", c_int32), ("use_graph", c_int32), ("guess_nonzero", c_int32),
                ("iterations", c_int32), ("reason", c_int32),
                ("rnorm", c_double), ("rnorm0", c_double),
                ("time_operator", c_int32), ("op_samples", c_int32), ("op_ms", c_double),
                ("graph_launches", c_int32), ("direct_iterations", c_int32),
                ("op_dot_samples", c_int32), ("xchg_samples", c_int32), ("op_dot_ms", c_double), ("xchg_ms", c_double)]


class SicMgLevel(ctypes.Structure):
    _fields_ = [("prob", SicProblem), ("fixed", c_void_p), ("dinv", c_void_p), ("lambda_max", c_double),
                ("parent_a", c_void_p), ("parent_b", c_void_p), ("rst_ptr", c_void_p), ("rst_idx", c_void_p),
                ("children", c_void_p),
                ("x", c_void_p), ("b", c_void_p), ("r", c_void_p), ("d", c_void_p), ("t", c_void_p), ("pv", c_void_p),
                ("pc_ct", c_void_p), ("pc_geom", c_void_p), ("pc_dinv", c_void_p), ("pc_lidx", c_void_p), ("halo", c_void_p)]


class SicMgOpts(ctypes.Structure):
    _fields_ = [("nu", c_int32), ("coarse_its", c_int32), ("smooth_lo", c_double), ("coarse_lo", c_double),
                ("safety", c_double), ("power_its", c_int32), ("power_its_warm", c_int32), ("fused_coarse", c_int32),
                ("reserved", c_int32)]


SIC_MG_MAX_LEVELS = 8


class SicHeat(ctypes.Structure):
    _fields_ = [("n_cells", c_int32), ("cell_stride", c_int32), ("n_nodes", c_int32), ("n_tri", c_int32),
                ("conn", c_void_p), ("grad", c_void_p), ("vol", c_void_p), ("rho_cp", c_void_p), ("k", c_void_p),
                ("tri", c_void_p), ("tri_area", c_void_p), ("tri_h", c_void_p), ("tri_q", c_void_p), ("fixed", c_void_p),
                ("halo", c_void_p)]


class SicHalo(ctypes.Structure):
    _fields_ = [("n_ranks", c_int32), ("rank", c_int32), ("n_peers", c_int32), ("n_shared_total", c_int32),
                ("peer", c_int32 * SIC_MAX_PEERS), ("peer_off", c_int32 * (SIC_MAX_PEERS + 1)),
                ("idx", c_void_p), ("owner_w", c_void_p), ("send_buf", c_void_p), ("recv_buf", c_void_p),
                ("comm", c_void_p), ("p2p", c_void_p), ("tile_order", c_void_p), ("n_iface_tiles", c_int32),
                ("reserved", c_int32)]


class SicError(RuntimeError):
    pass


def declare(lib, single_gpu_only=False):
    """Attach the prototypes of include/safeincave_cuda.h to a loaded library handle."""
    lib.sic_last_error.restype = c_char_p
    lib.sic_abi_version.restype = c_int
    PP = POINTER(SicProblem)
    lib.sic_device_info.argtypes = [POINTER(c_int), POINTER(c_int), POINTER(c_int)]
    lib.sic_tangent.argtypes = [PP, c_double, c_double, c_void_p]
    lib.sic_elastic_tangent.argtypes = [PP, c_void_p]
    lib.sic_post.argtypes = [PP, c_void_p, c_double, c_double, c_double, c_int, c_void_p, c_void_p, c_void_p]
    lib.sic_post_blocks.argtypes = [c_int]
    lib.sic_commit.argtypes = [PP, c_double, c_double, c_void_p]
    lib.sic_commit_rates.argtypes = [PP, c_void_p]
    lib.sic_desai_initial_hardening.argtypes = [PP, c_int, c_double, c_void_p, c_void_p]
    lib.sic_apply.argtypes = [PP, c_void_p, c_void_p, c_void_p, c_void_p]
    PH = POINTER(SicHalo)
    lib.sic_residual0.argtypes = [PP, c_void_p, c_void_p, c_void_p, c_void_p, PH, c_void_p]
    lib.sic_block_jacobi.argtypes = [PP, c_void_p, c_void_p, PH, c_void_p]
    lib.sic_neumann.argtypes = [c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]
    lib.sic_ksp_workspace_doubles.argtypes = [c_int, c_int]
    lib.sic_ksp_workspace_doubles.restype = c_int64
    lib.sic_ksp_solve.argtypes = [PP, POINTER(SicKsp), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, PH, c_void_p]
    lib.sic_guess_workspace_doubles.argtypes = [c_int]
    lib.sic_guess_workspace_doubles.restype = c_int64
    lib.sic_guess_extrapolate.argtypes = [c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, PH, c_void_p,
                                          POINTER(c_double), c_void_p]
    lib.sic_fp64_peak.argtypes = [POINTER(c_double), c_void_p]
    PL, PO = POINTER(SicMgLevel), POINTER(SicMgOpts)
    lib.sic_mg_workspace_doubles.argtypes = [c_int, c_int]
    lib.sic_mg_workspace_doubles.restype = c_int64
    lib.sic_mg_fused_coarse_launches.restype = c_int64
    lib.sic_mg_graph_captures.restype = c_int64
    lib.sic_mg_set_fused_exchange.argtypes = [c_int]
    lib.sic_mg_set_fused_exchange.restype = None
    lib.sic_mg_fused_exchange_launches.restype = c_int64
    lib.sic_mg_setup.argtypes = [PL, c_int, PO, c_void_p, c_void_p]
    lib.sic_mg_solve.argtypes = [PL, c_int, PO, POINTER(SicKsp), c_void_p, c_void_p, c_void_p, c_void_p]
    lib.sic_mg_vcycle.argtypes = [PL, c_int, PO, c_void_p, c_void_p]
    lib.sic_node_volumes.argtypes = [PP, c_void_p, PH, c_void_p]
    lib.sic_pq_fields.argtypes = [PP, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, PH, c_void_p]
    PHT = POINTER(SicHeat)
    lib.sic_heat_workspace_doubles.argtypes = [c_int]
    lib.sic_heat_workspace_doubles.restype = c_int64
    lib.sic_heat_step.argtypes = [PHT, c_double, c_void_p, c_void_p, POINTER(SicKsp), c_void_p, c_void_p]
    lib.sic_heat_cell_mean.argtypes = [PHT, c_void_p, c_void_p, c_void_p]
    if not single_gpu_only:
        lib.sic_comm_unique_id.argtypes = [c_void_p]
        lib.sic_comm_init.argtypes = [c_void_p, c_int, c_int, POINTER(c_void_p)]
        lib.sic_comm_destroy.argtypes = [c_void_p]
        lib.sic_halo_sum.argtypes = [PH, c_void_p, c_int, c_void_p]
        lib.sic_allreduce_sum.argtypes = [c_void_p, c_void_p, c_int, c_void_p]
        lib.sic_p2p_create.argtypes = [c_int, c_int, c_int, POINTER(c_void_p), c_void_p]
        lib.sic_p2p_connect.argtypes = [c_void_p, c_void_p]
        lib.sic_p2p_destroy.argtypes = [c_void_p]
        lib.sic_p2p_error.argtypes = [c_void_p]
        lib.sic_exchange.argtypes = [PH, c_void_p, c_int, c_void_p, c_int, c_void_p]


_lib = None


def load():
    """Load the CUDA library (once).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise SicError(
            f"{LIB_PATH} not found: build it with `python -m safeincave_b200.build` "
            "(there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    missing = [s for s in EXPORTS if not hasattr(lib, s)]
    if missing:
        raise SicError(f"libsafeincave_cuda.so lacks symbols {missing}")
    lib.sic_abi_version.restype = c_int
    if lib.sic_abi_version() != SIC_ABI_VERSION:
        raise SicError("libsafeincave_cuda.so ABI version mismatch; rebuild")
    declare(lib)
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().sic_last_error().decode(errors="replace")
        raise SicError(f"{what} failed ({rc}): {msg}")

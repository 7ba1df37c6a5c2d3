"""Builds compile-time variants of the library next to the product's one (libsafeincave_cuda_<name>.so, git-ignored)
for A/B timing on the GPU box: scripts/gpu_variants.sh copies each over libsafeincave_cuda.so and runs a timing script."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from safeincave_b200 import build as B  # noqa: E402

VARIANTS = {
    "t_reg_mb3": ["-DSIC_TAN_SMEM_G=0", "-DSIC_TAN_UNROLL=0", "-DSIC_TAN_MINBLOCKS=3"],     # round 1's structure
    "t_smem_mb3": ["-DSIC_TAN_SMEM_G=1", "-DSIC_TAN_UNROLL=0", "-DSIC_TAN_MINBLOCKS=3"],
    "t_smem_mb4": ["-DSIC_TAN_SMEM_G=1", "-DSIC_TAN_UNROLL=0", "-DSIC_TAN_MINBLOCKS=4"],
    "t_smem_mb5": ["-DSIC_TAN_SMEM_G=1", "-DSIC_TAN_UNROLL=0", "-DSIC_TAN_MINBLOCKS=5"],
    "t_reg_unroll_mb2": ["-DSIC_TAN_SMEM_G=0", "-DSIC_TAN_UNROLL=1", "-DSIC_TAN_MINBLOCKS=2"],
    "t_smem_unroll_mb3": ["-DSIC_TAN_SMEM_G=1", "-DSIC_TAN_UNROLL=1", "-DSIC_TAN_MINBLOCKS=3"],
    "t_smem_unroll_mb4": ["-DSIC_TAN_SMEM_G=1", "-DSIC_TAN_UNROLL=1", "-DSIC_TAN_MINBLOCKS=4"],
    # resident CTAs per SM of the compressed preconditioner operator (k_mg_ebe_pc)
    "pc_mb4": ["-DSIC_PC_MINBLOCKS=4"],
    "pc_mb5": ["-DSIC_PC_MINBLOCKS=5"],
    "pc_mb7": ["-DSIC_PC_MINBLOCKS=7"],
    "ebedot_mb4": ["-DSIC_EBE_DOT_MINBLOCKS=4"],     # exact Krylov operator compiled for 4 resident CTAs per SM
    "geom_soa": ["-DSIC_EBE_GEOM_TILES=0"],          # exact tile kernels read conn / grad / vol from the SoA rows
    "pcmath_f64": ["-DSIC_PC_F32MATH=0"],           # stress / forces / staging of k_mg_ebe_pc in FP64 (the default is FP32)
    "pcmath_f32_mb7": ["-DSIC_PC_F32MATH=1", "-DSIC_PC_MINBLOCKS=7"],
    "pcmath_f32_mb8": ["-DSIC_PC_F32MATH=1", "-DSIC_PC_MINBLOCKS=8"],
    "pc_mb8": ["-DSIC_PC_MINBLOCKS=8"],
}

if __name__ == "__main__":
    names = sys.argv[1:] or list(VARIANTS)
    with ThreadPoolExecutor(4) as ex:
        for lib in ex.map(lambda n: B.build(defines=VARIANTS[n], suffix="_" + n), names):
            print(lib)

"""ctypes loader for libb200zk.so (the C ABI declared in include/b200zk.h).

The library has no CPU fallback: if the shared object is missing, or no CUDA device is
present when a compute entry point is called, the call raises.
"""
from __future__ import annotations

import ctypes as C
import re
from pathlib import Path

PKG_ROOT = Path(__file__).resolve().parent.parent
REPO_ROOT = PKG_ROOT.parent
import os

# B200ZK_LIB_PATH: an experimental build (build.py --variant) for A/B measurements
LIB_PATH = Path(os.environ["B200ZK_LIB_PATH"]) if os.environ.get("B200ZK_LIB_PATH") else PKG_ROOT / "lib" / "libb200zk.so"
HEADER_PATH = REPO_ROOT / "include" / "b200zk.h"
ABI_VERSION = 1


class B200zkError(RuntimeError):
    pass


_lib = None


def header_symbols() -> list:
    """Every function name include/b200zk.h declares."""
    text = HEADER_PATH.read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(b200zk_[a-z0-9_]+)\s*\(", text)))


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise B200zkError(
            f"{LIB_PATH} not found: build it with `python anon-aadhaar-halo2_b200/build.py` "
            "(there is no CPU fallback)")
    lib = C.CDLL(str(LIB_PATH))
    u64p = C.POINTER(C.c_uint64)
    vp = C.c_void_p
    sz = C.c_size_t
    u32 = C.c_uint32
    u64 = C.c_uint64
    sig = {
        "b200zk_init": ([C.c_int], C.c_int),
        "b200zk_shutdown": ([], C.c_int),
        "b200zk_last_error": ([], C.c_char_p),
        "b200zk_abi_version": ([], u32),
        "b200zk_ntt": ([vp, u32, vp], C.c_int),
        "b200zk_intt": ([vp, u32, vp, vp], C.c_int),
        "b200zk_coeff_to_extended": ([vp, u32, vp, u32, vp, vp], C.c_int),
        "b200zk_extended_to_coeff": ([vp, u32, vp, vp, vp, vp, sz], C.c_int),
        "b200zk_divide_by_vanishing": ([vp, u32, vp, u32], C.c_int),
        "b200zk_ntt_many": ([vp, sz, sz, u32, vp], C.c_int),
        "b200zk_intt_many": ([vp, sz, sz, u32, vp, vp], C.c_int),
        "b200zk_coeff_to_extended_many": ([vp, sz, vp, sz, sz, u32, u32, vp, vp], C.c_int),
        "b200zk_ntt_dev": ([vp, sz, sz, u32, vp, vp, vp], C.c_int),
        "b200zk_coeff_to_extended_dev": ([vp, sz, vp, sz, sz, u32, u32, vp, vp, vp], C.c_int),
        "b200zk_extended_to_coeff_dev": ([vp, u32, vp, vp, vp, vp, u32, vp, sz, vp], C.c_int),
        "b200zk_coeff_to_coset_dev": ([vp, sz, vp, sz, sz, u32, vp, vp, vp], C.c_int),
        "b200zk_extended_coset_slice_dev": ([vp, sz, vp, sz, sz, u32, u32, u32, vp], C.c_int),
        "b200zk_extended_coset_interleave_dev": ([vp, vp, u32, u32, u32, vp], C.c_int),
        "b200zk_ntt4_first_pass_scatter_dev": ([vp, u32, u32, vp, u32, u32, vp, sz, sz, vp], C.c_int),
        "b200zk_ntt4_twiddle_scatter_dev": ([vp, u32, u32, vp, u32, u32, vp, sz, sz, vp], C.c_int),
        "b200zk_ntt4_gather_rows_dev": ([vp, vp, u32, u32, u32, vp], C.c_int),
        "b200zk_msm_g1": ([vp, vp, sz, vp], C.c_int),
        "b200zk_bases_register": ([vp, sz, u64p], C.c_int),
        "b200zk_bases_register_ex": ([vp, sz, C.c_int, u64p], C.c_int),
        "b200zk_msm_g1_registered_many": ([u64, vp, sz, sz, sz, vp], C.c_int),
        "b200zk_msm_g1_registered_dev": ([u64, vp, sz, sz, sz, vp, vp], C.c_int),
        "b200zk_bases_evict": ([u64], C.c_int),
        "b200zk_msm_g1_registered": ([u64, vp, sz, vp], C.c_int),
        "b200zk_msm_g1_dev": ([vp, vp, sz, vp, vp], C.c_int),
        "b200zk_msm_g1_dev_async": ([vp, vp, sz, vp, vp], C.c_int),
        "b200zk_g1_sum": ([vp, sz, vp], C.c_int),
        "b200zk_dev_alloc": ([sz, u64p], C.c_int),
        "b200zk_dev_free": ([u64], C.c_int),
        "b200zk_dev_view": ([u64, sz, sz, u64p], C.c_int),
        "b200zk_dev_upload": ([u64, sz, vp, sz], C.c_int),
        "b200zk_dev_download": ([u64, sz, vp, sz], C.c_int),
        "b200zk_dev_ptr": ([u64], vp),
        "b200zk_quotient_graph": ([vp, vp, u64, u64], C.c_int),
        "b200zk_quotient_permutation": ([vp, u64, vp, vp, vp, u32, vp, u32, u32, u32, u64, u64, u64, vp, vp, vp], C.c_int),
        "b200zk_quotient_lookup": ([vp, u64, u64, u64, u64, u64, u64, u64, u64], C.c_int),
        "b200zk_batch_invert": ([vp, sz], C.c_int),
        "b200zk_batch_invert_dev": ([vp, sz, vp], C.c_int),
        "b200zk_prefix_product_dev": ([vp, sz, vp, sz, sz, sz, vp, vp], C.c_int),
        "b200zk_permutation_product_dev": ([vp, vp, u32, u32, u32, vp, vp, vp, vp, u32, vp, vp, vp], C.c_int),
        "b200zk_lookup_product_dev": ([vp, vp, vp, vp, u32, u32, vp, vp, u32, vp, vp, vp], C.c_int),
        "b200zk_permute_expression_pair_dev": ([vp, vp, sz, u32, u32, u32, vp, vp, vp, sz, vp], C.c_int),
        "b200zk_linear_combination_dev": ([vp, vp, u32, sz, vp, vp], C.c_int),
        "b200zk_eval_polynomial_dev": ([vp, sz, sz, sz, vp, vp, vp], C.c_int),
        "b200zk_kate_division_dev": ([vp, sz, vp, vp, vp], C.c_int),
        "b200zk_g1_to_bytes": ([vp, sz, vp], C.c_int),
        "b200zk_g1_affine_to_bytes": ([vp, sz, vp], C.c_int),
        "b200zk_g1_to_evm_bytes": ([vp, sz, vp], C.c_int),
        "b200zk_g1_affine_from_bytes": ([vp, sz, vp], C.c_int),
        "b200zk_gen_scalars_dev": ([vp, sz, u64, sz], C.c_int),
        "b200zk_gen_points_dev": ([vp, sz, u64, sz], C.c_int),
        "b200zk_modmul_peak": ([u32, C.POINTER(C.c_double)], C.c_int),
        "b200zk_kernel_launches": ([], u64),
        "b200zk_msm_profile": ([C.c_int], C.c_int),
        "b200zk_host_register": ([vp, sz], C.c_int),
        "b200zk_host_unregister": ([vp], C.c_int),
        "b200zk_host_alloc": ([sz, C.POINTER(C.c_void_p)], C.c_int),
        "b200zk_host_free": ([vp], C.c_int),
        "b200zk_msm_tune": ([u32, u32, u32], C.c_int),
        "b200zk_msm_upload_pipeline": ([u32, sz], C.c_int),
        "b200zk_msm_upload_ranges": ([sz, u32, C.c_double, vp, vp], C.c_int),
        "b200zk_field_op": ([u32, u32, vp, vp, sz, vp], C.c_int),
        "b200zk_ntt_transfer_pipeline": ([u32, u32], C.c_int),
        "b200zk_ntt_tune": ([u32], C.c_int),
        "b200zk_stream_release": ([vp], C.c_int),
        "b200zk_mirror_enable": ([sz], C.c_int),
        "b200zk_mirror_invalidate": ([vp, sz], C.c_int),
        "b200zk_mirror_stats": ([u64p], C.c_int),
        "b200zk_msm_last_stages": ([C.POINTER(C.c_float), C.c_int, u64p], C.c_int),
    }
    for name, (argtypes, restype) in sig.items():
        try:
            fn = getattr(lib, name)
        except AttributeError:
            continue  # reported by tests/test_abi.py
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.b200zk_abi_version() != ABI_VERSION:
        raise B200zkError("libb200zk.so ABI version mismatch: rebuild the library")
    _lib = lib
    return lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().b200zk_last_error()
        raise B200zkError(msg.decode() if msg else f"b200zk error {rc}")

"""ctypes front end of libb200zk_witness.so (include/b200zk_witness.h): the witness integers of the
reference's `BigUintConfig::pow_mod_fixed_exp` (reference src/big_uint/chip.rs:454-490), computed
natively on all host cores for a batch of RSA verifications."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from ._lib import PKG_ROOT, B200zkError

LIB_PATH = PKG_ROOT / "lib" / "libb200zk_witness.so"
_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise B200zkError(f"{LIB_PATH} not found: build it with `python anon-aadhaar-halo2_b200/build.py`")
        lib = C.CDLL(str(LIB_PATH))
        lib.b200zk_witness_mul_mod_words.argtypes, lib.b200zk_witness_mul_mod_words.restype = [C.c_uint32], C.c_size_t
        lib.b200zk_witness_pow_steps.argtypes, lib.b200zk_witness_pow_steps.restype = [C.c_uint64], C.c_uint32
        lib.b200zk_witness_pow_words.argtypes, lib.b200zk_witness_pow_words.restype = [C.c_uint64, C.c_uint32], C.c_size_t
        lib.b200zk_witness_pow_mod_fixed_exp.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p]
        lib.b200zk_witness_pow_mod_fixed_exp.restype = C.c_int
        lib.b200zk_witness_words_to_fr.argtypes = [C.c_void_p, C.c_uint32, C.c_size_t, C.c_int, C.c_void_p]
        lib.b200zk_witness_words_to_fr.restype = C.c_int
        _lib = lib
    return _lib


@dataclass
class MulModRecord:
    """Views into one mul_mod record (include/b200zk_witness.h)."""
    a: np.ndarray
    b: np.ndarray
    q: np.ndarray
    r: np.ndarray
    ab: np.ndarray      # (2L - 1, 3)
    qn: np.ndarray      # (2L - 1, 3)
    carry: np.ndarray   # (2L - 1, 2)
    c: np.ndarray       # (2L - 1,)


def split_record(rec: np.ndarray, num_limbs: int) -> MulModRecord:
    L, M = num_limbs, 2 * num_limbs - 1
    o = [0, L, 2 * L, 3 * L, 4 * L, 4 * L + 3 * M, 4 * L + 6 * M, 4 * L + 8 * M, 4 * L + 9 * M]
    return MulModRecord(rec[o[0]:o[1]], rec[o[1]:o[2]], rec[o[2]:o[3]], rec[o[3]:o[4]], rec[o[4]:o[5]].reshape(M, 3),
                        rec[o[5]:o[6]].reshape(M, 3), rec[o[6]:o[7]].reshape(M, 2), rec[o[7]:o[8]])


def pow_mod_fixed_exp(base: np.ndarray, modulus: np.ndarray, e: int = 65537, threads: int = 0):
    """base, modulus: (count, num_limbs) uint64 little-endian limbs.  Returns (records, result):
    records (count, steps, words_per_record) and result (count, num_limbs) = base^e mod modulus."""
    lib = load()
    base = np.ascontiguousarray(base, dtype=np.uint64)
    modulus = np.ascontiguousarray(modulus, dtype=np.uint64)
    assert base.shape == modulus.shape and base.ndim == 2
    count, L = base.shape
    steps, rec, per = lib.b200zk_witness_pow_steps(e), lib.b200zk_witness_mul_mod_words(L), lib.b200zk_witness_pow_words(e, L)
    out = np.zeros((count, per), dtype=np.uint64)
    rc = lib.b200zk_witness_pow_mod_fixed_exp(base.ctypes.data, modulus.ctypes.data, e, L, count, threads, out.ctypes.data)
    if rc == 2:
        raise B200zkError("b200zk_witness: base >= modulus (the chip asserts x < n, reference src/chip.rs:88)")
    if rc:
        raise B200zkError(f"b200zk_witness_pow_mod_fixed_exp failed ({rc})")
    return out[:, : steps * rec].reshape(count, steps, rec), out[:, steps * rec:]


def words_to_fr(words: np.ndarray, threads: int = 0) -> np.ndarray:
    """(count, width) canonical little-endian integers below r -> (count, 4) `bn256::Fr` Montgomery limbs."""
    words = np.ascontiguousarray(words, dtype=np.uint64)
    if words.ndim == 1:
        words = words.reshape(-1, 1)
    out = np.zeros((words.shape[0], 4), dtype=np.uint64)
    rc = load().b200zk_witness_words_to_fr(words.ctypes.data, words.shape[1], words.shape[0], threads, out.ctypes.data)
    if rc:
        raise B200zkError(f"b200zk_witness_words_to_fr failed ({rc})")
    return out

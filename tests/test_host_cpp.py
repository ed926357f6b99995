"""The C++ host mirror (anon-aadhaar-halo2_b200/host/halo2_b200.hpp): constants on the CPU,
the whole commit -> iNTT -> coset NTT -> divide -> inverse coset NTT -> commit chain on the GPU."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import bn254 as bn
from oracle import c_oracle as co
from oracle import halo2_cpu as h
from util import fr1, jac_affine

ROOT = Path(__file__).resolve().parent.parent
HOST = ROOT / "anon-aadhaar-halo2_b200" / "host"


def build_example() -> Path:
    import build as zkbuild  # anon-aadhaar-halo2_b200/build.py

    zkbuild.build()
    subprocess.run(["make", "-s", "-C", str(HOST)], check=True)
    return HOST / "example_prover_path"


def parse(out: str) -> dict:
    d = {}
    key = None
    for line in out.splitlines():
        parts = line.split()
        if not parts:
            continue
        if line.startswith("  "):
            d[f"{key}.{parts[0]}"] = [int(x, 16) for x in parts[1:]]
        else:
            key = parts[0]
            d.setdefault(key, [])
            d[key].append(parts[1:])
    return d


def limbs(words) -> np.ndarray:
    return np.array([int(w, 16) for w in words], dtype=np.uint64)


def digest(a: np.ndarray) -> int:
    hsh = 1469598103934665603
    for l in a.reshape(-1).tolist():
        hsh ^= l
        hsh = hsh * 1099511628211 % (1 << 64)
    return hsh


@pytest.mark.parametrize("k,j", [(4, 4), (15, 5), (10, 9)])
def test_cpp_domain_constants_match_oracle(k, j):
    exe = build_example()
    out = parse(subprocess.run([str(exe), "--constants", str(k), str(j)], check=True, capture_output=True, text=True).stdout)
    o = h.EvaluationDomain(j, k)
    for name in ("omega", "omega_inv", "extended_omega", "extended_omega_inv", "g_coset", "g_coset_inv",
                 "ifft_divisor", "extended_ifft_divisor"):
        assert bn.fr_array_to_canonical(limbs(out[name][0])[None, :])[0] == getattr(o, name), name
    assert int(out["extended_k"][0][0]) == o.extended_k
    tev = [bn.fr_array_to_canonical(limbs(t)[None, :])[0] for t in out["t"]]
    assert tev == o.t_evaluations


@pytest.mark.gpu
def test_cpp_prover_path_matches_oracle(tmp_path):
    exe = build_example()
    k, j = 11, 4
    n = 1 << k
    col, gl = co.gen_scalars(41, n), co.gen_points(42, n)
    (tmp_path / "s.bin").write_bytes(col.tobytes())
    (tmp_path / "b.bin").write_bytes(gl.tobytes())
    res = subprocess.run([str(exe), str(k), str(j), str(tmp_path / "s.bin"), str(tmp_path / "b.bin")],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    out = parse(res.stdout)
    o = h.EvaluationDomain(j, k)
    L = fr1
    coeff = co.ifft(col, L(o.omega_inv), k, L(o.ifft_divisor))
    ext = co.coeff_to_extended(coeff, k, o.extended_k, L(o.extended_omega), L(bn.FR_ZETA))
    hdiv = co.divide_by_vanishing(ext, o.extended_k, bn.fr_array_from_canonical(o.t_evaluations))
    back = co.extended_to_coeff(ext, o.extended_k, L(o.extended_omega_inv), L(o.extended_ifft_divisor), L(bn.FR_ZETA),
                                o.n * o.quotient_poly_degree)
    for name, arr in (("lagrange_to_coeff", coeff), ("coeff_to_extended", ext), ("divide_by_vanishing_poly", hdiv),
                      ("extended_to_coeff", back)):
        size, dg = out[name][0]
        assert int(size) == arr.shape[0] and int(dg, 16) == digest(arr), name
        assert out[f"{name}.first"] == arr[0].tolist() and out[f"{name}.last"] == arr[-1].tolist()
    assert jac_affine(limbs(out["commit_lagrange"][0])) == jac_affine(co.best_multiexp(col, gl))
    assert jac_affine(limbs(out["commit"][0])) == jac_affine(co.best_multiexp(coeff, gl))
    assert jac_affine(limbs(out["best_multiexp"][0])) == jac_affine(co.best_multiexp(coeff, gl))
    assert out["best_fft_roundtrip"][0] == ["1"] and out["length_assert"][0] == ["1"]

"""The hot-path call sequence of one `create_proof` at the shape of the reference's
RSA-SHA256 circuit, on synthetic columns (SURVEY.md section 8 d, "Config 3 shape").

The reference never runs the real prover and its circuits cannot be compiled
here (F4), so the witness is replaced by seeded random columns of the same *shape*:
k = 15 (reference src/lib.rs:444), constraint degree 4 => extended k = 17, about 112 advice
columns, 24 lookups, 84 fixed columns, 115 permutation columns (58 grand products of chunk 2),
3 quotient pieces.  What is timed is exactly the sequence of `ParamsKZG::commit_lagrange`,
`lagrange_to_coeff`, `coeff_to_extended`, `evaluate_h`, `divide_by_vanishing_poly`,
`extended_to_coeff` and `commit` calls `create_proof` makes ([DEP] halo2_proofs
src/plonk/prover.rs @ v2023_01_20), batched per prover phase, all device-resident.  Witness
synthesis, transcripts, grand-product construction and the SHPLONK opening are not part of
the hot path and are not timed.
"""
from __future__ import annotations

import ctypes as C
import time
from dataclasses import dataclass

import numpy as np

from ._lib import check, load
from .api import EvaluationDomain, FR_ZETA, _ptr, fr_limbs, host_alloc_fr, host_free
from .quotient import FR_DELTA, DeviceColumn, EnvC, FlatGraph, _handles, _ptr32


@dataclass
class Shape:
    k: int = 15
    degree: int = 4                 # cs.degree(): extended k = k + 2, chunk = degree - 2 = 2
    advice: int = 112
    gate_columns: int = 80          # halo2-base FlexGate columns: q * (a0 + a1 * a2 - a3)
    lookups: int = 24
    fixed: int = 84
    instance: int = 2
    permutation_columns: int = 115
    blinding_factors: int = 5

    @property
    def permutation_sets(self) -> int:
        chunk = self.degree - 2
        return -(-self.permutation_columns // chunk)

    @property
    def lagrange_columns(self) -> int:
        """Columns committed in Lagrange basis and then taken to coefficient form."""
        return self.advice + 3 * self.lookups + self.permutation_sets


RSA_SHA256 = Shape()
SMALL = Shape(k=8, advice=12, gate_columns=8, lookups=3, fixed=10, instance=1, permutation_columns=9)

# b200zk_src kinds / b200zk_calc ops (include/b200zk.h)
_CONST, _INTER, _FIXED, _ADVICE, _INSTANCE, _CHALL, _BETA, _GAMMA, _THETA, _Y, _PREV = range(11)
_ADD, _SUB, _MUL, _SQUARE, _DOUBLE, _NEGATE, _HORNER, _STORE = range(8)


def gate_graph(shape: Shape) -> FlatGraph:
    """`GraphEvaluator` of the custom gates, flattened as upstream would produce it: one
    FlexGate-style constraint q_c * (a_c + a_c[+1] * a_c[+2] - a_c[+3]) per gate column, each
    sub-expression its own intermediate, then the y-Horner over all constraints."""
    rotations = np.array([0, 1, 2, 3], dtype=np.int32)
    calcs, parts = [], []
    t = 0
    for c in range(shape.gate_columns):
        q = c % shape.fixed
        calcs.append([_MUL, t, _ADVICE, c, 1, _ADVICE, c, 2, 0, 0])
        calcs.append([_ADD, t + 1, _ADVICE, c, 0, _INTER, t, 0, 0, 0])
        calcs.append([_SUB, t + 2, _INTER, t + 1, 0, _ADVICE, c, 3, 0, 0])
        calcs.append([_MUL, t + 3, _FIXED, q, 0, _INTER, t + 2, 0, 0, 0])
        parts.append([_INTER, t + 3, 0])
        t += 4
    calcs.append([_HORNER, t, _PREV, 0, 0, _Y, 0, 0, 0, len(parts)])
    return FlatGraph(np.zeros((1, 4), np.uint64), rotations, np.array(calcs, np.uint32),
                     np.array(parts, np.uint32), t + 1)


def lookup_graph(shape: Shape, j: int) -> FlatGraph:
    """(compressed_input + beta) * (compressed_table + gamma) for lookup j: a two-expression
    input (theta-compressed) against a two-column table."""
    a0 = shape.gate_columns + (j % max(1, shape.advice - shape.gate_columns))
    a1 = (a0 + 1) % shape.advice
    f0, f1 = (shape.fixed - 1 - j) % shape.fixed, (shape.fixed - 2 - j) % shape.fixed
    calcs = [
        [_HORNER, 0, _CONST, 0, 0, _THETA, 0, 0, 0, 2],      # ((0 * theta) + in0) * theta + in1
        [_HORNER, 1, _CONST, 0, 0, _THETA, 0, 0, 2, 2],
        [_ADD, 2, _INTER, 0, 0, _BETA, 0, 0, 0, 0],
        [_ADD, 3, _INTER, 1, 0, _GAMMA, 0, 0, 0, 0],
        [_MUL, 4, _INTER, 2, 0, _INTER, 3, 0, 0, 0],
    ]
    parts = [[_ADVICE, a0, 0], [_ADVICE, a1, 0], [_FIXED, f0, 0], [_FIXED, f1, 0]]
    return FlatGraph(np.zeros((1, 4), np.uint64), np.array([0], np.int32), np.array(calcs, np.uint32),
                     np.array(parts, np.uint32), 5)


class ProverHotPath:
    """Device-resident state of one proof's hot path + `run()`."""

    def __init__(self, shape: Shape = RSA_SHA256, seed: int = 0x5EED0000, sync=None):
        self.shape, self.lib = shape, load()
        self.sync = sync or (lambda: None)
        lib = self.lib
        s = shape
        self.domain = d = EvaluationDomain(s.degree, s.k)
        self.n, self.N = d.n, d.extended_len()
        n, N = self.n, self.N
        # ---- ParamsKZG: synthetic SRS (g and g_lagrange share one table here)
        pts = DeviceColumn(2 * n)                       # n points = 2n field elements
        check(lib.b200zk_gen_points_dev(C.c_void_p(pts.ptr), n, seed, 0))
        hp = pts.to_host().reshape(n, 8)
        pts.free()
        h = C.c_uint64(0)
        check(lib.b200zk_bases_register(_ptr(hp), n, C.byref(h)))
        self.h_bases = h.value
        # ---- proving key: extended-domain fixed / sigma / l_0 / l_last / l_active columns
        self.pk_cols = DeviceColumn((s.fixed + s.permutation_columns + 3) * N)
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.pk_cols.ptr), self.pk_cols.n, seed + 1, 0))
        views = [DeviceColumn.view(self.pk_cols, i * N, N) for i in range(s.fixed + s.permutation_columns + 3)]
        self.fixed = views[: s.fixed]
        self.sigma = views[s.fixed: s.fixed + s.permutation_columns]
        self.l0, self.l_last, self.l_active = views[-3:]
        # ---- witness-shaped columns (Lagrange basis), one allocation per prover phase
        self.n_lag = s.lagrange_columns
        self.lag = DeviceColumn(self.n_lag * n)
        self.instance_coeff = DeviceColumn(s.instance * n)
        self.ext = DeviceColumn((self.n_lag + s.instance) * N)
        self.ext_views = [DeviceColumn.view(self.ext, i * N, N) for i in range(self.n_lag + s.instance)]
        self.values = DeviceColumn(N)
        self.table = DeviceColumn(N)
        self.h_coeff = DeviceColumn(n * (s.degree - 1))
        self.points = DeviceColumn((self.n_lag + 1 + s.degree - 1) * 3)   # 12 limbs = 3 field elements each
        self.seed = seed
        # the advice columns come from witness synthesis on the CPU: a page-locked host copy that
        # run() uploads inside its timed region (the other Lagrange columns are derived on the device)
        self.host_advice = host_alloc_fr(s.advice * n)
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.lag.ptr), s.advice * n, seed + 2, 0))
        check(lib.b200zk_dev_download(self.lag.handle, 0, _ptr(self.host_advice), s.advice * n))
        self.gates = gate_graph(s)
        self.lookup_graphs = [lookup_graph(s, j) for j in range(s.lookups)]
        rnd = np.random.Generator(np.random.PCG64(seed))
        self.scalars = {nm: fr_limbs(int.from_bytes(rnd.bytes(31), "little")) for nm in ("beta", "gamma", "theta", "y")}
        perm_cols = [("advice", i % s.advice) for i in range(s.permutation_columns - s.instance)]
        perm_cols += [("instance", i) for i in range(s.instance)]
        self.perm_kind = np.array([{"fixed": 2, "advice": 3, "instance": 4}[k] for k, _ in perm_cols], np.uint32)
        self.perm_index = np.array([i for _, i in perm_cols], np.uint32)

    # column order inside `lag` / `ext`: advice | permuted input, permuted table (per lookup) |
    # permutation products | lookup products
    def _env(self):
        s = self.shape
        advice = self.ext_views[: s.advice]
        instance = self.ext_views[self.n_lag: self.n_lag + s.instance]
        keep = [_handles(self.fixed), _handles(advice), _handles(instance), np.zeros((1, 4), np.uint64)]
        env = EnvC()
        env.fixed, env.n_fixed = keep[0].ctypes.data, len(self.fixed)
        env.advice, env.n_advice = keep[1].ctypes.data, len(advice)
        env.instance, env.n_instance = keep[2].ctypes.data, len(instance)
        env.challenges, env.n_challenges = keep[3].ctypes.data, 0
        for name in ("beta", "gamma", "theta", "y"):
            getattr(env, name)[:] = [int(x) for x in self.scalars[name]]
        env.k, env.ext_k = self.domain.k, self.domain.extended_k
        env.range_begin, env.range_len = 0, 0
        env._keep = keep
        return env

    def _commit(self, ptr: int, count: int, out_index: int, length: int = None) -> None:
        length = length or self.n
        out = self.points.ptr + out_index * 96
        check(self.lib.b200zk_msm_g1_registered_dev(self.h_bases, C.c_void_p(ptr), length, count, length,
                                                    C.c_void_p(out), None))

    def run(self) -> dict:
        """One proof's hot path; returns wall-clock milliseconds per stage (device synchronised
        at every stage boundary)."""
        lib, s, d, n, N = self.lib, self.shape, self.domain, self.n, self.N
        t = {}
        marks = [time.perf_counter()]

        def mark(name):
            self.sync()
            marks.append(time.perf_counter())
            t[name] = 1e3 * (marks[-1] - marks[-2])

        # fresh witness-shaped columns (not timed: stands in for synthesis)
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.lag.ptr), self.lag.n, self.seed + 2, 0))
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.instance_coeff.ptr), self.instance_coeff.n, self.seed + 3, 0))
        self.sync()
        marks[0] = time.perf_counter()
        # ---- the advice columns arrive from the host (witness synthesis)
        check(lib.b200zk_dev_upload(self.lag.handle, 0, _ptr(self.host_advice), s.advice * n))
        mark("upload_advice")
        # ---- commitments in Lagrange basis, one batch per prover phase
        col = 0
        for count in (s.advice, 2 * s.lookups, s.permutation_sets + s.lookups):
            self._commit(self.lag.ptr + col * n * 32, count, col)
            col += count
        mark("commit_lagrange")
        # ---- lagrange_to_coeff, all columns
        check(lib.b200zk_ntt_dev(C.c_void_p(self.lag.ptr), n, self.n_lag, d.k, _ptr(d.omega_inv),
                                 _ptr(d.ifft_divisor), None))
        mark("lagrange_to_coeff")
        # ---- coeff_to_extended, all columns + instances
        check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.lag.ptr), n, C.c_void_p(self.ext.ptr), N, self.n_lag,
                                               d.k, d.extended_k, _ptr(d.extended_omega), _ptr(d.g_coset), None))
        check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.instance_coeff.ptr), n,
                                               C.c_void_p(self.ext.ptr + self.n_lag * N * 32), N, s.instance, d.k,
                                               d.extended_k, _ptr(d.extended_omega), _ptr(d.g_coset), None))
        mark("coeff_to_extended")
        # ---- evaluate_h
        env = self._env()
        g = self.gates.as_c()
        check(lib.b200zk_quotient_graph(C.byref(g), C.byref(env), 0, self.values.handle))
        mark("quotient_gates")
        lk0 = s.advice
        pp0 = s.advice + 2 * s.lookups
        lp0 = pp0 + s.permutation_sets
        products = self.ext_views[pp0: pp0 + s.permutation_sets]
        sig, ph = _handles(self.sigma), _handles(products)
        zeta, delta = fr_limbs(FR_ZETA), fr_limbs(FR_DELTA)
        check(lib.b200zk_quotient_permutation(
            C.byref(env), self.values.handle, _ptr32(self.perm_kind), _ptr32(self.perm_index), _ptr(sig),
            len(self.perm_kind), _ptr(ph), len(products), s.degree - 2, s.blinding_factors, self.l0.handle,
            self.l_last.handle, self.l_active.handle, _ptr(d.extended_omega), _ptr(zeta), _ptr(delta)))
        mark("quotient_permutation")
        for j in range(s.lookups):
            lg = self.lookup_graphs[j].as_c()
            check(lib.b200zk_quotient_graph(C.byref(lg), C.byref(env), 0, self.table.handle))
            check(lib.b200zk_quotient_lookup(C.byref(env), self.values.handle, self.table.handle,
                                             self.ext_views[lp0 + j].handle, self.ext_views[lk0 + 2 * j].handle,
                                             self.ext_views[lk0 + 2 * j + 1].handle, self.l0.handle,
                                             self.l_last.handle, self.l_active.handle))
        mark("quotient_lookups")
        # ---- h(X) = numerator / (X^n - 1), back to coefficients, commit the pieces
        t_ev = DeviceColumn.from_host(d.t_evaluations)
        keep = n * (s.degree - 1)
        check(lib.b200zk_extended_to_coeff_dev(C.c_void_p(self.values.ptr), d.extended_k, _ptr(d.extended_omega_inv),
                                               _ptr(d.extended_ifft_divisor), _ptr(d.g_coset), C.c_void_p(t_ev.ptr),
                                               d.t_evaluations.shape[0], C.c_void_p(self.h_coeff.ptr), keep, None))
        mark("divide_and_extended_to_coeff")
        self._commit(self.h_coeff.ptr, s.degree - 1, self.n_lag + 1)
        mark("commit_h_pieces")
        self.commitments = self.points.to_host()           # every commitment of the proof back on the host
        mark("download_commitments")
        t_ev.free()
        t["total"] = sum(t.values())
        return t

    # ------------------------------------------------------------ the same proof, one host-pointer call at a time
    # What an *untouched* `create_proof` does through the patched halo2_proofs bodies
    # (rust/halo2_proofs_patch/arithmetic_patch.rs): every polynomial is a host `Vec<Fr>`, every call is
    # synchronous and takes / returns host memory.  Call sequence:
    #   commit_lagrange(p) for every advice / permuted / product column          b200zk_msm_g1_registered
    #   lagrange_to_coeff(p) for each of them                                     b200zk_intt
    #   coeff_to_extended(z) for the permutation products (kept by `Committed`)   b200zk_coeff_to_extended
    #   evaluate_h(pk, advice, instance, .., lookups, permutations)               its replaced body: uploads of the
    #       coefficient columns and product cosets it is handed, device transforms + quotient kernels, one download
    #   divide_by_vanishing_poly, extended_to_coeff, commit(h piece) x (d - 1)    the host-pointer calls
    # With `mirror=True` the library keeps device mirrors of the host polynomials (b200zk_mirror_enable), so a
    # polynomial is uploaded by the first call that sees it and found in HBM by the later ones.
    def prepare_percall(self, pinned: bool = True) -> None:
        s, n, N = self.shape, self.n, self.N
        alloc = host_alloc_fr if pinned else (lambda count: np.zeros((count, 4), dtype=np.uint64))
        self._percall_pinned = pinned
        self.h_lag = alloc(self.n_lag * n)
        self.h_instance = alloc(s.instance * n)
        self.h_prod_ext = alloc(s.permutation_sets * N)
        self.h_values = alloc(N)
        self.h_hcoeff = alloc(n * (s.degree - 1))
        self.h_points = np.zeros((self.n_lag + s.degree - 1, 12), dtype=np.uint64)

    def run_percall(self, mirror: bool = False, mirror_bytes: int = 8 << 30) -> dict:
        lib, s, d, n, N = self.lib, self.shape, self.domain, self.n, self.N
        if not hasattr(self, "h_lag"):
            self.prepare_percall()
        t = {}
        marks = [time.perf_counter()]

        def mark(name):
            marks.append(time.perf_counter())          # every call below is synchronous
            t[name] = 1e3 * (marks[-1] - marks[-2])

        # fresh witness-shaped columns on the host (not timed: stands in for synthesis)
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.lag.ptr), self.lag.n, self.seed + 2, 0))
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.instance_coeff.ptr), self.instance_coeff.n, self.seed + 3, 0))
        check(lib.b200zk_dev_download(self.lag.handle, 0, _ptr(self.h_lag), self.lag.n))
        check(lib.b200zk_dev_download(self.instance_coeff.handle, 0, _ptr(self.h_instance), self.instance_coeff.n))
        if mirror:
            check(lib.b200zk_mirror_enable(mirror_bytes))
            check(lib.b200zk_mirror_invalidate(None, 0))   # a new proof: no mirror survives from the set-up above
        else:
            check(lib.b200zk_mirror_enable(0))
        col = lambda a, c, width: C.c_void_p(a.ctypes.data + c * width * 32)
        pp0 = s.advice + 2 * s.lookups
        lp0 = pp0 + s.permutation_sets
        self.sync()
        marks[0] = time.perf_counter()
        # ---- ParamsKZG::commit_lagrange, one polynomial per call
        for c in range(self.n_lag):
            check(lib.b200zk_msm_g1_registered(self.h_bases, col(self.h_lag, c, n), n, _ptr(self.h_points[c])))
        mark("commit_lagrange")
        # ---- EvaluationDomain::lagrange_to_coeff, in place on each host polynomial
        for c in range(self.n_lag):
            check(lib.b200zk_intt(col(self.h_lag, c, n), d.k, _ptr(d.omega_inv), _ptr(d.ifft_divisor)))
        mark("lagrange_to_coeff")
        # ---- permutation::Argument::commit keeps the extended form of every product polynomial
        for i in range(s.permutation_sets):
            check(lib.b200zk_coeff_to_extended(col(self.h_lag, pp0 + i, n), d.k, col(self.h_prod_ext, i, N), d.extended_k,
                                               _ptr(d.extended_omega), _ptr(d.g_coset)))
        mark("coeff_to_extended_products")
        # ---- Evaluator::evaluate_h (replaced body): stage what it is handed, extend, evaluate, return one column
        for c in list(range(pp0)) + list(range(lp0, self.n_lag)):
            check(lib.b200zk_dev_upload(self.lag.handle, c * n, col(self.h_lag, c, n), n))
        for c in range(s.instance):
            check(lib.b200zk_dev_upload(self.instance_coeff.handle, c * n, col(self.h_instance, c, n), n))
        for i in range(s.permutation_sets):
            check(lib.b200zk_dev_upload(self.ext.handle, (pp0 + i) * N, col(self.h_prod_ext, i, N), N))
        for first, count in ((0, pp0), (lp0, self.n_lag - lp0)):
            check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.lag.ptr + first * n * 32), n,
                                                   C.c_void_p(self.ext.ptr + first * N * 32), N, count, d.k, d.extended_k,
                                                   _ptr(d.extended_omega), _ptr(d.g_coset), None))
        check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.instance_coeff.ptr), n,
                                               C.c_void_p(self.ext.ptr + self.n_lag * N * 32), N, s.instance, d.k,
                                               d.extended_k, _ptr(d.extended_omega), _ptr(d.g_coset), None))
        self._quotient_kernels()
        check(lib.b200zk_dev_download(self.values.handle, 0, _ptr(self.h_values), N))
        mark("evaluate_h")
        # ---- vanishing::Argument::construct
        check(lib.b200zk_divide_by_vanishing(_ptr(self.h_values), d.extended_k, _ptr(d.t_evaluations),
                                             d.t_evaluations.shape[0]))
        keep = n * (s.degree - 1)
        check(lib.b200zk_extended_to_coeff(_ptr(self.h_values), d.extended_k, _ptr(d.extended_omega_inv),
                                           _ptr(d.extended_ifft_divisor), _ptr(d.g_coset), _ptr(self.h_hcoeff), keep))
        mark("divide_and_extended_to_coeff")
        for j in range(s.degree - 1):
            check(lib.b200zk_msm_g1_registered(self.h_bases, col(self.h_hcoeff, j, n), n, _ptr(self.h_points[self.n_lag + j])))
        mark("commit_h_pieces")
        t["total"] = 1e3 * (marks[-1] - marks[0])
        if mirror:
            st = (C.c_uint64 * 4)()
            check(lib.b200zk_mirror_stats(st))
            self.mirror_stats = {"hits": int(st[0]), "misses": int(st[1]), "resident_bytes": int(st[2]), "evictions": int(st[3])}
            check(lib.b200zk_mirror_invalidate(None, 0))   # the pool of device blocks stays for the next proof
        return t

    def percall_counts(self) -> dict:
        s = self.shape
        return {"b200zk_msm_g1_registered": self.n_lag + s.degree - 1, "b200zk_intt": self.n_lag,
                "b200zk_coeff_to_extended": s.permutation_sets,
                "b200zk_dev_upload (inside evaluate_h)": self.n_lag + s.instance,
                "b200zk_divide_by_vanishing": 1, "b200zk_extended_to_coeff": 1}

    def _quotient_kernels(self) -> None:
        """The gate, permutation and lookup blocks of evaluate_h over the extended columns (as in run())."""
        lib, s, d = self.lib, self.shape, self.domain
        env = self._env()
        g = self.gates.as_c()
        check(lib.b200zk_quotient_graph(C.byref(g), C.byref(env), 0, self.values.handle))
        lk0 = s.advice
        pp0 = s.advice + 2 * s.lookups
        lp0 = pp0 + s.permutation_sets
        products = self.ext_views[pp0: pp0 + s.permutation_sets]
        sig, ph = _handles(self.sigma), _handles(products)
        zeta, delta = fr_limbs(FR_ZETA), fr_limbs(FR_DELTA)
        check(lib.b200zk_quotient_permutation(
            C.byref(env), self.values.handle, _ptr32(self.perm_kind), _ptr32(self.perm_index), _ptr(sig),
            len(self.perm_kind), _ptr(ph), len(products), s.degree - 2, s.blinding_factors, self.l0.handle,
            self.l_last.handle, self.l_active.handle, _ptr(d.extended_omega), _ptr(zeta), _ptr(delta)))
        for j in range(s.lookups):
            lg = self.lookup_graphs[j].as_c()
            check(lib.b200zk_quotient_graph(C.byref(lg), C.byref(env), 0, self.table.handle))
            check(lib.b200zk_quotient_lookup(C.byref(env), self.values.handle, self.table.handle,
                                             self.ext_views[lp0 + j].handle, self.ext_views[lk0 + 2 * j].handle,
                                             self.ext_views[lk0 + 2 * j + 1].handle, self.l0.handle,
                                             self.l_last.handle, self.l_active.handle))

    # ------------------------------------------------------------ the same proof with the phases overlapped
    def run_overlapped(self, torch, hi_stream, lo_stream) -> dict:
        """The stages of run() as a prover that owns both sides would issue them: the commitments on a
        high-priority stream, the transforms of the same columns (lagrange_to_coeff on a copy, then
        coeff_to_extended) on a lower-priority stream, whose CTAs fill the SM slots the commitments'
        latency-bound bucket reductions leave idle; the quotient waits for both.  Same results as
        run() (tests/test_prover_shape_gpu.py); returns the wall-clock total only, since the stages no
        longer have separate times."""
        lib, s, d, n, N = self.lib, self.shape, self.domain, self.n, self.N
        hi, lo = C.c_void_p(hi_stream.cuda_stream), C.c_void_p(lo_stream.cuda_stream)
        if not hasattr(self, "coef"):
            self.coef = DeviceColumn(self.n_lag * n)
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.lag.ptr), self.lag.n, self.seed + 2, 0))
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.instance_coeff.ptr), self.instance_coeff.n, self.seed + 3, 0))
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        check(lib.b200zk_dev_upload(self.lag.handle, 0, _ptr(self.host_advice), s.advice * n))
        lag_t = self._torch_view(torch, self.lag, 0, self.n_lag * n)
        coef_t = self._torch_view(torch, self.coef, 0, self.n_lag * n)
        with torch.cuda.stream(lo_stream):
            coef_t.copy_(lag_t, non_blocking=True)
        check(lib.b200zk_ntt_dev(C.c_void_p(self.coef.ptr), n, self.n_lag, d.k, _ptr(d.omega_inv), _ptr(d.ifft_divisor), lo))
        check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.coef.ptr), n, C.c_void_p(self.ext.ptr), N, self.n_lag,
                                               d.k, d.extended_k, _ptr(d.extended_omega), _ptr(d.g_coset), lo))
        check(lib.b200zk_coeff_to_extended_dev(C.c_void_p(self.instance_coeff.ptr), n,
                                               C.c_void_p(self.ext.ptr + self.n_lag * N * 32), N, s.instance, d.k,
                                               d.extended_k, _ptr(d.extended_omega), _ptr(d.g_coset), lo))
        col = 0
        for count in (s.advice, 2 * s.lookups, s.permutation_sets + s.lookups):
            out = self.points.ptr + col * 96
            check(lib.b200zk_msm_g1_registered_dev(self.h_bases, C.c_void_p(self.lag.ptr + col * n * 32), n, count, n,
                                                   C.c_void_p(out), hi))
            col += count
        torch.cuda.synchronize()
        # ---- evaluate_h and the quotient's way back, as in run()
        env = self._env()
        g = self.gates.as_c()
        check(lib.b200zk_quotient_graph(C.byref(g), C.byref(env), 0, self.values.handle))
        lk0 = s.advice
        pp0 = s.advice + 2 * s.lookups
        lp0 = pp0 + s.permutation_sets
        products = self.ext_views[pp0: pp0 + s.permutation_sets]
        sig, ph = _handles(self.sigma), _handles(products)
        zeta, delta = fr_limbs(FR_ZETA), fr_limbs(FR_DELTA)
        check(lib.b200zk_quotient_permutation(
            C.byref(env), self.values.handle, _ptr32(self.perm_kind), _ptr32(self.perm_index), _ptr(sig),
            len(self.perm_kind), _ptr(ph), len(products), s.degree - 2, s.blinding_factors, self.l0.handle,
            self.l_last.handle, self.l_active.handle, _ptr(d.extended_omega), _ptr(zeta), _ptr(delta)))
        for j in range(s.lookups):
            lg = self.lookup_graphs[j].as_c()
            check(lib.b200zk_quotient_graph(C.byref(lg), C.byref(env), 0, self.table.handle))
            check(lib.b200zk_quotient_lookup(C.byref(env), self.values.handle, self.table.handle,
                                             self.ext_views[lp0 + j].handle, self.ext_views[lk0 + 2 * j].handle,
                                             self.ext_views[lk0 + 2 * j + 1].handle, self.l0.handle,
                                             self.l_last.handle, self.l_active.handle))
        t_ev = DeviceColumn.from_host(d.t_evaluations)
        keep = n * (s.degree - 1)
        check(lib.b200zk_extended_to_coeff_dev(C.c_void_p(self.values.ptr), d.extended_k, _ptr(d.extended_omega_inv),
                                               _ptr(d.extended_ifft_divisor), _ptr(d.g_coset), C.c_void_p(t_ev.ptr),
                                               d.t_evaluations.shape[0], C.c_void_p(self.h_coeff.ptr), keep, None))
        self._commit(self.h_coeff.ptr, s.degree - 1, self.n_lag + 1)
        self.commitments = self.points.to_host()
        torch.cuda.synchronize()
        total = 1e3 * (time.perf_counter() - t0)
        t_ev.free()
        return {"total": total}

    # ------------------------------------------------------------ one proof over several GPUs
    # SURVEY.md section 8 e: commitments and iNTTs are dealt by column, the coefficient columns are
    # all-gathered (the one bulk exchange), every rank extends and evaluates only "its" cosets of the
    # extended domain (b200zk_coeff_to_coset_dev: 1/P of a coeff_to_extended per coset), the evaluated
    # cosets are gathered on rank 0, which interleaves them and finishes h(X).
    class _CudaView:
        """Zero-copy torch view of library-owned device memory (for the NCCL collectives)."""

        def __init__(self, ptr: int, n_int64: int):
            self.__cuda_array_interface__ = {"shape": (n_int64,), "typestr": "<i8", "data": (ptr, False), "version": 2}

    def _torch_view(self, torch, col: DeviceColumn, offset_elems: int, n_elems: int):
        return torch.as_tensor(self._CudaView(col.ptr + offset_elems * 32, n_elems * 4), device="cuda")

    def prepare_sharded(self, world: int, rank: int) -> None:
        from .sharding import shard_range
        lib, s, d, n, N = self.lib, self.shape, self.domain, self.n, self.N
        self.world, self.rank = world, rank
        self.parts = 1 << (d.extended_k - d.k)
        self.my_cosets = list(range(rank, self.parts, world))
        self.cosets_per_rank = -(-self.parts // world)
        self.col_b, self.col_e = shard_range(self.n_lag, rank, world)
        self.maxc = -(-self.n_lag // world)
        # every rank holds the Lagrange columns (the witness); `coef_mine` is its share of them in coefficient form
        self.lag_pad = DeviceColumn((self.n_lag + self.maxc) * n)
        self.coef_mine = DeviceColumn(self.maxc * n)
        self.coef_all = DeviceColumn(world * self.maxc * n)
        self.cos = DeviceColumn((world * self.maxc + s.instance) * n)
        self.cos_views = [DeviceColumn.view(self.cos, i * n, n) for i in range(world * self.maxc + s.instance)]
        self.h_mine = DeviceColumn(self.cosets_per_rank * n)
        self.h_all = DeviceColumn(world * self.cosets_per_rank * n)
        n_pk = s.fixed + s.permutation_columns + 3
        self.pk_coset = {}
        for q in self.my_cosets:
            c = DeviceColumn(n_pk * n)
            check(lib.b200zk_extended_coset_slice_dev(C.c_void_p(self.pk_cols.ptr), N, C.c_void_p(c.ptr), n, n_pk, d.k,
                                                      d.extended_k, q, None))
            self.pk_coset[q] = (c, [DeviceColumn.view(c, i * n, n) for i in range(n_pk)])
        # global column index -> position in the gathered buffer
        self.col_pos = []
        for r in range(world):
            b, e = shard_range(self.n_lag, r, world)
            self.col_pos += [r * self.maxc + (cidx - b) for cidx in range(b, e)]
        from .api import FR_MODULUS
        self.coset_gen = [fr_limbs(FR_ZETA * pow(d.extended_omega_int, q, FR_MODULUS) % FR_MODULUS) for q in range(self.parts)]
        # ---- the commitments are dealt by *load*: a rank that evaluates cosets of h(X), and rank 0 which also finishes
        # h(X), gets fewer columns to commit, so that all ranks finish together.  Cost model in ms (per column, per coset,
        # the finish on rank 0); rebalance() replaces it with what the last run measured.
        self.cost = {"column": [0.125] * world, "fixed": [3.3 * len(range(r, self.parts, world)) + (1.3 if r == 0 else 0.0)
                                                           for r in range(world)]}
        self._deal_commits()

    def _deal_commits(self) -> None:
        """Column counts x_r with fixed_r + column_r * x_r equal for all ranks (sharding.deal_by_load)."""
        from .sharding import deal_by_load
        self.commit_ranges = deal_by_load(self.n_lag, self.cost["column"], self.cost["fixed"])

    def rebalance(self, torch, dist, t: dict) -> None:
        """Replace the cost model by the last run's measurements (all ranks) and deal the commitments again."""
        world = self.world
        b, e = self.commit_ranges[self.rank]
        mine = torch.tensor([t["commit_own_columns"] / max(1, e - b) if e > b else 0.0,
                             t["all_gather_coefficient_columns"] + t["coset_ntt_and_quotient_own_cosets"] + t.get("finish_h_on_rank0", 0.0)],
                            dtype=torch.float64, device="cuda")
        allv = torch.zeros(2 * world, dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_gather_into_tensor(allv, mine)
        else:
            allv = mine
        v = allv.cpu().tolist()
        rates = [v[2 * r] for r in range(world) if v[2 * r] > 0]
        default = sum(rates) / len(rates) if rates else 0.125
        self.cost = {"column": [v[2 * r] if v[2 * r] > 0 else default for r in range(world)],
                     "fixed": [v[2 * r + 1] for r in range(world)]}
        self._deal_commits()

    def run_sharded(self, torch, dist=None) -> dict:
        """The same proof as run(), spread over `world` ranks (prepare_sharded first).  world = 1 exercises the whole
        coset path on one GPU.  `torch` supplies the NCCL collectives and the device copies around library-owned memory.

        Schedule: (1) every rank takes its even share of the Lagrange columns to coefficient form (a copy: the Lagrange
        columns are still to be committed) and the coefficient columns are all-gathered — only the coset owners wait for
        that; (2) the ranks that own cosets
        of the extended domain extend and evaluate them; the all-gather of the evaluated cosets is issued
        asynchronously, so ranks without cosets do not wait for it; (3) every rank commits its load-weighted share of
        the Lagrange columns; (4) rank 0, which has the fewest columns, finishes h(X) (interleave, divide,
        extended_to_coeff, commit the pieces); (5) the commitments are gathered."""
        lib, s, d, n, N = self.lib, self.shape, self.domain, self.n, self.N
        world, rank = self.world, self.rank
        t, marks = {}, [time.perf_counter()]

        def mark(name, sync=True):
            if sync:
                self.sync()
            marks.append(time.perf_counter())
            t[name] = 1e3 * (marks[-1] - marks[-2])

        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.lag_pad.ptr), self.n_lag * n, self.seed + 2, 0))
        check(lib.b200zk_gen_scalars_dev(C.c_void_p(self.instance_coeff.ptr), self.instance_coeff.n, self.seed + 3, 0))
        self.sync()
        if world > 1:
            dist.barrier()
        marks[0] = time.perf_counter()
        # ---- (1) own share -> coefficient form, all-gather
        b, e = self.col_b, self.col_e
        if e > b:
            self._torch_view(torch, self.coef_mine, 0, (e - b) * n).copy_(self._torch_view(torch, self.lag_pad, b * n, (e - b) * n))
            self.sync()
            check(lib.b200zk_ntt_dev(C.c_void_p(self.coef_mine.ptr), n, e - b, d.k, _ptr(d.omega_inv), _ptr(d.ifft_divisor), None))
        mark("lagrange_to_coeff_own_columns")
        coef_work = None
        if world > 1:
            # only the ranks that own a coset read the gathered coefficient columns: the others take part in the
            # collective but go straight on to their commitments (the copy engines and a few SMs do the gathering)
            coef_work = dist.all_gather_into_tensor(self._torch_view(torch, self.coef_all, 0, world * self.maxc * n),
                                                    self._torch_view(torch, self.coef_mine, 0, self.maxc * n), async_op=True)
            if self.my_cosets:
                coef_work.wait()
                coef_work = None
        else:
            self._torch_view(torch, self.coef_all, 0, self.maxc * n).copy_(self._torch_view(torch, self.coef_mine, 0, self.maxc * n))
        # (a device-wide synchronisation would wait for the gathering: a rank without cosets has nothing to wait for here)
        mark("all_gather_coefficient_columns", sync=bool(self.my_cosets) or world == 1)
        # ---- (2) own cosets of the extended domain
        env_scal = self.scalars
        zeta_delta = fr_limbs(FR_DELTA)
        for slot, q in enumerate(self.my_cosets):
            g = self.coset_gen[q]
            ncols = world * self.maxc
            check(lib.b200zk_coeff_to_coset_dev(C.c_void_p(self.coef_all.ptr), n, C.c_void_p(self.cos.ptr), n, ncols, d.k,
                                                _ptr(d.omega), _ptr(g), None))
            check(lib.b200zk_coeff_to_coset_dev(C.c_void_p(self.instance_coeff.ptr), n,
                                                C.c_void_p(self.cos.ptr + ncols * n * 32), n, s.instance, d.k,
                                                _ptr(d.omega), _ptr(g), None))
            col = lambda cidx: self.cos_views[self.col_pos[cidx]]
            advice = [col(i) for i in range(s.advice)]
            instance = self.cos_views[ncols: ncols + s.instance]
            pkc, pkv = self.pk_coset[q]
            fixed, sigma = pkv[: s.fixed], pkv[s.fixed: s.fixed + s.permutation_columns]
            l0, l_last, l_active = pkv[-3:]
            keep = [_handles(fixed), _handles(advice), _handles(instance), np.zeros((1, 4), np.uint64)]
            env = EnvC()
            env.fixed, env.n_fixed = keep[0].ctypes.data, len(fixed)
            env.advice, env.n_advice = keep[1].ctypes.data, len(advice)
            env.instance, env.n_instance = keep[2].ctypes.data, len(instance)
            env.challenges, env.n_challenges = keep[3].ctypes.data, 0
            for name in ("beta", "gamma", "theta", "y"):
                getattr(env, name)[:] = [int(x) for x in env_scal[name]]
            env.k, env.ext_k = d.k, d.k
            env.range_begin, env.range_len = 0, 0
            values = DeviceColumn.view(self.h_mine, slot * n, n)
            gg = self.gates.as_c()
            check(lib.b200zk_quotient_graph(C.byref(gg), C.byref(env), 0, values.handle))
            lk0, pp0 = s.advice, s.advice + 2 * s.lookups
            lp0 = pp0 + s.permutation_sets
            products = [col(pp0 + i) for i in range(s.permutation_sets)]
            sig, ph = _handles(sigma), _handles(products)
            check(lib.b200zk_quotient_permutation(
                C.byref(env), values.handle, _ptr32(self.perm_kind), _ptr32(self.perm_index), _ptr(sig),
                len(self.perm_kind), _ptr(ph), len(products), s.degree - 2, s.blinding_factors, l0.handle,
                l_last.handle, l_active.handle, _ptr(d.omega), _ptr(g), _ptr(zeta_delta)))
            for j in range(s.lookups):
                lg = self.lookup_graphs[j].as_c()
                check(lib.b200zk_quotient_graph(C.byref(lg), C.byref(env), 0, self.table.handle))
                check(lib.b200zk_quotient_lookup(C.byref(env), values.handle, self.table.handle, col(lp0 + j).handle,
                                                 col(lk0 + 2 * j).handle, col(lk0 + 2 * j + 1).handle, l0.handle,
                                                 l_last.handle, l_active.handle))
            values.free()
        mark("coset_ntt_and_quotient_own_cosets", sync=bool(self.my_cosets) or world == 1)
        # the evaluated cosets travel to rank 0 while everybody commits: asynchronous, so a rank without cosets (it
        # arrives here at once) does not wait for the ranks that have some
        h_work = None
        if world > 1:
            h_work = dist.all_gather_into_tensor(self._torch_view(torch, self.h_all, 0, world * self.cosets_per_rank * n),
                                                 self._torch_view(torch, self.h_mine, 0, self.cosets_per_rank * n), async_op=True)
        # ---- (3) commitments of the Lagrange columns, dealt by load
        cb, ce = self.commit_ranges[rank]
        if ce > cb:
            self._commit(self.lag_pad.ptr + cb * n * 32, ce - cb, cb)
        mark("commit_own_columns")
        # ---- (4) h(X) on rank 0
        if coef_work is not None:
            coef_work.wait()
        if h_work is not None:
            h_work.wait()
        mark("gather_h_cosets")
        if rank == 0:
            for q in range(self.parts):
                owner, slot = q % world, q // world
                srcbuf = self.h_all if world > 1 else self.h_mine
                off = (owner * self.cosets_per_rank + slot) * n if world > 1 else slot * n
                check(lib.b200zk_extended_coset_interleave_dev(C.c_void_p(srcbuf.ptr + off * 32), C.c_void_p(self.values.ptr),
                                                               d.k, d.extended_k, q, None))
            t_ev = DeviceColumn.from_host(d.t_evaluations)
            keep_n = n * (s.degree - 1)
            check(lib.b200zk_extended_to_coeff_dev(C.c_void_p(self.values.ptr), d.extended_k, _ptr(d.extended_omega_inv),
                                                   _ptr(d.extended_ifft_divisor), _ptr(d.g_coset), C.c_void_p(t_ev.ptr),
                                                   d.t_evaluations.shape[0], C.c_void_p(self.h_coeff.ptr), keep_n, None))
            self._commit(self.h_coeff.ptr, s.degree - 1, self.n_lag + 1)
            self.sync()
            t_ev.free()
        mark("finish_h_on_rank0")
        # ---- (5) every commitment on every rank (the transcript needs them on rank 0)
        if world > 1:
            width = max(ce_ - cb_ for cb_, ce_ in self.commit_ranges)
            mine = torch.zeros(width * 12, dtype=torch.int64, device="cuda")
            if ce > cb:
                mine[: (ce - cb) * 12].copy_(self._torch_view(torch, self.points, cb * 3, (ce - cb) * 3))
            allp = torch.zeros(world * width * 12, dtype=torch.int64, device="cuda")
            dist.all_gather_into_tensor(allp, mine)
            for r, (rb, re_) in enumerate(self.commit_ranges):
                if re_ > rb and r != rank:
                    self._torch_view(torch, self.points, rb * 3, (re_ - rb) * 3).copy_(allp[r * width * 12: r * width * 12 + (re_ - rb) * 12])
        mark("gather_commitments")
        if world > 1:
            dist.barrier()
        t["total"] = 1e3 * (time.perf_counter() - marks[0])
        return t

    def counts(self) -> dict:
        s = self.shape
        return {"commit_lagrange": s.lagrange_columns, "commit": s.degree - 1, "lagrange_to_coeff": s.lagrange_columns,
                "coeff_to_extended": s.lagrange_columns + s.instance, "extended_to_coeff": 1,
                "quotient_columns_read": s.fixed + s.advice + s.instance + 2 * s.permutation_columns +
                                         s.permutation_sets + 3 * s.lookups + 3,
                "k": s.k, "extended_k": self.domain.extended_k}

    def close(self) -> None:
        self.lib.b200zk_bases_evict(self.h_bases)
        if getattr(self, "host_advice", None) is not None:
            host_free(self.host_advice)
            self.host_advice = None
        self.lib.b200zk_mirror_enable(0)
        if getattr(self, "_percall_pinned", False):
            for name in ("h_lag", "h_instance", "h_prod_ext", "h_values", "h_hcoeff"):
                host_free(getattr(self, name))
            self._percall_pinned = False
        for v in self.ext_views + self.fixed + self.sigma + [self.l0, self.l_last, self.l_active]:
            v.free()
        for q, (col, views) in getattr(self, "pk_coset", {}).items():
            for v in views:
                v.free()
            col.free()
        for v in getattr(self, "cos_views", []):
            v.free()
        for name in ("lag_pad", "coef_mine", "coef_all", "cos", "h_mine", "h_all", "coef"):
            if hasattr(self, name):
                getattr(self, name).free()
        for c in (self.pk_cols, self.lag, self.instance_coeff, self.ext, self.values, self.table, self.h_coeff,
                  self.points):
            c.free()

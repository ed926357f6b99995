"""End-to-end harness: a complete KZG/SHPLONK proof of the reference's SquareCircuit
(reference src/signal.rs:19-76, the circuit its Solidity verifier was generated for), assembled by
the *callers* of the hot path — keygen, `create_proof`, the EVM transcript and the SHPLONK
opening, all of which SURVEY.md section 8 leaves outside the product — from the outputs of a
backend that supplies the hot-path functions:

    OracleBackend  : oracle/ (CPU restatement), used to pin the harness itself and as the
                     byte-for-byte comparison target;
    DeviceBackend  : the product, through the same ctypes mirror every GPU parity test uses
                     (commit / commit_lagrange, lagrange_to_coeff, coeff_to_extended, evaluate_h,
                     divide_by_vanishing_poly, extended_to_coeff, permutation products,
                     eval_polynomial, kate_division, linear combinations).

The proof bytes go to oracle/sol_verifier.py, the transliteration of the reference's
solidity_verifier_contract/contract.sol.  The call order, the blinding draws and the query order
follow upstream `plonk::create_proof` / `multiopen::shplonk::ProverSHPLONK::create_proof`
([DEP] halo2_proofs @ v2023_01_20) as far as the contract pins them
(SURVEY.md appendix B); the blinding RNG is this file's own seeded stream (the reference's Rust RNG
cannot be run here), the same stream for every backend, so proofs are comparable byte for byte.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from oracle import bn254 as bn
from oracle import bn254_pairing as pairing
from oracle import c_oracle as co
from oracle import halo2_cpu as h2
from oracle import prover_steps_cpu as steps_cpu
from oracle import quotient_cpu as qc
from oracle import sol_verifier
from oracle.keccak import keccak256

R = bn.R
F = bn.fr_array_from_canonical
BLINDING_FACTORS = 5                 # cs.blinding_factors() for this circuit; contract.sol:400-408
DEGREE = 3                           # cs.degree(): s * a0 * a0 and the permutation argument
CHUNK_LEN = DEGREE - 2
PERM_COLUMNS = [("advice", 0), ("advice", 1), ("instance", 0)]     # src/signal.rs:32-34 enable_equality order


class Rng:
    """The fixed-seed blinding stream (splitmix64 words reduced mod r, oracle/bn254.py)."""

    def __init__(self, seed: int):
        self.seed, self.i = seed, 0

    def fr(self) -> int:
        v = bn.fr_array_to_canonical(bn.seeded_fr_mont_limbs(self.seed, 1, self.i))[0]
        self.i += 1
        return v


# ------------------------------------------------------------------------------ EVM transcript
class EvmTranscript:
    """The prover side of contract.sol:89-112 / :223-304: the buffer starts with vk_digest and the
    instances, points are absorbed as 64 big-endian bytes, scalars as 32; a challenge is
    keccak256(buffer) mod r and the buffer restarts from the 32-byte hash; a squeeze with nothing
    absorbed since the last one hashes hash || 0x01."""

    def __init__(self):
        self.buf = bytearray()
        self.proof = bytearray()
        self.fresh = False               # True right after a squeeze

    def common_scalar(self, s: int):
        self.buf += int(s).to_bytes(32, "big")
        self.fresh = False

    def write_scalar(self, s: int):
        self.common_scalar(s)
        self.proof += int(s).to_bytes(32, "big")

    def write_point(self, p):
        assert p is not None, "the EVM encoding has no identity point (contract.sol:80-82)"
        b = bn.g1_to_evm_bytes(p)
        self.buf += b
        self.proof += b
        self.fresh = False

    def squeeze(self) -> int:
        data = bytes(self.buf) + (b"\x01" if self.fresh else b"")
        hsh = keccak256(data)
        self.buf = bytearray(hsh)
        self.fresh = True
        return int.from_bytes(hsh, "big") % R


# ------------------------------------------------------------------------------ local KZG setup
@dataclass
class Srs:
    k: int
    g: np.ndarray               # (n, 8) affine Montgomery: [s^i] G
    g_lagrange: np.ndarray      # (n, 8): [L_i(s)] G
    g2: tuple
    s_g2: tuple


def setup(k: int, seed: int = 0x5EED) -> Srs:
    """`ParamsKZG::setup(k, rng)` (poly/kzg/commitment.rs): the group elements are determined by
    the secret s; L_i(s) = w^i (s^n - 1) / (n (s - w^i))."""
    n = 1 << k
    s = Rng(seed).fr()
    powers = [1] * n
    for i in range(1, n):
        powers[i] = powers[i - 1] * s % R
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (bn.FR_S - k), R)
    roots = [1] * n
    for i in range(1, n):
        roots[i] = roots[i - 1] * omega % R
    inv = steps_cpu.batch_invert([(s - w) % R for w in roots])
    common = (pow(s, n, R) - 1) * pow(n, -1, R) % R
    lag = [common * w % R * d % R for w, d in zip(roots, inv)]
    return Srs(k, co.g1_generator_mul(F(powers)), co.g1_generator_mul(F(lag)), pairing.G2_GEN,
               pairing.g2_mul(pairing.G2_GEN, s))


# ------------------------------------------------------------------------------ backends
class OracleBackend:
    """The hot-path functions from oracle/ (values are Python ints, points affine tuples)."""
    name = "oracle"

    def __init__(self, srs: Srs):
        self.srs = srs
        self.domain = h2.EvaluationDomain(DEGREE, srs.k)

    def commit(self, coeffs):
        return bn.g1_jacobian_limbs_to_affine(co.best_multiexp(F(coeffs), self.srs.g[:len(coeffs)]))

    def commit_lagrange(self, values):
        return bn.g1_jacobian_limbs_to_affine(co.best_multiexp(F(values), self.srs.g_lagrange[:len(values)]))

    def lagrange_to_coeff(self, a): return self.domain.lagrange_to_coeff(list(a))
    def coeff_to_extended(self, a): return self.domain.coeff_to_extended(list(a))
    def extended_to_coeff(self, a): return self.domain.extended_to_coeff(list(a))
    def divide_by_vanishing_poly(self, a): return self.domain.divide_by_vanishing_poly(list(a))

    def permutation_products(self, values, sigma, beta, gamma, blinds):
        return steps_cpu.permutation_products(values, sigma, CHUNK_LEN, self.domain.omega, beta, gamma,
                                              BLINDING_FACTORS, blinds)

    def load_pk(self, pk):
        return pk

    def evaluate_h(self, pk, advice_coeff, instance_coeff, product_coeff, y, beta, gamma, theta):
        d = self.domain
        env = {"fixed": pk.fixed_cosets, "advice": [d.coeff_to_extended(c) for c in advice_coeff],
               "instance": [d.coeff_to_extended(c) for c in instance_coeff], "challenges": [],
               "beta": beta, "gamma": gamma, "theta": theta, "y": y}
        perm = qc.PermutationData(PERM_COLUMNS, pk.permutation_cosets, [d.coeff_to_extended(c) for c in product_coeff],
                                  CHUNK_LEN, BLINDING_FACTORS)
        return qc.evaluate_h(d, qc.custom_gates_graph(square_gates()), env, pk.l0, pk.l_last, pk.l_active_row, perm, [])

    def eval_polynomial(self, poly, x): return steps_cpu.eval_polynomial(poly, x)
    def kate_division(self, poly, b): return steps_cpu.kate_division(poly, b)

    def linear_combination(self, polys, coeffs):
        out = [0] * len(polys[0])
        for p, c in zip(polys, coeffs):
            out = [(o + c * v) % R for o, v in zip(out, p)]
        return out


class DeviceBackend:
    """The same functions from the product (b200zk over the C ABI); conversions only."""
    name = "device"

    def __init__(self, srs: Srs, zk):
        self.zk, self.srs = zk, srs
        self.domain = zk.EvaluationDomain(DEGREE, srs.k)
        self.params = zk.ParamsKZG(srs.g, srs.g_lagrange)
        self.omega = h2.EvaluationDomain(DEGREE, srs.k).omega
        self._pk = None

    @staticmethod
    def _ints(a): return bn.fr_array_to_canonical(a)

    def commit(self, coeffs):
        return bn.g1_jacobian_limbs_to_affine(self.params.commit(F(coeffs)))

    def commit_lagrange(self, values):
        return bn.g1_jacobian_limbs_to_affine(self.params.commit_lagrange(F(values)))

    def lagrange_to_coeff(self, a): return self._ints(self.domain.lagrange_to_coeff(F(a)))
    def coeff_to_extended(self, a): return self._ints(self.domain.coeff_to_extended(F(a)))
    def extended_to_coeff(self, a): return self._ints(self.domain.extended_to_coeff(F(a)))
    def divide_by_vanishing_poly(self, a): return self._ints(self.domain.divide_by_vanishing_poly(F(a)))

    def permutation_products(self, values, sigma, beta, gamma, blinds):
        z = self.zk.permutation_products([F(v) for v in values], [F(s) for s in sigma], CHUNK_LEN, self.srs.k,
                                         F([beta])[0], F([gamma])[0], BLINDING_FACTORS,
                                         np.stack([F(b) for b in blinds]))
        return [self._ints(zz) for zz in z]

    def load_pk(self, pk):
        zk = self.zk
        col = lambda ints: zk.DeviceColumn.from_host(F(ints))
        self._pk = zk.ProvingKeyCosets(
            fixed_cosets=[col(c) for c in pk.fixed_cosets], l0=col(pk.l0), l_last=col(pk.l_last),
            l_active_row=col(pk.l_active_row), permutation_cosets=[col(c) for c in pk.permutation_cosets],
            permutation_columns=PERM_COLUMNS, degree=DEGREE, blinding_factors=BLINDING_FACTORS)
        self._evaluator = zk.Evaluator(zk.FlatGraph(**qc.custom_gates_graph(square_gates()).to_flat()), [])
        return pk

    def evaluate_h(self, pk, advice_coeff, instance_coeff, product_coeff, y, beta, gamma, theta):
        one = lambda v: F([v])[0]
        out = self._evaluator.evaluate_h(self.domain, self._pk, [F(c) for c in advice_coeff],
                                         [F(c) for c in instance_coeff], np.zeros((0, 4), dtype=np.uint64),
                                         one(y), one(beta), one(gamma), one(theta), (), [F(c) for c in product_coeff])
        return self._ints(out)

    def eval_polynomial(self, poly, x):
        return self._ints(self.zk.eval_polynomial(F(poly), F([x])[0])[None, :])[0]

    def kate_division(self, poly, b):
        return self._ints(self.zk.kate_division(F(poly), F([b])[0]))

    def linear_combination(self, polys, coeffs):
        return self._ints(self.zk.linear_combination([F(p) for p in polys], [F([c])[0] for c in coeffs]))

    def close(self):
        self.params.close()


# ------------------------------------------------------------------------------ circuit + keygen
def square_gates():
    """src/signal.rs:36-42: s * (signal_hash_square - signal_hash * signal_hash); the simple
    selector becomes fixed column 0 (contract.sol:443-452 reads it as f_0)."""
    a0, a1, f0 = qc.Advice(0), qc.Advice(1), qc.Fixed(0)
    return [f0 * (a1 - a0 * a0)]


@dataclass
class Assignment:
    """What `synthesize` (src/signal.rs:50-76) leaves in the columns.  The reference assigns one
    row; `signal_hashes` may hold more (one square per row) to give larger domains real content,
    and `copies` may tie cells together (the reference's commented-out constrain_instance,
    src/signal.rs:72-73, is the case [(("advice", 1, 0), ("instance", 0, 0))])."""
    k: int
    signal_hashes: list
    instances: list
    copies: list = field(default_factory=list)

    @property
    def n(self): return 1 << self.k
    @property
    def usable_rows(self): return self.n - (BLINDING_FACTORS + 1)


@dataclass
class ProvingKey:
    k: int
    omega: int
    fixed_values: list
    fixed_polys: list
    fixed_cosets: list
    permutations: list          # sigma columns, Lagrange basis
    permutation_polys: list
    permutation_cosets: list
    l0: list
    l_last: list
    l_active_row: list
    fixed_commitments: list
    permutation_commitments: list
    vk_digest: int


def _permutation_mapping(asg: Assignment):
    """permutation/keygen.rs Assembly: every cell starts as its own cycle; a copy constraint joins
    two cycles.  mapping[col][row] = the next cell of the cycle."""
    n = asg.n
    index = {c: i for i, c in enumerate(PERM_COLUMNS)}
    mapping = [[(c, r) for r in range(n)] for c in range(len(PERM_COLUMNS))]
    aux = [[(c, r) for r in range(n)] for c in range(len(PERM_COLUMNS))]
    sizes = [[1] * n for _ in PERM_COLUMNS]
    for (lk, li, lr), (rk, ri, rr) in asg.copies:
        lc, rc = index[(lk, li)], index[(rk, ri)]
        assert lr < asg.usable_rows and rr < asg.usable_rows
        lcyc, rcyc = aux[lc][lr], aux[rc][rr]
        if lcyc == rcyc:
            continue
        if sizes[lcyc[0]][lcyc[1]] < sizes[rcyc[0]][rcyc[1]]:
            lcyc, rcyc = rcyc, lcyc
            (lc, lr), (rc, rr) = (rc, rr), (lc, lr)
        sizes[lcyc[0]][lcyc[1]] += sizes[rcyc[0]][rcyc[1]]
        i, j = rc, rr
        while True:
            aux[i][j] = lcyc
            i, j = mapping[i][j]
            if (i, j) == (rc, rr):
                break
        mapping[lc][lr], mapping[rc][rr] = mapping[rc][rr], mapping[lc][lr]
    return mapping


def keygen(be, asg: Assignment) -> ProvingKey:
    """keygen_vk + keygen_pk for this circuit: the selector column, the permutation polynomials
    sigma_c(w^r) = delta^c' w^r' and l_0 / l_last / l_active_row, each through the backend's
    lagrange_to_coeff / coeff_to_extended / commit_lagrange."""
    n, k = asg.n, asg.k
    omega = pow(bn.FR_ROOT_OF_UNITY, 1 << (bn.FR_S - k), R)
    assert len(asg.signal_hashes) <= asg.usable_rows and len(asg.instances) <= asg.usable_rows
    selector = [1 if i < len(asg.signal_hashes) else 0 for i in range(n)]
    omega_pows = [1] * n
    for i in range(1, n):
        omega_pows[i] = omega_pows[i - 1] * omega % R
    delta_pows = [pow(bn.FR_DELTA, c, R) for c in range(len(PERM_COLUMNS))]
    mapping = _permutation_mapping(asg)
    permutations = [[delta_pows[mapping[c][r][0]] * omega_pows[mapping[c][r][1]] % R for r in range(n)]
                    for c in range(len(PERM_COLUMNS))]
    to_ext = lambda values: be.coeff_to_extended(be.lagrange_to_coeff(values))
    rows = lambda rs: [1 if i in rs else 0 for i in range(n)]
    l0 = to_ext(rows({0}))
    l_last = to_ext(rows({n - BLINDING_FACTORS - 1}))
    l_blind = to_ext(rows(set(range(n - BLINDING_FACTORS, n))))
    l_active = [(1 - a - b) % R for a, b in zip(l_last, l_blind)]
    fixed_polys = [be.lagrange_to_coeff(selector)]
    perm_polys = [be.lagrange_to_coeff(p) for p in permutations]
    fixed_comm = [be.commit_lagrange(selector)]
    perm_comm = [be.commit_lagrange(p) for p in permutations]
    digest_src = k.to_bytes(4, "big") + b"".join(bn.g1_to_evm_bytes(p) for p in fixed_comm + perm_comm)
    pk = ProvingKey(k, omega, [selector], fixed_polys, [be.coeff_to_extended(p) for p in fixed_polys], permutations,
                    perm_polys, [be.coeff_to_extended(p) for p in perm_polys], l0, l_last, l_active, fixed_comm,
                    perm_comm, int.from_bytes(keccak256(digest_src), "big") % R)
    return be.load_pk(pk)


def vk_code(pk: ProvingKey, srs: Srs, num_instances: int) -> bytes:
    return sol_verifier.encode_vk(pk.vk_digest, num_instances, pk.k, pk.omega, bn.G1_GEN, srs.g2, srs.s_g2,
                                  pk.fixed_commitments, pk.permutation_commitments, BLINDING_FACTORS)


# ------------------------------------------------------------------------------ create_proof
def _interpolate(points, evals):
    """lagrange_interpolate: coefficients (low to high) of the polynomial through the pairs."""
    m = len(points)
    out = [0] * m
    for i in range(m):
        num = [1]
        den = 1
        for j in range(m):
            if j == i:
                continue
            num = [(a - points[j] * b) % R for a, b in zip([0] + num, num + [0])]
            den = den * (points[i] - points[j]) % R
        scale = evals[i] * pow(den, -1, R) % R
        for t in range(m):
            out[t] = (out[t] + num[t] * scale) % R
    return out


def _poly_eval(coeffs, x):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def create_proof(be, pk: ProvingKey, asg: Assignment, rng: Rng, trace: dict | None = None,
                 witness_error: int = 0) -> bytes:
    """`plonk::create_proof` for one SquareCircuit instance with the SHPLONK prover and the EVM
    transcript; returns the 0x460 proof bytes of contract.sol:221."""
    n, k = asg.n, asg.k
    omega = pk.omega
    omega_inv = pow(omega, -1, R)
    tr = EvmTranscript()
    tr.common_scalar(pk.vk_digest)                                    # vk.hash_into
    for v in asg.instances:                                           # QUERY_INSTANCE = false: scalars only
        tr.common_scalar(v % R)
    instance_values = [[v % R for v in asg.instances] + [0] * (n - len(asg.instances))]
    instance_polys = [be.lagrange_to_coeff(instance_values[0])]

    # ---- advice columns: witness, then random blinding rows, commit_lagrange each
    advice_values = [[0] * n, [0] * n]
    for i, v in enumerate(asg.signal_hashes):
        advice_values[0][i] = v % R
        advice_values[1][i] = (v * v + (witness_error if i == 0 else 0)) % R
    for col in advice_values:
        for row in range(asg.usable_rows, n):
            col[row] = rng.fr()
    _advice_blinds = [rng.fr() for _ in advice_values]               # Blind::new(rng): drawn, unused by KZG
    for col in advice_values:
        tr.write_point(be.commit_lagrange(col))
    advice_polys = [be.lagrange_to_coeff(col) for col in advice_values]

    theta = tr.squeeze()
    beta = tr.squeeze()
    gamma = tr.squeeze()

    # ---- permutation argument: grand products, blinded, committed
    n_sets = -(-len(PERM_COLUMNS) // CHUNK_LEN)
    blinds = [[rng.fr() for _ in range(BLINDING_FACTORS)] for _ in range(n_sets)]
    perm_values = advice_values + instance_values
    z = be.permutation_products(perm_values, pk.permutations, beta, gamma, blinds)
    _z_blinds = [rng.fr() for _ in z]
    for zz in z:
        tr.write_point(be.commit_lagrange(zz))
    z_polys = [be.lagrange_to_coeff(zz) for zz in z]

    # ---- vanishing argument: random polynomial
    random_poly = [rng.fr() for _ in range(n)]
    _random_blind = rng.fr()
    tr.write_point(be.commit(random_poly))

    y = tr.squeeze()

    # ---- quotient: evaluate_h on the extended coset, divide by X^n - 1, back to coefficients
    h_ext = be.evaluate_h(pk, advice_polys, instance_polys, z_polys, y, beta, gamma, theta)
    h_coeff = be.extended_to_coeff(be.divide_by_vanishing_poly(h_ext))
    assert len(h_coeff) == n * (DEGREE - 1)
    h_pieces = [h_coeff[i:i + n] for i in range(0, len(h_coeff), n)]
    _h_blinds = [rng.fr() for _ in h_pieces]
    for piece in h_pieces:
        tr.write_point(be.commit(piece))

    x = tr.squeeze()
    x_next = x * omega % R
    x_last = x * pow(omega_inv, BLINDING_FACTORS + 1, R) % R
    xn = pow(x, n, R)

    # ---- evaluations, in upstream's order (contract.sol:284-294 reads them back in the same one)
    ev = be.eval_polynomial
    advice_evals = [ev(p, x) for p in advice_polys]
    fixed_evals = [ev(p, x) for p in pk.fixed_polys]
    random_eval = ev(random_poly, x)
    sigma_evals = [ev(p, x) for p in pk.permutation_polys]
    z_evals = []
    for s, p in enumerate(z_polys):
        z_evals.append([ev(p, x), ev(p, x_next)] + ([ev(p, x_last)] if s + 1 < len(z_polys) else []))
    for v in advice_evals + fixed_evals + [random_eval] + sigma_evals + [e for zs in z_evals for e in zs]:
        tr.write_scalar(v)

    # ---- SHPLONK (multiopen/shplonk/prover.rs): rotation sets in order of first appearance
    h_poly = be.linear_combination(h_pieces, [pow(xn, i, R) for i in range(len(h_pieces))])
    h_eval = ev(h_poly, x)
    set0 = [(p, [e]) for p, e in zip(advice_polys, advice_evals)]
    set0 += [(p, [e]) for p, e in zip(pk.fixed_polys, fixed_evals)]
    set0 += [(p, [e]) for p, e in zip(pk.permutation_polys, sigma_evals)]
    set0 += [(h_poly, [h_eval]), (random_poly, [random_eval])]
    rotation_sets = [([x], set0),
                     ([x, x_next, x_last], [(p, e) for p, e in zip(z_polys[:-1], z_evals[:-1])]),
                     ([x, x_next], [(z_polys[-1], z_evals[-1])])]
    super_points = [x, x_next, x_last]

    zeta = tr.squeeze()          # upstream's y
    nu = tr.squeeze()            # upstream's v

    def combine(polys, low_degree):
        """sum_j zeta^j (P_j(X) - low_j(X))"""
        powers = [pow(zeta, j, R) for j in range(len(polys))]
        out = be.linear_combination(polys, powers)
        for pw, low in zip(powers, low_degree):
            for t, c in enumerate(low):
                out[t] = (out[t] - pw * c) % R
        return out

    quotients = []
    interpolated = []
    for points, members in rotation_sets:
        lows = [_interpolate(points, evals) for _, evals in members]
        interpolated.append(lows)
        n_x = combine([p for p, _ in members], lows)
        for pt in points:                                             # div_by_vanishing
            n_x = be.kate_division(n_x, pt)
        quotients.append(n_x + [0] * (n - len(n_x)))
    w_poly = be.linear_combination(quotients, [pow(nu, i, R) for i in range(len(quotients))])
    tr.write_point(be.commit(w_poly))

    mu = tr.squeeze()            # upstream's u
    z_diffs, contributions = [], []
    for (points, members), lows in zip(rotation_sets, interpolated):
        z_i = 1
        for pt in super_points:
            if pt not in points:
                z_i = z_i * (mu - pt) % R
        z_diffs.append(z_i)
        contributions.append(combine([p for p, _ in members], [[_poly_eval(low, mu)] for low in lows]))
    zt_eval = 1
    for pt in super_points:
        zt_eval = zt_eval * (mu - pt) % R
    l_x = be.linear_combination(contributions + [w_poly],
                                [pow(nu, i, R) * zd % R for i, zd in enumerate(z_diffs)] + [-zt_eval % R])
    assert ev(l_x, mu) == 0, "SHPLONK linearisation does not vanish at u"
    z0_inv = pow(z_diffs[0], -1, R)
    w2_poly = [c * z0_inv % R for c in be.kate_division(l_x, mu)]
    tr.write_point(be.commit(w2_poly + [0] * (n - len(w2_poly))))

    if trace is not None:
        trace.update(theta=theta, beta=beta, gamma=gamma, y=y, x=x, zeta=zeta, nu=nu, mu=mu, h_coeff=h_coeff,
                     advice_polys=advice_polys, z_polys=z_polys)
    assert len(tr.proof) == sol_verifier.PROOF_LEN
    return bytes(tr.proof)

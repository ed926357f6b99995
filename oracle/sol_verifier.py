"""ORACLE (test infrastructure, never on the product path): the reference's Solidity verifier,
solidity_verifier_contract/contract.sol (`Halo2Verifier.verifyProof`, SquareCircuit shape, SHPLONK,
EVM Keccak transcript), transliterated statement by statement onto a small model of the EVM
(word-addressed memory, calldata, KECCAK256 and the precompiles 0x05-0x08).

This is the one artefact in /root/reference that fixes what a *valid proof* on this path is: the
transcript order and wire layout (contract.sol:223-304), the quotient identity with its y-folding
order (:439-511), the quotient-piece recombination (:513-534), the SHPLONK opening equation with
its rotation sets and challenge powers (:536-779) and the final pairing (:813-820).  A proof
assembled from this repository's MSM / NTT / quotient outputs is accepted here only if those
outputs are the commitments, evaluations and quotient the protocol defines; that is the end-to-end
pin of tests/test_square_proof_*.py.  The verifying key the contract reads with `extcodecopy` is
not in the reference: `encode_vk` lays a locally generated one out in the same 0x3a0
bytes (:14-35 and the commitments the final MSM reads at 0x0720-0x0800, :727-733).

Every block below cites the contract lines it restates; memory pointers keep the contract's values
so the two can be read side by side.  Nothing is reshaped: the scratch-memory aliasing of the
original (0x00-0xe0 used for both EC operands and batch-inversion inputs) is reproduced as is.
"""
from __future__ import annotations

from . import bn254 as bn
from . import bn254_pairing as pairing
from .keccak import keccak256

Q = bn.Q   # contract.sol:210
R = bn.R   # contract.sol:211

# contract.sol:6-66
PROOF_LEN_CPTR = 0x64
PROOF_CPTR = 0x84
NUM_INSTANCE_CPTR = 0x04e4
INSTANCE_CPTR = 0x0504
FIRST_QUOTIENT_X_CPTR = 0x0204
LAST_QUOTIENT_X_CPTR = 0x0244
VK_MPTR = 0x0480
VK_DIGEST_MPTR = 0x0480
NUM_INSTANCES_MPTR = 0x04a0
K_MPTR = 0x04c0
N_INV_MPTR = 0x04e0
OMEGA_MPTR = 0x0500
OMEGA_INV_MPTR = 0x0520
OMEGA_INV_TO_L_MPTR = 0x0540
HAS_ACCUMULATOR_MPTR = 0x0560
ACC_OFFSET_MPTR = 0x0580
NUM_ACC_LIMBS_MPTR = 0x05a0
NUM_ACC_LIMB_BITS_MPTR = 0x05c0
G1_X_MPTR = 0x05e0
G1_Y_MPTR = 0x0600
G2_X_1_MPTR = 0x0620
G2_X_2_MPTR = 0x0640
G2_Y_1_MPTR = 0x0660
G2_Y_2_MPTR = 0x0680
NEG_S_G2_X_1_MPTR = 0x06a0
NEG_S_G2_X_2_MPTR = 0x06c0
NEG_S_G2_Y_1_MPTR = 0x06e0
NEG_S_G2_Y_2_MPTR = 0x0700
CHALLENGE_MPTR = 0x0820
THETA_MPTR = 0x0820
BETA_MPTR = 0x0840
GAMMA_MPTR = 0x0860
Y_MPTR = 0x0880
X_MPTR = 0x08a0
ZETA_MPTR = 0x08c0
NU_MPTR = 0x08e0
MU_MPTR = 0x0900
ACC_LHS_X_MPTR = 0x0920
ACC_LHS_Y_MPTR = 0x0940
ACC_RHS_X_MPTR = 0x0960
ACC_RHS_Y_MPTR = 0x0980
X_N_MPTR = 0x09a0
X_N_MINUS_1_INV_MPTR = 0x09c0
L_LAST_MPTR = 0x09e0
L_BLIND_MPTR = 0x0a00
L_0_MPTR = 0x0a20
INSTANCE_EVAL_MPTR = 0x0a40
QUOTIENT_EVAL_MPTR = 0x0a60
QUOTIENT_X_MPTR = 0x0a80
QUOTIENT_Y_MPTR = 0x0aa0
G1_SCALAR_MPTR = 0x0ac0
PAIRING_LHS_X_MPTR = 0x0ae0
PAIRING_LHS_Y_MPTR = 0x0b00
PAIRING_RHS_X_MPTR = 0x0b20
PAIRING_RHS_Y_MPTR = 0x0b40

PROOF_LEN = 0x0460          # contract.sol:221
VK_LEN = 0x03a0             # contract.sol:307
DELTA = 4131629893567559867359510883348571134090853742863529169391034518566172092834   # contract.sol:440

W = 1 << 256


class Revert(Exception):
    pass


class _Evm:
    """Just enough of the EVM for contract.sol: byte memory, calldata, the precompiles."""

    def __init__(self, calldata: bytes, vk_code: bytes):
        self.mem = bytearray(0x4000)
        self.calldata = calldata
        self.vk_code = vk_code

    # -- memory / calldata
    def mload(self, p):
        return int.from_bytes(self.mem[p:p + 32], "big")

    def mstore(self, p, v):
        self.mem[p:p + 32] = (v % W).to_bytes(32, "big")

    def mstore8(self, p, v):
        self.mem[p] = v & 0xFF

    def calldataload(self, p):
        chunk = self.calldata[p:p + 32]
        return int.from_bytes(chunk + b"\x00" * (32 - len(chunk)), "big")

    def keccak256(self, p, n):
        return int.from_bytes(keccak256(bytes(self.mem[p:p + n])), "big")

    def extcodecopy(self, dst, off, n):
        chunk = self.vk_code[off:off + n]
        self.mem[dst:dst + n] = chunk + b"\x00" * (n - len(chunk))

    # -- precompiles; each returns the `success` word of STATICCALL
    def _g1_in(self, p):
        x, y = self.mload(p), self.mload(p + 32)
        if x >= Q or y >= Q:
            return False, None
        if x == 0 and y == 0:
            return True, None
        if not bn.g1_is_on_curve((x, y)):
            return False, None
        return True, (x, y)

    def _g1_out(self, p, pt):
        self.mstore(p, pt[0] if pt else 0)
        self.mstore(p + 32, pt[1] if pt else 0)

    def staticcall(self, addr, in_ptr, in_len, out_ptr, out_len):
        if addr == 0x05:       # MODEXP (EIP-198), as called at contract.sol:129-135
            assert in_len == 0xc0 and out_len == 0x20
            lb, le, lm = self.mload(in_ptr), self.mload(in_ptr + 0x20), self.mload(in_ptr + 0x40)
            assert lb == le == lm == 0x20
            b, e, m = self.mload(in_ptr + 0x60), self.mload(in_ptr + 0x80), self.mload(in_ptr + 0xa0)
            self.mstore(out_ptr, pow(b, e, m) if m else 0)
            return 1
        if addr == 0x06:       # ECADD (EIP-196)
            assert in_len == 0x80 and out_len == 0x40
            ok1, p1 = self._g1_in(in_ptr)
            ok2, p2 = self._g1_in(in_ptr + 0x40)
            if not (ok1 and ok2):
                return 0
            self._g1_out(out_ptr, bn.g1_add(p1, p2))
            return 1
        if addr == 0x07:       # ECMUL (EIP-196)
            assert in_len == 0x60 and out_len == 0x40
            ok, p1 = self._g1_in(in_ptr)
            if not ok:
                return 0
            s = self.mload(in_ptr + 0x40)
            self._g1_out(out_ptr, bn.g1_mul(p1, s % R) if p1 else None)
            return 1
        if addr == 0x08:       # ECPAIRING (EIP-197): (G1, G2) pairs, Fq2 words in (imaginary, real) order
            assert in_len % 0xc0 == 0 and out_len == 0x20
            pairs = []
            for off in range(in_ptr, in_ptr + in_len, 0xc0):
                ok, p1 = self._g1_in(off)
                if not ok:
                    return 0
                xi, xr, yi, yr = (self.mload(off + 0x40 + 0x20 * j) for j in range(4))
                if max(xi, xr, yi, yr) >= Q:
                    return 0
                q2 = None if xi == xr == yi == yr == 0 else ((xr, xi), (yr, yi))
                if not pairing.g2_is_on_curve(q2):
                    return 0
                if q2 is not None and pairing.g2_mul(q2, R - 1) != pairing.g2_neg(q2):
                    return 0                                   # not in the order-r subgroup
                pairs.append((p1, q2))
            self.mstore(out_ptr, 1 if pairing.pairing_check(pairs) else 0)
            return 1
        raise AssertionError(f"unexpected precompile {addr:#x}")


def encode_calldata(proof: bytes, instances) -> bytes:
    """ABI encoding of verifyProof(address vk, bytes proof, uint256[] instances) with the offsets
    the contract hard-codes (contract.sol:6-9): proof length at 0x64, proof at 0x84, instance count
    at 0x04e4, instances at 0x0504."""
    assert len(proof) == PROOF_LEN
    head = b"\x00" * 4 + (0).to_bytes(32, "big") + (0x60).to_bytes(32, "big") + (0x60 + 0x20 + PROOF_LEN).to_bytes(32, "big")
    body = len(proof).to_bytes(32, "big") + proof
    tail = len(instances).to_bytes(32, "big") + b"".join(int(v).to_bytes(32, "big") for v in instances)
    data = head + body + tail
    assert len(head) == PROOF_LEN_CPTR and len(head + body) == NUM_INSTANCE_CPTR
    return data


def encode_vk(vk_digest: int, num_instances: int, k: int, omega: int, g1, g2, s_g2, fixed_comms, permutation_comms,
              blinding_factors: int = 5) -> bytes:
    """The 0x3a0 bytes `extcodecopy(vk, VK_MPTR, 0x00, 0x03a0)` (contract.sol:307) loads: the 21
    words of contract.sol:14-35, then the fixed-column and permutation commitments the final MSM
    reads back from 0x0720 upwards (:727-733).  No accumulator."""
    n_inv = pow(1 << k, -1, R)
    omega_inv = pow(omega, -1, R)
    neg_s_g2 = pairing.g2_neg(s_g2)
    words = [vk_digest, num_instances, k, n_inv, omega, omega_inv, pow(omega_inv, blinding_factors + 1, R),
             0, 0, 0, 0, g1[0], g1[1],
             g2[0][1], g2[0][0], g2[1][1], g2[1][0],
             neg_s_g2[0][1], neg_s_g2[0][0], neg_s_g2[1][1], neg_s_g2[1][0]]
    for pt in list(fixed_comms) + list(permutation_comms):
        words += [pt[0], pt[1]]
    blob = b"".join(int(w).to_bytes(32, "big") for w in words)
    assert len(blob) == VK_LEN, "this contract is generated for 1 fixed + 3 permutation commitments"
    return blob


def verify_proof(vk_code: bytes, proof: bytes, instances) -> bool:
    """contract.sol:68-826.  True iff the contract would return 1; False where it reverts."""
    evm = _Evm(encode_calldata(proof, instances), vk_code)
    try:
        return _verify(evm)
    except Revert:
        return False


def _verify(evm: _Evm) -> bool:
    mload, mstore, calldataload = evm.mload, evm.mstore, evm.calldataload
    q, r = Q, R

    def addmod(a, b, m): return (a + b) % m
    def mulmod(a, b, m): return (a * b) % m
    def sub(a, b): return (a - b) % W

    # contract.sol:73-87
    def read_ec_point(success, proof_cptr, hash_mptr):
        x = calldataload(proof_cptr)
        y = calldataload(proof_cptr + 0x20)
        ret0 = success and x < q
        ret0 = ret0 and y < q
        ret0 = ret0 and mulmod(y, y, q) == addmod(mulmod(x, mulmod(x, x, q), q), 3, q)
        mstore(hash_mptr, x)
        mstore(hash_mptr + 0x20, y)
        return ret0, proof_cptr + 0x40, hash_mptr + 0x40

    # contract.sol:89-99
    def squeeze_challenge(challenge_mptr, hash_mptr):
        h = evm.keccak256(0x00, hash_mptr)
        mstore(challenge_mptr, h % r)
        mstore(0x00, h)
        return challenge_mptr + 0x20, 0x20

    # contract.sol:101-112
    def squeeze_challenge_cont(challenge_mptr):
        evm.mstore8(0x20, 0x01)
        h = evm.keccak256(0x00, 0x21)
        mstore(challenge_mptr, h % r)
        mstore(0x00, h)
        return challenge_mptr + 0x20

    # contract.sol:114-159
    def batch_invert(success, mptr_start, mptr_end):
        gp_mptr = mptr_end
        gp = mload(mptr_start)
        mptr = mptr_start + 0x20
        while mptr < sub(mptr_end, 0x20):
            gp = mulmod(gp, mload(mptr), r)
            mstore(gp_mptr, gp)
            mptr += 0x20
            gp_mptr += 0x20
        gp = mulmod(gp, mload(mptr), r)
        mstore(gp_mptr, 0x20)
        mstore(gp_mptr + 0x20, 0x20)
        mstore(gp_mptr + 0x40, 0x20)
        mstore(gp_mptr + 0x60, gp)
        mstore(gp_mptr + 0x80, sub(r, 2))
        mstore(gp_mptr + 0xa0, r)
        ret = success and bool(evm.staticcall(0x05, gp_mptr, 0xc0, gp_mptr, 0x20))
        all_inv = mload(gp_mptr)
        first_mptr = mptr_start
        second_mptr = first_mptr + 0x20
        gp_mptr = sub(gp_mptr, 0x20)
        while second_mptr < mptr:
            inv = mulmod(all_inv, mload(gp_mptr), r)
            all_inv = mulmod(all_inv, mload(mptr), r)
            mstore(mptr, inv)
            mptr = sub(mptr, 0x20)
            gp_mptr = sub(gp_mptr, 0x20)
        inv_first = mulmod(all_inv, mload(second_mptr), r)
        inv_second = mulmod(all_inv, mload(first_mptr), r)
        mstore(first_mptr, inv_first)
        mstore(second_mptr, inv_second)
        return ret

    # contract.sol:161-188
    def ec_add_acc(success, x, y):
        mstore(0x40, x)
        mstore(0x60, y)
        return success and bool(evm.staticcall(0x06, 0x00, 0x80, 0x00, 0x40))

    def ec_mul_acc(success, scalar):
        mstore(0x40, scalar)
        return success and bool(evm.staticcall(0x07, 0x00, 0x60, 0x00, 0x40))

    def ec_add_tmp(success, x, y):
        mstore(0xc0, x)
        mstore(0xe0, y)
        return success and bool(evm.staticcall(0x06, 0x80, 0x80, 0x80, 0x40))

    def ec_mul_tmp(success, scalar):
        mstore(0xc0, scalar)
        return success and bool(evm.staticcall(0x07, 0x80, 0x60, 0x80, 0x40))

    # contract.sol:190-207
    def ec_pairing(success, lhs_x, lhs_y, rhs_x, rhs_y):
        mstore(0x00, lhs_x)
        mstore(0x20, lhs_y)
        mstore(0x40, mload(G2_X_1_MPTR))
        mstore(0x60, mload(G2_X_2_MPTR))
        mstore(0x80, mload(G2_Y_1_MPTR))
        mstore(0xa0, mload(G2_Y_2_MPTR))
        mstore(0xc0, rhs_x)
        mstore(0xe0, rhs_y)
        mstore(0x100, mload(NEG_S_G2_X_1_MPTR))
        mstore(0x120, mload(NEG_S_G2_X_2_MPTR))
        mstore(0x140, mload(NEG_S_G2_Y_1_MPTR))
        mstore(0x160, mload(NEG_S_G2_Y_2_MPTR))
        ret = success and bool(evm.staticcall(0x08, 0x00, 0x180, 0x00, 0x20))
        return ret and bool(mload(0x00))

    success = True

    # ---- contract.sol:216-352: transcript, challenges, calldata checks
    evm.extcodecopy(VK_MPTR, 0x00, 0x40)
    success = success and calldataload(PROOF_LEN_CPTR) == PROOF_LEN
    num_instances = mload(NUM_INSTANCES_MPTR)
    success = success and num_instances == calldataload(NUM_INSTANCE_CPTR)
    mstore(0x00, mload(VK_DIGEST_MPTR))
    hash_mptr = 0x20
    instance_cptr = INSTANCE_CPTR
    instance_cptr_end = instance_cptr + 0x20 * num_instances
    while instance_cptr < instance_cptr_end:
        instance = calldataload(instance_cptr)
        success = success and instance < r
        mstore(hash_mptr, instance)
        instance_cptr += 0x20
        hash_mptr += 0x20

    proof_cptr = PROOF_CPTR
    challenge_mptr = CHALLENGE_MPTR

    # Phase 1 (:248-259): two advice commitments; theta, beta, gamma
    proof_cptr_end = proof_cptr + 0x80
    while proof_cptr < proof_cptr_end:
        success, proof_cptr, hash_mptr = read_ec_point(success, proof_cptr, hash_mptr)
    challenge_mptr, hash_mptr = squeeze_challenge(challenge_mptr, hash_mptr)
    challenge_mptr = squeeze_challenge_cont(challenge_mptr)
    challenge_mptr = squeeze_challenge_cont(challenge_mptr)

    # Phase 2 (:261-270): three permutation products + the vanishing argument's random polynomial; y
    proof_cptr_end = proof_cptr + 0x0100
    while proof_cptr < proof_cptr_end:
        success, proof_cptr, hash_mptr = read_ec_point(success, proof_cptr, hash_mptr)
    challenge_mptr, hash_mptr = squeeze_challenge(challenge_mptr, hash_mptr)

    # Phase 3 (:272-281): two quotient pieces; x
    proof_cptr_end = proof_cptr + 0x80
    while proof_cptr < proof_cptr_end:
        success, proof_cptr, hash_mptr = read_ec_point(success, proof_cptr, hash_mptr)
    challenge_mptr, hash_mptr = squeeze_challenge(challenge_mptr, hash_mptr)

    # Evaluations (:283-294)
    proof_cptr_end = proof_cptr + 0x01e0
    while proof_cptr < proof_cptr_end:
        ev = calldataload(proof_cptr)
        success = success and ev < r
        mstore(hash_mptr, ev)
        proof_cptr += 0x20
        hash_mptr += 0x20

    # Batch opening proof (:296-304)
    challenge_mptr, hash_mptr = squeeze_challenge(challenge_mptr, hash_mptr)       # zeta
    challenge_mptr = squeeze_challenge_cont(challenge_mptr)                        # nu
    success, proof_cptr, hash_mptr = read_ec_point(success, proof_cptr, hash_mptr)  # W
    challenge_mptr, hash_mptr = squeeze_challenge(challenge_mptr, hash_mptr)       # mu
    success, proof_cptr, hash_mptr = read_ec_point(success, proof_cptr, hash_mptr)  # W'

    evm.extcodecopy(VK_MPTR, 0x00, VK_LEN)                                          # :307

    # :309-349: accumulator limbs carried in the instances (absent for this circuit)
    if mload(HAS_ACCUMULATOR_MPTR):
        num_limbs = mload(NUM_ACC_LIMBS_MPTR)
        num_limb_bits = mload(NUM_ACC_LIMB_BITS_MPTR)
        cptr = INSTANCE_CPTR + mload(ACC_OFFSET_MPTR) * 0x20
        lhs_y_off = num_limbs * 0x20
        rhs_x_off = lhs_y_off * 2
        rhs_y_off = lhs_y_off * 3
        lhs_x = calldataload(cptr)
        lhs_y = calldataload(cptr + lhs_y_off)
        rhs_x = calldataload(cptr + rhs_x_off)
        rhs_y = calldataload(cptr + rhs_y_off)
        cptr_end = cptr + 0x20 * num_limbs
        shift = num_limb_bits
        while cptr < cptr_end:
            cptr += 0x20
            lhs_x = (lhs_x + (calldataload(cptr) << shift)) % W
            lhs_y = (lhs_y + (calldataload(cptr + lhs_y_off) << shift)) % W
            rhs_x = (rhs_x + (calldataload(cptr + rhs_x_off) << shift)) % W
            rhs_y = (rhs_y + (calldataload(cptr + rhs_y_off) << shift)) % W
            shift += num_limb_bits
        success = success and lhs_x < q and lhs_y < q
        success = success and mulmod(lhs_y, lhs_y, q) == addmod(mulmod(lhs_x, mulmod(lhs_x, lhs_x, q), q), 3, q)
        success = success and rhs_x < q and rhs_y < q
        success = success and mulmod(rhs_y, rhs_y, q) == addmod(mulmod(rhs_x, mulmod(rhs_x, rhs_x, q), q), 3, q)
        mstore(ACC_LHS_X_MPTR, lhs_x)
        mstore(ACC_LHS_Y_MPTR, lhs_y)
        mstore(ACC_RHS_X_MPTR, rhs_x)
        mstore(ACC_RHS_Y_MPTR, rhs_y)

    if not success:                                                                 # :354-357
        raise Revert()

    # ---- contract.sol:359-437: x^n, Lagrange evaluations, instance evaluation
    k = mload(K_MPTR)
    x = mload(X_MPTR)
    x_n = x
    for _ in range(k):
        x_n = mulmod(x_n, x_n, r)
    omega = mload(OMEGA_MPTR)
    mptr = X_N_MPTR
    mptr_end = mptr + 0x20 * (mload(NUM_INSTANCES_MPTR) + 6)
    if mload(NUM_INSTANCES_MPTR) == 0:
        mptr_end += 0x20
    pow_of_omega = mload(OMEGA_INV_TO_L_MPTR)
    while mptr < mptr_end:
        mstore(mptr, addmod(x, sub(r, pow_of_omega), r))
        pow_of_omega = mulmod(pow_of_omega, omega, r)
        mptr += 0x20
    x_n_minus_1 = addmod(x_n, sub(r, 1), r)
    mstore(mptr_end, x_n_minus_1)
    success = batch_invert(success, X_N_MPTR, mptr_end + 0x20)

    mptr = X_N_MPTR
    l_i_common = mulmod(x_n_minus_1, mload(N_INV_MPTR), r)
    pow_of_omega = mload(OMEGA_INV_TO_L_MPTR)
    while mptr < mptr_end:
        mstore(mptr, mulmod(l_i_common, mulmod(mload(mptr), pow_of_omega, r), r))
        pow_of_omega = mulmod(pow_of_omega, omega, r)
        mptr += 0x20

    l_blind = mload(X_N_MPTR + 0x20)
    l_i_cptr = X_N_MPTR + 0x40
    l_i_cptr_end = X_N_MPTR + 0xc0
    while l_i_cptr < l_i_cptr_end:
        l_blind = addmod(l_blind, mload(l_i_cptr), r)
        l_i_cptr += 0x20

    instance_eval = 0
    instance_cptr = INSTANCE_CPTR
    instance_cptr_end = instance_cptr + 0x20 * mload(NUM_INSTANCES_MPTR)
    while instance_cptr < instance_cptr_end:
        instance_eval = addmod(instance_eval, mulmod(mload(l_i_cptr), calldataload(instance_cptr), r), r)
        instance_cptr += 0x20
        l_i_cptr += 0x20

    x_n_minus_1_inv = mload(mptr_end)
    l_last = mload(X_N_MPTR)
    l_0 = mload(X_N_MPTR + 0xc0)
    mstore(X_N_MPTR, x_n)
    mstore(X_N_MINUS_1_INV_MPTR, x_n_minus_1_inv)
    mstore(L_LAST_MPTR, l_last)
    mstore(L_BLIND_MPTR, l_blind)
    mstore(L_0_MPTR, l_0)
    mstore(INSTANCE_EVAL_MPTR, instance_eval)

    # ---- contract.sol:439-511: the quotient identity, folded with y
    delta = DELTA
    y = mload(Y_MPTR)
    f_0 = calldataload(0x02c4)
    a_1 = calldataload(0x02a4)
    a_0 = calldataload(0x0284)
    var0 = mulmod(a_0, a_0, r)
    var1 = sub(r, var0)
    var2 = addmod(a_1, var1, r)
    var3 = mulmod(f_0, var2, r)
    quotient_eval_numer = var3                                                       # :443-452
    l_0 = mload(L_0_MPTR)                                                            # :453-457
    ev = addmod(l_0, sub(r, mulmod(l_0, calldataload(0x0364), r)), r)
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    perm_z_last = calldataload(0x0424)                                               # :458-462
    ev = mulmod(mload(L_LAST_MPTR), addmod(mulmod(perm_z_last, perm_z_last, r), sub(r, perm_z_last), r), r)
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    ev = mulmod(mload(L_0_MPTR), addmod(calldataload(0x03c4), sub(r, calldataload(0x03a4)), r), r)   # :463-466
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    ev = mulmod(mload(L_0_MPTR), addmod(calldataload(0x0424), sub(r, calldataload(0x0404)), r), r)   # :467-470
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    # :471-484 first permutation column (advice 0)
    gamma = mload(GAMMA_MPTR)
    beta = mload(BETA_MPTR)
    lhs = calldataload(0x0384)
    rhs = calldataload(0x0364)
    lhs = mulmod(lhs, addmod(addmod(calldataload(0x0284), mulmod(beta, calldataload(0x0304), r), r), gamma, r), r)
    mstore(0x00, mulmod(beta, mload(X_MPTR), r))
    rhs = mulmod(rhs, addmod(addmod(calldataload(0x0284), mload(0x00), r), gamma, r), r)
    mstore(0x00, mulmod(mload(0x00), delta, r))
    left_sub_right = addmod(lhs, sub(r, rhs), r)
    ev = addmod(left_sub_right, sub(r, mulmod(left_sub_right, addmod(mload(L_LAST_MPTR), mload(L_BLIND_MPTR), r), r)), r)
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    # :485-496 second permutation column (advice 1)
    lhs = calldataload(0x03e4)
    rhs = calldataload(0x03c4)
    lhs = mulmod(lhs, addmod(addmod(calldataload(0x02a4), mulmod(beta, calldataload(0x0324), r), r), gamma, r), r)
    rhs = mulmod(rhs, addmod(addmod(calldataload(0x02a4), mload(0x00), r), gamma, r), r)
    mstore(0x00, mulmod(mload(0x00), delta, r))
    left_sub_right = addmod(lhs, sub(r, rhs), r)
    ev = addmod(left_sub_right, sub(r, mulmod(left_sub_right, addmod(mload(L_LAST_MPTR), mload(L_BLIND_MPTR), r), r)), r)
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    # :497-507 third permutation column (the instance column, evaluated by the verifier)
    lhs = calldataload(0x0444)
    rhs = calldataload(0x0424)
    lhs = mulmod(lhs, addmod(addmod(mload(INSTANCE_EVAL_MPTR), mulmod(beta, calldataload(0x0344), r), r), gamma, r), r)
    rhs = mulmod(rhs, addmod(addmod(mload(INSTANCE_EVAL_MPTR), mload(0x00), r), gamma, r), r)
    left_sub_right = addmod(lhs, sub(r, rhs), r)
    ev = addmod(left_sub_right, sub(r, mulmod(left_sub_right, addmod(mload(L_LAST_MPTR), mload(L_BLIND_MPTR), r), r)), r)
    quotient_eval_numer = addmod(mulmod(quotient_eval_numer, y, r), ev, r)
    quotient_eval = mulmod(quotient_eval_numer, mload(X_N_MINUS_1_INV_MPTR), r)      # :510-511
    mstore(QUOTIENT_EVAL_MPTR, quotient_eval)

    # ---- contract.sol:516-534: h = h_0 + x^n h_1 (Horner from the last piece)
    mstore(0x00, calldataload(LAST_QUOTIENT_X_CPTR))
    mstore(0x20, calldataload(LAST_QUOTIENT_X_CPTR + 0x20))
    x_n = mload(X_N_MPTR)
    cptr = sub(LAST_QUOTIENT_X_CPTR, 0x40)
    cptr_end = sub(FIRST_QUOTIENT_X_CPTR, 0x40)
    while cptr_end < cptr:
        success = ec_mul_acc(success, x_n)
        success = ec_add_acc(success, calldataload(cptr), calldataload(cptr + 0x20))
        cptr = sub(cptr, 0x40)
    mstore(QUOTIENT_X_MPTR, mload(0x00))
    mstore(QUOTIENT_Y_MPTR, mload(0x20))

    # ---- contract.sol:537-779: SHPLONK.  Opening points x w^-6, x, x w (:539-553)
    x = mload(X_MPTR)
    omega = mload(OMEGA_MPTR)
    omega_inv = mload(OMEGA_INV_MPTR)
    x_pow_of_omega = mulmod(x, omega, r)
    mstore(0x02c0, x_pow_of_omega)
    mstore(0x02a0, x)
    x_pow_of_omega = mulmod(x, omega_inv, r)
    for _ in range(5):
        x_pow_of_omega = mulmod(x_pow_of_omega, omega_inv, r)
    mstore(0x0280, x_pow_of_omega)
    # :552-580  mu - point_i, the set vanishing value and the set differences
    mu = mload(MU_MPTR)
    mptr, mptr_end, point_mptr = 0x02e0, 0x0340, 0x0280
    while mptr < mptr_end:
        mstore(mptr, addmod(mu, sub(r, mload(point_mptr)), r))
        mptr += 0x20
        point_mptr += 0x20
    s = mload(0x0300)
    mstore(0x0340, s)
    diff = mload(0x02e0)
    diff = mulmod(diff, mload(0x0320), r)
    mstore(0x0360, diff)
    mstore(0x00, diff)
    diff = 1
    mstore(0x0380, diff)
    diff = mload(0x02e0)
    mstore(0x03a0, diff)
    # :581-587  barycentric weights of set 0 = {x}
    coeff = 1
    coeff = mulmod(coeff, mload(0x0300), r)
    mstore(0x20, coeff)
    # :588-606  set 1 = {x w^-6, x, x w}
    point_0, point_1, point_2 = mload(0x0280), mload(0x02a0), mload(0x02c0)
    coeff = addmod(point_0, sub(r, point_1), r)
    coeff = mulmod(coeff, addmod(point_0, sub(r, point_2), r), r)
    coeff = mulmod(coeff, mload(0x02e0), r)
    mstore(0x40, coeff)
    coeff = addmod(point_1, sub(r, point_0), r)
    coeff = mulmod(coeff, addmod(point_1, sub(r, point_2), r), r)
    coeff = mulmod(coeff, mload(0x0300), r)
    mstore(0x60, coeff)
    coeff = addmod(point_2, sub(r, point_0), r)
    coeff = mulmod(coeff, addmod(point_2, sub(r, point_1), r), r)
    coeff = mulmod(coeff, mload(0x0320), r)
    mstore(0x80, coeff)
    # :607-617  set 2 = {x, x w}
    coeff = addmod(point_1, sub(r, point_2), r)
    coeff = mulmod(coeff, mload(0x0300), r)
    mstore(0xa0, coeff)
    coeff = addmod(point_2, sub(r, point_1), r)
    coeff = mulmod(coeff, mload(0x0320), r)
    mstore(0xc0, coeff)
    # :618-633
    success = batch_invert(success, 0, 0xe0)
    diff_0_inv = mload(0x00)
    mstore(0x0360, diff_0_inv)
    mptr, mptr_end = 0x0380, 0x03c0
    while mptr < mptr_end:
        mstore(mptr, mulmod(mload(mptr), diff_0_inv, r))
        mptr += 0x20
    # :634-660  r_eval of set 0: a_0, a_1, f_0, sigma_0..2, h, random  (powers of zeta, Horner)
    coeff = mload(0x20)
    zeta = mload(ZETA_MPTR)
    r_eval = mulmod(coeff, calldataload(0x02e4), r)
    r_eval = mulmod(r_eval, zeta, r)
    r_eval = addmod(r_eval, mulmod(coeff, mload(QUOTIENT_EVAL_MPTR), r), r)
    cptr, cptr_end = 0x0344, 0x02e4
    while cptr_end < cptr:
        r_eval = addmod(mulmod(r_eval, zeta, r), mulmod(coeff, calldataload(cptr), r), r)
        cptr = sub(cptr, 0x20)
    cptr, cptr_end = 0x02c4, 0x0264
    while cptr_end < cptr:
        r_eval = addmod(mulmod(r_eval, zeta, r), mulmod(coeff, calldataload(cptr), r), r)
        cptr = sub(cptr, 0x20)
    mstore(0x03c0, r_eval)
    # :661-673  set 1: z_0, z_1
    r_eval = 0
    r_eval = addmod(r_eval, mulmod(mload(0x40), calldataload(0x0404), r), r)
    r_eval = addmod(r_eval, mulmod(mload(0x60), calldataload(0x03c4), r), r)
    r_eval = addmod(r_eval, mulmod(mload(0x80), calldataload(0x03e4), r), r)
    r_eval = mulmod(r_eval, zeta, r)
    r_eval = addmod(r_eval, mulmod(mload(0x40), calldataload(0x03a4), r), r)
    r_eval = addmod(r_eval, mulmod(mload(0x60), calldataload(0x0364), r), r)
    r_eval = addmod(r_eval, mulmod(mload(0x80), calldataload(0x0384), r), r)
    r_eval = mulmod(r_eval, mload(0x0380), r)
    mstore(0x03e0, r_eval)
    # :674-681  set 2: z_2
    r_eval = 0
    r_eval = addmod(r_eval, mulmod(mload(0xa0), calldataload(0x0424), r), r)
    r_eval = addmod(r_eval, mulmod(mload(0xc0), calldataload(0x0444), r), r)
    r_eval = mulmod(r_eval, mload(0x03a0), r)
    mstore(0x0400, r_eval)
    # :682-697  sums of the barycentric weights
    mstore(0x0420, mload(0x20))
    total = mload(0x40)
    total = addmod(total, mload(0x60), r)
    total = addmod(total, mload(0x80), r)
    mstore(0x0440, total)
    total = mload(0xa0)
    total = addmod(total, mload(0xc0), r)
    mstore(0x0460, total)
    # :698-729
    mptr, mptr_end, sum_mptr = 0x00, 0x60, 0x0420
    while mptr < mptr_end:
        mstore(mptr, mload(sum_mptr))
        mptr += 0x20
        sum_mptr += 0x20
    success = batch_invert(success, 0, 0x60)
    r_eval = mulmod(mload(0x40), mload(0x0400), r)
    sum_inv_mptr, sum_inv_mptr_end, r_eval_mptr = 0x20, 0x60, 0x03e0
    while sum_inv_mptr < sum_inv_mptr_end:                    # the pointer walks down and wraps: `lt` is unsigned
        r_eval = mulmod(r_eval, mload(NU_MPTR), r)
        r_eval = addmod(r_eval, mulmod(mload(sum_inv_mptr), mload(r_eval_mptr), r), r)
        sum_inv_mptr = sub(sum_inv_mptr, 0x20)
        r_eval_mptr = sub(r_eval_mptr, 0x20)
    mstore(G1_SCALAR_MPTR, sub(r, r_eval))
    # :731-779  the left-hand side of the pairing
    zeta = mload(ZETA_MPTR)
    nu = mload(NU_MPTR)
    mstore(0x00, calldataload(0x01c4))
    mstore(0x20, calldataload(0x01e4))
    success = ec_mul_acc(success, zeta)
    success = ec_add_acc(success, mload(QUOTIENT_X_MPTR), mload(QUOTIENT_Y_MPTR))
    ptr, ptr_end = 0x07e0, 0x06e0
    while ptr_end < ptr:
        success = ec_mul_acc(success, zeta)
        success = ec_add_acc(success, mload(ptr), mload(ptr + 0x20))
        ptr = sub(ptr, 0x40)
    success = ec_mul_acc(success, zeta)
    success = ec_add_acc(success, calldataload(0xc4), calldataload(0xe4))
    success = ec_mul_acc(success, zeta)
    success = ec_add_acc(success, calldataload(0x84), calldataload(0xa4))
    mstore(0x80, calldataload(0x0144))
    mstore(0xa0, calldataload(0x0164))
    success = ec_mul_tmp(success, zeta)
    success = ec_add_tmp(success, calldataload(0x0104), calldataload(0x0124))
    success = ec_mul_tmp(success, mulmod(nu, mload(0x0380), r))
    success = ec_add_acc(success, mload(0x80), mload(0xa0))
    nu = mulmod(nu, mload(NU_MPTR), r)
    mstore(0x80, calldataload(0x0184))
    mstore(0xa0, calldataload(0x01a4))
    success = ec_mul_tmp(success, mulmod(nu, mload(0x03a0), r))
    success = ec_add_acc(success, mload(0x80), mload(0xa0))
    mstore(0x80, mload(G1_X_MPTR))
    mstore(0xa0, mload(G1_Y_MPTR))
    success = ec_mul_tmp(success, mload(G1_SCALAR_MPTR))
    success = ec_add_acc(success, mload(0x80), mload(0xa0))
    mstore(0x80, calldataload(0x0464))
    mstore(0xa0, calldataload(0x0484))
    success = ec_mul_tmp(success, sub(r, mload(0x0340)))
    success = ec_add_acc(success, mload(0x80), mload(0xa0))
    mstore(0x80, calldataload(0x04a4))
    mstore(0xa0, calldataload(0x04c4))
    success = ec_mul_tmp(success, mload(MU_MPTR))
    success = ec_add_acc(success, mload(0x80), mload(0xa0))
    mstore(PAIRING_LHS_X_MPTR, mload(0x00))
    mstore(PAIRING_LHS_Y_MPTR, mload(0x20))
    mstore(PAIRING_RHS_X_MPTR, calldataload(0x04a4))
    mstore(PAIRING_RHS_Y_MPTR, calldataload(0x04c4))

    # ---- contract.sol:783-810: fold in the accumulator (absent here)
    if mload(HAS_ACCUMULATOR_MPTR):
        mstore(0x00, mload(ACC_LHS_X_MPTR))
        mstore(0x20, mload(ACC_LHS_Y_MPTR))
        mstore(0x40, mload(ACC_RHS_X_MPTR))
        mstore(0x60, mload(ACC_RHS_Y_MPTR))
        mstore(0x80, mload(PAIRING_LHS_X_MPTR))
        mstore(0xa0, mload(PAIRING_LHS_Y_MPTR))
        mstore(0xc0, mload(PAIRING_RHS_X_MPTR))
        mstore(0xe0, mload(PAIRING_RHS_Y_MPTR))
        challenge = evm.keccak256(0x00, 0x100) % r
        success = ec_mul_acc(success, challenge)
        success = ec_add_acc(success, mload(PAIRING_LHS_X_MPTR), mload(PAIRING_LHS_Y_MPTR))
        mstore(PAIRING_LHS_X_MPTR, mload(0x00))
        mstore(PAIRING_LHS_Y_MPTR, mload(0x20))
        mstore(0x00, mload(ACC_RHS_X_MPTR))
        mstore(0x20, mload(ACC_RHS_Y_MPTR))
        success = ec_mul_acc(success, challenge)
        success = ec_add_acc(success, mload(PAIRING_RHS_X_MPTR), mload(PAIRING_RHS_Y_MPTR))
        mstore(PAIRING_RHS_X_MPTR, mload(0x00))
        mstore(PAIRING_RHS_Y_MPTR, mload(0x20))

    # ---- contract.sol:813-826
    success = ec_pairing(success, mload(PAIRING_LHS_X_MPTR), mload(PAIRING_LHS_Y_MPTR),
                         mload(PAIRING_RHS_X_MPTR), mload(PAIRING_RHS_Y_MPTR))
    if not success:
        raise Revert()
    return True

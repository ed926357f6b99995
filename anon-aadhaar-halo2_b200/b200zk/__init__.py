"""b200zk — host-side mirror of the halo2_proofs hot-path interface over the C ABI.

Names, argument meaning and error behaviour follow the upstream Rust items the C ABI
replaces ([DEP] halo2_proofs 0.2.0 @ v2023_01_20, reference Cargo.lock:469-471):

    arithmetic::best_fft            -> best_fft(a, omega, log_n)
    arithmetic::best_multiexp       -> best_multiexp(coeffs, bases)
    poly::EvaluationDomain          -> EvaluationDomain(j, k).{lagrange_to_coeff, ...}
    poly::kzg::commitment::ParamsKZG -> ParamsKZG(g, g_lagrange).{commit, commit_lagrange}

Arrays are numpy ``uint64`` in the wire layout of halo2curves (4 LE limbs, Montgomery):
Fr vectors are (n, 4), G1Affine vectors (n, 8), a G1 result (12,).  Like upstream, bad
lengths are programming errors: they raise ``AssertionError`` (Rust ``assert_eq!``).
"""
from ._lib import B200zkError, LIB_PATH, header_symbols, load, check  # noqa: F401
from .api import (  # noqa: F401
    EvaluationDomain,
    ParamsKZG,
    best_fft,
    best_multiexp,
    fresh,
    g1_affine_from_bytes,
    g1_affine_to_bytes,
    g1_sum,
    g1_to_bytes,
    g1_to_evm_bytes,
    host_alloc_fr,
    host_free,
    init,
    kernel_launches,
    modmul_peak,
    pinned,
    shutdown,
)
from .quotient import (  # noqa: F401,E402
    DeviceColumn,
    Evaluator,
    FlatGraph,
    LookupCommitted,
    ProvingKeyCosets,
)
from .prover_steps import (  # noqa: F401,E402
    batch_invert,
    eval_polynomial,
    eval_polynomial_many,
    kate_division,
    linear_combination,
    lookup_products,
    permutation_products,
    permute_expression_pairs,
)

"""sonic_b200 — host-side mirror of the reference's interface for the prover hot path.

The names follow sdiehl/sonic (`SRS.new`, `commitPoly`, `openPoly`, `prove`, `hscProve`,
`ArithCircuit`, `Assignment`, `GateWeights`); every call goes through the C ABI of
``libsonic_b200.so`` (include/sonic_b200.h), exactly as the Haskell shim in
INTEGRATION.md would.  There is no CPU implementation behind these functions: importing
works anywhere, but the first call raises if the CUDA library or a GPU is missing.
"""
from .capi import SonicError, lib, init, shutdown, set_option, last_timing_ms, launch_count, device_count  # noqa: F401
from .api import (  # noqa: F401
    ArithCircuit,
    Assignment,
    GateWeights,
    HscProof,
    Proof,
    RndOracle,
    SRS,
    commitPoly,
    hscProve,
    hscProveBiV,
    msm,
    msm_partial,
    g1_sum,
    openPoly,
    pcv_fold,
    prove_shard,
    prove_batch,
    prove_combine,
    prove,
    prove_bytes,
    R_MODULUS,
)

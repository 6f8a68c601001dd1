"""quill_zkvm_b200 -- B200-native proving hot path of the Quill zkVM (KZG-commit MSM + sumcheck prover).

Host-side mirror of the reference's Rust API over the C ABI in include/quill_b200.h.  All compute runs in
hand-written sm_100a CUDA kernels inside libquill_b200.so; there is no CPU fallback.
"""
from ._lib import QuillError, load  # noqa: F401
from .api import (  # noqa: F401
    Context, DeviceBuffer, EvaluationClaim, KZG, KZGOpeningProof, MLEvalProof, SRS, SumcheckProof, Transcript,
    VirtualPolyExpr, VirtualPolynomialStore, ZeroCheckProof, fast_eq_eval_hypercube, logup_denominators,
)

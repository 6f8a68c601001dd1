"""ctypes loader for libquill_b200.so (the C ABI declared in include/quill_b200.h).

There is no CPU fallback: if the shared library is missing this raises, and without a CUDA device
`Context()` raises QuillError(QZ_ERR_NO_DEVICE).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libquill_b200.so")

QZ_OK, QZ_ERR_INVALID_ARG, QZ_ERR_DEGREE, QZ_ERR_CUDA, QZ_ERR_NCCL, QZ_ERR_EXPR, QZ_ERR_NO_DEVICE, QZ_ERR_ALLOC = range(8)
QZ_MAX_ROUND_COEFFS = 33

# every symbol include/quill_b200.h declares (tests check the .so exports each one)
SYMBOLS = [
    "qz_ctx_create", "qz_ctx_destroy", "qz_status_str", "qz_last_error", "qz_ctx_sync", "qz_kernel_launches",
    "qz_dev_alloc", "qz_dev_free", "qz_dev_trim", "qz_dev_upload", "qz_dev_download", "qz_dev_random_fr",
    "qz_transcript_new", "qz_transcript_append_bytes", "qz_transcript_draw_challenge", "qz_transcript_draw_fr",
    "qz_transcript_append_fr", "qz_transcript_append_g1", "qz_g1_serialize",
    "qz_srs_upload", "qz_srs_generate", "qz_srs_precompute", "qz_srs_free", "qz_srs_len", "qz_srs_download",
    "qz_msm", "qz_kzg_commit", "qz_kzg_open", "qz_mlpcs_open", "qz_mlpcs_open_begin", "qz_mlpcs_open_finish",
    "qz_compute_s_polynomial",
    "qz_sumcheck_prove", "qz_zerocheck_prove", "qz_eq_table", "qz_logup_denominators",
    "qz_comm_unique_id", "qz_comm_init", "qz_comm_peer_memory", "qz_msm_sharded", "qz_msm_split", "qz_sumcheck_prove_sharded", "qz_zerocheck_prove_sharded", "qz_comm_allgather_host", "qz_comm_resync",
    "qz_last_elapsed_ms", "qz_last_stat", "qz_msm_accumulate_stats", "qz_bench_imad", "qz_bench_fp_mul",
    "qz_test_field_op", "qz_test_fold", "qz_test_mid_plan", "qz_test_g1_add", "qz_test_g1_mul",
]

_lib = None


class QuillError(RuntimeError):
    def __init__(self, status: int, detail: str = ""):
        self.status = status
        super().__init__(f"quill_b200 status {status}: {detail}")


def load():
    """Load the shared library (once).  Raises OSError with a build hint when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("QZ_LIB_PATH", LIB_PATH)  # A/B measurements of alternative builds (tools/_libs/); default: in-tree
    if not os.path.exists(path):
        raise OSError(f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(make -C quill_zkvm_b200/csrc).  There is no CPU fallback.")
    # NCCL is bound lazily with dlopen inside the library; point it at torch's bundled copy when present
    if "QZ_NCCL_LIB" not in os.environ:
        try:
            import nvidia.nccl as _n  # type: ignore
            cand = os.path.join(list(_n.__path__)[0], "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["QZ_NCCL_LIB"] = cand
        except Exception:
            pass
    lib = C.CDLL(path)
    vp, sz, i32, u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
    lib.qz_ctx_create.argtypes = [i32, vp, C.POINTER(vp)]
    lib.qz_ctx_destroy.argtypes = [vp]
    lib.qz_ctx_destroy.restype = None
    lib.qz_status_str.argtypes = [i32]
    lib.qz_status_str.restype = C.c_char_p
    lib.qz_last_error.argtypes = [vp]
    lib.qz_last_error.restype = C.c_char_p
    lib.qz_ctx_sync.argtypes = [vp]
    lib.qz_kernel_launches.argtypes = [vp]
    lib.qz_kernel_launches.restype = u64
    lib.qz_dev_alloc.argtypes = [vp, sz, C.POINTER(vp)]
    lib.qz_dev_free.argtypes = [vp, vp]
    lib.qz_dev_trim.argtypes = [vp]
    lib.qz_dev_upload.argtypes = [vp, vp, vp, sz]
    lib.qz_dev_download.argtypes = [vp, vp, vp, sz]
    lib.qz_dev_random_fr.argtypes = [vp, vp, sz, u64]
    lib.qz_transcript_new.argtypes = [vp, sz, vp]
    lib.qz_transcript_new.restype = None
    lib.qz_transcript_append_bytes.argtypes = [vp, vp, sz]
    lib.qz_transcript_append_bytes.restype = None
    lib.qz_transcript_draw_challenge.argtypes = [vp, vp, sz]
    lib.qz_transcript_draw_challenge.restype = None
    lib.qz_transcript_draw_fr.argtypes = [vp, vp, vp]
    lib.qz_transcript_append_fr.argtypes = [vp, vp, vp]
    lib.qz_transcript_append_g1.argtypes = [vp, vp, vp]
    lib.qz_g1_serialize.argtypes = [vp, vp, vp]
    lib.qz_srs_upload.argtypes = [vp, vp, sz, C.POINTER(vp)]
    lib.qz_srs_generate.argtypes = [vp, vp, vp, sz, C.POINTER(vp)]
    lib.qz_srs_precompute.argtypes = [vp, vp, i32]
    lib.qz_srs_free.argtypes = [vp]
    lib.qz_srs_free.restype = None
    lib.qz_srs_len.argtypes = [vp]
    lib.qz_srs_len.restype = sz
    lib.qz_srs_download.argtypes = [vp, vp, sz, sz, vp]
    lib.qz_msm.argtypes = [vp, vp, vp, sz, i32, vp]
    lib.qz_kzg_commit.argtypes = [vp, vp, vp, sz, i32, vp]
    lib.qz_kzg_open.argtypes = [vp, vp, vp, sz, i32, vp, vp, vp]
    lib.qz_mlpcs_open.argtypes = [vp, vp, vp, sz, i32, vp, sz, vp, vp, vp, vp]
    lib.qz_mlpcs_open_begin.argtypes = [vp, vp, vp, sz, i32, vp, sz, vp, vp, C.POINTER(vp), C.POINTER(sz)]
    lib.qz_mlpcs_open_finish.argtypes = [vp, vp, vp, sz, i32, vp, sz, vp, vp]
    lib.qz_compute_s_polynomial.argtypes = [vp, vp, sz, vp, sz, vp]
    sc = [vp, sz, sz, vp, i32, vp, sz, vp, sz, vp, vp, sz, vp, vp, vp, vp]
    lib.qz_sumcheck_prove.argtypes = sc
    lib.qz_sumcheck_prove_sharded.argtypes = sc
    lib.qz_zerocheck_prove.argtypes = [vp, sz, sz, vp, i32, vp, sz, vp, sz, vp, sz, vp, vp, vp, vp, vp]
    lib.qz_zerocheck_prove_sharded.argtypes = lib.qz_zerocheck_prove.argtypes
    lib.qz_eq_table.argtypes = [vp, sz, vp, vp, i32]
    lib.qz_logup_denominators.argtypes = [vp, sz, sz, vp, i32, vp, sz, vp, sz, vp, sz, vp, vp, i32]
    lib.qz_comm_unique_id.argtypes = [vp]
    lib.qz_comm_init.argtypes = [vp, vp, i32, i32]
    lib.qz_comm_peer_memory.argtypes = [vp]
    lib.qz_msm_sharded.argtypes = [vp, vp, vp, sz, i32, vp]
    lib.qz_msm_split.argtypes = [vp, vp, vp, sz, i32, vp]
    lib.qz_comm_allgather_host.argtypes = [vp, vp, vp, sz]
    lib.qz_comm_resync.argtypes = [vp]
    lib.qz_last_elapsed_ms.argtypes = [vp, i32]
    lib.qz_last_elapsed_ms.restype = C.c_float
    lib.qz_last_stat.argtypes = [vp, i32]
    lib.qz_last_stat.restype = C.c_double
    lib.qz_msm_accumulate_stats.argtypes = [vp, i32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
    lib.qz_bench_imad.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.qz_bench_fp_mul.argtypes = [vp, i32, C.POINTER(C.c_double)]
    lib.qz_test_field_op.argtypes = [vp, i32, i32, vp, vp, vp, sz]
    lib.qz_test_fold.argtypes = [vp, vp, vp, vp, vp, sz]
    lib.qz_test_mid_plan.argtypes = [C.c_uint64, i32, i32, i32, C.c_uint32, i32, vp, vp, vp, vp]
    lib.qz_test_g1_add.argtypes = [vp, vp, vp, vp, sz]
    lib.qz_test_g1_mul.argtypes = [vp, vp, vp, vp, sz]
    _lib = lib
    return lib

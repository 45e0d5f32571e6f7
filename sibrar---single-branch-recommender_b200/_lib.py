"""ctypes binding of ``csrc/libsibrar_b200.so`` (the C ABI declared in ``include/sibrar_b200.h``).

There is no CPU fallback: if the shared library is missing and cannot be built, or a kernel call fails, this
module raises.  All pointers handed to the library are raw device pointers of caller-owned torch tensors.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
# SBR_LIB_PATH: another build of the same library (the timeline build `make -C csrc stamps`: scripts/step_timeline.py)
LIB_PATH = os.environ.get("SBR_LIB_PATH") or os.path.join(CSRC, "libsibrar_b200.so")

ACT = {None: 0, "none": 0, "relu": 1, "tanh": 2, "sigmoid": 3, "selu": 4}
SRC_TABLE, SRC_CATEGORICAL, SRC_TAG = 0, 1, 2
LOSS = {"bpr": 0, "bce": 1, "sampled_softmax": 2}

c_i64, c_i32, c_f32, c_vp, c_u64 = C.c_int64, C.c_int32, C.c_float, C.c_void_p, C.c_uint64


class GemmEpilogue(C.Structure):
    _fields_ = [("bias", c_vp), ("act", C.c_int), ("out_bf16", c_vp), ("ld_bf16", c_i64), ("out_f32", c_vp),
                ("ld_f32", c_i64), ("colstats", c_vp), ("colstats_sum_only", C.c_int), ("colstats_rows", C.c_int), ("actgrad_y", c_vp), ("ld_actgrad", c_i64),
                ("actgrad_act", C.c_int), ("transpose_out", C.c_int), ("atomic_out", C.c_int), ("split_k", C.c_int),
                ("split_stride", c_i64), ("alpha", c_f32)]


class ModalitySrc(C.Structure):
    _fields_ = [("kind", C.c_int), ("remap", c_vp), ("table", c_vp), ("grad", c_vp), ("codes", c_vp),
                ("max_tags", c_i32), ("pad_id", c_i32), ("n_table_rows", c_i64), ("key_base", c_i64)]


class BnInline(C.Structure):
    _fields_ = [("z", c_vp), ("mean_invstd", c_vp), ("gamma", c_vp), ("beta", c_vp), ("sums", c_vp)]


class Mlp2Layer(C.Structure):
    _fields_ = [("w_bf16", c_vp), ("ldw", c_i64), ("bias", c_vp), ("in_f", C.c_int), ("out_f", C.c_int),
                ("act", C.c_int)]


class Mlp2Desc(C.Structure):
    _fields_ = [("srcs", c_vp), ("n_mods", C.c_int), ("idx", c_vp), ("mods", c_vp), ("n_idx", c_i64), ("k", C.c_int),
                ("C", C.c_int), ("normalize", C.c_int), ("p_drop", c_f32), ("seed", c_u64), ("step_dev", c_vp),
                ("keep_mask", c_vp), ("err_flag", c_vp), ("n_layers", C.c_int), ("layers", Mlp2Layer * 2)]


class Mlp2BnTail(C.Structure):
    _fields_ = [("counter", c_vp), ("eps", c_f32), ("momentum", c_f32), ("mean_invstd", c_vp), ("running_mean", c_vp),
                ("running_var", c_vp), ("num_batches_tracked", c_vp)]


class Mlp2Bn(C.Structure):
    _fields_ = [("mean_invstd", c_vp), ("gamma", c_vp), ("sums", c_vp), ("n_replicas", C.c_int), ("dgamma", c_vp),
                ("dbeta", c_vp)]


class RefTable(C.Structure):
    _fields_ = [("stamp", c_vp), ("pos", c_vp), ("list", c_vp), ("count", c_vp), ("seg_first", c_vp),
                ("seg_list", c_vp), ("seg_count", c_vp)]


class AdamTensor(C.Structure):
    _fields_ = [("param", c_vp), ("grad", c_vp), ("exp_avg", c_vp), ("exp_avg_sq", c_vp), ("shadow_bf16", c_vp),
                ("numel", c_i64), ("cols", c_i64), ("shadow_ld", c_i64)]


class McComm(C.Structure):
    """sbr_mc_comm_t (include/sibrar_b200.h): symmetric buffers of the fused all-reduce + optimizer kernel"""
    _fields_ = [("flat_grads", c_vp), ("mc_grads", c_vp), ("sum_local", c_vp), ("sum_mc", c_vp), ("total", c_i64),
                ("peer_flags_dev", c_vp), ("flags_local", c_vp), ("state", c_vp), ("world", C.c_int), ("rank", C.c_int)]


_PROTOS = {
    "sbr_gemm_bf16": [c_vp, c_i64, C.c_int, c_vp, c_i64, C.c_int, c_i64, c_i64, c_i64, C.POINTER(GemmEpilogue), c_vp],
    "sbr_gemm_bits_bf16": [c_vp, c_i64, c_vp, c_i64, C.c_int, c_i64, c_i64, c_i64, c_vp, c_vp],
    "sbr_splitk_reduce": [c_vp, C.c_int, c_i64, c_i64, c_i64, c_i64, c_vp, C.c_int, c_vp, c_i64, C.c_int, c_vp, c_i64,
                          c_vp],
    "sbr_cast_f32_to_bf16": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp],
    "sbr_transpose_f32_to_bf16": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp],
    "sbr_transpose_f32": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp],
    "sbr_csr_to_dense_bf16": [c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp],
    "sbr_spmm_csr": [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, C.c_int, c_vp, c_i64, C.c_int, C.c_int, c_vp,
                     c_i64, c_vp, C.c_int, c_vp],
    "sbr_spmm_csr_bf16": [c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, c_i64, c_vp, C.c_int, c_vp, c_i64, C.c_int, C.c_int,
                          c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp],
    "sbr_mark_referenced": [c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp],
    "sbr_gather_rows_bf16": [c_vp, c_i64, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp],
    "sbr_spmm_scatter_wgrad": [c_vp, c_vp, c_vp, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, c_vp, c_i64, c_vp],
    "sbr_transpose_add_f32": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, c_vp],
    "sbr_sample_modalities": [c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_u64, c_vp, c_vp],
    "sbr_tick": [c_vp, c_vp],
    "sbr_step_begin": [c_vp, c_vp, c_vp, c_i64, c_vp],
    "sbr_row_gather_fwd": [c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_f32, c_u64, c_vp, c_vp,
                           c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp],
    "sbr_row_gather_bwd": [c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_f32, c_u64, c_vp, c_vp,
                           c_vp, c_i64, c_vp],
    "sbr_gather_plan": [c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_int, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "sbr_row_gather_bwd_segmented": [c_vp, C.c_int, c_i64, c_vp, c_vp, c_vp, c_i64, C.c_int, C.c_int, c_f32, c_u64,
                                     c_vp, c_vp, c_vp, c_i64, C.c_int, c_vp, c_vp],
    "sbr_tag_bag_fwd": [c_vp, C.c_int, C.c_int32, c_vp, c_i64, C.c_int, c_vp, c_vp],
    "sbr_tag_bag_bwd": [c_vp, C.c_int, C.c_int32, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp],
    "sbr_actgrad_colsum": [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, c_i64, c_i64, c_vp, c_i64, c_vp, c_i64, c_vp,
                           C.c_int, c_vp],
    "sbr_bn_finalize": [c_vp, C.c_int, c_i64, C.c_int, c_f32, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp],
    "sbr_gemm_colstats_rows": [c_i64, c_i64],
    "sbr_bn_apply": [c_vp, c_i64, c_vp, c_vp, c_vp, C.c_int, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, c_vp],
    "sbr_bn_eval_coeffs": [c_vp, c_vp, C.c_int, c_f32, c_vp, c_vp],
    "sbr_bn_bwd_reduce": [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, C.c_int, c_vp, c_vp],
    "sbr_bn_bwd_apply": [c_vp, c_i64, c_vp, c_vp, c_i64, C.c_int, c_vp, c_i64, c_vp, c_vp, c_vp, C.c_int, c_i64, C.c_int,
                         c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp],
    "sbr_score_loss": [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                       c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "sbr_score_bwd": [c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_vp, c_vp],
    "sbr_score_loss_bn": [c_vp, c_vp, c_vp, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, c_f32, c_vp, c_vp, c_vp,
                          c_vp, C.c_int, c_vp],
    "sbr_infonce": [c_vp, c_i64, c_i64, C.c_int, c_f32, c_f32, c_vp, c_vp, C.c_int, c_vp, c_vp],
    "sbr_infonce_split": [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp],
    "sbr_infonce_lse": [c_vp, c_i64, c_f32, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp],
    "sbr_infonce_weights": [c_vp, c_i64, c_vp, c_vp, c_vp, c_vp],
    "sbr_clamp_min_fwd": [c_vp, c_i64, c_f32, c_vp, c_vp],
    "sbr_clamp_min_bwd": [c_vp, c_i64, c_vp, c_vp],
    "sbr_logit_bias_fwd": [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "sbr_logit_bias_bwd": [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp],
    "sbr_aggregate": [c_vp, c_i64, C.c_int, C.c_int, C.c_int, c_vp, c_vp, c_i64, c_vp],
    "sbr_mlp2_colstats_rows": [c_i64],
    "sbr_mlp2_fwd": [C.POINTER(Mlp2Desc), c_i64, C.c_int, c_vp, c_i64, c_vp, C.c_int, c_vp],
    "sbr_mlp2_fwd_bn": [C.POINTER(Mlp2Desc), c_i64, C.c_int, c_vp, c_i64, c_vp, C.c_int, C.POINTER(Mlp2BnTail), c_vp],
    "sbr_mlp2_bwd": [C.POINTER(Mlp2Desc), c_i64, C.c_int, c_vp, c_i64, c_vp, c_i64, C.POINTER(Mlp2Bn),
                     C.POINTER(c_vp), C.POINTER(c_vp), c_vp, c_i64, c_vp],
    "sbr_mlp2_trace_read": [c_vp, C.c_int],
    "sbr_mlp2_cta_times_read": [c_vp, C.c_int],
    "sbr_adam_step": [c_vp, C.c_int, c_i64, c_vp, c_vp, c_f32, c_f32, c_f32, c_f32, c_f32, C.c_int, c_vp, c_f32,
                      c_vp],
    "sbr_adam_step_mc": [c_vp, C.c_int, c_i64, c_vp, c_vp, c_f32, c_f32, c_f32, c_f32, c_f32, C.c_int, c_vp, c_f32,
                         C.c_int, C.POINTER(McComm), C.c_int, c_vp],
    "sbr_topk_workspace_bytes": [c_i64, c_i64, C.c_int, C.c_int, C.c_int, C.POINTER(c_i64)],
    "sbr_topk_scores_masked": [c_vp, c_i64, c_vp, c_i64, c_i64, c_i64, C.c_int, c_vp, c_vp, C.c_int, C.c_int, c_i32,
                               c_vp, c_vp, c_i64, c_vp],
    "sbr_topk_merge": [c_vp, C.c_int, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp],
    "sbr_metrics_at_k": [c_vp, c_i64, C.c_int, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_i64, c_vp],
    "sbr_sample_batch": [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_u64, c_vp, c_vp, c_vp, c_vp],
    "sbr_sample_negatives": [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_vp, c_i64, C.c_int, C.c_int, c_u64,
                             c_vp, c_vp, c_vp, c_vp],
    "sbr_sample_epoch_batch": [c_vp, c_vp, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_i64, c_i64, C.c_int, c_u64, c_vp,
                               c_vp, c_vp, c_vp],
}

EXPORTS = sorted(list(_PROTOS) + ["sbr_last_error", "sbr_version"])


def build(verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a into ``csrc/libsibrar_b200.so`` (nvcc cross-compiles without a GPU)."""
    jobs = str(max(1, min(16, os.cpu_count() or 1)))
    r = subprocess.run(["make", "-C", CSRC, "-j", jobs], capture_output=True, text=True)
    if verbose or r.returncode != 0:
        print(r.stdout[-4000:])
        print(r.stderr[-4000:])
    if r.returncode != 0:
        raise RuntimeError("building libsibrar_b200.so failed (see output above)")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            build()
        _lib = C.CDLL(LIB_PATH)
        _lib.sbr_last_error.restype = C.c_char_p
        _lib.sbr_version.restype = C.c_int
        for name, args in _PROTOS.items():
            fn = getattr(_lib, name)
            fn.argtypes = args
            fn.restype = C.c_int
    return _lib


class SbrError(RuntimeError):
    pass


# kernels launched per C-ABI call (for the "gpu_launches" figure of bench.py)
KERNELS_PER_CALL = {"sbr_infonce": 2, "sbr_topk_workspace_bytes": 0, "sbr_gather_plan": 3}
_launches = [0]


def reset_launch_counter():
    _launches[0] = 0


def launch_counter() -> int:
    return _launches[0]


def call(name: str, *args):
    _launches[0] += KERNELS_PER_CALL.get(name, 1)
    rc = getattr(lib(), name)(*args)
    if rc != 0:
        raise SbrError(f"{name} failed ({rc}): {lib().sbr_last_error().decode()}")


def ptr(t):
    """raw device pointer of a torch tensor (None -> NULL)"""
    return None if t is None else t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream

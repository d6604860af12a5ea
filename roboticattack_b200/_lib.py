"""ctypes binding of ``libvla_b200.so`` (declared in ``include/vla_b200.h``).

There is no CPU fallback: if the library is missing or a call fails, the caller gets an exception.
"""
from __future__ import annotations

import ctypes
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_longlong, c_size_t, c_uint8,
                    c_void_p)
from pathlib import Path

_LIB = None
import os as _os

# VLA_LIB_PATH: load another build of the library (A/B measurements of compile-time variants)
LIB_PATH = Path(_os.environ.get("VLA_LIB_PATH") or (Path(__file__).resolve().parent / "libvla_b200.so"))

FE_WARP, FE_PASTE20, FE_FIX, FE_NONE = 0, 1, 2, 3
LOSS_UADA, LOSS_UADA_DDP, LOSS_UPA, LOSS_CE, LOSS_NEG_CE = 0, 1, 2, 3, 4
OPT_ADAMW, OPT_PGD = 0, 1
S_LOSS, S_CE, S_AUX0, S_AUX1, S_UAD, S_NTOK, S_NACT, S_GRAD_MEAN, NUM_SCALARS = 0, 1, 2, 3, 4, 5, 6, 7, 8
FLAG_FORWARD_ONLY = 1
STEP_NO_GRAPH, STEP_NO_UPDATE = 1, 2
COMM_ID_BYTES = 128
ABI_VERSION = 3


class VLAError(RuntimeError):
    pass


class LossParams(Structure):
    _fields_ = [("kind", c_int), ("mse_weight", c_float), ("alpha", c_float), ("belta", c_float), ("ce_scale", c_float)]


class StepParams(Structure):
    _fields_ = [("ph", c_int), ("pw", c_int), ("fe_mode", c_int), ("loss", LossParams), ("opt_kind", c_int),
                ("lr", c_float), ("beta1", c_float), ("beta2", c_float), ("eps", c_float), ("clip_l1", c_float), ("flags", c_int)]


class GemmEpilogue(Structure):
    """vla_gemm_epilogue (parity tests of the fused GEMM epilogues)."""
    _fields_ = [("bias", c_void_p), ("gamma", c_void_p), ("resid", c_void_p), ("ldr", c_int64), ("act", c_int),
                ("preact_out", c_void_p), ("out_f32", c_int), ("out_group", c_int), ("out_stride", c_int), ("out_offset", c_int),
                ("resid_mod", c_int), ("aux_mode", c_int), ("aux", c_void_p), ("ldaux", c_int64), ("pair_mode", c_int),
                ("rope_cos", c_void_p), ("rope_sin", c_void_p), ("rope_L", c_int), ("rope_cols", c_int), ("act_out", c_void_p),
                ("ld_act", c_int64), ("delta_out", c_void_p), ("delta_L", c_int), ("w_constant", c_int)]


class Config(Structure):
    _fields_ = [("img", c_int), ("patch", c_int),
                ("dino_dim", c_int), ("dino_depth", c_int), ("dino_heads", c_int), ("dino_mlp", c_int),
                ("dino_prefix", c_int), ("dino_layerscale", c_int),
                ("sig_dim", c_int), ("sig_depth", c_int), ("sig_heads", c_int), ("sig_mlp", c_int),
                ("sig_prefix", c_int), ("sig_layerscale", c_int),
                ("vit_ln_eps", c_float),
                ("llm_hidden", c_int), ("llm_layers", c_int), ("llm_heads", c_int), ("llm_ffn", c_int), ("vocab", c_int),
                ("rms_eps", c_float),
                ("norm_mean", (c_float * 3) * 2), ("norm_std", (c_float * 3) * 2)]


# name -> (restype, argtypes); every symbol include/vla_b200.h declares
SIGNATURES = {
    "vla_last_error": (c_char_p, []),
    "vla_abi_version": (c_int, []),
    "vla_launch_count": (c_longlong, []),
    "vla_profile_gemm_begin": (c_int, []),
    "vla_gemm_set_mode": (c_int, [c_int, c_int]),
    "vla_gemm_set_autotune": (c_int, [c_int]),
    "vla_profile_gemm_end": (c_int, [POINTER(ctypes.c_double), POINTER(ctypes.c_double), POINTER(c_int)]),
    "vla_patch_frontend_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, POINTER(c_float), c_void_p]),
    "vla_patch_sim_paste": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                    c_void_p]),
    "vla_patch_frontend_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                       c_int, c_int, POINTER(c_float), c_void_p]),
    "vla_loss_head": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, POINTER(LossParams), c_void_p, c_void_p,
                              c_void_p, c_void_p, c_void_p]),
    "vla_patch_update": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_float, c_float,
                                 c_float, c_int, c_float, c_float, c_void_p, c_void_p]),
    "vla_gemm_bf16_tn": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                 c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p]),
    "vla_gemm_bf16_tn_ex": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                                    POINTER(GemmEpilogue), c_void_p]),
    "vla_gemv_bf16": (c_int, [c_void_p, c_int64, c_void_p, c_float, c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int,
                              c_void_p, c_int64, c_int, c_int, c_void_p]),
    "vla_attention_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vla_layernorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float,
                                  c_void_p]),
    "vla_layernorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int,
                                  c_void_p]),
    "vla_rmsnorm_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_float, c_void_p]),
    "vla_rmsnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vla_attention_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "vla_attention_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                  c_int, c_int, c_int, c_void_p]),
    "vla_attention_bwd_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                       c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "vla_attention_set_impl": (c_int, [c_int]),
    "vla_rope_inplace": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_void_p]),
    "vla_swiglu_fwd": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vla_swiglu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    "vla_gelu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    "vla_engine_create": (c_int, [POINTER(Config), POINTER(c_void_p)]),
    "vla_engine_destroy": (None, [c_void_p]),
    "vla_engine_weight_bytes": (c_size_t, [c_void_p]),
    "vla_engine_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "vla_engine_set_buffers": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_size_t, c_int, c_int]),
    "vla_engine_load_weight": (c_int, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "vla_engine_weights_ready": (c_int, [c_void_p]),
    "vla_engine_set_rope": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vla_engine_set_batch": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "vla_engine_set_placements": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p]),
    "vla_engine_num_supervised": (c_int, [c_void_p]),
    "vla_engine_full_vocab_pred": (c_int, [c_void_p, c_void_p, c_void_p]),
    "vla_engine_set_single_stream": (c_int, [c_void_p, c_int]),
    "vla_fwd_bwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(LossParams), c_void_p, c_void_p,
                            c_void_p, c_int, c_void_p]),
    "vla_engine_debug_tap": (c_int64, [c_void_p, c_char_p, c_void_p, c_int64, c_void_p]),
    "vla_engine_decode_greedy": (c_int, [c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "vla_comm_unique_id": (c_int, [c_void_p]),
    "vla_comm_create": (c_int, [c_void_p, c_int, c_int, POINTER(c_void_p)]),
    "vla_comm_destroy": (None, [c_void_p]),
    "vla_comm_world": (c_int, [c_void_p]),
    "vla_allreduce_patch_grad": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "vla_engine_set_step_state": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "vla_engine_drop_graphs": (c_int, [c_void_p]),
    "vla_engine_get_step_state": (c_int, [c_void_p, POINTER(c_int), POINTER(c_int)]),
    "vla_attack_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, POINTER(StepParams), c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "vla_graph_replays": (c_longlong, []),
    "vla_graph_kernel_nodes": (c_int, [c_void_p]),
}


def lib():
    """Load (once) and return the ctypes handle; raises if the CUDA extension has not been built."""
    global _LIB
    if _LIB is None:
        if not LIB_PATH.exists():
            raise VLAError(
                f"{LIB_PATH} is missing: build it with `python -m roboticattack_b200.build` "
                "(there is no CPU fallback for the attack hot path)")
        handle = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name)   # AttributeError if the header and the library diverge
            fn.restype = res
            fn.argtypes = args
        _LIB = handle
    return _LIB


def check(rc: int, what: str = "call"):
    if rc != 0:
        msg = lib().vla_last_error().decode("utf-8", "replace")
        raise VLAError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def norm_array(mean, std):
    flat = [float(v) for s in mean for v in s] + [float(v) for s in std for v in s]
    return (c_float * 12)(*flat)

"""ctypes binding of libvilbert_b200.so (the C ABI declared in include/vilbert_b200.h).

The shared library is built in-tree by ``__graft_entry__.build()`` / ``csrc/Makefile``.  There is
deliberately NO fallback: if the library is missing, or no sm_100 GPU is visible when an engine is
created, the call raises -- the reference-facing API never silently runs on CPU or through PyTorch ops.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvilbert_b200.so")

VB200_ABI_VERSION = 2
F32, F16, BF16 = 0, 1, 2

OUT_VIL_PREDICTION = 1 << 0
OUT_VIL_PREDICTION_GQA = 1 << 1
OUT_VIL_LOGIT = 1 << 2
OUT_VIL_BINARY_PREDICTION = 1 << 3
OUT_VIL_TRI_PREDICTION = 1 << 4
OUT_VISION_PREDICTION = 1 << 5
OUT_VISION_LOGIT = 1 << 6
OUT_LINGUISIC_PREDICTION = 1 << 7
OUT_LINGUISIC_LOGIT = 1 << 8
OUT_TASK_HEADS = 0x15F
OUT_ALL = 0x1FF
OUT_ATTENTION = 1 << 9


class Tensor(C.Structure):
    _fields_ = [("name", C.c_char_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * 2), ("data", C.c_void_p)]


class Inputs(C.Structure):
    _fields_ = [("batch", C.c_int32), ("n_tokens", C.c_int32), ("n_regions", C.c_int32),
                ("question", C.c_void_p), ("features", C.c_void_p), ("spatials", C.c_void_p),
                ("segment_ids", C.c_void_p), ("input_mask", C.c_void_p), ("image_mask", C.c_void_p),
                ("co_attention_mask", C.c_void_p), ("task_tokens", C.c_void_p)]


OUTPUT_FIELDS = ["vil_prediction", "vil_prediction_gqa", "vil_logit", "vil_binary_prediction",
                 "vil_tri_prediction", "vision_prediction", "vision_logit", "linguisic_prediction",
                 "linguisic_logit", "sequence_output_t", "sequence_output_v", "pooled_output", "attention_probs"]


class RegionInputs(C.Structure):
    """vb200_region_inputs: custom_prediction()'s detector output (worker.py:422-455) instead of built tensors."""
    _fields_ = [("batch", C.c_int32), ("n_tokens", C.c_int32), ("n_boxes", C.c_int32),
                ("question", C.c_void_p), ("segment_ids", C.c_void_p), ("input_mask", C.c_void_p),
                ("task_tokens", C.c_void_p), ("box_features", C.c_void_p), ("boxes", C.c_void_p),
                ("image_wh", C.c_void_p), ("num_boxes", C.c_void_p), ("spatials_out", C.c_void_p)]


class Outputs(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in OUTPUT_FIELDS]


class Options(C.Structure):
    _fields_ = [("device", C.c_int32), ("num_labels", C.c_int32), ("use_cuda_graph", C.c_int32),
                ("use_pdl", C.c_int32), ("strict", C.c_int32), ("act_fp16", C.c_int32),
                ("fused_layernorm", C.c_int32), ("split_fp32", C.c_int32), ("ln_fold", C.c_int32), ("max_plans", C.c_int32)]


# every symbol include/vilbert_b200.h declares (tests/test_cabi.py checks the list against the header)
EXPORTS = ["vb200_abi_version", "vb200_create", "vb200_destroy", "vb200_last_error", "vb200_forward", "vb200_forward_slot",
           "vb200_forward_host", "vb200_forward_host_slot", "vb200_plan_info", "vb200_model_dim", "vb200_set_option", "vb200_timeline", "vb200_profile_ops",
           "vb200_attention_layout", "vb200_forward_regions", "vb200_encode_text", "vb200_encode_image", "vb200_forward_cached",
           "vb200_linear", "vb200_linear_split", "vb200_linear_ln", "vb200_linear_chain", "vb200_layernorm", "vb200_layernorm_split", "vb200_attention_f32",
           "vb200_self_attention", "vb200_co_attention"]

_lib = None


class VilbertB200Error(RuntimeError):
    pass


def load():
    """Load the shared library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VilbertB200Error(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C vilbert-multi-task_b200/csrc`).  There is no CPU/PyTorch fallback.")
    lib = C.CDLL(LIB_PATH)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise VilbertB200Error(f"{LIB_PATH} does not export {name}")
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    lib.vb200_abi_version.restype = C.c_int
    lib.vb200_create.argtypes = [C.c_char_p, i64, C.POINTER(Tensor), C.POINTER(Options), C.POINTER(vp)]
    lib.vb200_destroy.argtypes = [vp]
    lib.vb200_last_error.argtypes = [vp]
    lib.vb200_last_error.restype = C.c_char_p
    lib.vb200_forward.argtypes = [vp, C.POINTER(Inputs), C.POINTER(Outputs), u32, vp]
    lib.vb200_forward_slot.argtypes = [vp, C.POINTER(Inputs), C.POINTER(Outputs), u32, i32, vp]
    lib.vb200_forward_host.argtypes = [vp, C.POINTER(Inputs), C.POINTER(Outputs), u32, vp]
    lib.vb200_forward_host_slot.argtypes = [vp, C.POINTER(Inputs), C.POINTER(Outputs), u32, i32, i32, vp]
    lib.vb200_plan_info.argtypes = [vp, i32, i32, i32, u32, C.POINTER(i64), C.POINTER(C.c_double)]
    lib.vb200_model_dim.argtypes = [vp, C.c_char_p, C.POINTER(i64)]
    lib.vb200_set_option.argtypes = [vp, C.c_char_p, i64]
    lib.vb200_timeline.argtypes = [vp, i32, i32, i32, u32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64)]
    lib.vb200_profile_ops.argtypes = [vp, i32, i32, i32, u32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(C.c_double),
                                      C.POINTER(C.c_double), C.POINTER(i32)]
    lib.vb200_attention_layout.argtypes = [vp, i32, i32, i32, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i64), C.POINTER(i64)]
    lib.vb200_forward_regions.argtypes = [vp, C.POINTER(RegionInputs), C.POINTER(Outputs), u32, i32, vp]
    lib.vb200_encode_text.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.vb200_encode_image.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.vb200_forward_cached.argtypes = [vp, i32, i32, i32, vp, i32, vp, vp, vp, vp, i32, vp, vp, vp, C.POINTER(Outputs), u32, i32, vp]
    lib.vb200_linear_split.argtypes = [vp, i64, vp, i64, vp, i32, vp, i64, vp, i64, i64, i64, i64, vp]
    lib.vb200_linear_ln.argtypes = [vp, i64, vp, i64, vp, i32, vp, i32, vp, vp, i64, vp, i32, vp, vp, vp, i32, f32, i32, vp, i64, vp, i64,
                                    i64, i64, i64, i32, vp]
    lib.vb200_linear_chain.argtypes = [vp, i64, vp, i64, vp, i32, vp, i64, vp, i64, vp, vp, i64, vp, i64, i64, i64, i64, i64, i32, vp, vp]
    lib.vb200_layernorm_split.argtypes = [vp, i64, vp, i64, vp, vp, f32, vp, i64, vp, i64, i64, i64, vp]
    lib.vb200_attention_f32.argtypes = [vp, i64, vp, vp, i64, i32, vp, i32, i32, i32, i32, i32, vp, i64, i32, vp, vp]
    lib.vb200_linear.argtypes = [vp, i64, vp, i64, vp, vp, i64, vp, vp, f32, i32, vp, i64, vp, i64,
                                 i64, i64, i64, i32, i32, i32, i32, vp, vp]
    lib.vb200_layernorm.argtypes = [vp, i64, vp, i64, vp, vp, f32, vp, i64, vp, i64, i64, i64, i32, vp]
    lib.vb200_self_attention.argtypes = [vp, i64, i32, vp, vp, i64, i32, i32, i32, i32, i32, vp]
    lib.vb200_co_attention.argtypes = [vp, i64, vp, i64, i32, vp, vp, vp, i64, vp, i64, i32, i32, i32, i32, i32, i32, vp]
    for name in EXPORTS:
        if name not in ("vb200_last_error",):
            getattr(lib, name).restype = C.c_int
    if lib.vb200_abi_version() != VB200_ABI_VERSION:
        raise VilbertB200Error("libvilbert_b200.so ABI version mismatch: rebuild the extension")
    _lib = lib
    return lib


def check(rc: int, handle=None):
    if rc != 0:
        msg = load().vb200_last_error(handle)
        raise VilbertB200Error(f"vilbert_b200 error {rc}: {msg.decode() if msg else '?'}")

"""ctypes binding of libdeft_b200.so (the C ABI declared in include/deft_b200.h).

The product path has NO fallback: if the shared library is missing, or the device is not a
B200 (sm_100), importing/calling raises -- it never routes to PyTorch eager or to the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DEFT_B200_LIB: an alternative build of the same library (A/B experiments: tools/), never a different implementation
LIB_PATH = os.environ.get("DEFT_B200_LIB") or os.path.join(_HERE, "libdeft_b200.so")

c_fp = C.c_void_p   # device pointers travel as integers
c_int_p = C.POINTER(C.c_int)
c_ptr_p = C.POINTER(C.c_void_p)

_SIGS = {
    "deft_version": (C.c_int, []),
    "deft_last_error": (C.c_char_p, []),
    "deft_device_ok": (C.c_int, []),
    "deft_launch_count": (C.c_ulonglong, []),
    "deft_set_pdl": (None, [C.c_int]),
    "deft_simt_fallback_count": (C.c_ulonglong, []),
    "deft_dcn_patch_launch_count": (C.c_ulonglong, []),
    "deft_debug_set_dcn_patch": (None, [C.c_int]),
    "deft_debug_set_engine_log": (None, [C.c_int]),
    "deft_debug_set_trace": (None, [c_fp]),
    "deft_get_pdl": (C.c_int, []),
    "deft_tc_tile_n": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "deft_tc_blob_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "deft_tc_pack_weights": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp]),
    "deft_tc_scratch_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "deft_nchw_to_nhwc": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "deft_nhwc_to_nchw": (C.c_int, [c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "deft_conv2d_bn_act": (C.c_int, [c_ptr_p, c_int_p, c_int_p, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, C.c_int, c_fp, C.c_int, c_fp, c_fp,
                                     c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint, c_fp, C.c_size_t,
                                     c_fp]),
    "deft_stem_conv7x7": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_int, c_fp]),
    "deft_maxpool2x2": (C.c_int, [c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "deft_dcnv2_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "deft_dcnv2_forward": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_int, c_fp, C.c_int, c_fp,
                                     C.c_int, c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_uint, c_fp, C.c_size_t, c_fp]),
    "deft_upsample_add": (C.c_int, [c_fp, C.c_int, c_fp, c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp]),
    "deft_heads_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "deft_heads_forward": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_int, c_fp, c_fp, c_int_p,
                                     C.c_int, c_ptr_p, C.c_uint, c_fp, C.c_size_t, c_fp]),
    "deft_sigmoid_output": (C.c_int, [c_fp, C.c_size_t, c_fp, C.c_size_t, C.c_float, c_fp]),
    "deft_decode_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "deft_decode": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                              c_fp, c_fp, c_fp, C.c_size_t, c_fp]),
    "deft_gather_feat": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, C.c_int, c_fp, c_fp]),
    "deft_boxes_to_centres": (C.c_int, [c_fp, C.c_int, C.c_float, C.c_float, c_fp, c_fp]),
    "deft_afe_extract_workspace_bytes": (C.c_size_t, [C.c_int]),
    "deft_afe_extract": (C.c_int, [c_ptr_p, c_int_p, c_int_p, c_int_p, c_int_p, c_ptr_p, c_ptr_p, c_int_p, c_fp, C.c_int, c_fp,
                                   C.c_int, c_fp, C.c_size_t, c_fp]),
    "deft_afe_affinity_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "deft_afe_affinity": (C.c_int, [c_fp, c_int_p, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                    c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_float, c_ptr_p, c_int_p, c_fp, C.c_uint, c_fp, C.c_size_t, c_fp]),
    "deft_afe_pairs_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "deft_afe_project": (C.c_int, [c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_ptr_p, c_int_p, c_fp, c_fp, C.c_uint, c_fp]),
    "deft_afe_affinity_cached": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp,
                                           c_fp, c_fp, c_fp, C.c_float, c_ptr_p, c_int_p, c_fp, C.c_uint, c_fp, C.c_size_t, c_fp]),
    "deft_afe_forward_scores": (C.c_int, [c_fp, c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_float,
                                          c_ptr_p, c_int_p, C.c_float, c_fp, C.c_uint, c_fp, C.c_size_t, c_fp]),
    "deft_afe_loss_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "deft_afe_loss": (C.c_int, [c_fp, c_fp, c_fp, c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_size_t, c_fp]),
    "deft_track_nodes_per_track": (C.c_int, []),
    "deft_memory_push": (C.c_int, [c_fp, c_fp, C.c_int, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp]),
    "deft_track_similarity": (C.c_int, [c_fp, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_int, C.c_int, c_fp, C.c_int, C.c_int,
                                        C.c_int, C.c_double, C.c_double, c_fp, c_fp, c_fp]),
    "deft_track_append": (C.c_int, [c_fp, c_fp, C.c_int, c_fp, c_fp, c_fp, c_fp]),
    "deft_count_above": (C.c_int, [c_fp, C.c_int, C.c_float, c_fp, c_fp]),
    "deft_assign_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "deft_fuse_and_assign": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, c_fp, C.c_double, C.c_double, c_fp, c_fp, c_fp, C.c_size_t,
                                       c_fp]),
    "deft_stem_conv7x7_u8": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), c_fp, c_fp, c_fp, c_fp, C.c_int,
                                       c_fp, c_fp]),
    "deft_preprocess_warp": (C.c_int, [c_fp, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float), c_fp, C.c_int,
                                       C.c_int, c_fp]),
    "deft_dcnv2_backward_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "deft_dcnv2_backward": (C.c_int, [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp, C.c_int, C.c_int, c_fp, C.c_int, c_fp,
                                      C.c_int, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_size_t, c_fp]),
    "deft_lstm_step": (C.c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, C.c_int, C.c_int, C.c_int,
                                 c_fp, c_fp, c_fp, c_fp]),
}

FLAG_RELU = 1
FLAG_TENSORCORE = 2
FLAG_TF32X1 = 4

_lib = None


class DeftError(RuntimeError):
    pass


def lib():
    """The loaded library; raises loudly when the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DeftError(f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                            "(deft_b200 has no CPU / eager fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)     # AttributeError if the .so does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def exported_symbols():
    return list(_SIGS.keys())


def check(rc, what=""):
    if rc != 0:
        msg = lib().deft_last_error().decode()
        raise DeftError(f"{what} failed (rc={rc}): {msg}")


def require_device(t):
    """Every product entry point calls this on its input tensor."""
    if not t.is_cuda:
        raise DeftError("deft_b200 runs on a B200 only: got a CPU tensor (there is no CPU fallback)")
    if not lib().deft_device_ok():
        raise DeftError(lib().deft_last_error().decode())


def float_array(vals):
    return (C.c_float * len(vals))(*[float(v) for v in vals])


def int_array(vals):
    return (C.c_int * len(vals))(*vals)


def ptr_array(vals):
    return (C.c_void_p * len(vals))(*vals)


def stream_ptr(device=None):
    """The current torch stream of `device` (a torch.device / index / tensor; default: the current device).  Kernels
    must run on a stream of the device that owns their pointers -- callers pass the tensor they launch on."""
    import torch
    if device is not None and hasattr(device, "device"):
        device = device.device
    return torch.cuda.current_stream(device).cuda_stream


class on_device:
    """`with on_device(t):` -- make t's device current for the enclosed C-ABI calls (cudaFuncSetAttribute, launches and
    cuTensorMapEncode act on the CURRENT device) and give the stream of that device.  A no-op when it already is."""
    __slots__ = ("dev", "prev")

    def __init__(self, t):
        import torch
        d = t.device if hasattr(t, "device") else t
        self.dev = d.index if d.index is not None else torch.cuda.current_device()
        self.prev = None

    def __enter__(self):
        import torch
        cur = torch.cuda.current_device()
        if cur != self.dev:
            self.prev = cur
            torch.cuda.set_device(self.dev)
        return torch.cuda.current_stream(self.dev).cuda_stream

    def __exit__(self, *exc):
        if self.prev is not None:
            import torch
            torch.cuda.set_device(self.prev)
        return False

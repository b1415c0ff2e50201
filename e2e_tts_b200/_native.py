"""ctypes binding of include/e2e_tts_b200.h.  There is no fallback: if the CUDA library cannot be loaded the
import fails loudly (the product path must never silently run on anything else)."""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

from .build import LIB_PATH, build_native

E2E_MAX_UPSAMPLES = 8
E2E_MAX_KERNELS = 8
E2E_MAX_DILATIONS = 8
E2E_OPERAND_BF16 = 0
E2E_OPERAND_FP16 = 1


class VocConfig(ctypes.Structure):
    _fields_ = [
        ("in_channels", c_int32),
        ("upsample_initial_channel", c_int32),
        ("resblock", c_int32),
        ("num_upsamples", c_int32),
        ("upsample_rates", c_int32 * E2E_MAX_UPSAMPLES),
        ("upsample_kernel_sizes", c_int32 * E2E_MAX_UPSAMPLES),
        ("num_kernels", c_int32),
        ("resblock_kernel_sizes", c_int32 * E2E_MAX_KERNELS),
        ("num_dilations", c_int32 * E2E_MAX_KERNELS),
        ("resblock_dilation_sizes", (c_int32 * E2E_MAX_DILATIONS) * E2E_MAX_KERNELS),
        ("istft_n_fft", c_int32),
    ]


# every symbol include/e2e_tts_b200.h declares: (restype, argtypes)
SYMBOLS = {
    "e2e_voc_create": (c_int, [POINTER(VocConfig), POINTER(c_void_p)]),
    "e2e_voc_destroy": (None, [c_void_p]),
    "e2e_voc_load_layer": (c_int, [c_void_p, c_char_p, POINTER(c_float), c_int64, POINTER(c_float), c_int64]),
    "e2e_voc_missing_layers": (c_int, [c_void_p]),
    "e2e_voc_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32]),
    "e2e_voc_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                c_void_p, c_size_t, c_void_p]),
    "e2e_voc_forward_pcm16": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                      c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "e2e_voc_forward_spec": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                     c_void_p, c_void_p, c_size_t, c_void_p]),
    "e2e_istft_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_int32, c_int32, c_void_p,
                                  c_void_p]),
    "e2e_voc_set_profile_events": (c_int, [c_void_p, c_void_p, c_void_p]),
    "e2e_voc_set_operand_dtype": (c_int, [c_void_p, c_int32]),
    "e2e_voc_operand_dtype": (c_int, [c_void_p]),
    "e2e_voc_hop": (c_int, [c_void_p]),
    "e2e_voc_launches_per_forward": (c_int, [c_void_p]),
    "e2e_voc_last_forward_was_graph": (c_int, [c_void_p]),
    "e2e_resblock_create": (c_int, [c_int32, c_int32, c_int32, POINTER(c_int32), c_int32, POINTER(c_void_p)]),
    "e2e_resblock_destroy": (None, [c_void_p]),
    "e2e_resblock_load_layer": (c_int, [c_void_p, c_char_p, POINTER(c_float), c_int64, POINTER(c_float), c_int64]),
    "e2e_resblock_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32]),
    "e2e_resblock_forward": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int32, c_int32, c_void_p,
                                     c_void_p, c_size_t, c_void_p]),
    "e2e_postnet_create": (c_int, [c_int32, c_int32, c_int32, c_int32, POINTER(c_void_p)]),
    "e2e_postnet_destroy": (None, [c_void_p]),
    "e2e_postnet_load_layer": (c_int, [c_void_p, c_int32, POINTER(c_float), c_int64, POINTER(c_float), c_int64]),
    "e2e_postnet_workspace_bytes": (c_size_t, [c_void_p, c_int32, c_int32]),
    "e2e_postnet_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_size_t,
                                    c_void_p]),
    "e2e_mel_create": (c_int, [c_int32, c_int32, c_int32, c_int32, POINTER(c_float), POINTER(c_void_p)]),
    "e2e_mel_destroy": (None, [c_void_p]),
    "e2e_mel_num_frames": (c_int64, [c_void_p, c_int64]),
    "e2e_mel_forward": (c_int, [c_void_p, c_void_p, c_int32, c_int64, c_int64, c_void_p, c_void_p, c_void_p,
                                c_void_p]),
    "e2e_last_error_string": (c_char_p, []),
    "e2e_version_string": (c_char_p, []),
}

_lib = None


class NativeError(RuntimeError):
    pass


def lib() -> ctypes.CDLL:
    """Load (building first if the sources are newer and nvcc is present) and type the library."""
    global _lib
    if _lib is not None:
        return _lib
    # E2E_TTS_B200_LIB: load another in-tree build of the same sources (A/B experiments on the GPU box)
    path = os.environ.get("E2E_TTS_B200_LIB") or build_native()
    if not os.path.exists(path):
        raise ImportError("e2e_tts_b200: CUDA library %s is missing and could not be built" % LIB_PATH)
    handle = ctypes.CDLL(path)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(handle, name)  # AttributeError here = header / library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().e2e_last_error_string()
        raise NativeError("%s failed (code %d): %s" % (what, rc, (msg or b"").decode("utf-8", "replace")))

"""ctypes binding of libmtts.so (include/mtts.h).  No torch types cross this boundary.

The library is built in-tree by `matcha_tts_b200.build.build()` (nvcc, sm_100a).  There is no
Python / CPU implementation behind it: if the shared object is missing, import of the compute
path fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmtts.so")

MTTS_SOLVER_EULER = 0
MTTS_SOLVER_MIDPOINT = 1


class MttsConfig(C.Structure):
    _fields_ = [("in_channels", C.c_int), ("out_channels", C.c_int), ("channels", C.c_int),
                ("heads", C.c_int), ("head_dim", C.c_int), ("n_mid_blocks", C.c_int)]


class MttsTextConfig(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("n_vocab", "n_feats", "n_channels", "filter_channels", "n_heads", "n_layers", "kernel_size",
                                       "prenet", "filter_channels_dp", "kernel_size_dp", "n_spks", "spk_emb_dim")]


class MttsVocConfig(C.Structure):
    _fields_ = [("num_mels", C.c_int), ("upsample_initial_channel", C.c_int), ("n_ups", C.c_int), ("upsample_rates", C.c_int * 4),
                ("upsample_kernel_sizes", C.c_int * 4), ("n_resblocks", C.c_int), ("resblock_kernel_sizes", C.c_int * 3),
                ("n_dilations", C.c_int), ("resblock_dilation_sizes", (C.c_int * 3) * 3)]


class MttsError(RuntimeError):
    pass


_lib = None

# name -> (restype, argtypes); every symbol declared in include/mtts.h
SIGNATURES = {
    "mtts_create": (C.c_int, [C.POINTER(MttsConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "mtts_destroy": (None, [C.c_void_p]),
    "mtts_last_error": (C.c_char_p, []),
    "mtts_version": (C.c_char_p, []),
    "mtts_num_weights": (C.c_int, [C.c_void_p]),
    "mtts_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "mtts_weight_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "mtts_weight_arena_bytes": (C.c_size_t, [C.c_void_p]),
    "mtts_set_weight_arena": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mtts_load_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "mtts_weights_loaded": (C.c_int, [C.c_void_p]),
    "mtts_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "mtts_estimator_forward": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                                                          C.c_void_p]),
    "mtts_euler_solve": (C.c_int, [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p, C.c_size_t,
                                                                     C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "mtts_set_chains": (C.c_int, [C.c_void_p, C.c_int]),
    "mtts_set_lanes": (C.c_int, [C.c_void_p, C.c_int]),
    "mtts_debug_lane_grid": (C.c_int, [C.c_void_p, C.c_int, C.c_int]),
    "mtts_release_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mtts_last_launch_count": (C.c_int, [C.c_void_p]),
    "mtts_debug_profile_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtts_debug_profile_end": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int),
                                         C.POINTER(C.c_double)]),
    "mtts_debug_set_timeline": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "mtts_debug_set_tail_timeline": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtts_debug_set_launch_limit": (C.c_int, [C.c_void_p, C.c_int]),
    "mtts_debug_set_tile_timeline": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtts_debug_buffer_offset": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_char_p]),
    "mtts_text_create": (C.c_int, [C.POINTER(MttsTextConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "mtts_text_destroy": (None, [C.c_void_p]),
    "mtts_text_num_weights": (C.c_int, [C.c_void_p]),
    "mtts_text_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "mtts_text_weight_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "mtts_text_weight_arena_bytes": (C.c_size_t, [C.c_void_p]),
    "mtts_text_set_weight_arena": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mtts_text_load_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "mtts_text_weights_loaded": (C.c_int, [C.c_void_p]),
    "mtts_text_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "mtts_text_release_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mtts_text_encoder_forward": (C.c_int, [C.c_void_p] + [C.c_void_p] * 6 + [C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p]),
    "mtts_text_last_launch_count": (C.c_int, [C.c_void_p]),
    "mtts_text_debug_set_launch_limit": (C.c_int, [C.c_void_p, C.c_int]),
    "mtts_text_debug_buffer_offset": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "mtts_voc_create": (C.c_int, [C.POINTER(MttsVocConfig), C.c_int, C.POINTER(C.c_void_p)]),
    "mtts_voc_destroy": (None, [C.c_void_p]),
    "mtts_voc_num_weights": (C.c_int, [C.c_void_p]),
    "mtts_voc_weight_name": (C.c_char_p, [C.c_void_p, C.c_int]),
    "mtts_voc_weight_numel": (C.c_int64, [C.c_void_p, C.c_int]),
    "mtts_voc_weight_arena_bytes": (C.c_size_t, [C.c_void_p]),
    "mtts_voc_set_weight_arena": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "mtts_voc_load_weight": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_void_p]),
    "mtts_voc_weights_loaded": (C.c_int, [C.c_void_p]),
    "mtts_voc_hop_length": (C.c_int, [C.c_void_p]),
    "mtts_voc_workspace_bytes": (C.c_size_t, [C.c_void_p, C.c_int, C.c_int]),
    "mtts_voc_release_workspace": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "mtts_voc_generator_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p]),
    "mtts_voc_last_launch_count": (C.c_int, [C.c_void_p]),
    "mtts_voc_debug_set_launch_limit": (C.c_int, [C.c_void_p, C.c_int]),
    "mtts_voc_debug_buffer_offset": (C.c_int64, [C.c_void_p, C.c_int, C.c_int, C.c_char_p]),
    "mtts_voc_debug_profile_begin": (C.c_int, [C.c_void_p, C.c_void_p]),
    "mtts_voc_debug_profile_end": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "mtts_stft_frames": (C.c_int, [C.c_int]),
    "mtts_stft_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "mtts_stft_magnitude": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "mtts_denoiser_forward": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_float, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int,
                                        C.c_void_p]),
    "mtts_debug_gemm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p]),
}


def load():
    """dlopen libmtts.so and declare every prototype.  Raises if the library was not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MttsError(
            f"{LIB_PATH} not found: the CUDA extension is not built. Run `python -c 'import __graft_entry__ as g; "
            "g.build()'` (needs nvcc); there is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)            # AttributeError if the .so does not export the symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        msg = load().mtts_last_error()
        raise MttsError(f"libmtts error {code}: {msg.decode() if msg else '?'}")

"""ctypes binding of the C-ABI library (include/kvq_b200.h).  There is NO fallback: if libkvq_b200.so is missing
or a call fails, a RuntimeError carrying kvq_last_error_string() is raised."""
import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libkvq_b200.so")
MAX_STAGES = 4


class KvqSwinConfig(ctypes.Structure):
    _fields_ = [("embed_dim", c_int32), ("num_stages", c_int32), ("depths", c_int32 * MAX_STAGES),
                ("num_heads", c_int32 * MAX_STAGES), ("window", c_int32 * 3), ("frag_bias", c_int32 * MAX_STAGES),
                ("head_hidden", c_int32), ("ln_eps", c_float), ("split_weights", c_int32),
                ("resized_window", c_int32 * 3)]


class KvqResNetConfig(ctypes.Structure):
    _fields_ = [("layers", c_int32 * 4), ("feat3d_dim", c_int32), ("head", c_int32)]


class KvqSlowFastConfig(ctypes.Structure):
    _fields_ = [("depths", c_int32 * 4), ("alpha", c_int32), ("slow_pool", c_int32 * 3), ("fast_pool", c_int32 * 3)]


class KvqClipConfig(ctypes.Structure):
    _fields_ = [("width", c_int32), ("heads", c_int32), ("layers", c_int32), ("patch", c_int32),
                ("adapter_from", c_int32)]


STAGE_HOOK = ctypes.CFUNCTYPE(c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_void_p)

_I3 = c_int32 * 3
_F3 = c_float * 3

# name -> (restype, argtypes); mirrors include/kvq_b200.h one to one
PROTOTYPES = {
    "kvq_swin3d_num_weights": (c_int, [POINTER(KvqSwinConfig)]),
    "kvq_swin3d_workspace_bytes": (c_size_t, [POINTER(KvqSwinConfig), c_int, c_int, c_int, c_int]),
    "kvq_swin3d_forward": (c_int, [POINTER(KvqSwinConfig), POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int,
                                   c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_swin3d_forward_x16": (c_int, [POINTER(KvqSwinConfig), POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int,
                                       c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_swin3d_forward_hooked": (c_int, [POINTER(KvqSwinConfig), POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int,
                                          c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, STAGE_HOOK, c_void_p]),
    "kvq_pack_split_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "kvq_cast_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_attn_table_len": (c_int, [c_int, c_int, c_int]),
    "kvq_pack_bias_table": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "kvq_linear_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "kvq_linear_resid_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kvq_mlp_fused": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "kvq_ln_window": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_int, c_int, c_int, c_int, c_int,
                              POINTER(c_int32), POINTER(c_int32), c_void_p]),
    "kvq_window_rows": (c_int64, [c_int, c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32)]),
    "kvq_window_row_map": (c_int64, [c_int, c_int, c_int, POINTER(c_int32), POINTER(c_int32), c_int, c_void_p, c_void_p]),
    "kvq_window_attention_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int32),
                                                        POINTER(c_int32)]),
    "kvq_window_attention": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_int, POINTER(c_int32), POINTER(c_int32), c_void_p, c_size_t, c_int,
                                     c_void_p]),
    "kvq_vqa_head_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "kvq_vqa_head": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                             c_void_p, c_size_t, c_void_p]),
    "kvq_fragment_gather_u8": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_int, POINTER(c_float), POINTER(c_float), c_void_p]),
    "kvq_qrs_select_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_int, c_int, c_void_p]),
    "kvq_contrique_num_weights": (c_int, []),
    "kvq_contrique_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "kvq_contrique_forward": (c_int, [POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_clip_num_weights": (c_int, [POINTER(KvqClipConfig)]),
    "kvq_clip_workspace_bytes": (c_size_t, [POINTER(KvqClipConfig), c_int, c_int, c_int]),
    "kvq_clip_visual_forward": (c_int, [POINTER(KvqClipConfig), POINTER(c_void_p), c_int, c_void_p, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_mha_f16": (c_int, [c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                            ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, c_int, c_int, c_int, c_float, c_void_p]),
    "kvq_cdm_mix": (c_int, [c_void_p] * 10 + [c_int, c_int, c_int, c_void_p]),
    "kvq_small_linear_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kvq_blend_f16_f32": (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_resnet_num_weights": (c_int, [POINTER(KvqResNetConfig)]),
    "kvq_resnet_feature_dim": (c_int, [POINTER(KvqResNetConfig)]),
    "kvq_simplevqa_workspace_bytes": (c_size_t, [POINTER(KvqResNetConfig), c_int, c_int, c_int, c_int]),
    "kvq_simplevqa_forward": (c_int, [POINTER(KvqResNetConfig), POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int,
                                      c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_slowfast_num_weights": (c_int, [POINTER(KvqSlowFastConfig)]),
    "kvq_slowfast_workspace_bytes": (c_size_t, [POINTER(KvqSlowFastConfig), c_int, c_int, c_int, c_int, c_int]),
    "kvq_slowfast_forward": (c_int, [POINTER(KvqSlowFastConfig), POINTER(c_void_p), c_int, c_void_p, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_slow_frame_indices": (c_int, [c_int, c_int, POINTER(c_int32), c_int]),
    "kvq_pack_pathway_slow_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "kvq_conv_gemm_f16": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                  c_int, c_int, c_int, c_void_p]),
    "kvq_im2col_cl_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_int32),
                                  POINTER(c_int32), POINTER(c_int32), c_int, c_void_p]),
    "kvq_im2col_stem_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, POINTER(c_int32),
                                    POINTER(c_int32), POINTER(c_int32), c_int, c_void_p]),
    "kvq_conv_implicit_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                      c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), c_int, c_int,
                                      c_int, c_void_p]),
    "kvq_conv_image_kblocks": (c_int, [c_int, c_int]),
    "kvq_conv_narrow_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int,
                                    c_int, c_int, POINTER(c_int32), POINTER(c_int32), POINTER(c_int32), c_int, c_int,
                                    c_void_p]),
    "kvq_stem_weight_rows": (c_int, [c_int]),
    "kvq_stem_conv_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                  c_void_p]),
    "kvq_maxpool_hw_f16": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "kvq_pool_stats_f16": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "kvq_rowdot_mean_f32": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "kvq_resize_aa_taps": (c_int, [c_int, c_int]),
    "kvq_resize_aa_weights": (c_int, [c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "kvq_resize_view_workspace_bytes": (c_size_t, [c_int] * 10),
    "kvq_resize_view_u8": (c_int, [c_void_p] + [c_int] * 11 + [c_float, POINTER(c_float), POINTER(c_float), c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "kvq_resize_view_bilinear_u8": (c_int, [c_void_p] + [c_int] * 11 + [c_float, POINTER(c_float), POINTER(c_float), c_void_p,
                                            c_void_p, c_void_p]),
    "kvq_launch_count": (ctypes.c_longlong, []),
    "kvq_profile_enable": (None, [c_int]),
    "kvq_profile_num_categories": (c_int, []),
    "kvq_profile_category_name": (c_char_p, [c_int]),
    "kvq_profile_collect": (c_int, [POINTER(c_float), POINTER(c_int), c_int]),
    "kvq_debug_attn_timers": (c_int, [POINTER(ctypes.c_ulonglong), c_int]),
    "kvq_last_error_string": (c_char_p, []),
    "kvq_build_info": (c_char_p, []),
}

_lib = None


def load():
    """Load libkvq_b200.so (built by `__graft_entry__.build()` / `make -C csrc`)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"kvq_b200: native library not found at {LIB_PATH}; run `python -c 'import __graft_entry__ as g; "
                f"g.build()'` (there is no CPU or PyTorch fallback for this path)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)        # AttributeError here = header / library mismatch
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().kvq_last_error_string().decode()


def check(rc, what):
    if rc != 0:
        raise RuntimeError(f"kvq_b200.{what} failed (code {rc}): {last_error()}")


def i3(v):
    return _I3(*[int(t) for t in v])


def f3(v):
    return _F3(*[float(t) for t in v])

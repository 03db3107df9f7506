"""ctypes binding of libsga_b200.so (include/sga_b200.h).  There is NO CPU fallback: if the library is missing
or fails to load, every entry point raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsga_b200.so")

c_float_p = ctypes.c_void_p  # device pointers travel as integers


class SgaConfig(ctypes.Structure):
    _fields_ = [
        ("n_images", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("num_filters", ctypes.c_int32), ("n_ref", ctypes.c_int32), ("operand", ctypes.c_int32),
        ("lmbda", ctypes.c_float), ("clip_eps", ctypes.c_float),
        ("lr", ctypes.c_double), ("beta1", ctypes.c_double), ("beta2", ctypes.c_double), ("adam_eps", ctypes.c_double),
        ("seed", ctypes.c_uint64),
        ("first_image", ctypes.c_int32), ("variant", ctypes.c_int32),
    ]


class SgaWeights(ctypes.Structure):
    _fields_ = [
        ("ga_kernel", ctypes.c_void_p * 4), ("ga_bias", ctypes.c_void_p * 4), ("ga_beta", ctypes.c_void_p * 3),
        ("ga_gamma", ctypes.c_void_p * 3),
        ("gs_kernel", ctypes.c_void_p * 4), ("gs_bias", ctypes.c_void_p * 4), ("gs_beta", ctypes.c_void_p * 3),
        ("gs_gamma", ctypes.c_void_p * 3),
        ("ha_kernel", ctypes.c_void_p * 3), ("ha_bias", ctypes.c_void_p * 2),
        ("hs_kernel", ctypes.c_void_p * 3), ("hs_bias", ctypes.c_void_p * 3),
        ("eb_matrix", ctypes.c_void_p * 4), ("eb_bias", ctypes.c_void_p * 4), ("eb_factor", ctypes.c_void_p * 3),
    ]


OPERAND_F32 = 0
OPERAND_BF16 = 1
VARIANT_SGA = 0          # sga.py
VARIANT_BITS_BACK = 1    # bb_sga.py
BB_STATS_PER_IMAGE = 6
STATS_HEADER = 3
STATS_PER_IMAGE = 5
PROF_LABELS = 24
PROF_NAMES = {0: "elementwise", 1: "igdn_fwd", 2: "igdn_bwd", 3: "l3_fwd", 4: "l3_dgrad",
              8: "hs0_fwd", 9: "hs1_fwd", 10: "hs2_fwd", 11: "hs0_dgrad", 12: "hs1_dgrad", 13: "hs2_dgrad",
              16: "gs0_fwd", 17: "gs1_fwd", 18: "gs2_fwd", 19: "gs0_dgrad", 20: "gs1_dgrad", 21: "gs2_dgrad"}

# every symbol include/sga_b200.h declares
EXPORTS = [
    "sga_last_error", "sga_version", "sga_create", "sga_destroy", "sga_workspace_bytes", "sga_latent_shape",
    "sga_bind_weights", "sga_encode_init", "sga_grad", "sga_step", "sga_loop", "sga_eval_hard", "sga_reconstruction",
    "sga_launch_count", "sga_profile_enable", "sga_profile_read", "sga_op_soft_round",
    "sga_op_eb_likelihood", "sga_op_gauss_likelihood", "sga_op_adam", "sga_op_conv", "sga_op_gdn", "sga_op_igdn_bwd",
    "sga_bb_hyper_init", "sga_bb_grad", "sga_bb_step", "sga_bb_eval",
    "sga_eval_tilde", "sga_codec_create", "sga_codec_destroy", "sga_codec_create_host", "sga_codec_z_centers", "sga_codec_bin",
    "sga_rc_encode", "sga_rc_decode",
]

_lib = None


class SgaError(RuntimeError):
    pass


def load():
    """Load the shared library (once). Raises if it has not been built: the product has no CPU path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise SgaError("libsga_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                       "or `python %s/build.py`. There is no CPU fallback." % (LIB_PATH, HERE))
    lib = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, f32, u64, f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_float, ctypes.c_uint64,
                                   ctypes.c_double)
    lib.sga_last_error.restype = ctypes.c_char_p
    lib.sga_last_error.argtypes = []
    lib.sga_version.restype = ctypes.c_int
    lib.sga_create.argtypes = [ctypes.POINTER(SgaConfig), ctypes.POINTER(vp)]
    lib.sga_destroy.argtypes = [vp]
    lib.sga_workspace_bytes.restype = ctypes.c_size_t
    lib.sga_workspace_bytes.argtypes = [vp]
    lib.sga_latent_shape.argtypes = [vp] + [ctypes.POINTER(i32)] * 4
    lib.sga_bind_weights.argtypes = [vp, ctypes.POINTER(SgaWeights), vp]
    lib.sga_encode_init.argtypes = [vp, vp, vp, vp, vp]
    lib.sga_grad.argtypes = [vp, vp, vp, vp, f32, i32, vp, vp, vp, vp, vp, vp]
    lib.sga_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, vp, vp, vp, vp]
    lib.sga_loop.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ctypes.POINTER(f32), i32, i32, vp, vp]
    lib.sga_eval_hard.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sga_reconstruction.argtypes = [vp, vp, vp]
    lib.sga_launch_count.restype = ctypes.c_longlong
    lib.sga_launch_count.argtypes = []
    lib.sga_profile_enable.argtypes = [vp, i32]
    lib.sga_profile_read.argtypes = [vp, ctypes.POINTER(f32), ctypes.POINTER(i32), i32]
    lib.sga_op_soft_round.argtypes = [vp, vp, i64, f32, f32, u64, i32, i32, vp, vp, vp]
    lib.sga_op_eb_likelihood.argtypes = [vp, i64, i32, vp * 4, vp * 4, vp * 3, vp, vp, vp]
    lib.sga_op_gauss_likelihood.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    lib.sga_op_adam.argtypes = [vp, vp, vp, vp, i64, f64, f64, f64, f64, i32, vp]
    lib.sga_op_conv.argtypes = [i32, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, i32, i32, vp, vp]
    lib.sga_op_gdn.argtypes = [i32, i32, vp, vp, vp, i64, i32, vp, vp, vp]
    lib.sga_op_igdn_bwd.argtypes = [i32, vp, vp, vp, vp, i64, i32, vp, vp]
    lib.sga_bb_hyper_init.argtypes = [vp, vp, vp, vp, vp]
    lib.sga_bb_grad.argtypes = [vp, vp, vp, vp, vp, f32, i32, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.sga_bb_step.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, f32, i32, i32, vp, vp, i32, f64, vp, vp]
    lib.sga_bb_eval.argtypes = [vp, vp, vp, vp, vp, i32, vp, vp, vp]
    lib.sga_eval_tilde.argtypes = [vp, vp, vp, vp, vp, vp]
    lib.sga_codec_create.argtypes = [vp, ctypes.POINTER(vp), vp]
    lib.sga_codec_destroy.argtypes = [vp]
    lib.sga_codec_create_host.argtypes = [i32, vp, vp, vp, i32, vp, vp, vp, ctypes.POINTER(vp)]
    lib.sga_codec_z_centers.argtypes = [vp, vp, i32]
    lib.sga_codec_bin.argtypes = [vp, vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, vp]
    lib.sga_rc_encode.argtypes = [vp, i32, i64, vp, vp, vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]
    lib.sga_rc_decode.argtypes = [vp, i32, i64, vp, vp, ctypes.c_uint64, vp]
    for name in EXPORTS:
        if name not in ("sga_last_error", "sga_workspace_bytes", "sga_launch_count"):
            getattr(lib, name).restype = ctypes.c_int
    _lib = lib
    return lib


def check(status):
    if status != 0:
        raise SgaError("libsga_b200: status %d: %s" % (status, load().sga_last_error().decode("utf-8", "replace")))

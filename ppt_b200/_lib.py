"""ctypes loader for libppt_b200.so.  Fails loudly: no fallbacks."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PPT_B200_LIB") or os.path.join(_HERE, "libppt_b200.so")  # override: A/B builds
ABI_VERSION = 2

_c = ctypes
_p, _i, _f, _i64 = _c.c_void_p, _c.c_int, _c.c_float, _c.c_int64



class EncoderBn(_c.Structure):
    """ppt_encoder_bn_t (include/ppt_b200.h): device pointers into the Encoder's BatchNorm / conv tensors."""
    _fields_ = [("conv1_weight", _p), ("conv1_bias", _p),
                ("bn1_weight", _p), ("bn1_bias", _p), ("bn1_running_mean", _p), ("bn1_running_var", _p),
                ("bn1_num_batches_tracked", _p),
                ("bn2_weight", _p), ("bn2_bias", _p), ("bn2_running_mean", _p), ("bn2_running_var", _p),
                ("bn2_num_batches_tracked", _p),
                ("momentum", _f), ("eps", _f)]


# name -> (restype, argtypes); must list every function declared in include/ppt_b200.h
SIGNATURES = {
    "ppt_abi_version": (_i, []),
    "ppt_strerror": (_c.c_char_p, [_i]),
    "ppt_spatial_index_bytes": (_i64, [_i, _i]),
    "ppt_spatial_index_build": (_i, [_p, _p, _i, _i, _p]),
    "ppt_fps": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _p]),
    "ppt_square_distance": (_i, [_p, _p, _p, _i, _i, _i, _p]),
    "ppt_knn": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ppt_knn_group": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ppt_ball_query": (_i, [_p, _p, _p, _f, _i, _i, _i, _i, _p]),
    "ppt_gather": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "ppt_group_concat": (_i, [_p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _p]),
    "ppt_three_nn": (_i, [_p, _p, _p, _p, _i, _i, _i, _p]),
    "ppt_three_interpolate": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ppt_three_interpolate_grad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _p]),
    "ppt_graph_feature": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ppt_graph_feature_grad": (_i, [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p]),
    "ppt_edge_gn_workspace_bytes": (_i64, [_i, _i]),
    "ppt_edge_gn_max_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _f, _f, _p]),
    "ppt_sa_mlp_packed_bytes": (_i64, [_i, _i, _i, _i]),
    "ppt_sa_mlp_workspace_bytes": (_i64, [_i64, _i, _i, _i, _i]),
    "ppt_sa_mlp_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ppt_fp_mlp_packed_bytes": (_i64, [_i, _i, _i]),
    "ppt_fp_mlp_workspace_bytes": (_i64, [_i64, _i, _i, _i]),
    "ppt_fp_mlp_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    "ppt_encoder_packed_bytes": (_i64, [_i]),
    "ppt_encoder_workspace_bytes": (_i64, [_i64, _i]),
    "ppt_encoder_forward": (_i, [_p, _p, _p, _p, _p, _i64, _i, _p]),
    "ppt_encoder_forward_phases": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _p]),
    "ppt_encoder_forward_ex": (_i, [_p, _p, _p, _p, _p, _i64, _i, _i, _i, _p, _p]),
    "ppt_encoder_train_workspace_bytes": (_i64, [_i64, _i]),
    "ppt_encoder_forward_train": (_i, [_p, _p, _c.POINTER(EncoderBn), _p, _p, _p, _i64, _i, _p]),
    "ppt_posembed_packed_bytes": (_i64, [_i]),
    "ppt_tokenizer_workspace_bytes": (_i64, [_i64, _i]),
    "ppt_tokenizer_forward": (_i, [_p, _p, _p, _p, _p, _p, _p, _i64, _i, _i, _p]),
    "ppt_clock_probe": (_i, [_p, _i, _i64, _p]),
    "ppt_selftest_umma": (_i, [_p, _p, _p, _i, _i, _i, _p]),
}

_lib = None


class PptLibraryError(RuntimeError):
    pass


def load():
    """Returns the loaded library; raises PptLibraryError if it is missing or stale."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PptLibraryError(
            "%s not found: build it with `python -m ppt_b200.build` (needs nvcc; no CPU fallback exists)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise PptLibraryError("libppt_b200.so does not export %s; rebuild" % name) from e
        fn.restype = res
        fn.argtypes = args
    if lib.ppt_abi_version() != ABI_VERSION:
        raise PptLibraryError("libppt_b200.so ABI %d != expected %d; rebuild" % (lib.ppt_abi_version(), ABI_VERSION))
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().ppt_strerror(code).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, code))

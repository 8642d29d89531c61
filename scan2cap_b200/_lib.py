"""ctypes binding of libs2c.so -- the only way the Python host code reaches the CUDA kernels.

There is NO fallback: if the library is missing, unloadable or lacks a symbol, importing this
module (and therefore using any op of the package) raises immediately.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2c.so")

c_int, c_float, c_void_p, c_ll = ctypes.c_int, ctypes.c_float, ctypes.c_void_p, ctypes.c_longlong
P = c_void_p

# name -> argtypes, exactly include/s2c.h
SIGNATURES = {
    "s2c_furthest_point_sampling": [P, c_int, c_int, c_int, P, P, P],
    "s2c_gather_points": [P, P, c_int, c_int, c_int, c_int, P, P],
    "s2c_gather_points_grad": [P, P, c_int, c_int, c_int, c_int, P, P],
    "s2c_ball_query": [P, P, c_int, c_int, c_int, c_float, c_int, P, P, P],
    "s2c_group_points": [P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "s2c_group_points_grad": [P, P, c_int, c_int, c_int, c_int, c_int, P, P],
    "s2c_three_nn": [P, P, c_int, c_int, c_int, P, P, P],
    "s2c_three_interpolate": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "s2c_three_interpolate_grad": [P, P, P, c_int, c_int, c_int, c_int, P, P],
    "s2c_query_and_group": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_ll, c_float, c_int, c_int, c_int,
                            P, P, P],
    "s2c_mlp_layer_fwd": [P, c_ll, c_ll, c_int, P, P, P, c_int, P, c_ll, P, P, P],
    "s2c_mlp_layer_fwd_v2": [P, c_ll, c_ll, c_int, P, P, P, c_int, P, c_ll, P, P, P, P],
    "s2c_mlp_probe": [P, c_int],
    "s2c_mlp_layer_bwd_data": [P, c_ll, P, c_ll, c_ll, c_int, P, P, P, P, P, c_int, P, P, P, c_int, P, c_ll, P, P, P, c_ll,
                               P, P, P, P, P],
    "s2c_mlp_layer_bwd_weight": [P, c_ll, P, c_ll, P, P, P, P, c_ll, P, P, c_ll, c_int, c_int, P, c_ll, P],
    "s2c_pool_fwd": [P, c_ll, c_ll, c_int, c_int, P, P, P, P, P],
    "s2c_pool_bwd_stats": [P, P, P, c_ll, c_ll, c_int, c_int, P, P, P, P, P],
    "s2c_query_and_group_grid": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_ll, c_float, c_int, c_int, c_int, P, P,
                                 P, c_ll, P],
    "s2c_ball_query_grid_build": [P, c_int, c_int, c_float, P, c_ll, P],
    "s2c_query_and_group_grid_prebuilt": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_ll, c_float, c_int, c_int, c_int, P,
                                          P, P, c_ll, P],
    "s2c_bn_finalize": [P, P, c_ll, c_int, P, P, ctypes.c_double, ctypes.c_double, c_int, c_int, P, P, P, P, P, P, P, P],
    "s2c_bn_backward_coeffs": [P, P, P, P, P, c_ll, c_int, c_int, P, P, P, P, P, P],
    "s2c_group_rows_grad": [P, c_ll, c_int, c_int, P, c_int, c_ll, c_int, c_float, P, P],
    "s2c_mlp_layer_bwd_input": [P, c_ll, P, c_ll, c_ll, c_int, P, P, P, P, c_ll, c_int, P, c_ll, P, P, P],
    "s2c_col_sum": [P, c_ll, c_ll, c_int, P, P],
    "s2c_gemm_tn": [P, c_ll, P, c_ll, c_int, c_int, c_int, P, c_ll, P, P],
    "s2c_gemm": [P, c_ll, c_ll, P, c_ll, c_ll, P, c_int, c_int, c_int, c_int, P, c_ll, P],
    "s2c_caption_decode_fwd": [P, P],
    "s2c_caption_decode_bwd": [P, P],
    "s2c_detection_loss": [c_int] * 8 + [P, P, P, c_ll] + [P] * 22,
    "s2c_points_in_boxes_count": [P, c_ll, c_int, c_int, P, c_int, P, P],
    "s2c_nms3d": [P, P, P, P, c_int, c_int, ctypes.c_double, c_int, c_int, P, P],
    "s2c_knn_adjacency": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, ctypes.c_double, P, P, P],
    "s2c_adam_step": [P, P, P, P, c_ll, P, P, c_int, P, c_float, P],
    "s2c_edgeconv_fwd": [P, c_ll, c_int, P, P, P, c_ll, P, P, P, P, c_int, P, P, P, P, P, P],
    "s2c_edgeconv_bwd": [P, P, c_ll, c_int, P, P, P, c_ll, P, P, P, c_int, P, P, P, P, P, P, P, P, P],
}


class CaptionParams(ctypes.Structure):
    """struct s2c_caption_params of include/s2c.h (field order matters)."""
    _PTRS = ("pre_word pre_tgt mapped obj valid "
             "w_tdh w_ih1 w_hh1 b_ih1 b_hh1 w_hidd w_att w_lang b_lang w_ih2 w_hh2 b_ih2 b_hh2 "
             "u h1 r1 z1 n1 hn1 q probs att lang r2 z2 n2 hn2 h2 scores "
             "wt_tdh wt_ih1 wt_hh1 wt_hidd wt_lang wt_ih2 wt_hh2 "
             "d_h2 d_probs "
             "dgi2 dgh2 dlang datt dq dgi1 dgh1 du d_mapped d_obj d_watt").split()
    _fields_ = ([(n, c_int) for n in ("B", "T", "K", "E", "H", "F")] + [("ld_tdh", c_ll)] +
                [(n, c_void_p) for n in _PTRS] + [("dbg_ts", c_void_p), ("grid_bar", c_void_p)])


class S2CError(RuntimeError):
    pass


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "scan2cap_b200: %s not found -- build it with `python -m scan2cap_b200.build` "
            "(there is no CPU or PyTorch fallback)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    lib.s2c_version.restype = c_int
    lib.s2c_last_error.restype = ctypes.c_char_p
    lib.s2c_ball_query_grid_workspace_bytes.restype = c_ll
    lib.s2c_ball_query_grid_workspace_bytes.argtypes = [c_int, c_int]
    lib.s2c_edgeconv_workspace_bytes.restype = c_ll
    lib.s2c_edgeconv_workspace_bytes.argtypes = [c_ll, c_int, c_int, c_int]
    lib.s2c_query_and_group_grid_tune.restype = c_int
    lib.s2c_query_and_group_grid_tune.argtypes = [c_int]
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = c_int
    return lib


LIB = _load()


def check(rc, name):
    if rc != 0:
        msg = LIB.s2c_last_error().decode("utf-8", "replace")
        raise S2CError("%s failed (code %d): %s" % (name, rc, msg))


# ---- instrumentation used by bench.py: how many of OUR kernels were launched, and optional per-entry timing ----
LAUNCH_COUNT = 0
TIMING = None  # None, or {entry_name: [(start_event, end_event), ...]} filled while set (torch.cuda.Event pairs)


def call(name, *args):
    global LAUNCH_COUNT
    LAUNCH_COUNT += 1
    if TIMING is not None and name in TIMING:
        import torch
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        check(getattr(LIB, name)(*args), name)
        b.record()
        TIMING[name].append((a, b))
        return
    check(getattr(LIB, name)(*args), name)

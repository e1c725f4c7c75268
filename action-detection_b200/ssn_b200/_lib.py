"""ctypes binding of libssn_b200.so (C ABI declared in include/ssnb.h).

The library is the product: there is no CPU or PyTorch fallback.  If the shared object is missing
the import fails loudly; build it with `python -c "import __graft_entry__ as g; g.build()"` or
action-detection_b200/csrc/build.sh.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libssn_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError("libssn_b200.so not found at %s — build the CUDA extension first "
                      "(action-detection_b200/csrc/build.sh); there is no fallback path." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

EXACT_FP32, FAST_FP16, EXACT_TC = 0, 1, 2


class Config(C.Structure):
    _fields_ = [("in_channels", C.c_int32), ("frames", C.c_int32), ("precision", C.c_int32),
                ("training", C.c_int32), ("grad_scale", C.c_float), ("bn1_train", C.c_int32), ("reserved", C.c_int32 * 2)]


class HeadsCfg(C.Structure):
    _fields_ = [("n", C.c_int32), ("props_per_video", C.c_int32), ("num_class", C.c_int32),
                ("feat_dim", C.c_int32), ("feat_mult", C.c_int32), ("fg_per_video", C.c_int32),
                ("comp_group", C.c_int32), ("global_videos", C.c_int32), ("keep_neg", C.c_int32),
                ("comp_denom", C.c_float), ("comp_w", C.c_float), ("reg_w", C.c_float),
                ("loss_scale", C.c_float)]


_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
_ip = C.POINTER(C.c_int)
_pp = C.POINTER(C.c_void_p)

# name -> (restype, argtypes); every symbol include/ssnb.h declares
SIGNATURES = {
    "ssnb_create": (_i, [C.POINTER(Config), C.POINTER(_vp)]),
    "ssnb_destroy": (_i, [_vp]),
    "ssnb_last_error": (C.c_char_p, [_vp]),
    "ssnb_version": (C.c_char_p, []),
    "ssnb_num_convs": (_i, []),
    "ssnb_conv_info": (_i, [_i, _i, C.c_char_p, _i, _ip, _ip, _ip, _ip, _ip]),
    "ssnb_workspace_bytes": (_sz, [_vp]),
    "ssnb_set_workspace": (_i, [_vp, _vp, _sz]),
    "ssnb_pack_weights": (_i, [_vp, _pp, _pp, _pp, _pp, _pp, _pp, _vp]),
    "ssnb_set_bn1": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _f, _f]),
    "ssnb_backbone_fwd": (_i, [_vp, _vp, _vp, _vp]),
    "ssnb_backbone_bwd": (_i, [_vp, _vp, _pp, _pp, _vp]),
    "ssnb_backbone_bwd_range": (_i, [_vp, _vp, _pp, _pp, _i, _i, _vp]),
    "ssnb_bind_grads": (_i, [_vp, _pp, _pp]),
    "ssnb_set_grad_accumulate": (_i, [_vp, _i]),
    "ssnb_num_ops": (_i, [_vp]),
    "ssnb_op_info": (_i, [_vp, _i, C.c_char_p, _i, C.c_char_p, _i, C.c_char_p, _i]),
    "ssnb_value_shape": (_i, [_vp, C.c_char_p, _ip, _ip, _ip]),
    "ssnb_value_write": (_i, [_vp, C.c_char_p, _i, _vp, _vp]),
    "ssnb_value_read": (_i, [_vp, C.c_char_p, _i, _vp, _vp]),
    "ssnb_run_op": (_i, [_vp, _i, _i, _vp]),
    "ssnb_launch_count": (C.c_int64, [_vp]),
    "ssnb_global_launch_count": (C.c_int64, []),
    "ssnb_stpp_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _ip, _ip, _ip, _ip, _i, _i, _vp, _vp, _vp]),
    "ssnb_stpp_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _ip, _ip, _ip, _ip, _i, _i, _vp, _vp]),
    "ssnb_gpool_stpp_fwd": (_i, [_vp, _vp, _vp, _i, _i, _ip, _ip, _ip, _ip, _i, _i, _vp, _vp, _vp, _vp]),
    "ssnb_stpp_reorg": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp]),
    "ssnb_stpp_reorg_workspace_bytes": (_sz, [_i, _i]),
    "ssnb_stpp_reorg_prefix": (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _i, _i, _ip, _ip, _vp, _vp, _vp, _vp, _vp]),
    "ssnb_linear_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "ssnb_test_fc_cropmean": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "ssnb_linear_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ssnb_ohem_hinge_fwd": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "ssnb_ohem_hinge_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ssnb_classwise_reg_fwd": (_i, [_vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ssnb_classwise_reg_bwd": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "ssnb_heads_loss_workspace_bytes": (_sz, [C.POINTER(HeadsCfg)]),
    "ssnb_heads_loss_fwd_bwd": (_i, [C.POINTER(HeadsCfg)] + [_vp] * 25),
    "ssnb_grad_overflow": (_i, [_vp, _i]),
    "ssnb_timing_begin": (_i, [_vp]),
    "ssnb_timing_report": (C.c_char_p, []),
    "ssnb_detect_workspace_bytes": (_sz, [_i, _i]),
    "ssnb_detect_postprocess": (_i, [_vp, _vp, _vp, _vp, _i, _i, C.c_double, _i, _vp, _vp, _vp, _vp]),
    "ssnb_sgd_step": (_i, [_vp, _vp, _vp, _sz, _f, _f, _f, _f, _vp]),
    "ssnb_sgd_step_groups": (_i, [_vp, _vp, _vp, _sz, _vp, _vp, _vp, _i, _f, _f, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)          # AttributeError here == symbol missing from the .so
    _fn.restype = _res
    _fn.argtypes = _args


def check(rc, handle=None, what=""):
    if rc != 0:
        msg = lib.ssnb_last_error(handle)
        raise RuntimeError("libssn_b200 %s failed (code %d): %s" % (what, rc, (msg or b"").decode()))


def int_array(seq):
    return (C.c_int * len(seq))(*[int(v) for v in seq])


def ptr_array(tensors):
    return (C.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])

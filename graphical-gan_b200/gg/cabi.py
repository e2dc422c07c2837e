"""ctypes binding of include/gg_b200.h (libgg_b200.so).

This is the only place the host side touches native code.  Torch tensors are used purely as
device-memory containers: every call passes ``tensor.data_ptr()`` and the raw ``cudaStream_t`` of
torch's current stream.  There is no CPU fallback: if the shared library is missing, import of this
module raises, and every non-zero status from the library raises ``GGError``.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GG_LIB selects an alternative build of the same library (e.g. the -DGG_TIMELINE instrumented one for tools/)
LIB_PATH = os.environ.get("GG_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libgg_b200.so")


class GGError(RuntimeError):
    pass


if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libgg_b200.so not found at %s — build it with graphical-gan_b200/build.sh "
        "(or __graft_entry__.build()); there is no CPU fallback" % LIB_PATH)

lib = C.CDLL(LIB_PATH)

c_f = C.c_float
c_i = C.c_int
c_ll = C.c_longlong
c_p = C.c_void_p
c_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol include/gg_b200.h declares
SIGNATURES = {
    "gg_last_error": (C.c_char_p, []),
    "gg_version": (c_i, []),
    "gg_launch_count": (c_ll, []),
    "gg_reset_launch_count": (None, []),
    "gg_set_conv_backend": (c_i, [c_i]),
    "gg_get_conv_backend": (c_i, []),
    "gg_last_backend": (c_i, []),
    "gg_set_pdl": (c_i, [c_i]),
    "gg_set_null_launch": (c_i, [c_i]),
    "gg_set_tc_max_ctas": (c_i, [c_i]),
    "gg_set_tc_stages": (c_i, [c_i]),
    "gg_last_tc_info": (c_i, [C.POINTER(c_i)]),
    "gg_debug_set_small_buffer": (c_i, [c_p]),
    "gg_debug_small_wgrad_info": (c_i, [c_i] * 9 + [C.POINTER(c_i)]),
    "gg_conv2d_fwd": (c_i, [c_p, c_p, c_p, c_p] + [c_i] * 11 + [c_i, c_f, c_p, c_sz, c_p]),
    "gg_conv2d_dgrad": (c_i, [c_p, c_p, c_p, c_p] + [c_i] * 11 + [c_i, c_f, c_p, c_sz, c_p]),
    "gg_conv2d_wgrad": (c_i, [c_p, c_p, c_p] + [c_i] * 11 + [c_p, c_sz, c_p]),
    "gg_conv2d_dgrad_actgrad": (c_i, [c_p, c_p, c_p, c_p, c_i, c_f] + [c_i] * 11 + [c_p, c_sz, c_p]),
    "gg_conv2d_tc_supported": (c_i, [c_i] * 10),
    "gg_conv2d_stats_tiles": (c_i, [c_i] * 12),
    "gg_conv2d_bnstats": (c_i, [c_i, c_p, c_p, c_p, c_p, c_p] + [c_i] * 11 + [c_i, c_f, c_p, c_sz, c_p]),
    "gg_conv2d_wgrad_workspace": (c_sz, [c_i] * 9),
    "gg_conv2d_workspace": (c_sz, [c_i] * 10),
    "gg_gemm": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_i, c_f, c_p, c_sz, c_p]),
    "gg_gemm_workspace": (c_sz, [c_i, c_i, c_i]),
    "gg_bn_slices": (c_i, [c_i, c_i]),
    "gg_bn_stats": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "gg_bn_apply": (c_i, [c_p, c_p, c_i, c_f, c_p, c_p, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p]),
    "gg_bn_bwd_reduce": (c_i, [c_p] * 8 + [c_i, c_i, c_i, c_f, c_p]),
    "gg_bn_bwd_apply": (c_i, [c_p] * 8 + [c_i, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p]),
    "gg_bn_fold_partials": (c_i, [c_p, c_i, c_p, c_i, c_p]),
    "gg_bn_fused_supported": (c_i, [c_i, c_i]),
    "gg_bn_fwd_fused": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_f, c_p]),
    "gg_bn_bwd_fused": (c_i, [c_p] * 9 + [c_i, c_i, c_i, c_f, c_p]),
    "gg_unary": (c_i, [c_i, c_p, c_p, c_ll, c_f, c_f, c_p]),
    "gg_binary": (c_i, [c_i, c_p, c_p, c_p, C.POINTER(c_i), C.POINTER(c_i), C.POINTER(c_i), c_f, c_p]),
    "gg_reduce": (c_i, [c_i, c_p, c_p, c_i, c_i, c_i, c_p]),
    "gg_reduce_ws": (c_i, [c_i, c_p, c_p, c_i, c_i, c_i, c_p, c_sz, c_p]),
    "gg_reduce_workspace": (c_sz, [c_i, c_i, c_i]),
    "gg_softmax_fwd": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "gg_softmax_bwd": (c_i, [c_p, c_p, c_p, c_i, c_i, c_p]),
    "gg_transpose_b2d": (c_i, [c_p, c_p, c_i, c_i, c_i, c_p]),
    "gg_transpose_b2d_ex": (c_i, [c_p, c_p, c_i, c_i, c_i, c_ll, c_ll, c_p, c_i, c_f, c_p]),
    "gg_transpose4": (c_i, [c_p, c_p, C.POINTER(c_i), C.POINTER(c_i), c_p]),
    "gg_gather_rows": (c_i, [c_p, c_p, c_p, c_p, c_i, c_i, c_i, c_p]),
    "gg_ew_run": (c_i, [c_p, c_p]),
    "gg_ew_program_bytes": (c_i, []),
    "gg_copy2d": (c_i, [c_p, c_ll, c_p, c_ll, c_ll, c_ll, c_i, c_p]),
    "gg_fill": (c_i, [c_p, c_ll, c_f, c_p]),
    "gg_one_hot": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "gg_argmax": (c_i, [c_p, c_p, c_i, c_i, c_p]),
    "gg_cast_i32_f32": (c_i, [c_p, c_p, c_ll, c_f, c_f, c_p]),
    "gg_cast_u8_f32": (c_i, [c_p, c_p, c_ll, c_f, c_f, c_p]),
    "gg_cast_f32_i32": (c_i, [c_p, c_p, c_ll, c_p]),
    "gg_widen_u8_i32": (c_i, [c_p, c_p, c_ll, c_p]),
    "gg_add_n": (c_i, [C.POINTER(c_p), c_i, c_p, c_ll, c_p]),
    "gg_bce_mean": (c_i, [c_p, c_i, c_f, c_f, c_p, c_i, c_p]),
    "gg_bce_mean_grad": (c_i, [c_p, c_i, c_f, c_f, c_p, c_p, c_i, c_p]),
    "gg_dist_mean": (c_i, [c_p, c_p, c_ll, c_i, c_f, c_p, c_i, c_p]),
    "gg_gp_slope_penalty": (c_i, [c_p, c_i, c_i, c_f, c_p, c_p, c_p]),
    "gg_adam_multi": (c_i, [c_p, c_p, c_i, c_p, c_f, c_f, c_f, c_f, c_f, c_p]),
    "gg_adam_tick": (c_i, [c_p, c_f, c_f, c_p]),
    "gg_adam_apply": (c_i, [c_p, c_p, c_i, c_p, c_f, c_f, c_f, c_f, c_f, c_p]),
    "gg_rmsprop_multi": (c_i, [c_p, c_p, c_i, c_f, c_f, c_f, c_f, c_p]),
    "gg_pack_grads": (c_i, [c_p, c_p, c_i, c_p, c_p, c_i, c_p]),
    "gg_rng_tick": (c_i, [c_p, c_p]),
    "gg_rng_normal": (c_i, [c_p, c_ll, c_f, c_f, C.c_uint64, C.c_uint32, c_p, c_p]),
    "gg_rng_uniform": (c_i, [c_p, c_ll, c_f, c_f, C.c_uint64, C.c_uint32, c_p, c_p]),
    "gg_rng_categorical": (c_i, [c_p, c_i, c_p, c_i, C.c_uint64, C.c_uint32, c_p, c_p]),
    "gg_debug_set_buffer": (c_i, [c_p]),
    "gg_comm_buffer_bytes": (c_sz, [c_i]),
    "gg_comm_alloc": (c_i, [c_i, C.POINTER(c_p), c_p]),
    "gg_comm_open": (c_i, [c_p, C.POINTER(c_p)]),
    "gg_comm_alloc_bytes": (c_i, [c_sz, C.POINTER(c_p), c_p]),
    "gg_bn_dp_site_bytes": (c_sz, [c_i, c_i]),
    "gg_bn_fused_grid": (c_i, [c_i, c_i]),
    "gg_bn_fwd_fused_dp": (c_i, [c_p, c_p, c_p, c_f, c_p, c_p, c_p, c_i, c_i, c_i, c_f, C.POINTER(c_p), c_i, c_i, c_ll, c_p]),
    "gg_bn_bwd_fused_dp": (c_i, [c_p] * 9 + [c_i, c_i, c_i, c_f, C.POINTER(c_p), c_i, c_i, c_ll, c_p]),
    "gg_allreduce_small": (c_i, [c_p, c_p, c_i, C.POINTER(c_p), c_i, c_i, c_i, c_p, c_p]),
    "gg_trace_event_create": (c_i, [C.POINTER(c_p)]),
    "gg_trace_event_record": (c_i, [c_p, c_p]),
    "gg_trace_event_elapsed_us": (c_i, [c_p, c_p, C.POINTER(c_f)]),
    "gg_probe_tma_strided": (c_i, [c_p] + [c_i] * 11 + [c_p, c_p]),
    "gg_probe_umma_tf32": (c_i, [c_p, c_p, c_p, c_i, c_i, c_i, c_i, c_i, c_p]),
}

GG_EW_MAX_IN, GG_EW_MAX_OUT, GG_EW_MAX_INSTR, GG_EW_REGS = 12, 4, 40, 32


class EwInstr(C.Structure):
    """gg_ew_instr of include/gg_b200.h"""
    _fields_ = [("kind", c_i), ("op", c_i), ("dst", c_i), ("src0", c_i), ("src1", c_i), ("a", c_f), ("b", c_f)]


class EwProgram(C.Structure):
    """gg_ew_program of include/gg_b200.h (passed by pointer; the library copies it into the launch)"""
    _fields_ = [("n_in", c_i), ("n_out", c_i), ("n_instr", c_i), ("flat", c_i), ("reduce_op", c_i), ("dims", c_i * 4),
                ("inp", c_p * GG_EW_MAX_IN), ("in_is_int", c_i * GG_EW_MAX_IN), ("in_stride", (c_i * 4) * GG_EW_MAX_IN),
                ("out", c_p * GG_EW_MAX_OUT), ("out_reg", c_i * GG_EW_MAX_OUT), ("instr", EwInstr * GG_EW_MAX_INSTR)]


for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = the .so is stale w.r.t. the header
    _fn.restype = _res
    _fn.argtypes = _args

if lib.gg_ew_program_bytes() != C.sizeof(EwProgram):
    raise ImportError("gg_ew_program layout mismatch: library %d bytes, ctypes %d" % (lib.gg_ew_program_bytes(), C.sizeof(EwProgram)))

GG_ADAM_CHUNK = 4096

ACT = {None: 0, "none": 0, "relu": 1, "leaky": 2, "tanh": 3, "sigmoid": 4}
UNARY = {"copy": 0, "relu": 1, "leaky": 2, "tanh": 3, "sigmoid": 4, "exp": 5, "log": 6, "sqrt": 7, "square": 8,
         "neg": 9, "abs": 10, "affine": 11, "pow": 12, "rsqrt": 13, "recip": 14, "bce": 15, "clip": 16, "sign": 17,
         "softsign": 18, "divc": 19, "rdivc": 20}
BINARY = {"add": 0, "sub": 1, "mul": 2, "div": 3, "max": 4, "min": 5, "relu_grad": 6, "leaky_grad": 7,
          "tanh_grad": 8, "sigmoid_grad": 9, "bce_grad": 10, "ge_mask": 11, "gt_mask": 12, "abs_grad": 13, "pow": 14}
REDUCE = {"sum": 0, "mean": 1, "max": 2}


def last_error():
    return lib.gg_last_error().decode("utf-8", "replace")


def check(rc, what=""):
    if rc != 0:
        raise GGError("%s failed with status %d: %s" % (what or "libgg_b200 call", rc, last_error()))


def call(name, *args):
    """Call an int-returning entry point and raise on a non-zero status."""
    check(getattr(lib, name)(*args), name)


def last_tc_info():
    """dict view of gg_last_tc_info"""
    out = (c_i * 8)()
    call("gg_last_tc_info", out)
    return dict(zip(("mode", "tiles", "splits", "n_tile", "stages", "cluster", "smem", "m_tiles"), list(out)))


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def int4(vals):
    return (c_i * 4)(*vals)

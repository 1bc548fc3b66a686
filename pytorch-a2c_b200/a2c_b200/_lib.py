"""ctypes binding of liba2c_b200.so (include/a2c_b200.h).

There is no CPU fallback: importing this module without the built library, or
calling a kernel without a CUDA device, raises.  The library is built in-tree
by ``pytorch-a2c_b200/build.py`` (``__graft_entry__.build()``).
"""
import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "liba2c_b200.so")


class A2CError(RuntimeError):
    pass


class StateView(C.Structure):
    _fields_ = [("base", C.c_void_p), ("dtype", C.c_int), ("fs", C.c_int), ("T", C.c_int),
                ("reserved", C.c_int), ("stride_r", C.c_longlong), ("stride_t", C.c_longlong),
                ("stride_c", C.c_longlong), ("pflag", C.c_void_p),
                ("pflag_stride_r", C.c_longlong)]


class LossCfg(C.Structure):
    _fields_ = [("pi_coef", C.c_float), ("val_coef", C.c_float), ("entr_coef", C.c_float),
                ("norm_advs", C.c_int), ("use_nstep", C.c_int), ("n_global", C.c_longlong)]


class ConvShape(C.Structure):
    _fields_ = [("Cin", C.c_int), ("Cout", C.c_int), ("H", C.c_int), ("W", C.c_int),
                ("OH", C.c_int), ("OW", C.c_int), ("K", C.c_int), ("stride", C.c_int),
                ("pad", C.c_int)]


class OptimCfg(C.Structure):
    _fields_ = [("kind", C.c_int), ("zero_grad", C.c_int), ("lr", C.c_float),
                ("max_norm", C.c_float), ("alpha", C.c_float), ("eps", C.c_float),
                ("beta1", C.c_float), ("beta2", C.c_float), ("step", C.c_longlong)]


_p, _i, _ll, _f = C.c_void_p, C.c_int, C.c_longlong, C.c_float
_SV, _CS = C.POINTER(StateView), C.POINTER(ConvShape)

# name -> argtypes (every function returns int unless listed in _RESTYPES)
SIGNATURES = {
    "a2c_abi_version": [],
    "a2c_last_error": [],
    "a2c_launch_count": [],
    "a2c_obs_prep_stack": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "a2c_obs_prep_stack_host": [_p, _p, _p, _p, _i, _i, _i, _i, _p],
    "a2c_obs_ingest_bytes_per_frame": [_i],
    "a2c_obs_ingest_rows_async": [_p, _p, _ll, _i, _p],
    "a2c_obs_prep_stack_rows": [_p, _p, _p, _p, _i, _i, _i, _i, _i, _p],
    "a2c_prep_other": [_p, _p, _i, _ll, _i, _i, _i, _p],
    "a2c_ring_carry": [_p, _p, _i, _i, _i, _p],
    "a2c_states_materialize": [_SV, _p, _i, _i, _p],
    "a2c_sample_action": [_p, _p, _p, _i, _i, _i, _p],
    "a2c_delta_scan": [_p, _p, _p, _p, _p, _p, _p, _i, _i, _f, _f, _p],
    "a2c_discount": [_p, _p, _p, _p, _ll, _f, _p],
    "a2c_discount2": [_p, _p, _p, _p, _p, _p, _ll, _f, _f, _p],
    "a2c_loss_fwd_bwd": [_p, _p, _p, _p, _p, _p, _p, C.POINTER(LossCfg), _p, _p, _p, _ll, _i, _p],
    "a2c_gemm": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _p, _i, _i, _i, _i, _i, _i, _p, _ll, _p],
    "a2c_gemm_masked": [_p, _ll, _ll, _p, _ll, _ll, _p, _ll, _p, _p, _i, _i, _i, _i, _i, _i, _p, _ll, _p],
    "a2c_linear_state_fwd": [_SV, _p, _p, _p, _i, _i, _i, _p, _ll, _p],
    "a2c_linear_state_wgrad": [_SV, _p, _p, _i, _i, _p, _ll, _p],
    "a2c_colsum": [_p, _p, _ll, _i, _p],
    "a2c_relu_bwd": [_p, _p, _p, _ll, _p],
    "a2c_conv2d_fwd": [_p, _SV, _p, _p, _p, _i, _CS, _i, _i, _p],
    "a2c_conv2d_bwd_data": [_p, _p, _p, _p, _i, _CS, _i, _p],
    "a2c_conv2d_bwd_data_relu": [_p, _p, _p, _p, _p, _p, _i, _CS, _i, _p, _ll, _p],
    "a2c_conv2d_fwd_relu_bits": [_p, _SV, _p, _p, _p, _p, _i, _CS, _i, _p],
    "a2c_conv2d_bwd_data_relu_bits": [_p, _p, _p, _p, _p, _p, _p, _i, _CS, _i, _p, _ll, _p],
    "a2c_channel_sum": [_p, _p, _p, _i, _i, _i, _p, _ll, _p],
    "a2c_conv2d_last_path": [],
    "a2c_conv2d_fused_mask": [_CS, _i, _i],
    "a2c_conv2d_bwd_weight_workspace": [_CS],
    "a2c_conv2d_bwd_weight": [_p, _SV, _p, _p, _p, _p, _i, _CS, _p, _ll, _i, _p],
    "a2c_im2col": [_SV, _p, _p, _i, _CS, _p],
    "a2c_col2im_nhwc": [_p, _p, _p, _i, _CS, _p],
    "a2c_layernorm_fwd": [_p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_layernorm_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_gru_gates1": [_p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_gru_gates2": [_p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_gru_gates2_bwd": [_p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_gru_gates1_bwd": [_p, _p, _p, _p, _p, _p, _p, _ll, _i, _p],
    "a2c_grad_sumsq": [_p, _p, _ll, _p],
    "a2c_clip_optim_step": [_p, _p, _p, _p, _p, C.POINTER(OptimCfg), _p, _ll, _p],
}
_RESTYPES = {"a2c_last_error": C.c_char_p, "a2c_launch_count": C.c_longlong,
             "a2c_obs_ingest_bytes_per_frame": C.c_longlong,
             "a2c_conv2d_bwd_weight_workspace": C.c_longlong}
_NO_STATUS = {"a2c_abi_version", "a2c_last_error", "a2c_launch_count", "a2c_obs_ingest_bytes_per_frame",
              "a2c_conv2d_bwd_weight_workspace", "a2c_conv2d_fused_mask", "a2c_conv2d_last_path"}

_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise A2CError(
            "liba2c_b200.so is not built (%s). Run `python pytorch-a2c_b200/build.py`; "
            "there is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)     # AttributeError if the symbol is missing
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, C.c_int)
    _lib = lib
    return lib


# bench.py sets PROFILE = {} to time calls in place: {tag: [(start_evt, end_evt, work), ...]}
# with CUDA events recorded on the launching (current) stream around the C-ABI call.
PROFILE = None


def call(name, *args, tag=None, work=0):
    """Calls an entry point; raises A2CError with the library's message on failure.
    tag / work (flops or bytes the launch performs) feed the optional profiler."""
    lib = load()
    if PROFILE is not None and tag is not None:
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = getattr(lib, name)(*args)
        e1.record()
        PROFILE.setdefault(tag, []).append((e0, e1, work))
    else:
        rc = getattr(lib, name)(*args)
    if name in _NO_STATUS:
        return rc
    if rc != 0:
        msg = lib.a2c_last_error()
        raise A2CError("%s failed (%d): %s" % (name, rc, msg.decode() if msg else "?"))
    return 0


def launch_count():
    return int(load().a2c_launch_count())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t):
    """Device (or pinned host) pointer of a tensor; None -> NULL."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


def require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise A2CError("a2c_b200 kernels need CUDA tensors (got %s); there is no CPU "
                           "fallback" % t.device)

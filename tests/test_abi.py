"""CPU-only: the C-ABI library loads and exports every symbol include/a2c_b200.h
declares; the product refuses to run without a GPU instead of falling back."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "a2c_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(a2c_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for must in ("a2c_obs_prep_stack", "a2c_delta_scan", "a2c_loss_fwd_bwd", "a2c_gemm",
                 "a2c_conv2d_fwd", "a2c_conv2d_bwd_data", "a2c_conv2d_bwd_weight",
                 "a2c_clip_optim_step", "a2c_sample_action"):
        assert must in syms


def test_library_exports_every_declared_symbol():
    from a2c_b200 import _lib
    assert os.path.isfile(_lib.LIB_PATH), "run python pytorch-a2c_b200/build.py"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), "missing export: " + name
    assert lib.a2c_abi_version() == 3
    # the ctypes table binds exactly the declared entry points
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    _lib.load()


def test_argument_errors_are_reported_not_thrown():
    from a2c_b200 import _lib
    lib = _lib.load()
    rc = lib.a2c_delta_scan(None, None, None, None, None, None, None, 4, 5, 0.99, 0.97, None)
    assert rc == -1
    assert b"null" in lib.a2c_last_error()


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from a2c_b200 import _lib, models, updater, utils
    with pytest.raises(_lib.A2CError):
        net = models.FCModel([2, 80, 80], 3, h_size=8)
        updater.Updater(net, {"is_discrete": True, "lr": 1e-4, "optim_type": "RMSprop"})
    with pytest.raises(_lib.A2CError):
        utils.discount(torch.ones(3), torch.zeros(3), 0.5)
    net = models.FCModel([2, 80, 80], 3, h_size=8)
    with pytest.raises(_lib.A2CError):
        net(torch.zeros(1, 2, 80, 80))


def test_fused_mask_query_is_host_logic():
    """a2c_conv2d_fused_mask tells the caller which backward entries apply the ReLU mask
    themselves (bit 0 data, bit 1 weight) and whether the weight gradient also yields the bias
    gradient for free (bit 2); pure host logic, so it runs without a GPU."""
    from a2c_b200 import _lib
    lib = _lib.load()

    def shape(cin, cout, h, k, s, p):
        oh = (h - k + 2 * p) // s + 1
        return _lib.ConvShape(cin, cout, h, h, oh, oh, k, s, p)

    def q(sh, prec, view=0):
        return lib.a2c_conv2d_fused_mask(ctypes.byref(sh), prec, view)

    assert q(shape(16, 24, 80, 3, 1, 1), 0) == 3            # FFMA kernels always take the mask
    assert q(shape(16, 24, 80, 3, 1, 1), 1) == 3            # ConvModel layers 2-4: shift-GEMM both ways
    assert q(shape(24, 32, 80, 3, 2, 1), 1) == 7            # 24 input channels: a free operand slot carries the bias row
    assert q(shape(32, 64, 40, 3, 2, 1), 1) == 3            # 32 input channels: no free slot
    assert q(shape(2, 16, 80, 3, 1, 1), 1, view=1) == 7     # first layer from the uint8 ring: packed rows have a free bias slot
    assert q(shape(48, 64, 10, 3, 2, 1), 1) == 7            # GRUModel's 10x10 layer: scalar producers, 32 + 16 channel chunks
    assert q(shape(16, 32, 19, 4, 2, 0), 1) == 0            # A3C's 4x4 s2 layer: pre-masked gradient
    assert q(shape(16, 24, 80, 3, 1, 1), 2) == 0            # single-pass TF32 paths want the masked dY


def test_relu_bits_entries_check_their_arguments():
    """The 1-bit ReLU mask entries reject a null bit plane and channel counts that are not multiples
    of 8 before touching the device (host-side checks: runs without a GPU)."""
    from a2c_b200 import _lib
    lib = _lib.load()
    sh = _lib.ConvShape(16, 12, 8, 8, 8, 8, 3, 1, 1)
    buf = (ctypes.c_float * 4)()
    p = ctypes.cast(buf, ctypes.c_void_p)
    rc = lib.a2c_conv2d_fwd_relu_bits(p, None, p, p, p, p, 1, ctypes.byref(sh), 1, None)
    assert rc == -1 and b"multiple of 8" in lib.a2c_last_error()
    sh = _lib.ConvShape(16, 16, 8, 8, 8, 8, 3, 1, 1)
    rc = lib.a2c_conv2d_fwd_relu_bits(p, None, p, p, p, None, 1, ctypes.byref(sh), 1, None)
    assert rc == -1
    rc = lib.a2c_conv2d_bwd_data_relu_bits(p, None, p, p, None, p, None, 1, ctypes.byref(sh), 1, None, 0, None)
    assert rc == -1 and b"Xbits" in lib.a2c_last_error()

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
    assert lib.a2c_abi_version() == 1
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

"""Policy networks with the reference's constructor signatures, attribute names
and state_dict keys (a2c/models.py), executed by the sm_100a kernels of
liba2c_b200.so.

The torch ``nn.Linear`` / ``nn.Conv2d`` / ``nn.LayerNorm`` objects below are
parameter containers only (same names, shapes and default initialisation as the
reference, so its checkpoints load with ``strict=True``); their ``forward`` is
never called.  ``forward`` accepts the reference's fp32 ``(B, fs, 80, 80)``
states, uint8 states, or an ``ops.View`` over the stage-1 plane ring.

Out of scope (NotImplementedError): ``bnorm=True`` and continuous actions
(SURVEY.md section 2 rows 7 and 9).
"""
import numpy as np
import torch
import torch.nn as nn

from . import ops
from .ops import (ConvStackFn, GRUCellFn, GRUSequenceFn, Im2colStackFn, LayerNormFn, LinearFn, LinearStateFn,
                  View, nhwc_linear_weight)


def _as_view(x):
    if isinstance(x, View):
        return x
    if not torch.is_tensor(x):
        raise TypeError("states must be a tensor or an ops.View")
    if x.dtype not in (torch.uint8, torch.float32):
        x = x.float()
    return View.of_states(x.contiguous())


def _check(bnorm, is_discrete):
    if bnorm:
        raise NotImplementedError("bnorm=True is out of scope (unused by every config; "
                                  "the reference's ConvModel bnorm path is dimensionally broken)")
    if not is_discrete:
        raise NotImplementedError("continuous actions are out of scope (Pong is discrete)")


def _linear(x, lin, relu=False, x_relu=False):
    return LinearFn.apply(x, lin.weight, lin.bias, relu, x_relu)


class _PolicyBase(nn.Module):
    def req_grads(self, calc_bool):
        """a2c/models.py:354-364 -- requires_grad switch for every parameter."""
        for p in self.parameters():
            p.requires_grad = calc_bool

    def activation_bytes_per_state(self):
        """Rough device bytes one state keeps alive through forward + backward (fp32 layer
        outputs, and about as much again for the gradients in flight); sizes the Updater's
        micro-batches.  ConvModel: 4 x 332,800 x 2.6 = 3.4 MB."""
        floats, size = 0, 80
        for blk in getattr(self, "convs", []):
            conv = blk[0]
            size = (size - conv.kernel_size[0] + 2 * conv.padding[0]) // conv.stride[0] + 1
            floats += conv.out_channels * size * size
        for m in self.modules():
            if isinstance(m, nn.Linear):
                floats += m.out_features
        floats += 8 * getattr(self, "h_size", 0) if getattr(self, "is_recurrent", False) else 0
        return int(4 * 2.6 * floats) + 4096

    @staticmethod
    def _conv_block(cin, cout, ksize, stride, padding):
        # Sequential(Conv2d, ReLU): index 0 carries the parameters ("convs.i.0.weight")
        return nn.Sequential(nn.Conv2d(cin, cout, ksize, stride=stride, padding=padding), nn.ReLU())

    def _features(self, view):
        """Conv + ReLU stack as one autograd node; the Linear layer that consumes the flattened
        features applies the last ReLU's mask in its input-gradient GEMM (``x_relu=True`` at the
        call sites below), hence ``premasked=True``."""
        geom = tuple((blk[0].stride[0], blk[0].padding[0]) for blk in self.convs)
        params = []
        for blk in self.convs:
            params += [blk[0].weight, blk[0].bias]
        x = ConvStackFn.apply(view, True, geom, *params)
        return x.reshape(x.shape[0], -1)


class FCModel(_PolicyBase):
    """a2c/models.py:367-405."""

    def __init__(self, input_shape, output_space, h_size=200, bnorm=False, is_discrete=True,
                 **kwargs):
        super().__init__()
        _check(bnorm, is_discrete)
        self.is_discrete = is_discrete
        self.is_recurrent = False
        self.flat_size = int(np.prod(input_shape[-3:]))
        self.input_shape = input_shape
        self.output_space = output_space
        self.h_size = h_size
        self.base = nn.Sequential(nn.Linear(self.flat_size, h_size), nn.ReLU(),
                                  nn.Linear(h_size, h_size))
        self.action_out = nn.Linear(h_size, output_space)
        self.value_out = nn.Sequential(nn.LayerNorm(h_size), nn.Linear(h_size, 1), nn.Linear(1, 1))

    def forward(self, x):
        v = _as_view(x)
        fx = LinearStateFn.apply(v, self.base[0].weight, self.base[0].bias, True)
        fx = _linear(fx, self.base[2])
        pi = _linear(fx, self.action_out)
        ln = self.value_out[0]
        val = LayerNormFn.apply(fx, ln.weight, ln.bias)
        val = _linear(val, self.value_out[1])
        val = _linear(val, self.value_out[2])
        return val, pi


class A3CModel(_PolicyBase):
    """a2c/models.py:7-90.  The value head reads the DETACHED embedding (:85);
    ``emb_bnorm`` exists but is unused with bnorm off (:60-62)."""

    def __init__(self, input_space, output_space, h_size=256, bnorm=False, is_discrete=True,
                 **kwargs):
        super().__init__()
        _check(bnorm, is_discrete)
        self.is_recurrent = False
        self.input_space = input_space
        self.output_space = output_space
        self.h_size = h_size
        self.is_discrete = is_discrete
        self.convs = nn.ModuleList([])
        self.conv1 = self._conv_block(input_space[-3], 16, 8, 4, 0)
        self.convs.append(self.conv1)
        self.conv2 = self._conv_block(16, 32, 4, 2, 0)
        self.convs.append(self.conv2)
        self.features = nn.Sequential(*self.convs)
        s = input_space[-1]
        s = (s - 8) // 4 + 1
        s = (s - 4) // 2 + 1
        self.flat_size = 32 * s * s
        self.proj_matrx = nn.Linear(self.flat_size, h_size)
        self.emb_bnorm = nn.BatchNorm1d(h_size)
        self.pi = nn.Linear(h_size, output_space)
        self.value = nn.Linear(h_size, 1)

    def encoder(self, state):
        view = _as_view(state)
        if ops.get_precision() >= 1:
            # 8x8 s4 / 4x4 s2 layers have no shift-GEMM form: explicit patch matrices + tcgen05 GEMMs,
            # activations channels-last; the projection weight is permuted instead of the activations
            geom = tuple((blk[0].stride[0], blk[0].padding[0]) for blk in self.convs)
            params = []
            for blk in self.convs:
                params += [blk[0].weight, blk[0].bias]
            feats = Im2colStackFn.apply(view, geom, *params)
            cout = self.convs[-1][0].out_channels
            w = nhwc_linear_weight(self.proj_matrx.weight, cout, self.flat_size // cout)
            return LinearFn.apply(feats, w, self.proj_matrx.bias, False, True)
        return _linear(self._features(view), self.proj_matrx, x_relu=True)

    def policy(self, state_emb, bnorm=False):
        if bnorm:
            _check(True, True)
        return _linear(state_emb.detach(), self.value), _linear(state_emb, self.pi)

    def forward(self, x, bnorm=False):
        return self.policy(self.encoder(x), bnorm=bnorm)


class ConvModel(_PolicyBase):
    """a2c/models.py:177-304."""

    def __init__(self, input_space, output_space, h_size=288, bnorm=False, is_discrete=True,
                 **kwargs):
        super().__init__()
        _check(bnorm, is_discrete)
        self.is_recurrent = False
        self.input_space = input_space
        self.output_space = output_space
        self.h_size = h_size
        self.bnorm = bnorm
        self.is_discrete = is_discrete
        self.convs = nn.ModuleList([])
        size = input_space[-1]
        for cin, cout, stride in ((input_space[-3], 16, 1), (16, 24, 1), (24, 32, 2), (32, 64, 2)):
            self.convs.append(self._conv_block(cin, cout, 3, stride, 1))
            size = (size - 3 + 2) // stride + 1
        self.features = nn.Sequential(*self.convs)
        self.flat_size = 64 * size * size
        conv_h = 2000
        self.resize_emb = nn.Sequential(nn.Linear(self.flat_size, conv_h), nn.ReLU())
        self.pi = nn.Sequential(nn.Linear(conv_h, h_size), nn.ReLU(),
                                nn.Linear(h_size, output_space))
        self.value = nn.Sequential(nn.Linear(conv_h, h_size), nn.ReLU(), nn.Linear(h_size, 1))

    def emb_net(self, state):
        return _linear(self._features(_as_view(state)), self.resize_emb[0], relu=True, x_relu=True)

    def policy(self, state_emb):
        pi = _linear(_linear(state_emb, self.pi[0], relu=True), self.pi[2])
        value = _linear(_linear(state_emb, self.value[0], relu=True), self.value[2])
        return value, pi

    def forward(self, x):
        return self.policy(self.emb_net(x))


class GRU(nn.Module):
    """a2c/models.py:426-476: parameter layout W_x (3, x, h), W_h (3, h, h),
    b (3, 1, h); right-multiplication."""

    def __init__(self, x_size=256, h_size=256, layer_norm=False, **kwargs):
        super().__init__()
        self.x_size = x_size
        self.h_size = h_size
        self.n_state_vars = 1
        std = 1 / float(np.sqrt(h_size))
        self.W_x = nn.Parameter(torch.normal(torch.zeros(3, x_size, h_size), std=std))
        self.W_h = nn.Parameter(torch.normal(torch.zeros(3, h_size, h_size), std=std))
        self.b = nn.Parameter(torch.zeros(3, 1, h_size))

    def forward(self, x, old_h, keep=None):
        hnew, hnext = GRUCellFn.apply(x, old_h, self.W_x, self.W_h, self.b, keep)
        return hnew if keep is None else (hnew, hnext)


class GRUModel(_PolicyBase):
    """a2c/models.py:546-677 (activation "lerelu" resolves to plain ReLU, :688-689)."""

    def __init__(self, input_space, output_space, h_size=288, bnorm=False, is_discrete=True,
                 **kwargs):
        super().__init__()
        _check(bnorm, is_discrete)
        self.is_recurrent = True
        self.input_space = input_space
        self.output_space = output_space
        self.h_size = h_size
        self.bnorm = bnorm
        self.is_discrete = is_discrete
        self.convs = nn.ModuleList([])
        size = input_space[-1]
        for cin, cout, stride in ((input_space[-3], 16, 1), (16, 24, 2), (24, 32, 2), (32, 48, 2),
                                  (48, 64, 2)):
            self.convs.append(self._conv_block(cin, cout, 3, stride, 1))
            size = (size - 3 + 2) // stride + 1
        self.features = nn.Sequential(*self.convs)
        self.flat_size = 64 * size * size
        self.resize_emb = nn.Sequential(nn.Linear(self.flat_size, h_size), nn.ReLU())
        self.gru = GRU(x_size=h_size, h_size=h_size)
        self.pi = nn.Linear(h_size, output_space)
        self.value = nn.Linear(h_size, 1)

    def emb_net(self, mdp_state):
        return _linear(self._features(_as_view(mdp_state)), self.resize_emb[0], relu=True, x_relu=True)

    def policy(self, h):
        return _linear(h, self.value), _linear(h, self.pi)

    def forward(self, x, old_h):
        h = self.gru(self.emb_net(x), old_h)
        val, pi = self.policy(h)
        return val, pi, h

    def bptt_forward(self, states, h0, dones, R, T):
        """Same result as T calls of forward with h *= (1-done) in between
        (a2c/updater.py:139-169), but the h-independent part (convs, embedding)
        runs once over all R*T states and only the GRU cell is stepped.
        Returns (vals (N,), logits (N, A)) rollout-major."""
        H = self.h_size
        emb = self.emb_net(states)                                  # (N, H), n = r*T + t
        emb_t = emb.view(R, T, H).transpose(0, 1).contiguous()      # time-major
        keep = (1.0 - dones.view(R, T)).t().contiguous()            # (T, R)
        hall = GRUSequenceFn.apply(emb_t, h0.contiguous(), self.gru.W_x, self.gru.W_h, self.gru.b,
                                   keep).view(T * R, H)
        val, pi = self.policy(hall)
        A = pi.shape[-1]
        vals = val.view(T, R).t().reshape(-1)
        logits = pi.view(T, R, A).transpose(0, 1).reshape(R * T, A)
        return vals, logits


MODEL_ALIASES = {"fc": "FCModel", "conv": "ConvModel", "a3c": "A3CModel", "gru": "GRUModel"}


def get_model_class(name):
    """``hyps['model']`` class name (a2c/training.py:78) or README alias (README.md:74)."""
    name = MODEL_ALIASES.get(name, name)
    cls = globals().get(name)
    if cls is None or not isinstance(cls, type) or not issubclass(cls, nn.Module):
        raise KeyError("unknown model %r" % name)
    return cls

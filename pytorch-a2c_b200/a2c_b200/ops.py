"""Host-side wrappers over the C ABI: buffer management (torch tensors are only
device memory + streams here) and the autograd glue that lets the models run
under ``loss.backward()`` like the reference's nn.Modules do.

Every function launches kernels of liba2c_b200.so on the current CUDA stream.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ConvShape, LossCfg, OptimCfg, StateView, call, ptr, require_cuda, stream_ptr

PLANE = 6400

# GEMM / conv contraction mode: 0 = fp32 FFMA, 1 = 3xTF32 on tcgen05 (fp32 parity, the
# default), 2 = single-pass TF32 on tcgen05 (stated tolerance 2e-3).  A2C_PRECISION overrides.
import os as _os
_PRECISION = 1


def set_precision(p):
    global _PRECISION
    p = int(p)
    if p not in (0, 1, 2):
        raise ValueError("precision must be 0 (fp32 FFMA), 1 (3xTF32 on tcgen05, fp32 parity) or "
                         "2 (single-pass TF32 on tcgen05); got %r" % (p,))
    _PRECISION = p


set_precision(_os.environ.get("A2C_PRECISION", "1"))     # validated like any other caller


def get_precision():
    return _PRECISION


_WS = {}


def workspace(device, nbytes=64 << 20):
    """Per-device scratch for split-K / conv weight-gradient partials."""
    key = (device.index if device.index is not None else torch.cuda.current_device())
    w = _WS.get(key)
    if w is None or w.numel() < nbytes:
        w = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WS[key] = w
    return w


# --------------------------------------------------------------------------
# stage 1
# --------------------------------------------------------------------------

class View:
    """Python handle of an a2c_state_view; keeps the backing tensors alive."""

    def __init__(self, base, fs, T, R, stride_r, stride_t, stride_c, pflag=None, pflag_stride_r=0):
        require_cuda(base, pflag)
        assert base.dtype in (torch.uint8, torch.float32)
        self.base, self.pflag = base, pflag
        self.fs, self.T, self.R = fs, T, R
        self.N = R * T
        self.c = StateView(base.data_ptr(), 0 if base.dtype == torch.uint8 else 1, fs, T, 0,
                           stride_r, stride_t, stride_c,
                           pflag.data_ptr() if pflag is not None else None, pflag_stride_r)

    @property
    def device(self):
        return self.base.device

    @staticmethod
    def of_states(states):
        """(N, fs, 80, 80) contiguous u8 / f32 tensor -> view with T = 1."""
        assert states.dim() == 4 and states.shape[2:] == (80, 80) and states.is_contiguous()
        N, fs = states.shape[:2]
        return View(states, fs, 1, N, fs * PLANE, fs * PLANE, PLANE)

    def slots(self, r0, r1):
        """View of rollout slots [r0, r1) (all their steps): a micro-batch of the update."""
        assert 0 <= r0 < r1 <= self.R
        base = self.base.view(-1)[r0 * self.c.stride_r:]
        pf = None
        if self.pflag is not None:
            pf = self.pflag.view(-1)[r0 * self.c.pflag_stride_r:]
        return View(base, self.fs, self.T, r1 - r0, self.c.stride_r, self.c.stride_t,
                    self.c.stride_c, pf, self.c.pflag_stride_r)

    def materialize(self, dtype=torch.float32):
        out = torch.empty((self.N, self.fs, 80, 80), dtype=dtype, device=self.device)
        call("a2c_states_materialize", C.byref(self.c), ptr(out),
             0 if dtype == torch.uint8 else 1, self.R, stream_ptr())
        return out


class PlaneRing:
    """HBM plane ring of stage 1: planes u8 [R, T+fs, 6400] + reset flags.

    State t of slot r is planes[r, t : t+fs]; ``push_frames`` preps raw frames
    into it (a2c/preprocessing.py:11-17 + a2c/utils.py:26-43), ``carry`` makes
    the last state the first of the next rollout (a2c/runner.py:246)."""

    def __init__(self, R, T, fs, device):
        self.R, self.T, self.fs = R, T, fs
        self.P = T + fs
        self.planes = torch.zeros((R, self.P, PLANE), dtype=torch.uint8, device=device)
        self.pflag = torch.zeros((R, self.P), dtype=torch.uint8, device=device)
        self.started = False

    def push_frames(self, frames, reset, first_frame_index=0, variant=0):
        """frames u8 [R, F, 210, 160, 3], reset u8 [R, F]; frame f lands at
        trajectory index first_frame_index + f.  Device or pinned-host inputs."""
        R, F = frames.shape[:2]
        assert R == self.R and frames.shape[2:] == (210, 160, 3) and frames.dtype == torch.uint8
        assert reset.shape == (R, F) and reset.dtype == torch.uint8
        assert frames.is_contiguous() and reset.is_contiguous()
        if F == 0:
            return
        j0 = self.fs - 1 + first_frame_index
        if frames.is_cuda:
            require_cuda(reset)
            call("a2c_obs_prep_stack", ptr(frames), ptr(reset), ptr(self.planes), ptr(self.pflag),
                 R, F, self.P, j0, variant, stream_ptr(),
                 tag="a2c_obs_prep_stack|bytes|%dx%d" % (R, F), work=44800 * R * F)
        else:
            if not (frames.is_pinned() and reset.is_pinned()):
                raise _lib.A2CError("host frames / reset must be pinned (tensor.pin_memory())")
            call("a2c_obs_prep_stack_host", ptr(frames), ptr(reset), ptr(self.planes),
                 ptr(self.pflag), R, F, self.P, j0, stream_ptr())
        self.started = True

    def push_rows(self, rows, reset, first_frame_index=0, layout=0):
        """Stage 1 from a device staging buffer of FrameIngest: rows u8 [R, F, bytes_per_frame]
        in staging ``layout`` (0: the 80 selected rows, 1: the cropped band, 2: whole frames),
        reset u8 [R, F] on the device."""
        require_cuda(rows, reset)
        R, F = rows.shape[:2]
        assert R == self.R and rows.dtype == torch.uint8 and rows.dim() == 3
        assert rows.shape[2] == call("a2c_obs_ingest_bytes_per_frame", layout)
        assert reset.shape == (R, F) and reset.dtype == torch.uint8
        assert rows.is_contiguous() and reset.is_contiguous()
        if F == 0:
            return
        call("a2c_obs_prep_stack_rows", ptr(rows), ptr(reset), ptr(self.planes), ptr(self.pflag),
             R, F, self.P, self.fs - 1 + first_frame_index, layout, stream_ptr(),
             tag="a2c_obs_prep_stack_rows|bytes|%dx%d" % (R, F), work=44800 * R * F)
        self.started = True

    def carry(self):
        call("a2c_ring_carry", ptr(self.planes), ptr(self.pflag), self.R, self.P, self.fs,
             stream_ptr())

    def view(self, T=None, t0=0):
        """States t0 .. t0+T-1 of every slot (default: the T rollout states)."""
        T = self.T if T is None else T
        base = self.planes.view(-1)[t0 * PLANE:]
        pf = self.pflag.view(-1)[t0:]
        return View(base, self.fs, T, self.R, self.P * PLANE, PLANE, PLANE, pf, self.P)

    def state_at(self, t):
        """View of the single state index t of every slot (e.g. t = T: bootstrap state)."""
        return self.view(T=1, t0=t)


class FrameIngest:
    """Pipelined host -> device ingest of whole rollouts for stage 1.

    ``prefetch(frames, reset)`` enqueues, on a private copy stream, ONE asynchronous copy from
    pinned host memory into a device staging buffer, so it runs under the previous update's
    kernels; ``commit(ring)`` makes the compute stream wait for that copy and runs stage 1 from
    the staging buffer.  ``layout`` picks what crosses PCIe (see a2c_obs_ingest_rows_async):
    0 = only the 80 selected rows of each frame (38 % of the bytes, 480-byte DMA rows: 49 GB/s
    measured, the default -- least time on the link, which bounds configs like C3), 1 = the
    cropped band of each frame (76 %, one 76 KB run per frame, 55 GB/s), 2 = whole frames.  ``depth``
    staging buffers are cycled; a buffer is overwritten only after the stage-1 launch that read
    it has finished (event on the compute stream)."""

    def __init__(self, R, F, device, depth=2, layout=0):
        self.R, self.F, self.device, self.layout = R, F, device, layout
        self.stream = torch.cuda.Stream(device=device)
        self.frame_bytes = int(call("a2c_obs_ingest_bytes_per_frame", layout))
        self.rows = [torch.empty((R, F, self.frame_bytes), dtype=torch.uint8, device=device)
                     for _ in range(depth)]
        self.reset = [torch.empty((R, F), dtype=torch.uint8, device=device) for _ in range(depth)]
        self.copied = [torch.cuda.Event() for _ in range(depth)]
        self.consumed = [None] * depth
        self.head = 0            # next buffer to fill
        self.tail = 0            # next buffer to commit
        self.pending = 0
        self.bytes_per_prefetch = R * F * (self.frame_bytes + 1)

    def prefetch(self, frames_host, reset_host):
        assert self.pending < len(self.rows), "FrameIngest: all staging buffers are in flight"
        if not (frames_host.is_pinned() and reset_host.is_pinned()):
            raise _lib.A2CError("host frames / reset must be pinned (tensor.pin_memory())")
        assert frames_host.shape == (self.R, self.F, 210, 160, 3) and frames_host.is_contiguous()
        assert reset_host.shape == (self.R, self.F) and reset_host.dtype == torch.uint8
        k = self.head
        if self.consumed[k] is not None:
            self.stream.wait_event(self.consumed[k])
        with torch.cuda.stream(self.stream):
            call("a2c_obs_ingest_rows_async", ptr(frames_host), ptr(self.rows[k]),
                 self.R * self.F, self.layout, C.c_void_p(self.stream.cuda_stream))
            self.reset[k].copy_(reset_host, non_blocking=True)
            self.copied[k].record(self.stream)
        self.head = (k + 1) % len(self.rows)
        self.pending += 1

    def commit(self, ring, first_frame_index=0):
        assert self.pending > 0, "FrameIngest.commit without a prefetch"
        k = self.tail
        cur = torch.cuda.current_stream()
        cur.wait_event(self.copied[k])
        ring.push_rows(self.rows[k], self.reset[k], first_frame_index, self.layout)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.consumed[k] = ev
        self.tail = (k + 1) % len(self.rows)
        self.pending -= 1


# --------------------------------------------------------------------------
# acting / stage 2 / stage 3
# --------------------------------------------------------------------------

def sample_action(x, u, is_prob=False):
    require_cuda(x, u)
    x = x.contiguous().float()
    u = u.contiguous().float()
    B, A = x.shape
    out = torch.empty(B, dtype=torch.int64, device=x.device)
    call("a2c_sample_action", ptr(x), ptr(u), ptr(out), B, A, 1 if is_prob else 0, stream_ptr())
    return out


def delta_scan(rewards, dones, values, gamma, lambda_, want_stats=True):
    """In-place post-rollout fix-up of rewards/dones + deltas, advs, returns.
    rewards, dones f32 [R*T] (or [R, T]); values f32 [R, T+1]."""
    require_cuda(rewards, dones, values)
    R, T1 = values.shape
    T = T1 - 1
    assert rewards.numel() == R * T and dones.numel() == R * T
    assert rewards.is_contiguous() and dones.is_contiguous() and values.is_contiguous()
    dev = rewards.device
    deltas = torch.empty(R * T, dtype=torch.float32, device=dev)
    advs = torch.empty_like(deltas)
    rets = torch.empty_like(deltas)
    stats = torch.zeros(4, dtype=torch.float64, device=dev) if want_stats else None
    call("a2c_delta_scan", ptr(rewards), ptr(dones), ptr(values), ptr(deltas), ptr(advs), ptr(rets),
         ptr(stats), R, T, float(gamma), float(gamma * lambda_), stream_ptr(),
         tag="a2c_delta_scan|bytes|%dx%d" % (R, T), work=28 * R * T)
    return deltas, advs, rets, stats


def discount(array, dones, factor, stats=None):
    require_cuda(array, dones)
    a = array.reshape(-1).contiguous().float()
    d = dones.reshape(-1).contiguous().float()
    out = torch.empty_like(a)
    call("a2c_discount", ptr(a), ptr(d), ptr(out), ptr(stats), a.numel(), float(factor),
         stream_ptr())
    return out


def discount2(deltas, rewards, dones, gamma_lambda, gamma, stats=None):
    """advs = discount(deltas, dones, gamma*lambda) and (rewards not None) returns =
    discount(rewards, dones, gamma) in one launch; stats f64 [4] accumulates their sums."""
    require_cuda(deltas, rewards, dones)
    d = deltas.reshape(-1).contiguous().float()
    dn = dones.reshape(-1).contiguous().float()
    advs = torch.empty_like(d)
    r = rets = None
    if rewards is not None:
        r = rewards.reshape(-1).contiguous().float()
        rets = torch.empty_like(d)
    call("a2c_discount2", ptr(d), ptr(r), ptr(dn), ptr(advs), ptr(rets), ptr(stats), d.numel(),
         float(gamma_lambda), float(gamma), stream_ptr(),
         tag="a2c_discount2|bytes|%d" % d.numel(), work=(12 + 8 * (rewards is not None)) * d.numel())
    return advs, rets


def loss_fwd_bwd(logits, vals, actions, advs, returns, adv_stats, hyps, n_global=None,
                 ret_norm=None):
    require_cuda(logits, vals, actions, advs)
    N, A = logits.shape
    logits = logits.contiguous()
    vals = vals.reshape(-1).contiguous()
    actions = actions.reshape(-1).contiguous()
    assert actions.dtype == torch.int64
    cfg = LossCfg(float(hyps["pi_coef"]), float(hyps["val_coef"]), float(hyps["entr_coef"]),
                  1 if hyps["norm_advs"] else 0, 1 if hyps["use_nstep_rets"] else 0,
                  int(n_global if n_global is not None else N))
    dlogits = torch.empty_like(logits)
    dvals = torch.empty_like(vals)
    out = torch.empty(4, dtype=torch.float32, device=logits.device)
    call("a2c_loss_fwd_bwd", ptr(logits), ptr(vals), ptr(actions), ptr(advs), ptr(returns),
         ptr(adv_stats), ptr(ret_norm), C.byref(cfg), ptr(dlogits), ptr(dvals), ptr(out), N, A,
         stream_ptr(), tag="a2c_loss_fwd_bwd|bytes|%dx%d" % (N, A), work=(24 + 8 * A) * N)
    return dlogits, dvals, out


# --------------------------------------------------------------------------
# stage 4 primitives
# --------------------------------------------------------------------------

def gemm(A, sam, sak, B, sbk, sbn, M, N, K, out=None, ldc=None, bias=None, relu=False, beta=0,
         precision=None, c_offset=0, cmask=None):
    """out[M, N] (+)= A * B with explicit element strides (see a2c_gemm); cmask (laid out like
    out): out = cmask > 0 ? value : 0 (a2c_gemm_masked)."""
    dev = A.device
    if out is None:
        out = torch.empty((M, N), dtype=torch.float32, device=dev)
        ldc = N
    ws = workspace(dev)
    cptr = C.c_void_p(out.data_ptr() + 4 * c_offset)
    call("a2c_gemm_masked", ptr(A), sam, sak, ptr(B), sbk, sbn, cptr, ldc, ptr(bias), ptr(cmask), M, N, K,
         1 if relu else 0, beta, _PRECISION if precision is None else precision, ptr(ws),
         ws.numel(), stream_ptr(), tag="a2c_gemm|flops|%dx%dx%d" % (M, N, K), work=2.0 * M * N * K)
    return out


def colsum(X, M, N):
    out = torch.empty(N, dtype=torch.float32, device=X.device)
    call("a2c_colsum", ptr(X), ptr(out), M, N, stream_ptr())
    return out


def relu_bwd(dY, Y):
    dX = torch.empty_like(dY)
    call("a2c_relu_bwd", ptr(dY), ptr(Y), ptr(dX), dY.numel(), stream_ptr())
    return dX


class LinearFn(torch.autograd.Function):
    """y = relu?(x W^T + b); x [M, K], W [N, K] (nn.Linear layout).  ``x_relu``: x is the output
    of a ReLU whose mask the input gradient applies itself (dx = dy W * (x > 0)), so the layer
    below receives an already-masked gradient (see ConvStackFn)."""

    @staticmethod
    def forward(ctx, x, W, b, relu, x_relu=False):
        require_cuda(x, W, b)
        x = x.contiguous()
        M, K = x.shape
        N = W.shape[0]
        y = gemm(x, K, 1, W, 1, K, M, N, K, bias=b, relu=relu)
        ctx.save_for_backward(x, W, y if relu else None)
        ctx.relu = relu
        ctx.x_relu = x_relu
        ctx.has_bias = b is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W, y = ctx.saved_tensors
        dy = dy.contiguous()
        if ctx.relu:
            dy = relu_bwd(dy, y)
        M, K = x.shape
        N = W.shape[0]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            dx = gemm(dy, N, 1, W, K, 1, M, K, N,            # dy [M,N] * W [N,K]
                      cmask=x if ctx.x_relu else None)
        if ctx.needs_input_grad[1]:
            dW = gemm(dy, 1, N, x, K, 1, N, K, M)            # dy^T [N,M] * x [M,K]
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy, M, N)
        return dx, dW, db, None, None


class LinearStateFn(torch.autograd.Function):
    """First-layer linear reading a state view (u8 ring or f32 states)."""

    @staticmethod
    def forward(ctx, view, W, b, relu):
        require_cuda(W, b)
        N = W.shape[0]
        assert W.shape[1] == view.fs * PLANE and W.is_contiguous()
        y = torch.empty((view.N, N), dtype=torch.float32, device=W.device)
        ws = workspace(W.device)
        call("a2c_linear_state_fwd", C.byref(view.c), ptr(W), ptr(b), ptr(y), view.R, N,
             1 if relu else 0, ptr(ws), ws.numel(), stream_ptr(), tag="a2c_linear_state_fwd|flops|%dx%d" % (view.N, N),
             work=2.0 * view.N * N * view.fs * PLANE)
        ctx.view = view
        ctx.save_for_backward(W, y if relu else None)
        ctx.relu = relu
        return y

    @staticmethod
    def backward(ctx, dy):
        W, y = ctx.saved_tensors
        view = ctx.view
        dy = dy.contiguous()
        if ctx.relu:
            dy = relu_bwd(dy, y)
        N = W.shape[0]
        dW = db = None
        if ctx.needs_input_grad[1]:
            dW = torch.empty_like(W)
            ws = workspace(W.device)
            call("a2c_linear_state_wgrad", C.byref(view.c), ptr(dy), ptr(dW), view.R, N, ptr(ws),
                 ws.numel(), stream_ptr(), tag="a2c_linear_state_wgrad|flops|%dx%d" % (view.N, N),
                 work=2.0 * view.N * N * view.fs * PLANE)
        if ctx.needs_input_grad[2]:
            db = colsum(dy, view.N, N)
        return None, dW, db, None


_SLOW_PATH_WARNED = set()


def _warn_if_ffma(op, sig, precision):
    """One warning per (op, layer shape) when a tensor-core request ran on the FFMA kernels."""
    if precision >= 1 and (op, sig) not in _SLOW_PATH_WARNED and call("a2c_conv2d_last_path") == 2:
        _SLOW_PATH_WARNED.add((op, sig))
        import warnings
        warnings.warn("a2c_b200: %s of conv layer %s has no tcgen05 path at precision %d and ran on the "
                      "fp32 FFMA kernels" % (op, sig, precision), RuntimeWarning, stacklevel=3)


def conv_shape(Cin, Cout, H, W, K, stride, pad):
    OH = (H - K + 2 * pad) // stride + 1
    OW = (W - K + 2 * pad) // stride + 1
    return ConvShape(Cin, Cout, H, W, OH, OW, K, stride, pad)


class Conv2dFn(torch.autograd.Function):
    """y = relu?(conv2d(x, W) + b), NCHW fp32.  x is a tensor or a View (first layer)."""

    @staticmethod
    def forward(ctx, x, W, b, stride, pad, relu):
        require_cuda(W, b)
        Cout, Cin, K, _ = W.shape
        is_view = isinstance(x, View)
        if is_view:
            N, H, Wd = x.N, 80, 80
            assert x.fs == Cin
        else:
            require_cuda(x)
            x = x.contiguous()
            N, _, H, Wd = x.shape
        s = conv_shape(Cin, Cout, H, Wd, K, stride, pad)
        y = torch.empty((N, Cout, s.OH, s.OW), dtype=torch.float32, device=W.device)
        flops = 2.0 * N * s.OH * s.OW * Cout * Cin * K * K
        sig = "%d:%d>%d,%dx%d,k%ds%d" % (N, Cin, Cout, H, Wd, K, stride)
        call("a2c_conv2d_fwd", None if is_view else ptr(x), C.byref(x.c) if is_view else None,
             ptr(W), ptr(b), ptr(y), N, C.byref(s), 1 if relu else 0,
             _PRECISION, stream_ptr(),
             tag="a2c_conv2d_fwd|flops|" + sig, work=flops)
        ctx.flops, ctx.sig, ctx.precision = flops, sig, _PRECISION
        ctx.shape, ctx.relu, ctx.is_view, ctx.N = s, relu, is_view, N
        if is_view:
            ctx.view = x
            ctx.save_for_backward(W, y)
        else:
            ctx.save_for_backward(W, y, x)
        return y

    @staticmethod
    def backward(ctx, dy):
        s, N = ctx.shape, ctx.N
        if ctx.is_view:
            W, y = ctx.saved_tensors
            x = None
        else:
            W, y, x = ctx.saved_tensors
        dy = dy.contiguous()
        mask = y if ctx.relu else None
        prec = ctx.precision
        need_dx = not ctx.is_view and ctx.needs_input_grad[0]
        need_dw = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        if mask is not None:
            # the shift-GEMM kernels apply the ReLU mask while they stage dY; the other
            # tensor-core paths want the masked gradient (one extra elementwise pass)
            fused = call("a2c_conv2d_fused_mask", C.byref(s), prec, 1 if ctx.is_view else 0)
            if (need_dx and not fused & 1) or (need_dw and not fused & 2):
                dy = relu_bwd(dy, y)
                mask = None
        dx = dW = db = None
        if not ctx.is_view and ctx.needs_input_grad[0]:
            dx = torch.empty_like(x)
            call("a2c_conv2d_bwd_data", ptr(dy), ptr(mask), ptr(W), ptr(dx), N, C.byref(s), prec,
                 stream_ptr(), tag="a2c_conv2d_bwd_data|flops|" + ctx.sig, work=ctx.flops)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            dW = torch.empty_like(W)
            db = torch.empty(s.Cout, dtype=torch.float32, device=W.device)
            need = call("a2c_conv2d_bwd_weight_workspace", C.byref(s))
            ws = workspace(W.device, max(need, 64 << 20))
            call("a2c_conv2d_bwd_weight", ptr(x), C.byref(ctx.view.c) if ctx.is_view else None,
                 ptr(dy), ptr(mask), ptr(dW), ptr(db), N, C.byref(s), ptr(ws), ws.numel(), prec,
                 stream_ptr(), tag="a2c_conv2d_bwd_weight|flops|" + ctx.sig, work=ctx.flops)
        return dx, dW, db, None, None, None


class ConvStackFn(torch.autograd.Function):
    """A stack of Conv2d + ReLU layers (``convs`` of ConvModel / GRUModel / A3CModel,
    a2c/models.py:98,312,685) as ONE autograd node, so that the backward pass can apply every
    ReLU mask where the masked gradient is born:

      * the node receives the gradient of its last ReLU output ALREADY masked (``premasked``:
        the Linear layer above applied ``x > 0`` in its input-gradient GEMM), or masks it once;
      * layer l's input gradient kernel writes dX * (X > 0) -- X being layer l-1's ReLU output --
        and the per-channel sums of that product, i.e. layer l-1's masked output gradient and
        bias gradient (a2c_conv2d_bwd_data_relu);
      * so no weight- or input-gradient kernel reads a forward output as a mask, and the
        separate bias-gradient pass over dY disappears for all layers but the last.

    apply(x, premasked, (stride, pad) per layer..., W0, b0, W1, b1, ...) -> features [N, C, h, w].
    x: tensor or View (first layer reads the uint8 ring)."""

    @staticmethod
    def forward(ctx, x, premasked, geom, *params):
        L = len(geom)
        assert len(params) == 2 * L
        is_view = isinstance(x, View)
        acts, shapes, bitmasks = [], [], []
        cur = x
        for l in range(L):
            W, b = params[2 * l], params[2 * l + 1]
            require_cuda(W, b)
            Cout, Cin, K, _ = W.shape
            stride, pad = geom[l]
            if l == 0 and is_view:
                N, Hh, Ww = cur.N, 80, 80
                assert cur.fs == Cin
            else:
                cur = cur.contiguous()
                N, _, Hh, Ww = cur.shape
            s = conv_shape(Cin, Cout, Hh, Ww, K, stride, pad)
            y = torch.empty((N, Cout, s.OH, s.OW), dtype=torch.float32, device=W.device)
            flops = 2.0 * N * s.OH * s.OW * Cout * Cin * K * K
            sig = "%d:%d>%d,%dx%d,k%ds%d" % (N, Cin, Cout, Hh, Ww, K, stride)
            first_view = l == 0 and is_view
            # A layer whose ReLU mask will be applied "at the source" in backward (same rule as there: its
            # weight-gradient kernel has no free operand slot for the bias row) also leaves a 1-bit copy of
            # the mask: the masked epilogue of the layer above then reads one byte per 8 channels and
            # position instead of eight strided floats of Y.
            bits = None
            if l < L - 1 and Cout % 8 == 0 and any(ctx.needs_input_grad):
                fm = call("a2c_conv2d_fused_mask", C.byref(s), _PRECISION, 1 if first_view else 0)
                if not (fm & 4):
                    bits = torch.empty((N, Cout // 8, s.OH * s.OW), dtype=torch.uint8, device=W.device)
            if bits is not None:
                call("a2c_conv2d_fwd_relu_bits", None if first_view else ptr(cur),
                     C.byref(cur.c) if first_view else None, ptr(W), ptr(b), ptr(y), ptr(bits), N, C.byref(s),
                     _PRECISION, stream_ptr(), tag="a2c_conv2d_fwd|flops|" + sig, work=flops)
            else:
                call("a2c_conv2d_fwd", None if first_view else ptr(cur), C.byref(cur.c) if first_view else None,
                     ptr(W), ptr(b), ptr(y), N, C.byref(s), 1, _PRECISION, stream_ptr(),
                     tag="a2c_conv2d_fwd|flops|" + sig, work=flops)
            bitmasks.append(bits)
            _warn_if_ffma("forward", sig, _PRECISION)
            shapes.append((s, flops, sig))
            acts.append(y)
            cur = y
        ctx.shapes, ctx.N, ctx.L = shapes, N, L
        ctx.premasked, ctx.precision, ctx.is_view = premasked, _PRECISION, is_view
        ctx.view = x if is_view else None
        ctx.bitmasks = bitmasks
        ctx.save_for_backward(*(([] if is_view else [x]) + acts + list(params[0::2])))
        return acts[-1]

    @staticmethod
    def backward(ctx, dy):
        L, N, prec = ctx.L, ctx.N, ctx.precision
        saved = list(ctx.saved_tensors)
        x0 = None if ctx.is_view else saved.pop(0)
        acts, Ws = saved[:L], saved[L:]
        dy = dy.contiguous()
        if not ctx.premasked:
            dy = relu_bwd(dy, acts[-1])
        dev = dy.device
        grads = [None] * (2 * L)
        ws0 = workspace(dev)
        # Where is each ReLU mask applied?  Boundary l | l+1 (layer l's output gradient):
        #   at the source  -- layer l+1's input-gradient epilogue writes the masked gradient and its
        #                     channel sums (= layer l's bias gradient); costs one strided read of Y_l there;
        #   in the consumers -- layer l's weight- and input-gradient producers read Y_l as a mask.
        # Measured on B200: the source form wins exactly when layer l would
        # otherwise need a separate bias-sum pass over dY (its weight-gradient kernel has no free
        # operand slot for the bias row), so that is the rule.
        free_bias = []
        for l in range(L):
            s = ctx.shapes[l][0]
            bits = call("a2c_conv2d_fused_mask", C.byref(s), prec, 1 if (l == 0 and ctx.is_view) else 0)
            free_bias.append(bool(bits & 4))
        masked = True                   # is `dy` (layer l's output gradient) already masked?
        have_bias = False               # has layer l's bias gradient been produced by the layer above?
        dx0 = None
        for l in range(L - 1, -1, -1):
            s, flops, sig = ctx.shapes[l]
            W = Ws[l]
            xin = acts[l - 1] if l > 0 else x0
            mask = None if masked else acts[l]
            db = None
            if not have_bias:
                db = torch.empty(s.Cout, dtype=torch.float32, device=dev)
                if masked and not free_bias[l]:
                    # top layer without a free slot: channel sums of its (masked) output gradient
                    call("a2c_channel_sum", ptr(dy), None, ptr(db), N, s.Cout, s.OH * s.OW, ptr(ws0),
                         ws0.numel(), stream_ptr())
                    grads[2 * l + 1], db = db, None
                else:
                    grads[2 * l + 1] = db          # written by the weight-gradient call below
            dW = torch.empty_like(W)
            need = call("a2c_conv2d_bwd_weight_workspace", C.byref(s))
            ws = workspace(dev, max(need, 64 << 20))
            first_view = l == 0 and ctx.is_view
            call("a2c_conv2d_bwd_weight", None if first_view else ptr(xin),
                 C.byref(ctx.view.c) if first_view else None, ptr(dy), ptr(mask), ptr(dW), ptr(db), N,
                 C.byref(s), ptr(ws), ws.numel(), prec, stream_ptr(),
                 tag="a2c_conv2d_bwd_weight|flops|" + sig, work=flops)
            _warn_if_ffma("weight gradient", sig, prec)
            grads[2 * l] = dW
            if l > 0:
                dx = torch.empty_like(xin)
                at_source = not free_bias[l - 1]
                if at_source:
                    dbp = torch.empty(s.Cin, dtype=torch.float32, device=dev)
                    xbits = ctx.bitmasks[l - 1]
                    if xbits is not None:
                        call("a2c_conv2d_bwd_data_relu_bits", ptr(dy), ptr(mask), ptr(W), ptr(xin), ptr(xbits),
                             ptr(dx), ptr(dbp), N, C.byref(s), prec, ptr(ws), ws.numel(), stream_ptr(),
                             tag="a2c_conv2d_bwd_data|flops|" + sig, work=flops)
                    else:
                        call("a2c_conv2d_bwd_data_relu", ptr(dy), ptr(mask), ptr(W), ptr(xin), ptr(dx), ptr(dbp),
                             N, C.byref(s), prec, ptr(ws), ws.numel(), stream_ptr(),
                             tag="a2c_conv2d_bwd_data|flops|" + sig, work=flops)
                    grads[2 * (l - 1) + 1] = dbp
                else:
                    call("a2c_conv2d_bwd_data", ptr(dy), ptr(mask), ptr(W), ptr(dx), N, C.byref(s), prec,
                         stream_ptr(), tag="a2c_conv2d_bwd_data|flops|" + sig, work=flops)
                masked, have_bias = at_source, at_source
                dy = dx
            elif not ctx.is_view and ctx.needs_input_grad[0]:
                dx0 = torch.empty_like(x0)
                call("a2c_conv2d_bwd_data", ptr(dy), ptr(mask), ptr(W), ptr(dx0), N, C.byref(s), prec,
                     stream_ptr(), tag="a2c_conv2d_bwd_data|flops|" + sig, work=flops)
        return (dx0, None, None) + tuple(grads)


class Im2colStackFn(torch.autograd.Function):
    """Conv2d + ReLU stack for layers without a shift-GEMM form (A3CModel's 8x8 s4 and 4x4 s2
    convolutions, a2c/models.py:22-38, 98) as ONE autograd node on explicit patch matrices:

        col_l = im2col(input_l)             [N*OH*OW, Cin*K*K]   written to HBM once, kept for backward
        Y_l   = relu(col_l W_l^T + b_l)     [N*OH*OW, Cout]      channels-last = the next layer's input
        dW_l  = (col_l^T dY_l)^T            split-K tcgen05 GEMM over the N*OH*OW rows
        dY_{l-1} = col2im(dY_l W_l) * (Y_{l-1} > 0)

    Every contraction is a2c_gemm (3xTF32 on tcgen05 at precision 1); the layers are HBM-bound on
    the patch-matrix traffic.  The first layer reads the state view (uint8 ring) with the
    reference's weight flattening; later layers use channels-fastest columns and a permuted copy
    of the (tiny) weight.  Returns the last activation FLATTENED CHANNELS-LAST, [N, OH*OW*Cout]:
    the caller permutes the weight of the Linear layer that follows (``nhwc_linear_weight``)
    instead of transposing the activations.  The incoming gradient must already be masked by the
    last ReLU (LinearFn(..., x_relu=True) does that in its input-gradient GEMM)."""

    @staticmethod
    def forward(ctx, view, geom, *params):
        L = len(geom)
        assert isinstance(view, View) and len(params) == 2 * L
        dev = params[0].device
        N = view.N
        cols, acts, shapes, wflat = [], [], [], []
        Hh = Ww = 80
        cur = None
        for l in range(L):
            W, b = params[2 * l], params[2 * l + 1]
            require_cuda(W, b)
            Cout, Cin, K, _ = W.shape
            stride, pad = geom[l]
            s = conv_shape(Cin, Cout, Hh, Ww, K, stride, pad)
            Kc = Cin * K * K
            M = N * s.OH * s.OW
            if l == 0:
                assert view.fs == Cin
                col = torch.empty((M, Kc), dtype=torch.float32, device=dev)
                call("a2c_im2col", C.byref(view.c), None, ptr(col), N, C.byref(s), stream_ptr(),
                     tag="a2c_im2col|bytes|%d:%d,k%ds%d" % (N, Cin, K, stride), work=4.0 * M * Kc)
                wf = W.reshape(Cout, Kc)                       # k = (ci, ky, kx): the reference's order
            else:
                col = torch.empty((M, Kc), dtype=torch.float32, device=dev)
                call("a2c_im2col", None, ptr(cur), ptr(col), N, C.byref(s), stream_ptr(),
                     tag="a2c_im2col|bytes|%d:%d,k%ds%d" % (N, Cin, K, stride), work=8.0 * M * Kc)
                wf = W.permute(0, 2, 3, 1).reshape(Cout, Kc).contiguous()     # k = (ky, kx, ci)
            y = gemm(col, Kc, 1, wf, 1, Kc, M, Cout, Kc, bias=b, relu=True)    # [M, Cout] channels-last
            cols.append(col); acts.append(y); shapes.append(s); wflat.append(wf)
            cur, Hh, Ww = y, s.OH, s.OW
        ctx.shapes, ctx.N, ctx.L = shapes, N, L
        ctx.save_for_backward(*(cols + acts + wflat))
        return acts[-1].view(N, -1)

    @staticmethod
    def backward(ctx, dfeat):
        L, N = ctx.L, ctx.N
        saved = ctx.saved_tensors
        cols, acts, wflat = saved[:L], saved[L:2 * L], saved[2 * L:]
        dev = dfeat.device
        grads = [None] * (2 * L)
        dy = dfeat.contiguous().view(acts[-1].shape)           # already masked by the last ReLU
        for l in range(L - 1, -1, -1):
            s = ctx.shapes[l]
            Kc = s.Cin * s.K * s.K
            M = N * s.OH * s.OW
            # dW^T [Kc, Cout] = col^T dY  (few output tiles, K = M rows: split-K)
            dWt = gemm(cols[l], 1, Kc, dy, s.Cout, 1, Kc, s.Cout, M)
            if l == 0:
                grads[0] = dWt.t().reshape(s.Cout, s.Cin, s.K, s.K)
            else:
                grads[2 * l] = dWt.t().reshape(s.Cout, s.K, s.K, s.Cin).permute(0, 3, 1, 2)
            grads[2 * l + 1] = colsum(dy, M, s.Cout)
            if l > 0:
                dcol = gemm(dy, s.Cout, 1, wflat[l], Kc, 1, M, Kc, s.Cout)     # dY W  [M, Kc], channels fastest
                dx = torch.empty_like(acts[l - 1])
                call("a2c_col2im_nhwc", ptr(dcol), ptr(acts[l - 1]), ptr(dx), N, C.byref(s), stream_ptr(),
                     tag="a2c_col2im|bytes|%d:%d,k%ds%d" % (N, s.Cin, s.K, s.stride), work=4.0 * M * Kc)
                dy = dx
        return (None, None) + tuple(grads)


def nhwc_linear_weight(W, C_, P_):
    """Weight of a Linear layer that follows a channels-last flatten: column (pos*C + c) of the
    result is column (c*P + pos) of the reference-layout weight [out, C*P] (differentiable view ops;
    the parameter itself keeps the reference layout, so checkpoints interchange)."""
    out = W.shape[0]
    return W.view(out, C_, P_).transpose(1, 2).reshape(out, P_ * C_)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta):
        require_cuda(x, gamma, beta)
        x = x.contiguous()
        rows, H = x.shape
        y = torch.empty_like(x)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty_like(mean)
        call("a2c_layernorm_fwd", ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(mean), ptr(rstd), rows,
             H, stream_ptr())
        ctx.save_for_backward(x, gamma, mean, rstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma, mean, rstd = ctx.saved_tensors
        dy = dy.contiguous()
        rows, H = x.shape
        dx = torch.empty_like(x)
        dg = torch.empty_like(gamma)
        db = torch.empty_like(gamma)
        call("a2c_layernorm_bwd", ptr(dy), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx), ptr(dg),
             ptr(db), rows, H, stream_ptr())
        return dx, dg, db


class GRUCellFn(torch.autograd.Function):
    """The reference's hand-rolled GRU (a2c/models.py:465-476):
         z = sig(x Wx0 + h Wh0 + b0);  r = sig(x Wx1 + h Wh1 + b1)
         h' = z*h + (1-z)*tanh(x Wx2 + (r*h) Wh2 + b2)
    Returns (h', h'*keep) -- the second output is the masked carry of
    a2c/updater.py:164 (keep = 1 - done, or None)."""

    @staticmethod
    def forward(ctx, x, h, Wx, Wh, b, keep):
        require_cuda(x, h, Wx, Wh, b, keep)
        x, h = x.contiguous(), h.contiguous()
        B, H = h.shape
        X = x.shape[1]
        dev = x.device
        gx = torch.empty((B, 3 * H), dtype=torch.float32, device=dev)
        gh = torch.empty((B, 2 * H), dtype=torch.float32, device=dev)
        bflat = b.reshape(3, H)
        for g in range(3):      # x.mm(W_x[g]) + b[g]   (W_x[g] is [X, H]: right-multiply)
            gemm(x, X, 1, Wx[g], H, 1, B, H, X, out=gx, ldc=3 * H, bias=bflat[g], c_offset=g * H)
        for g in range(2):
            gemm(h, H, 1, Wh[g], H, 1, B, H, H, out=gh, ldc=2 * H, c_offset=g * H)
        z = torch.empty((B, H), dtype=torch.float32, device=dev)
        r, rh = torch.empty_like(z), torch.empty_like(z)
        call("a2c_gru_gates1", ptr(gx), ptr(gh), ptr(h), ptr(z), ptr(r), ptr(rh), B, H, stream_ptr())
        gc = gemm(rh, H, 1, Wh[2], H, 1, B, H, H)
        cand, hnew = torch.empty_like(z), torch.empty_like(z)
        hnext = torch.empty_like(z)
        call("a2c_gru_gates2", ptr(gx), ptr(gc), ptr(h), ptr(z), ptr(keep), ptr(cand), ptr(hnew),
             ptr(hnext), B, H, stream_ptr())
        ctx.save_for_backward(x, h, Wx, Wh, z, r, rh, cand, keep)
        return hnew, hnext

    @staticmethod
    def backward(ctx, dhnew, dhnext):
        x, h, Wx, Wh, z, r, rh, cand, keep = ctx.saved_tensors
        B, H = h.shape
        X = x.shape[1]
        dev = x.device
        dhnew = dhnew.contiguous() if dhnew is not None else None
        dhnext = dhnext.contiguous() if dhnext is not None else None
        dgx = torch.empty((B, 3 * H), dtype=torch.float32, device=dev)
        dgc = torch.empty((B, H), dtype=torch.float32, device=dev)
        dz, dh = torch.empty_like(dgc), torch.empty_like(dgc)
        call("a2c_gru_gates2_bwd", ptr(dhnew), ptr(dhnext), ptr(keep), ptr(h), ptr(z), ptr(cand),
             ptr(dgc), ptr(dz), ptr(dh), B, H, stream_ptr())
        drh = gemm(dgc, H, 1, Wh[2], 1, H, B, H, H)          # dgc * Wh2^T
        dgzr = torch.empty((B, 2 * H), dtype=torch.float32, device=dev)
        call("a2c_gru_gates1_bwd", ptr(drh), ptr(dz), ptr(h), ptr(z), ptr(r), ptr(dgzr), ptr(dh), B, H,
             stream_ptr())
        dWx = torch.empty_like(Wx)
        dWh = torch.empty_like(Wh)
        db = torch.empty((3, 1, H), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        # gate pre-activation grads: z, r live in dgzr[:, g*H:], candidate in dgc
        gsrc = [(dgzr, 2 * H, 0), (dgzr, 2 * H, H), (dgc, H, 0)]
        for g, (G, ld, off) in enumerate(gsrc):
            Gp = G.view(-1)[off:]
            # dx (+)= G_g * Wx[g]^T ; dh += G_g * Wh[g]^T (g < 2)
            gemm(Gp, ld, 1, Wx[g], 1, H, B, X, H, out=dx, ldc=X, beta=0 if g == 0 else 1)
            if g < 2:
                gemm(Gp, ld, 1, Wh[g], 1, H, B, H, H, out=dh, ldc=H, beta=1)
            # dWx[g] = x^T G_g ; dWh[g] = (h or r*h)^T G_g
            gemm(x, 1, X, Gp, ld, 1, X, H, B, out=dWx, ldc=H, c_offset=g * X * H)
            src = h if g < 2 else rh
            gemm(src, 1, H, Gp, ld, 1, H, H, B, out=dWh, ldc=H, c_offset=g * H * H)
        # bias grads: column sums of the three gate-gradient blocks, written in place
        dbf = db.view(-1)
        call("a2c_colsum", ptr(dgzr), ptr(dbf), B, 2 * H, stream_ptr())
        call("a2c_colsum", ptr(dgc), ptr(dbf[2 * H:]), B, H, stream_ptr())
        return dx, dh, dWx, dWh, db, None


class GRUSequenceFn(torch.autograd.Function):
    """T steps of the reference's GRU cell with the masked carry of BPTT
    (a2c/updater.py:139-169: ``h <- cell(x_t, h); h <- h * (1 - done_t)``) as ONE autograd node.

    Everything that does not depend on the recurrence is hoisted out of the time loop and runs as
    large GEMMs over all T*R rows: the input products x_t W_x + b (forward), and in backward the
    weight gradients dW_x = sum_t x_t^T G_t, dW_h = sum_t h_t^T G_t (one split-K GEMM each, K =
    T*R), the bias gradient and dx = G W_x^T.  The loop itself keeps 5 launches per step forward
    (h W_h0, h W_h1, gates, (r*h) W_h2, gates) and 6 backward, all on preallocated [T, R, *]
    buffers.  x [T, R, X] time-major, h0 [R, H], keep [T, R] = 1 - done -> hnew [T, R, H]."""

    @staticmethod
    def forward(ctx, x, h0, Wx, Wh, b, keep):
        require_cuda(x, h0, Wx, Wh, b, keep)
        x, h0, keep = x.contiguous(), h0.contiguous(), keep.contiguous()
        T, B, X = x.shape
        H = h0.shape[1]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        gx = torch.empty((T * B, 3 * H), **f32)
        bflat = b.reshape(3, H)
        xf = x.view(T * B, X)
        for g in range(3):      # x.mm(W_x[g]) + b[g] for every step at once
            gemm(xf, X, 1, Wx[g], H, 1, T * B, H, X, out=gx, ldc=3 * H, bias=bflat[g], c_offset=g * H)
        gx = gx.view(T, B, 3 * H)
        hprev = torch.empty((T + 1, B, H), **f32)        # hprev[t] = carry entering step t
        hprev[0].copy_(h0)
        z, r, rh, cand, hnew = (torch.empty((T, B, H), **f32) for _ in range(5))
        gh = torch.empty((B, 2 * H), **f32)
        gc = torch.empty((B, H), **f32)
        for t in range(T):
            h = hprev[t]
            for g in range(2):
                gemm(h, H, 1, Wh[g], H, 1, B, H, H, out=gh, ldc=2 * H, c_offset=g * H)
            call("a2c_gru_gates1", ptr(gx[t]), ptr(gh), ptr(h), ptr(z[t]), ptr(r[t]), ptr(rh[t]), B, H,
                 stream_ptr())
            gemm(rh[t], H, 1, Wh[2], H, 1, B, H, H, out=gc, ldc=H)
            call("a2c_gru_gates2", ptr(gx[t]), ptr(gc), ptr(h), ptr(z[t]), ptr(keep[t]), ptr(cand[t]),
                 ptr(hnew[t]), ptr(hprev[t + 1]), B, H, stream_ptr())
        ctx.save_for_backward(x, Wx, Wh, keep, hprev, z, r, rh, cand)
        return hnew

    @staticmethod
    def backward(ctx, dhnew):
        x, Wx, Wh, keep, hprev, z, r, rh, cand = ctx.saved_tensors
        T, B, X = x.shape
        H = hprev.shape[2]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        dhnew = dhnew.contiguous()
        dgzr = torch.empty((T, B, 2 * H), **f32)         # z / r pre-activation gradients of every step
        dgc = torch.empty((T, B, H), **f32)              # candidate pre-activation gradients
        dz = torch.empty((B, H), **f32)
        drh = torch.empty((B, H), **f32)
        dh = [torch.empty((B, H), **f32), torch.empty((B, H), **f32)]
        carry = None                                     # d loss / d hprev[t + 1]
        for t in range(T - 1, -1, -1):
            h, cur = hprev[t], dh[t & 1]
            call("a2c_gru_gates2_bwd", ptr(dhnew[t]), ptr(carry), ptr(keep[t]), ptr(h), ptr(z[t]),
                 ptr(cand[t]), ptr(dgc[t]), ptr(dz), ptr(cur), B, H, stream_ptr())
            gemm(dgc[t], H, 1, Wh[2], 1, H, B, H, H, out=drh, ldc=H)            # dgc * Wh2^T
            call("a2c_gru_gates1_bwd", ptr(drh), ptr(dz), ptr(h), ptr(z[t]), ptr(r[t]), ptr(dgzr[t]),
                 ptr(cur), B, H, stream_ptr())
            G = dgzr[t].view(-1)
            gemm(G, 2 * H, 1, Wh[0], 1, H, B, H, H, out=cur, ldc=H, beta=1)
            gemm(G[H:], 2 * H, 1, Wh[1], 1, H, B, H, H, out=cur, ldc=H, beta=1)
            carry = cur
        dh0 = carry
        # hoisted: parameter and input gradients over all T*B rows
        N = T * B
        xf, hf, rhf = x.view(N, X), hprev[:T].reshape(N, H), rh.view(N, H)
        Gzr, Gc = dgzr.view(-1), dgc.view(-1)
        gsrc = [(Gzr, 2 * H), (Gzr[H:], 2 * H), (Gc, H)]
        dWx, dWh = torch.empty_like(Wx), torch.empty_like(Wh)
        dx = torch.empty((N, X), **f32)
        for g, (Gp, ld) in enumerate(gsrc):
            gemm(Gp, ld, 1, Wx[g], 1, H, N, X, H, out=dx, ldc=X, beta=0 if g == 0 else 1)
            gemm(xf, 1, X, Gp, ld, 1, X, H, N, out=dWx, ldc=H, c_offset=g * X * H)
            gemm(hf if g < 2 else rhf, 1, H, Gp, ld, 1, H, H, N, out=dWh, ldc=H, c_offset=g * H * H)
        db = torch.empty((3, 1, H), **f32)
        dbf = db.view(-1)
        call("a2c_colsum", ptr(dgzr), ptr(dbf), N, 2 * H, stream_ptr())
        call("a2c_colsum", ptr(dgc), ptr(dbf[2 * H:]), N, H, stream_ptr())
        return dx.view(T, B, X), dh0, dWx, dWh, db, None


# --------------------------------------------------------------------------
# clip + optimizer
# --------------------------------------------------------------------------

def grad_sumsq(flat_grad, out=None):
    if out is None:
        out = torch.zeros(1, dtype=torch.float64, device=flat_grad.device)
    call("a2c_grad_sumsq", ptr(flat_grad), ptr(out), flat_grad.numel(), stream_ptr())
    return out


def clip_optim_step(param, grad, state1, state2, sumsq, kind, lr, max_norm, step=1, alpha=0.99,
                    eps=1e-8, beta1=0.9, beta2=0.999, zero_grad=True, norm_out=None):
    cfg = OptimCfg(kind, 1 if zero_grad else 0, float(lr), float(max_norm), float(alpha),
                   float(eps), float(beta1), float(beta2), int(step))
    call("a2c_clip_optim_step", ptr(param), ptr(grad), ptr(state1), ptr(state2), ptr(sumsq),
         C.byref(cfg), ptr(norm_out), param.numel(), stream_ptr(),
         tag="a2c_clip_optim_step|bytes|%d" % param.numel(),
         work=(24 if kind == 0 else 32) * param.numel())

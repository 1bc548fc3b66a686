"""prep_fxn callables (a2c/preprocessing.py), executed by the stage-1 kernels.

``pong_prep`` is the one on the hot path of every BASELINE config (fused with the frame stack
into the plane ring, ``pong_prep_stack``); the others are the reference's remaining
preprocessors as elementwise device kernels.  Every callable keeps
the reference's per-frame signature ``f(pic: ndarray[H, W, C] uint8) -> ndarray[1, ...]`` and has
a batched ``*_batch`` device entry (frames u8 [n, H, W, C] on the GPU -> device tensor).  The
policy networks of this package take 80 x 80 states, i.e. the Pong pipeline; the other
preprocessors are provided for the ``prep_fxn`` plug point, not wired to a model.
"""
import numpy as np
import torch

from . import ops
from ._lib import call, ptr, require_cuda, stream_ptr


def pong_prep(pic):
    """a2c/preprocessing.py:11-17 on the device: (210,160,3) uint8 frame ->
    (1,80,80) uint8 in {0,1}.  Unlike the reference it does not write through
    views of ``pic``.  Per-frame calls pay a host<->device round trip; the
    batched entry is ``pong_prep_stack``."""
    frame = torch.from_numpy(np.ascontiguousarray(pic, dtype=np.uint8)).cuda()
    ring = _scratch_ring(frame.device)
    ring.push_frames(frame.view(1, 1, 210, 160, 3), _ones(frame.device))
    return ring.planes[0, 0].view(1, 80, 80).cpu().numpy()


_RINGS = {}


def _scratch_ring(device):
    key = device.index
    if key not in _RINGS:
        _RINGS[key] = (ops.PlaneRing(1, 1, 1, device), torch.ones((1, 1), dtype=torch.uint8, device=device))
    return _RINGS[key][0]


def _ones(device):
    _scratch_ring(device)
    return _RINGS[device.index][1]


def pong_prep_stack(frames_u8, reset_mask, ring, first_frame_index=0):
    """Batched device entry: frames u8 [R, F, 210,160,3], reset u8 [R, F] are
    prepped and pushed into ``ring`` (ops.PlaneRing)."""
    ring.push_frames(frames_u8, reset_mask, first_frame_index)
    return ring


def _batch(frames, kind, out_shape, dtype):
    require_cuda(frames)
    assert frames.dtype == torch.uint8 and frames.dim() == 4 and frames.is_contiguous()
    n, H, W, Cc = frames.shape
    out = torch.empty(out_shape, dtype=dtype, device=frames.device)
    call("a2c_prep_other", ptr(frames), ptr(out), kind, n, H, W, Cc, stream_ptr())
    return out


def normalize_prep_batch(frames):
    """frames u8 [n, H, W, C] -> f64 [n, 1, H, W, C] (a2c/preprocessing.py:4-6 per frame)."""
    n, H, W, Cc = frames.shape
    return _batch(frames, 1, (n, 1, H, W, Cc), torch.float64)


def snake_prep_batch(frames):
    """frames u8 [n, H, W, 3] -> f32 [n, 1, H, W] (a2c/preprocessing.py:25-32)."""
    n, H, W, _ = frames.shape
    return _batch(frames, 2, (n, 1, H, W), torch.float32)


def breakout_prep_batch(frames):
    """frames u8 [n, 210, 160, 3] -> u8 [n, 1, 80, 72] (a2c/preprocessing.py:19-23)."""
    n, H, W, _ = frames.shape
    return _batch(frames, 3, (n, 1, 80, (W - 16 + 1) // 2), torch.uint8)


def _per_frame(batch_fn):
    def f(pic):
        pic = np.ascontiguousarray(pic)
        if pic.dtype != np.uint8:
            raise TypeError("%s expects a uint8 frame (got %s)" % (batch_fn.__name__, pic.dtype))
        if pic.ndim == 2:
            pic = pic[:, :, None]
        return batch_fn(torch.from_numpy(pic[None]).cuda())[0].cpu().numpy()
    return f


normalize_prep = _per_frame(normalize_prep_batch)
normalize_prep.__name__ = "normalize_prep"
snake_prep = _per_frame(snake_prep_batch)
snake_prep.__name__ = "snake_prep"
breakout_prep = _per_frame(breakout_prep_batch)
breakout_prep.__name__ = "breakout_prep"


def null_prep(pic):
    """a2c/preprocessing.py:8-9."""
    return pic[None]

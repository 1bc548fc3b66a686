"""prep_fxn callables (a2c/preprocessing.py).  Only ``pong_prep`` is on the
hot path of the BASELINE configs; the others are out of scope (SURVEY.md
section 8f item 3)."""
import numpy as np
import torch

from . import ops


def pong_prep(pic):
    """a2c/preprocessing.py:11-17 on the device: (210,160,3) uint8 frame ->
    (1,80,80) uint8 in {0,1}.  Unlike the reference it does not write through
    views of ``pic``.  Per-frame calls pay a host<->device round trip; the
    batched entry is ``pong_prep_stack``."""
    frame = torch.from_numpy(np.ascontiguousarray(pic, dtype=np.uint8)).cuda()
    ring = ops.PlaneRing(1, 1, 1, frame.device)
    reset = torch.ones((1, 1), dtype=torch.uint8, device=frame.device)
    ring.push_frames(frame.view(1, 1, 210, 160, 3), reset)
    return ring.planes[0, 0].view(1, 80, 80).cpu().numpy()


def pong_prep_stack(frames_u8, reset_mask, ring, first_frame_index=0):
    """Batched device entry: frames u8 [R, F, 210,160,3], reset u8 [R, F] are
    prepped and pushed into ``ring`` (ops.PlaneRing)."""
    ring.push_frames(frames_u8, reset_mask, first_frame_index)
    return ring


def _out_of_scope(name):
    def f(pic):
        raise NotImplementedError("%s is out of scope for the B200 hot path (no BASELINE config "
                                  "uses it)" % name)
    f.__name__ = name
    return f


breakout_prep = _out_of_scope("breakout_prep")
snake_prep = _out_of_scope("snake_prep")
normalize_prep = _out_of_scope("normalize_prep")
null_prep = _out_of_scope("null_prep")

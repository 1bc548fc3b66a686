"""Helpers with the reference's names (a2c/utils.py:4-79); ``discount`` and
``sample_action`` run on the device."""
from collections import deque  # noqa: F401  (re-exported for Runner users)

import numpy as np
import torch

from . import ops


def try_key(dict_, key, default):
    """a2c/utils.py:4-7."""
    if key in dict_:
        return dict_[key]
    return default


def cuda_if(tobj):
    """a2c/utils.py:9-12."""
    if torch.cuda.is_available():
        tobj = tobj.cuda()
    return tobj


def deque_maxmin(deq):
    """a2c/utils.py:14-24."""
    max_val, min_val = deq[0], deq[0]
    for v in deq:
        if v > max_val:
            max_val = v
        if v < min_val:
            min_val = v
    return max_val, min_val


def next_state(env, obs_deque, obs, reset):
    """a2c/utils.py:26-43 for a single env (host-side; the batched device path
    is ops.PlaneRing).  Kept so per-env callers keep working."""
    if reset:
        obs = env.reset()
        for _ in range(obs_deque.maxlen - 1):
            obs_deque.append(np.zeros(obs.shape))
    obs_deque.append(obs)
    return np.concatenate(obs_deque, axis=0)


def sample_action(pi, rand_nums=None):
    """a2c/utils.py:45-60 on the device: pi (..., A) probabilities -> float
    tensor of action indices (-1 where the running sum never reaches u), on the
    CPU like the reference's result.  ``rand_nums`` lets callers supply the
    uniform draws (default: torch.rand, as the reference)."""
    lead = pi.shape[:-1]
    if rand_nums is None:
        rand_nums = torch.rand(*lead)
    p2 = cuda_if(pi.reshape(-1, pi.shape[-1]).float())
    u = cuda_if(torch.as_tensor(rand_nums, dtype=torch.float32).reshape(-1))
    acts = ops.sample_action(p2, u, is_prob=True)
    return acts.float().reshape(lead).cpu()


def discount(array, dones, discount_factor):
    """a2c/utils.py:63-79 as one kernel launch; returns a CPU FloatTensor like
    the reference does (a2c/utils.py:74)."""
    a = cuda_if(torch.as_tensor(array, dtype=torch.float32))
    d = cuda_if(torch.as_tensor(dones, dtype=torch.float32))
    return ops.discount(a, d, discount_factor).cpu()

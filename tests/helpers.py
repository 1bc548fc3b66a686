"""Shared helpers for the parity tests: rebuild the seed-derived inputs a golden
fixture was generated from (tests/golden/make_golden.py) and compare compactly
stored tensors."""
import json
import os
from collections import OrderedDict

import numpy as np
import torch

from a2c_b200 import synth
from oracle import a2c_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SAMPLE = 4096


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name), allow_pickle=False))


def hyps_of(g):
    return json.loads(str(g["hyps_json"]))


def compact(a):
    """Same compaction as make_golden._compact."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    if a.size <= SAMPLE:
        return a.copy()
    stride = a.size // SAMPLE
    return np.concatenate([[np.float32(np.sqrt((a.astype(np.float64) ** 2).sum()))],
                           a[::stride][:SAMPLE]]).astype(np.float32)


def rel_err(a, b):
    """max |a-b| / max |b| (0 when both are all-zero)."""
    a = np.asarray(a, dtype=np.float64).reshape(-1)
    b = np.asarray(b, dtype=np.float64).reshape(-1)
    den = np.abs(b).max() if b.size else 0.0
    if den == 0:
        return float(np.abs(a).max()) if a.size else 0.0
    return float(np.abs(a - b).max() / den)


def compact_err(full, golden_compact):
    """Error of a full tensor against its compactly stored golden value: the
    L2 norm and the strided samples are compared separately (each relative to
    its own scale) so a large norm cannot hide a wrong sample."""
    full = np.asarray(full, dtype=np.float32).reshape(-1)
    if full.size <= SAMPLE:
        return rel_err(full, golden_compact)
    c = compact(full)
    return max(rel_err(c[:1], golden_compact[:1]), rel_err(c[1:], golden_compact[1:]))


def golden_grad_scale(g, prefix):
    """Largest gradient entry over all parameters of a golden update."""
    m = 0.0
    for k, v in g.items():
        if k.startswith(prefix):
            v = np.asarray(v).reshape(-1)
            m = max(m, float(np.abs(v[1:] if v.size > SAMPLE else v).max()))
    return m


def grad_err(full, golden_compact, scale, floor=1e-3):
    """Gradient parity metric.  Per tensor: max |a-b| / max(max |b|, floor*scale)
    where ``scale`` is the largest gradient entry of the whole model.  A tensor
    whose entries are all below floor*scale (e.g. a first-layer gradient that is
    a sum of cancelling terms 1000x smaller than the rest) is compared at
    floor*scale: fp32 with a different summation order cannot resolve better
    than ~1e-7 of the terms being cancelled.  Big tensors: norm and samples are
    checked separately."""
    full = np.asarray(full, dtype=np.float32).reshape(-1)
    c = compact(full)
    gc = np.asarray(golden_compact, dtype=np.float64).reshape(-1)
    if full.size > SAMPLE:
        norm_err = abs(float(c[0]) - gc[0]) / max(gc[0], floor * scale * np.sqrt(full.size))
        den = max(np.abs(gc[1:]).max(), floor * scale)
        return max(norm_err, float(np.abs(c[1:].astype(np.float64) - gc[1:]).max() / den))
    den = max(np.abs(gc).max() if gc.size else 0.0, floor * scale)
    return float(np.abs(c.astype(np.float64) - gc).max() / den) if den > 0 else 0.0


def golden_true_norm(g, prefix):
    """sqrt(sum ||g_i||^2) in float64 from the stored per-parameter gradients.
    NOTE: the reference's reported GradNorm (torch clip_grad_norm_, fp32 on CPU)
    deviates from this by up to ~2e-4 relative on multi-million-element tensors
    (measured: 12.8676 reported vs 12.8698 true for FCModel C1)."""
    tot = 0.0
    for k, v in g.items():
        if k.startswith(prefix):
            v = np.asarray(v, dtype=np.float64).reshape(-1)
            tot += v[0] ** 2 if v.size > SAMPLE else float((v ** 2).sum())
    return float(np.sqrt(tot))


def param_err_ok(p_after, golden_compact, step, tol=1e-5, step_frac=5e-2):
    """Post-step parameter parity: |p - p_ref| <= tol*max|p_ref| + step_frac*step.
    The second term exists because RMSprop/Adam's first steps are sign-like
    (step ~ lr*g/(|g|/sqrt(1-alpha)+eps)): for first-layer gradients of order
    1e-6 the eps term makes the update sensitive to the gradient's relative
    (not absolute) error, and those gradients are only known to ~1e-2 relative
    in fp32 (reference included).  The optimizer kernel itself is checked at
    5e-6 against torch.optim on identical gradients
    (test_fused_clip_optimizer_matches_torch); gradients are checked separately."""
    full = np.asarray(p_after, dtype=np.float32).reshape(-1)
    c = compact(full).astype(np.float64)
    gc = np.asarray(golden_compact, dtype=np.float64).reshape(-1)
    if full.size > SAMPLE:
        c, gc = c[1:], gc[1:]
    return float(np.abs(c - gc).max()) <= tol * float(np.abs(gc).max()) + step_frac * step


def grad_err_compact(a_compact, b_compact, scale, floor=1e-3):
    """grad_err between two compactly stored tensors (same compaction)."""
    a = np.asarray(a_compact, dtype=np.float64).reshape(-1)
    b = np.asarray(b_compact, dtype=np.float64).reshape(-1)
    if a.size == SAMPLE + 1:      # [norm, samples...]
        den = max(np.abs(b[1:]).max(), floor * scale)
        return max(abs(a[0] - b[0]) / max(b[0], 1e-30), float(np.abs(a[1:] - b[1:]).max() / den))
    den = max(np.abs(b).max() if b.size else 0.0, floor * scale)
    return float(np.abs(a - b).max() / den) if den > 0 else 0.0


def params_for(model, fs, h, A=3):
    shapes = O.param_shapes(model, fs=fs, h=h, A=A)
    return OrderedDict((k, torch.from_numpy(v)) for k, v in
                       synth.hashed_state_dict(shapes).items())


def small_trajectories(g, R, T):
    seed = int(g["seed"])
    return [synth.make_trajectory(R, T, seed=seed + 17 * k, first_rollout=(k == 0))
            for k in range(int(g["n_updates"]))]


def chain_frames(trajs):
    """Frames / reset flags of consecutive rollouts as the runner sees them:
    frame 0 of rollout k>0 is the bookmark, i.e. frame T of rollout k-1."""
    frames = [trajs[0]["frames"]]
    reset = [trajs[0]["reset"]]
    for t in trajs[1:]:
        frames.append(t["frames"][:, 1:])
        reset.append(t["reset"][:, 1:])
    return np.concatenate(frames, axis=1), np.concatenate(reset, axis=1)


def golden_states(g, k, N, fs):
    bits = np.unpackbits(g["u%d_states_bits" % k])[:N * fs * 6400]
    return bits.reshape(N, fs, 80, 80).astype(np.uint8)


def full_inputs(model, R, T, fs, h, seed):
    """Mirror of make_golden.full_inputs with the oracle's prep (bit-exact)."""
    traj = synth.make_trajectory(R, T, seed=seed)
    states = O.prep_and_stack(traj["frames"][:, :T], traj["reset"][:, :T], fs)
    N = R * T
    done = ((traj["reset"][:, 1:] != 0) | (traj["rewards"] != 0)).astype(np.float32)
    r2, d2, deltas = O.rollout_bookkeeping(traj["rewards"], done, traj["values"], 0.99)
    rng = np.random.default_rng(seed + 1)
    actions = torch.from_numpy(rng.integers(0, 3, size=N)).long()
    shared = {"states_u8": torch.from_numpy(states.reshape(N, fs, 80, 80)),
              "rewards": r2.reshape(-1).clone(), "dones": d2.reshape(-1).clone(),
              "deltas": deltas.reshape(-1).clone(), "actions": actions}
    if model == "GRUModel":
        shared["h_states"] = torch.from_numpy(synth.hashed_tensor((N, h), 999, 0.5)).float()
    return traj, shared

"""Deterministic synthetic Pong-shaped trajectories.

Host-side utility shared by the golden-fixture script, the parity tests and
bench.py.  Everything is produced by numpy's PCG64 integer streams so the same
seed gives the same bytes on every machine (fixtures store outputs only).

A trajectory for ``R`` rollout slots of ``T`` steps holds

  frames  u8  [R, T+1, 210, 160, 3]  frame t is the newest observation of state t
  reset   u8  [R, T+1]               reset[r, 0]: state 0 starts a fresh episode;
                                     reset[r, t+1]: env returned done at step t, so
                                     frame t+1 is what ``env.reset()`` produced
                                     (a2c/utils.py:37-40)
  rewards f32 [R, T]                 in {-1, 0, +1}, P(!=0) = 0.1
  u       f32 [R, T]                 the uniform draw ``sample_action`` would make
  values  f32 [R, T+1]               rollout-time V(s_t), V(s_T) is the bootstrap

Backgrounds 144 and 109 are over-represented so that all three branches of
pong_prep (a2c/preprocessing.py:14-16) are exercised.
"""
import numpy as np

FRAME_H, FRAME_W, FRAME_C = 210, 160, 3
FRAME_BYTES = FRAME_H * FRAME_W * FRAME_C
PLANE = 80
PLANE_BYTES = PLANE * PLANE


def make_frames(rng, n, p_noise=0.1):
    """n frames, background 144 or 109 (per frame), a fraction p_noise of the
    bytes uniform in 0..255 (so 0, 144 and 109 also occur as noise)."""
    bg = rng.integers(0, 2, size=(n, 1), dtype=np.uint8)
    bg = np.where(bg == 0, np.uint8(144), np.uint8(109)).astype(np.uint8)
    out = np.empty((n, FRAME_BYTES), dtype=np.uint8)
    # chunked so peak memory stays ~2x the output
    step = max(1, (64 << 20) // FRAME_BYTES)
    for s in range(0, n, step):
        e = min(n, s + step)
        noise = rng.integers(0, 256, size=(e - s, FRAME_BYTES), dtype=np.uint8)
        sel = rng.integers(0, 256, size=(e - s, FRAME_BYTES), dtype=np.uint8)
        thr = int(round(p_noise * 256))
        out[s:e] = np.where(sel < thr, noise, bg[s:e])
    return out.reshape(n, FRAME_H, FRAME_W, FRAME_C)


def make_trajectory(R, T, seed=1234, p_rew=0.1, p_reset=0.02, first_rollout=True,
                    with_frames=True):
    rng = np.random.default_rng(seed)
    traj = {}
    if with_frames:
        traj["frames"] = make_frames(rng, R * (T + 1)).reshape(
            R, T + 1, FRAME_H, FRAME_W, FRAME_C)
    sign = rng.integers(0, 2, size=(R, T)).astype(np.float32) * 2 - 1
    hit = rng.random(size=(R, T)) < p_rew
    traj["rewards"] = np.where(hit, sign, 0).astype(np.float32)
    reset = np.zeros((R, T + 1), dtype=np.uint8)
    reset[:, 1:] = rng.random(size=(R, T)) < p_reset
    reset[:, 0] = 1 if first_rollout else 0
    traj["reset"] = reset
    # 24-bit uniforms: exactly representable in fp32, same law as torch.rand
    traj["u"] = (rng.integers(0, 1 << 24, size=(R, T)).astype(np.float32)
                 / np.float32(1 << 24))
    traj["values"] = rng.standard_normal(size=(R, T + 1)).astype(np.float32)
    return traj


def hashed_tensor(shape, salt, scale):
    """Weights that are a pure function of (index, salt): exact in fp32, no RNG
    library involved, so the golden script (reference side) and the tests
    (oracle / CUDA side) rebuild identical parameters without storing them.
    ``scale`` should be a power of two."""
    n = int(np.prod(shape)) if len(shape) else 1
    i = np.arange(n, dtype=np.uint64)
    h = (i * np.uint64(2654435761) + np.uint64(salt) * np.uint64(40503)
         + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    v = ((h >> np.uint64(8)).astype(np.float64) / float(1 << 24) - 0.5) * 2.0
    return (v * scale).astype(np.float32).reshape(shape)


def hashed_state_dict(shapes, scale_for=None):
    """shapes: ordered {name: shape}.  Returns {name: fp32 ndarray}.  Weight
    scale is 2^-k with 2^-k <= 1/sqrt(fan_in) (keeps activations O(1))."""
    out = {}
    for salt, (name, shape) in enumerate(shapes.items()):
        shape = tuple(shape)
        if scale_for is not None and name in scale_for:
            scale = scale_for[name]
        elif len(shape) >= 2:
            fan_in = int(np.prod(shape[1:]))
            if name.startswith("gru.W"):
                fan_in = shape[1]
            scale = 2.0 ** np.floor(np.log2(1.0 / np.sqrt(fan_in)))
        elif name.endswith("weight") and len(shape) == 1:
            out[name] = (1.0 + hashed_tensor(shape, salt + 1, 2.0 ** -3)).astype(np.float32)
            continue
        else:
            scale = 2.0 ** -4
        out[name] = hashed_tensor(shape, salt + 1, scale)
    return out

"""CPU oracle for the A2C rollout-post-processing + update hot path.

TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs
of ``bench.py`` may import this module, and only as the checker / the reported
CPU baseline.  The product (``pytorch-a2c_b200/``) never imports it and raises
when its CUDA library is missing.

What it is: a restatement, in numpy / torch-CPU fp32, of the algorithm in
grantsrb/PyTorch-A2C for this path.  Each function cites the reference
file:line it follows (paths relative to the reference root).  Byte / index work
is numpy integer arithmetic; fp32 work uses torch-CPU ops in the reference's
association order, so results are bit-identical to the reference on the same
torch build wherever the reference is deterministic.

Parity pin: the reference ships no tests or golden vectors,
and its arithmetic lives in un-vendored, unpinned third-party packages
(``torch``, ``numpy``; requirements.txt:1-2).  The pin is therefore
``tests/golden/*.npz``: outputs of the UNMODIFIED reference modules imported
from /root/reference under torch 2.11.0 (CPU) / numpy 2.3.5 by
``tests/golden/make_golden.py`` (committed next to the fixtures).
``tests/test_oracle_golden.py`` checks every function here against those
fixtures, and against the live reference when it is mounted.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------
# stage 1: observation prep + frame stack (uint8, bit-exact)
# --------------------------------------------------------------------------

def pong_prep(pic):
    """a2c/preprocessing.py:11-17.  crop rows 35:195, every 2nd row / column,
    channel 0; 144 -> 0, 109 -> 0, anything else non-zero -> 1; leading axis.
    The reference does this in place on views of the caller's frame; the
    result is identical, we just do not mutate ``pic``."""
    v = np.ascontiguousarray(pic[35:195:2, ::2, 0])
    out = ((v != 0) & (v != 144) & (v != 109)).astype(np.uint8)
    return out[None]


def normalize_prep(pic):
    """a2c/preprocessing.py:4-6."""
    new_pic = 3 * (pic - 255 / 2) / (255 / 2)
    return new_pic[None]


def null_prep(pic):
    """a2c/preprocessing.py:8-9."""
    return pic[None]


def breakout_prep(pic):
    """a2c/preprocessing.py:19-23.  ``rgb2grey`` of the 2-D channel-0 array is the identity in
    the skimage releases that still export that name (rgb2gray returned 2-D input unchanged
    until it was removed); skimage is not installed here, so this step is restated, not run."""
    pic = pic[35:195, 8:-8]
    pic = pic[::2, ::2, 0]
    return np.ascontiguousarray(pic)[None]


def snake_prep(pic):
    """a2c/preprocessing.py:25-32."""
    new_pic = np.zeros(pic.shape[:2], dtype=np.float32)
    new_pic[:, :][pic[:, :, 0] == 1] = 1
    new_pic[:, :][pic[:, :, 0] == 255] = 1.5
    new_pic[:, :][pic[:, :, 1] == 255] = 0
    new_pic[:, :][pic[:, :, 2] == 255] = .33
    return new_pic[None]


class FrameStack:
    """a2c/utils.py:26-43 (``next_state``) with its ``deque(maxlen=fs)``
    (a2c/runner.py:145).  On reset: fs-1 zero planes then the reset
    observation; states list planes oldest first along axis 0."""

    def __init__(self, n_frame_stack):
        self.fs = n_frame_stack
        self.planes = []

    def push(self, obs, reset):
        if reset:
            self.planes = [np.zeros_like(obs) for _ in range(self.fs - 1)]
        self.planes.append(obs)
        self.planes = self.planes[-self.fs:]
        return np.concatenate(self.planes, axis=0)


def prep_and_stack(frames, reset, fs, carry=None):
    """Whole-trajectory form of the two functions above.

    frames u8 [R, F, 210, 160, 3]; reset u8 [R, F] (reset[r, f] != 0: frame f
    starts a fresh stack); carry: optional list (len R) of FrameStack objects
    holding the previous rollout's planes.  Returns states u8 [R, F, fs, 80, 80].
    """
    R, Fn = frames.shape[:2]
    out = np.zeros((R, Fn, fs, 80, 80), dtype=np.uint8)
    for r in range(R):
        st = carry[r] if carry is not None else FrameStack(fs)
        for f in range(Fn):
            out[r, f] = st.push(pong_prep(frames[r, f]), bool(reset[r, f]))
    return out


# --------------------------------------------------------------------------
# acting: inverse-CDF action sampling (index bit-exact given pi, u)
# --------------------------------------------------------------------------

def sample_action(pi, u):
    """a2c/utils.py:45-60 with the uniform draw ``u`` (``torch.rand`` at :54)
    passed in.  Sequential fp32 running sum over the action axis; the action is
    the first index whose running sum reaches u; -1 when none does."""
    pi = torch.as_tensor(pi, dtype=torch.float32)
    u = torch.as_tensor(u, dtype=torch.float32)
    cum = torch.zeros(pi.shape[:-1], dtype=torch.float32)
    act = torch.full(pi.shape[:-1], -1.0, dtype=torch.float32)
    for i in range(pi.shape[-1]):
        cum = cum + pi[..., i]
        hit = (cum >= u) & (act < 0)
        act = torch.where(hit, torch.full_like(act, float(i)), act)
    return act


def get_action(logits, u):
    """a2c/runner.py:94-97: softmax then sample_action, returned as int64."""
    probs = F.softmax(torch.as_tensor(logits, dtype=torch.float32), dim=-1)
    return sample_action(probs, u).to(torch.int64)


# --------------------------------------------------------------------------
# stage 2: rollout bookkeeping (deltas, bootstrap fold) and reverse scan
# --------------------------------------------------------------------------

def rollout_bookkeeping(rewards, dones, values, gamma):
    """a2c/runner.py:222-245 on recorded values, for R rollout slots at once.

    rewards f32 [R, T] raw env rewards; dones f32 [R, T] the stored done flag
    (already including the Pong pseudo-done of :213-214); values f32 [R, T+1]
    rollout-time V(s_t) with V(s_T) the bootstrap value (:237-241).
    Returns (rewards', dones', deltas) exactly as ``datas`` holds them after
    ``Runner.rollout``: bootstrap folded into the last reward and dones[end]=1
    when the last step was not done (:236-243), deltas per :228-232 and :244.
    All arithmetic fp32, gamma rounded to fp32, mul then add (no fma)."""
    r = torch.as_tensor(rewards, dtype=torch.float32).clone()
    d = torch.as_tensor(dones, dtype=torch.float32).clone()
    v = torch.as_tensor(values, dtype=torch.float32)
    T = r.shape[1]
    g = torch.tensor(gamma, dtype=torch.float32)
    deltas = torch.zeros_like(r)
    # :231  prev_rew + gamma*val*(1-prev_done) - prev_val
    deltas[:, :T - 1] = r[:, :T - 1] + (g * v[:, 1:T]) * (1 - d[:, :T - 1]) - v[:, :T - 1]
    notdone = d[:, T - 1] == 0
    boot = r[:, T - 1] + g * v[:, T]                       # :242
    r[:, T - 1] = torch.where(notdone, boot, r[:, T - 1])
    d[:, T - 1] = torch.where(notdone, torch.ones_like(d[:, T - 1]), d[:, T - 1])  # :243
    deltas[:, T - 1] = r[:, T - 1] - v[:, T - 1]           # :244
    return r, d, deltas


def discount(array, dones, discount_factor):
    """a2c/utils.py:63-79: flat reverse scan, reset where dones[i] == 1 BEFORE
    adding array[i]; fp32 with the factor rounded to fp32, mul then add."""
    a = np.asarray(torch.as_tensor(array, dtype=torch.float32).reshape(-1).numpy())
    d = np.asarray(torch.as_tensor(dones, dtype=torch.float32).reshape(-1).numpy())
    g = np.float32(discount_factor)
    out = np.zeros(len(a), dtype=np.float32)
    rs = np.float32(0)
    for i in range(len(a) - 1, -1, -1):
        if d[i] == 1:
            rs = np.float32(0)
        rs = np.float32(a[i] + np.float32(g * rs))
        out[i] = rs
    return torch.from_numpy(out)


# --------------------------------------------------------------------------
# stage 4: policy networks (functional; parameter names = reference state_dict)
# --------------------------------------------------------------------------

def param_shapes(model, fs=2, h=256, A=3):
    """Ordered {state_dict key: shape} of the trainable parameters, aliases
    removed (the reference registers the same conv blocks under ``convs.*``,
    ``features.*`` and, for A3CModel, ``conv1/conv2``; a2c/models.py:22-38,
    196-240, 565-621)."""
    s = OrderedDict()
    if model == "FCModel":                               # a2c/models.py:367-395
        s["base.0.weight"] = (h, fs * 6400); s["base.0.bias"] = (h,)
        s["base.2.weight"] = (h, h); s["base.2.bias"] = (h,)
        s["action_out.weight"] = (A, h); s["action_out.bias"] = (A,)
        s["value_out.0.weight"] = (h,); s["value_out.0.bias"] = (h,)
        s["value_out.1.weight"] = (1, h); s["value_out.1.bias"] = (1,)
        s["value_out.2.weight"] = (1, 1); s["value_out.2.bias"] = (1,)
    elif model == "A3CModel":                            # a2c/models.py:7-47
        s["convs.0.0.weight"] = (16, fs, 8, 8); s["convs.0.0.bias"] = (16,)
        s["convs.1.0.weight"] = (32, 16, 4, 4); s["convs.1.0.bias"] = (32,)
        s["proj_matrx.weight"] = (h, 2048); s["proj_matrx.bias"] = (h,)
        s["pi.weight"] = (A, h); s["pi.bias"] = (A,)
        s["value.weight"] = (1, h); s["value.bias"] = (1,)
    elif model == "ConvModel":                           # a2c/models.py:177-264
        chans = [fs, 16, 24, 32, 64]
        for i in range(4):
            s["convs.%d.0.weight" % i] = (chans[i + 1], chans[i], 3, 3)
            s["convs.%d.0.bias" % i] = (chans[i + 1],)
        s["resize_emb.0.weight"] = (2000, 25600); s["resize_emb.0.bias"] = (2000,)
        s["pi.0.weight"] = (h, 2000); s["pi.0.bias"] = (h,)
        s["pi.2.weight"] = (A, h); s["pi.2.bias"] = (A,)
        s["value.0.weight"] = (h, 2000); s["value.0.bias"] = (h,)
        s["value.2.weight"] = (1, h); s["value.2.bias"] = (1,)
    elif model == "GRUModel":                            # a2c/models.py:546-636
        chans = [fs, 16, 24, 32, 48, 64]
        for i in range(5):
            s["convs.%d.0.weight" % i] = (chans[i + 1], chans[i], 3, 3)
            s["convs.%d.0.bias" % i] = (chans[i + 1],)
        s["resize_emb.0.weight"] = (h, 1600); s["resize_emb.0.bias"] = (h,)
        s["gru.W_x"] = (3, h, h); s["gru.W_h"] = (3, h, h); s["gru.b"] = (3, 1, h)
        s["pi.weight"] = (A, h); s["pi.bias"] = (A,)
        s["value.weight"] = (1, h); s["value.bias"] = (1,)
    else:
        raise KeyError(model)
    return s


CONV_STRIDES = {"ConvModel": (1, 1, 2, 2), "GRUModel": (1, 2, 2, 2, 2)}


def is_recurrent(model):
    return model == "GRUModel"


def gru_cell(p, x, h):
    """a2c/models.py:465-476.  Right-multiplication x.mm(W); the candidate uses
    (r*h).mm(W_h[2]); h' = z*h + (1-z)*cand."""
    Wx, Wh, b = p["gru.W_x"], p["gru.W_h"], p["gru.b"]
    z = torch.sigmoid(x.mm(Wx[0]) + h.mm(Wh[0]) + b[0])
    r = torch.sigmoid(x.mm(Wx[1]) + h.mm(Wh[1]) + b[1])
    cand = torch.tanh(x.mm(Wx[2]) + (r * h).mm(Wh[2]) + b[2])
    return z * h + (1 - z) * cand


def model_forward(model, p, x, h=None):
    """Forward of the four policy networks with bnorm off.
    Returns (val [B,1], logits [B,A]) or (val, logits, h') for GRUModel."""
    if model == "FCModel":                               # a2c/models.py:396-405
        fx = x.reshape(len(x), -1)
        fx = F.relu(F.linear(fx, p["base.0.weight"], p["base.0.bias"]))
        fx = F.linear(fx, p["base.2.weight"], p["base.2.bias"])
        logits = F.linear(fx, p["action_out.weight"], p["action_out.bias"])
        v = F.layer_norm(fx, (fx.shape[-1],), p["value_out.0.weight"], p["value_out.0.bias"])
        v = F.linear(v, p["value_out.1.weight"], p["value_out.1.bias"])
        v = F.linear(v, p["value_out.2.weight"], p["value_out.2.bias"])
        return v, logits
    if model == "A3CModel":                              # a2c/models.py:59-90
        f = F.relu(F.conv2d(x, p["convs.0.0.weight"], p["convs.0.0.bias"], stride=4))
        f = F.relu(F.conv2d(f, p["convs.1.0.weight"], p["convs.1.0.bias"], stride=2))
        emb = F.linear(f.reshape(len(f), -1), p["proj_matrx.weight"], p["proj_matrx.bias"])
        logits = F.linear(emb, p["pi.weight"], p["pi.bias"])
        v = F.linear(emb.detach(), p["value.weight"], p["value.bias"])   # :85 detached
        return v, logits
    if model == "ConvModel":                             # a2c/models.py:276-304
        f = x
        for i, st in enumerate(CONV_STRIDES[model]):
            f = F.relu(F.conv2d(f, p["convs.%d.0.weight" % i], p["convs.%d.0.bias" % i],
                                stride=st, padding=1))
        e = F.relu(F.linear(f.reshape(len(f), -1), p["resize_emb.0.weight"],
                            p["resize_emb.0.bias"]))
        hp = F.relu(F.linear(e, p["pi.0.weight"], p["pi.0.bias"]))
        logits = F.linear(hp, p["pi.2.weight"], p["pi.2.bias"])
        hv = F.relu(F.linear(e, p["value.0.weight"], p["value.0.bias"]))
        v = F.linear(hv, p["value.2.weight"], p["value.2.bias"])
        return v, logits
    if model == "GRUModel":                              # a2c/models.py:648-677
        f = x
        for i, st in enumerate(CONV_STRIDES[model]):     # "lerelu" -> ReLU (:688-689)
            f = F.relu(F.conv2d(f, p["convs.%d.0.weight" % i], p["convs.%d.0.bias" % i],
                                stride=st, padding=1))
        e = F.relu(F.linear(f.reshape(len(f), -1), p["resize_emb.0.weight"],
                            p["resize_emb.0.bias"]))
        hn = gru_cell(p, e, h)
        logits = F.linear(hn, p["pi.weight"], p["pi.bias"])
        v = F.linear(hn, p["value.weight"], p["value.bias"])
        return v, logits, hn
    raise KeyError(model)


def bptt(model, p, states, h_states, dones, R, T):
    """a2c/updater.py:139-169: h0 = stored h at t=0 of every rollout; T
    sequential full-model steps; h *= (1-done_t) after each; outputs flattened
    rollout-major."""
    hs = h_states.view(R, T, -1)[:, 0]
    st = states.view(R, T, *states.shape[1:])
    ds = 1 - dones.view(R, T, 1)
    vals, logits = [], []
    for t in range(T):
        v, l, hs = model_forward(model, p, st[:, t], hs)
        hs = hs * ds[:, t]
        vals.append(v)
        logits.append(l.unsqueeze(1))
    vals = torch.cat(vals, dim=-1).reshape(-1)
    logits = torch.cat(logits, dim=1).reshape(-1, logits[0].shape[-1])
    return vals, logits


# --------------------------------------------------------------------------
# stage 3 + update: returns, normalisation, loss, backward, clip
# --------------------------------------------------------------------------

def a2c_loss(logits, vals, actions, advs, returns, hyps):
    """a2c/updater.py:101-106,119-127 (discrete branch).  advs / returns carry
    no gradient.  Negative actions index from the end, as torch indexing does
    for the reference's ``log_softs[arange, actions]``."""
    lsm = F.log_softmax(logits, dim=-1)
    idx = torch.arange(len(actions), device=logits.device)
    log_ps = lsm[idx, actions]
    entr_loss = -hyps["entr_coef"] * (lsm * F.softmax(logits, dim=-1)).sum(-1).mean()
    pi_loss = hyps["pi_coef"] * -(log_ps * advs).mean()
    val_loss = hyps["val_coef"] * F.mse_loss(vals.squeeze(), returns)
    loss = pi_loss + val_loss - entr_loss
    return loss, pi_loss, val_loss, entr_loss


class UpdaterOracle:
    """a2c/updater.py:9-137.  ``compute`` is update_model up to and including
    loss.backward() + clip_grad_norm_; ``step`` adds optimizer.step()."""

    def __init__(self, model, params, hyps, dtype=torch.float32, device="cpu"):
        # dtype=float64 evaluates the same algorithm in double (scans stay fp32,
        # they are bit-exact by construction): the yardstick for ill-conditioned sums.
        # device="cuda" (tests only, with float64) runs the SAME torch ops on the GPU's fp64
        # units so that full-size gradients can be compared tensor by tensor in seconds; its
        # output is pinned to the committed CPU-generated fp64 fixtures before it is trusted.
        self.model = model
        self.hyps = hyps
        self.dtype = dtype
        self.device = torch.device(device)
        self.p = OrderedDict((k, torch.as_tensor(v).to(self.device, dtype).clone()
                              .requires_grad_(True)) for k, v in params.items())
        self.optim = getattr(torch.optim, hyps["optim_type"])(
            list(self.p.values()), lr=hyps["lr"])            # :226-229
        self.ret_mean = None
        self.ret_std = None

    def forward(self, shared):
        hyps = self.hyps
        states = shared["states"].to(self.device, self.dtype)
        if is_recurrent(self.model):
            hs = shared["h_states"].to(self.device, self.dtype)
            if hyps["use_bptt"]:                             # :76
                return bptt(self.model, self.p, states, hs, shared["dones"].to(self.device, self.dtype),
                            hyps["n_rollouts"], hyps["n_tsteps"])
            v, l, _ = model_forward(self.model, self.p, states, hs)  # :78
            return v, l
        return model_forward(self.model, self.p, states)     # :80

    def compute(self, shared):
        hyps = self.hyps
        dones = shared["dones"].reshape(-1).cpu()
        advs = discount(shared["deltas"].cpu(), dones,
                        hyps["gamma"] * hyps["lambda_"]).to(self.device, self.dtype)  # :70
        vals, logits = self.forward(shared)
        if hyps["use_nstep_rets"]:                           # :83-84
            returns = advs + vals.detach().reshape(-1)
        else:                                                # :86-88
            returns = discount(shared["rewards"].cpu(), dones, hyps["gamma"]).to(self.device, self.dtype)
        raw_advs, raw_returns = advs.clone(), returns.clone()
        if hyps.get("norm_returns", False):                  # :89-96
            if self.ret_mean is None:
                self.ret_mean, self.ret_std = returns.mean(), returns.std()
            else:
                self.ret_mean = 0.01 * returns.mean() + 0.99 * self.ret_mean
                self.ret_std = 0.01 * returns.std() + 0.99 * self.ret_std
            returns = (returns - self.ret_mean) / (self.ret_std + 1e-6)
        if hyps["norm_advs"]:                                # :97-98 (unbiased std)
            advs = (advs - advs.mean()) / (advs.std() + 1e-6)
        loss, pi_loss, val_loss, entr_loss = a2c_loss(
            logits, vals, shared["actions"].reshape(-1).long().to(self.device), advs, returns, hyps)
        for q in self.p.values():
            q.grad = None
        loss.backward()                                      # :128
        grads = OrderedDict((k, (q.grad.clone() if q.grad is not None else None))
                            for k, q in self.p.items())
        with_grad = [q for q in self.p.values() if q.grad is not None]
        norm = torch.nn.utils.clip_grad_norm_(with_grad, hyps["max_norm"])  # :129
        return {
            "Loss": loss.item(), "Pi_Loss": pi_loss.item(), "ValLoss": val_loss.item(),
            "Entropy": entr_loss.item(), "GradNorm": float(norm),
            "advs": raw_advs, "returns": raw_returns, "norm_advs": advs.detach(),
            "vals": vals.detach().reshape(-1), "logits": logits.detach(),
            "grads": grads,
            "clipped_grads": OrderedDict((k, (q.grad.clone() if q.grad is not None else None))
                                         for k, q in self.p.items()),
        }

    def step(self, shared):
        out = self.compute(shared)
        self.optim.step()                                    # :131
        self.optim.zero_grad()                               # :132
        return out


# --------------------------------------------------------------------------
# whole rollout, batched over slots (a2c/runner.py:174-248 with a recorded env)
# --------------------------------------------------------------------------

def rollout_from_trajectory(traj, fs, gamma, values=None, actions=None):
    """Builds the post-``Runner.rollout`` ``datas`` dict from a recorded
    trajectory (see a2c_b200/synth.py).  Pong rule: stored done = env done or
    reward != 0 (:213-214), but only env done resets the stack (:212,227)."""
    frames, reset, rew = traj["frames"], traj["reset"], traj["rewards"]
    R, T = rew.shape
    states = prep_and_stack(frames, reset, fs)               # [R, T+1, fs, 80, 80]
    env_done = reset[:, 1:].astype(bool)
    done = (env_done | (rew != 0)).astype(np.float32)
    vals = traj["values"] if values is None else values
    r2, d2, deltas = rollout_bookkeeping(rew, done, vals, gamma)
    datas = {
        "states": torch.from_numpy(states[:, :T].reshape(R * T, fs, 80, 80)).float(),  # :199
        "rewards": r2.reshape(-1), "dones": d2.reshape(-1), "deltas": deltas.reshape(-1),
    }
    if actions is not None:
        datas["actions"] = torch.as_tensor(actions).reshape(-1).long()
    return datas, states


class RunnerOracle:
    """a2c/runner.py:147-248 for R rollout slots at once, against a recorded
    trajectory.  Per step: store state (:199) and h (:202), forward (:204-206),
    sample with the recorded uniform (:207), read reward / env-done, Pong
    pseudo-done (:212-214), zero h on done (:218-220), push the next frame
    through the stack with reset = env-done (:227).  Deltas, bootstrap and the
    forced dones[end]=1 follow ``rollout_bookkeeping``.  State and h carry over
    to the next rollout (:246-248).  The only intended difference from the
    reference is batching: forwards run on R states at once instead of one."""

    def __init__(self, model, fs, gamma, h_size):
        self.model, self.fs, self.gamma, self.h_size = model, fs, gamma, h_size
        self.stacks = None
        self.state = None
        self.h = None

    def rollout(self, p, traj):
        frames, reset, rew, u = traj["frames"], traj["reset"], traj["rewards"], traj["u"]
        R, T = rew.shape
        fs = self.fs
        if self.stacks is None:
            self.stacks = [FrameStack(fs) for _ in range(R)]
            self.state = np.stack([self.stacks[r].push(pong_prep(frames[r, 0]), True)
                                   for r in range(R)])
            if is_recurrent(self.model):
                self.h = torch.zeros(R, self.h_size)
        states = np.zeros((R, T, fs, 80, 80), dtype=np.uint8)
        h_states = torch.zeros(R, T, self.h_size) if self.h is not None else None
        vals = torch.zeros(R, T + 1)
        actions = torch.zeros(R, T, dtype=torch.int64)
        dones = np.zeros((R, T), dtype=np.float32)
        with torch.no_grad():
            for t in range(T):
                states[:, t] = self.state
                x = torch.from_numpy(self.state).float()
                if self.h is not None:
                    h_states[:, t] = self.h
                    v, logits, self.h = model_forward(self.model, p, x, self.h)
                else:
                    v, logits = model_forward(self.model, p, x)
                vals[:, t] = v.reshape(-1)
                actions[:, t] = get_action(logits, torch.from_numpy(u[:, t]))
                env_done = reset[:, t + 1] != 0
                done = env_done | (rew[:, t] != 0)
                dones[:, t] = done
                if self.h is not None:
                    self.h = self.h * torch.from_numpy(~done).float()[:, None]
                self.state = np.stack([
                    self.stacks[r].push(pong_prep(frames[r, t + 1]), bool(env_done[r]))
                    for r in range(R)])
            x = torch.from_numpy(self.state).float()
            if self.h is not None:
                vals[:, T] = model_forward(self.model, p, x, self.h)[0].reshape(-1)
            else:
                vals[:, T] = model_forward(self.model, p, x)[0].reshape(-1)
        r2, d2, deltas = rollout_bookkeeping(rew, dones, vals, self.gamma)
        datas = {"states": torch.from_numpy(states.reshape(R * T, fs, 80, 80)).float(),
                 "rewards": r2.reshape(-1), "dones": d2.reshape(-1),
                 "deltas": deltas.reshape(-1), "actions": actions.reshape(-1)}
        if h_states is not None:
            datas["h_states"] = h_states.reshape(R * T, self.h_size)
        return datas

"""Generate tests/golden/*.npz from the UNMODIFIED reference at /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py [small] [c2] [c3] [c4]

Inputs are synthetic and seed-derived (a2c_b200/synth.py) so fixtures store
outputs only.  Weights are ``synth.hashed_state_dict`` (a pure function of the
parameter index) loaded into the reference modules with load_state_dict.
Recorded environment: a fake env object exposing reset/step/get_action/render,
the only members a2c/runner.py:207-210,227 and a2c/utils.py:38 touch.
"""
import os
import sys
import time
from collections import OrderedDict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.join(ROOT, "pytorch-a2c_b200"))
sys.path.insert(0, ROOT)

import _ref_loader  # noqa: E402
from a2c_b200 import synth  # noqa: E402
from oracle import a2c_oracle as O  # noqa: E402  (only for param name lists / bookkeeping of full-size inputs)

ref = _ref_loader.load_reference()

BASE_HYPS = dict(gamma=0.99, lambda_=0.98, entr_coef=0.005, val_coef=0.5, pi_coef=1.0,
                 max_norm=0.5, lr=1e-4, optim_type="RMSprop", norm_advs=True,
                 use_nstep_rets=False, use_bptt=False, is_discrete=True,
                 env_type="Pong-v0", action_shift=1, render=False, h_size=256)


class RecordedEnv:
    """Plays back a synth trajectory slot through the reference's own prep_fxn."""

    def __init__(self, frames, reset, rewards, u):
        self.frames, self.reset_flags, self.rewards, self.u = frames, reset, rewards, u
        self.t = 0          # index of the frame that will be returned next
        self.actions = []

    def _obs(self):
        f = ref.preprocessing.pong_prep(self.frames[self.t].copy())
        self.t += 1
        return f

    def reset(self):
        return self._obs()

    def step(self, action):
        k = self.t - 1      # step index
        rew = float(self.rewards[k])
        done = bool(self.reset_flags[k + 1])
        if done:
            # next_state(reset=True) discards this obs and calls env.reset() (utils.py:37-38);
            # do not consume a frame for the discarded observation
            return np.zeros((1, 80, 80), dtype=np.uint8), rew, done, {}
        return self._obs(), rew, done, {}

    def get_action(self, logits):
        # SequentialEnvironment.get_action (runner.py:94-97) with the recorded uniform
        probs = torch.nn.functional.softmax(logits, dim=-1)
        k = self.t - 1
        st = torch.get_rng_state()
        # make torch.rand(1) inside sample_action return exactly u[k]
        with _forced_rand(float(self.u[k])):
            a = ref.utils.sample_action(probs.data)
        torch.set_rng_state(st)
        return int(a.item())

    def render(self):
        pass


class _forced_rand:
    def __init__(self, val):
        self.val = val

    def __enter__(self):
        self.orig = torch.rand
        val = self.val
        torch.rand = lambda *shape: torch.full(shape, val, dtype=torch.float32)

    def __exit__(self, *a):
        torch.rand = self.orig


class _Q:
    """size-1 queue stand-in for rew_q (runner.py:216)."""

    def __init__(self, v):
        self.v = v

    def get(self):
        return self.v

    def put(self, v):
        self.v = v


def build_ref_model(name, fs, h, A=3):
    cls = getattr(ref.models, name)
    net = cls([fs, 80, 80], A, h_size=h, bnorm=False, is_discrete=True)
    shapes = O.param_shapes(name, fs=fs, h=h, A=A)
    sd = synth.hashed_state_dict(shapes)
    full = net.state_dict()
    for k in full:
        base = k
        # aliases: features.* / conv1.* / conv2.* mirror convs.*
        if k.startswith("features."):
            base = "convs." + k[len("features."):]
        elif k.startswith("conv1."):
            base = "convs.0." + k[len("conv1."):]
        elif k.startswith("conv2."):
            base = "convs.1." + k[len("conv2."):]
        if base in sd:
            full[k] = torch.from_numpy(sd[base])
    net.load_state_dict(full)
    return net, shapes


def capture_update(updater, shared):
    """Runs the reference update_model and captures pre-clip grads through a
    wrapper around torch's clip_grad_norm_ (the reference file is untouched)."""
    cap = {}
    orig = torch.nn.utils.clip_grad_norm_

    def wrapped(params, max_norm, *a, **k):
        params = list(params)
        cap["grads"] = [None if p.grad is None else p.grad.detach().clone() for p in params]
        out = orig(params, max_norm, *a, **k)
        cap["clipped"] = [None if p.grad is None else p.grad.detach().clone() for p in params]
        return out

    torch.nn.utils.clip_grad_norm_ = wrapped
    try:
        info = updater.update_model(shared)
    finally:
        torch.nn.utils.clip_grad_norm_ = orig
    return info, cap


def named_unique_params(net, shapes):
    sd = dict(net.named_parameters())
    return OrderedDict((k, sd[k]) for k in shapes)


def run_runner_and_update(model, R, T, fs, h, seed, use_bptt=False, extra=None, n_updates=1):
    hyps = dict(BASE_HYPS)
    hyps.update(n_tsteps=T, n_rollouts=R, n_envs=R, n_frame_stack=fs, h_size=h,
                use_bptt=use_bptt, model=model)
    if extra:
        hyps.update(extra)
    torch.manual_seed(0)
    net, shapes = build_ref_model(model, fs, h)
    N = R * T
    datas = {"states": torch.zeros(N, fs, 80, 80), "deltas": torch.zeros(N),
             "rewards": torch.zeros(N), "dones": torch.zeros(N),
             "actions": torch.zeros(N).long()}
    if net.is_recurrent:
        datas["h_states"] = torch.zeros(N, h)
    out = {}
    trajs = [synth.make_trajectory(R, T, seed=seed + 17 * k, first_rollout=(k == 0))
             for k in range(n_updates)]
    runners = []
    rew_q = _Q(-1)
    for r in range(R):
        runner = ref.runner.Runner(datas, hyps, None, None, rew_q)
        runner.net = net
        runners.append(runner)
    updater = ref.updater.Updater(net, hyps)
    boot_vals = np.zeros((n_updates, R), dtype=np.float32)
    for k in range(n_updates):
        traj = trajs[k]
        for r in range(R):
            runner = runners[r]
            if k == 0:
                env = RecordedEnv(traj["frames"][r], traj["reset"][r], traj["rewards"][r],
                                  traj["u"][r])
                runner.env = env
                # Runner.run prologue (runner.py:160-168)
                runner.state_bookmark = ref.utils.next_state(env, runner.obs_deque, obs=None,
                                                             reset=True)
                runner.h_bookmark = torch.zeros(1, h) if net.is_recurrent else None
                runner.ep_rew = 0
            else:
                env = runner.env
                # continue the same env: frame 0 of this trajectory is the bookmark frame
                env.frames, env.reset_flags = traj["frames"][r], traj["reset"][r]
                env.rewards, env.u = traj["rewards"][r], traj["u"][r]
                env.t = 1
            for p in net.parameters():
                p.requires_grad = False
            runner.rollout(net, r, hyps)
        pre = "u%d_" % k
        for key, v in datas.items():
            if key == "states":
                out[pre + "states_bits"] = np.packbits(v.numpy().astype(np.uint8))
            else:
                out[pre + key] = v.numpy().copy()
        info, cap = capture_update(updater, datas)
        for key in ("Loss", "Pi_Loss", "ValLoss", "Entropy"):
            out[pre + key] = np.float64(info[key])
        out[pre + "GradNorm"] = np.float64(float(info["GradNorm"]))
        params = named_unique_params(net, shapes)
        plist = list(net.parameters())
        for name, p in params.items():
            idx = [i for i, q in enumerate(plist) if q is p][0]
            g = cap["grads"][idx]
            if g is not None:
                out[pre + "grad/" + name] = _compact(g.numpy())
            out[pre + "param_after/" + name] = _compact(p.detach().numpy())
    out["hyps_json"] = np.array(_json(hyps))
    out["seed"] = np.int64(seed)
    out["n_updates"] = np.int64(n_updates)
    return out


SAMPLE = 4096


def _compact(a):
    """Small tensors whole; big ones as (norm, first SAMPLE strided samples)."""
    a = np.asarray(a, dtype=np.float32).reshape(-1)
    if a.size <= SAMPLE:
        return a.copy()
    stride = a.size // SAMPLE
    return np.concatenate([[np.float32(np.sqrt((a.astype(np.float64) ** 2).sum()))],
                           a[::stride][:SAMPLE]]).astype(np.float32)


def _json(h):
    import json
    return json.dumps({k: v for k, v in h.items() if isinstance(v, (int, float, str, bool))})


def gen_small():
    out = {}
    # ---- pong_prep / next_state -------------------------------------------
    rng = np.random.default_rng(7)
    frames = synth.make_frames(rng, 6)
    frames[0][:] = 144
    frames[1][:] = 109
    frames[2][:] = 0
    preps = np.stack([ref.preprocessing.pong_prep(f.copy()) for f in frames])
    out["prep_frames_seed"] = np.int64(7)
    out["prep_out_bits"] = np.packbits(preps)
    # mutation check (the reference writes through the view)
    f = frames[3].copy()
    ref.preprocessing.pong_prep(f)
    out["prep_mutates_input"] = np.bool_(not np.array_equal(f, frames[3]))

    class E:
        def __init__(s):
            s.k = 100

        def reset(s):
            s.k += 1
            return np.full((1, 80, 80), s.k % 2, dtype=np.uint8)
    from collections import deque
    for fs in (2, 3):
        dq = deque(maxlen=fs)
        env = E()
        seq = []
        resets = [True, False, False, True, False, False, False]
        for i, rs in enumerate(resets):
            obs = np.full((1, 80, 80), (i * 7 + 3) % 2, dtype=np.uint8)
            obs[0, i, i] = 1 - obs[0, i, i]
            st = ref.utils.next_state(env, dq, obs, rs)
            seq.append(np.asarray(st, dtype=np.uint8))
        out["stack_fs%d_bits" % fs] = np.packbits(np.stack(seq))
    # ---- sample_action -----------------------------------------------------
    g = torch.Generator().manual_seed(11)
    logits = torch.randn(4096, 3, generator=g) * 2
    pi = torch.softmax(logits, -1)
    u = torch.rand(4096, generator=g)
    # edge rows: u above the fp32 running sum -> -1 ; u == 0 ; exact ties
    pi[0] = torch.tensor([0.25, 0.25, 0.25]); u[0] = 0.9
    pi[1] = torch.tensor([0.5, 0.25, 0.25]); u[1] = 0.5
    pi[2] = torch.tensor([0.0, 0.0, 1.0]); u[2] = 0.0
    pi[3] = torch.tensor([0.1, 0.2, 0.7]); u[3] = 0.3
    orig = torch.rand
    torch.rand = lambda *shape: u.clone().reshape(shape)
    try:
        acts = ref.utils.sample_action(pi)
    finally:
        torch.rand = orig
    out["sa_pi"], out["sa_u"], out["sa_act"] = pi.numpy(), u.numpy(), acts.numpy()
    # ---- discount ------------------------------------------------------------
    out["disc_known"] = ref.utils.discount(torch.tensor([1., 1, 1]), torch.tensor([0., 0, 1]),
                                           .5).numpy()
    arr = torch.randn(600, generator=g)
    dn = (torch.rand(600, generator=g) < 0.1).float()
    out["disc_arr"], out["disc_dones"] = arr.numpy(), dn.numpy()
    out["disc_out_g99"] = ref.utils.discount(arr, dn, 0.99).numpy()
    out["disc_out_gl"] = ref.utils.discount(arr, dn, 0.99 * 0.98).numpy()
    np.savez_compressed(os.path.join(HERE, "units.npz"), **out)
    print("units.npz written")

    # ---- Runner.rollout + Updater.update_model, small configs -----------------
    t0 = time.time()
    c1 = run_runner_and_update("FCModel", R=4, T=15, fs=2, h=256, seed=1234, n_updates=2)
    np.savez_compressed(os.path.join(HERE, "c1_fc.npz"), **c1)
    print("c1_fc.npz", time.time() - t0)
    for tag, model, kw in (
            ("a3c", "A3CModel", {}),
            ("conv", "ConvModel", {}),
            ("gru", "GRUModel", {}),
            ("gru_bptt", "GRUModel", {"use_bptt": True}),
            ("fc_adam_nstep", "FCModel", {"extra": {"optim_type": "Adam", "use_nstep_rets": True,
                                                    "norm_returns": True}}),
            ("fc_fs3", "FCModel", {"fs": 3, "extra": {"norm_advs": False}})):
        t0 = time.time()
        fs = kw.pop("fs", 2)
        o = run_runner_and_update(model, R=3, T=6, fs=fs, h=64 if model != "ConvModel" else 32,
                                  seed=4321, n_updates=2, **kw)
        np.savez_compressed(os.path.join(HERE, "small_%s.npz" % tag), **o)
        print("small_%s.npz" % tag, time.time() - t0)


def full_inputs(model, R, T, fs, h, seed):
    """Synthetic shared_data at BASELINE sizes.  States come from the
    reference's own pong_prep + next_state; rewards'/dones'/deltas from the
    (golden-pinned) bookkeeping restatement on N(0,1) rollout values."""
    from collections import deque
    traj = synth.make_trajectory(R, T, seed=seed)
    N = R * T
    states = torch.zeros(N, fs, 80, 80)

    class E:
        pass
    for r in range(R):
        dq = deque(maxlen=fs)
        env = E()
        for t in range(T):
            fr = ref.preprocessing.pong_prep(traj["frames"][r, t].copy())
            env.reset = lambda fr=fr: fr
            st = ref.utils.next_state(env, dq, fr, bool(traj["reset"][r, t]))
            states[r * T + t] = torch.FloatTensor(st)
    done = ((traj["reset"][:, 1:] != 0) | (traj["rewards"] != 0)).astype(np.float32)
    r2, d2, deltas = O.rollout_bookkeeping(traj["rewards"], done, traj["values"], 0.99)
    rng = np.random.default_rng(seed + 1)
    actions = torch.from_numpy(rng.integers(0, 3, size=N)).long()
    shared = {"states": states, "rewards": r2.reshape(-1).clone(), "dones": d2.reshape(-1).clone(),
              "deltas": deltas.reshape(-1).clone(), "actions": actions}
    if model == "GRUModel":
        shared["h_states"] = torch.from_numpy(
            synth.hashed_tensor((N, h), 999, 0.5)).float()
    return shared


def gen_full(tag, model, R, T, fs=2, h=256, seed=1234, **extra):
    hyps = dict(BASE_HYPS)
    hyps.update(n_tsteps=T, n_rollouts=R, n_envs=R, n_frame_stack=fs, h_size=h, model=model)
    hyps.update(extra)
    t0 = time.time()
    shared = full_inputs(model, R, T, fs, h, seed)
    print(tag, "inputs", time.time() - t0)
    net, shapes = build_ref_model(model, fs, h)
    torch.set_num_threads(os.cpu_count())
    updater = ref.updater.Updater(net, hyps)
    t0 = time.time()
    info, cap = capture_update(updater, shared)
    dt = time.time() - t0
    print(tag, "update_model", dt, info)
    out = {"hyps_json": np.array(_json(hyps)), "seed": np.int64(seed),
           "ref_update_seconds": np.float64(dt), "ref_threads": np.int64(torch.get_num_threads())}
    for key in ("Loss", "Pi_Loss", "ValLoss", "Entropy"):
        out[key] = np.float64(info[key])
    out["GradNorm"] = np.float64(float(info["GradNorm"]))
    out["advs"] = _compact(ref.utils.discount(shared["deltas"], shared["dones"],
                                              0.99 * 0.98).numpy())
    out["returns"] = _compact(ref.utils.discount(shared["rewards"], shared["dones"], 0.99).numpy())
    out["states_sum"] = np.float64(shared["states"].double().sum().item())
    params = named_unique_params(net, shapes)
    plist = list(net.parameters())
    for name, p in params.items():
        idx = [i for i, q in enumerate(plist) if q is p][0]
        g = cap["grads"][idx]
        if g is not None:
            out["grad/" + name] = _compact(g.numpy())
        out["param_after/" + name] = _compact(p.detach().numpy())
    np.savez_compressed(os.path.join(HERE, "full_%s.npz" % tag), **out)
    print("full_%s.npz written" % tag)


if __name__ == "__main__":
    what = sys.argv[1:] or ["small"]
    if "small" in what:
        gen_small()
    if "c3" in what:
        gen_full("c3_a3c", "A3CModel", R=1024, T=20)
    if "c4" in what:
        gen_full("c4_gru_bptt", "GRUModel", R=512, T=32, use_bptt=True)
    if "c2" in what:
        gen_full("c2_conv", "ConvModel", R=256, T=20)

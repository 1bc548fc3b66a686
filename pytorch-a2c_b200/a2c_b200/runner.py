"""Batched, device-resident Runner with the reference's constructor and method
names (a2c/runner.py:109-248).

The reference runs one OS process per environment, each doing batch-1 forwards
and per-step host<->device copies.  Here one Runner per GPU owns all of its
rank's R rollout slots: per timestep one stage-1 launch preps + stacks the R new
frames into the plane ring, one batched forward + one sampling kernel pick the R
actions, and at the end one stage-2 launch produces the post-rollout
rewards / dones / deltas (bootstrap folded, dones[end] = 1) exactly as
``Runner.rollout`` leaves them in ``datas``.

The environment side is a vectorised protocol (gym itself is out of scope):

    env.reset()            -> frames u8 [R, 210, 160, 3]
    env.step(actions i64 [R])
                           -> frames u8 [R, 210, 160, 3], rewards f32 [R], dones bool/u8 [R]
                              (for a done env the returned frame is the reset observation,
                               as next_state(reset=True) -> env.reset() in a2c/utils.py:37-38)
    env.uniforms(t)        -> optional, f32 [R] uniform draws for sampling (else torch.rand)

``RecordedVecEnv`` plays back a synthetic trajectory (a2c_b200/synth.py).
"""
import numpy as np
import torch

from . import ops
from .utils import try_key


class RecordedVecEnv:
    """Vector env over a recorded trajectory; frames may live on the device or
    in pinned host memory (then every step pays its H2D copy, like a real env)."""

    def __init__(self, traj, device, frames_on_device=True):
        self.device = device
        self.R, self.T = traj["rewards"].shape
        fr = torch.from_numpy(traj["frames"])
        self.frames = fr.to(device) if frames_on_device else fr.pin_memory()
        self.reset_flags = torch.from_numpy(traj["reset"]).to(device)
        self.rewards = torch.from_numpy(traj["rewards"]).to(device)
        self.u = torch.from_numpy(traj["u"]).to(device)
        self.t = 0

    def load(self, traj):
        """Continue with the next recorded rollout (its frame 0 is the bookmark)."""
        self.__init__(traj, self.device, self.frames.is_cuda)

    def reset(self):
        return self.frames[:, 0]

    def step(self, actions):
        t = self.t
        self.t += 1
        return self.frames[:, t + 1], self.rewards[:, t], self.reset_flags[:, t + 1]

    def uniforms(self, t):
        return self.u[:, t].contiguous()


class Runner:
    def __init__(self, datas, hyps, gate_q, stop_q, rew_q):
        self.hyps = hyps
        self.datas = datas
        self.gate_q = gate_q
        self.stop_q = stop_q
        self.rew_q = rew_q
        self.env = None
        self.ring = None
        self.h_bookmark = None
        self.ep_rew = None
        self._stage_frames = None
        self._stage_reset = None

    # ---- set-up -------------------------------------------------------------
    def attach(self, net, env):
        """Device-side equivalent of the prologue of Runner.run (a2c/runner.py:157-168)."""
        hyps = self.hyps
        self.net = net
        self.env = env
        dev = next(net.parameters()).device
        self.device = dev
        R, T, fs = env.R, hyps['n_tsteps'], hyps['n_frame_stack']
        self.R, self.T = R, T
        self.ring = ops.PlaneRing(R, T, fs, dev)
        frames0 = env.reset()
        ones = torch.ones((R, 1), dtype=torch.uint8, device=dev)
        self._push(frames0, ones, 0)
        if net.is_recurrent:
            self.h_bookmark = torch.zeros(R, net.h_size, device=dev)
        self.ep_rew = torch.zeros(R, device=dev)
        self.values = torch.zeros((R, T + 1), dtype=torch.float32, device=dev)
        for p in net.parameters():
            p.requires_grad = False

    def _push(self, frames, reset, index):
        R = self.R
        if frames.is_cuda:
            fr = frames.reshape(R, 1, 210, 160, 3)
            if not fr.is_contiguous():
                fr = fr.contiguous()
            self.ring.push_frames(fr, reset.to(torch.uint8).reshape(R, 1).contiguous(), index)
        else:
            # host frames: staged through ONE persistent pinned buffer (allocated once), which the
            # stage-1 kernel reads zero-copy; the _host entry synchronises before returning, so the
            # buffer is free again when the next step's frames arrive
            fr = frames.reshape(R, 1, 210, 160, 3)
            if not (fr.is_contiguous() and fr.is_pinned()):
                if self._stage_frames is None:
                    self._stage_frames = torch.empty((R, 1, 210, 160, 3), dtype=torch.uint8).pin_memory()
                self._stage_frames.copy_(fr)
                fr = self._stage_frames
            if self._stage_reset is None:
                self._stage_reset = torch.empty((R, 1), dtype=torch.uint8).pin_memory()
            self._stage_reset.copy_(reset.to(torch.uint8).reshape(R, 1))
            self.ring.push_frames(fr, self._stage_reset, index)

    def run(self, net):
        """a2c/runner.py:147-172: serve rollouts until the process is terminated."""
        if self.env is None:
            raise RuntimeError("call attach(net, vec_env) first: gym environments are out of "
                               "scope, the Runner steps a vectorised env (see module docstring)")
        while True:
            idx = self.gate_q.get()
            self.rollout(self.net, idx, self.hyps)
            self.stop_q.put(idx)

    # ---- the rollout ----------------------------------------------------------
    @torch.no_grad()
    def rollout(self, net, idx, hyps):
        """Fills ``datas`` for all R slots of this runner (slot r occupies
        [(idx*R + r)*T, (idx*R + r + 1)*T), i.e. idx selects a block of R slots)."""
        R, T, dev = self.R, self.T, self.device
        idx = 0 if idx is None else idx
        lo, hi = idx * R * T, (idx + 1) * R * T
        datas = self.datas
        rew_v = datas['rewards'][lo:hi].view(R, T)
        done_v = datas['dones'][lo:hi].view(R, T)
        act_v = datas['actions'][lo:hi].view(R, T)
        h = self.h_bookmark
        if getattr(self, "_needs_carry", False):
            self.ring.carry()
            self._needs_carry = False
        pong = "Pong" in hyps['env_type']
        actions_dev = torch.empty((R, T), dtype=torch.int64, device=dev)
        for t in range(T):
            view = self.ring.state_at(t)
            if h is not None:
                datas['h_states'][lo:hi].view(R, T, -1)[:, t] = h            # runner.py:202
                val, logits, h = net(view, h)
            else:
                val, logits = net(view)
            self.values[:, t] = val.reshape(-1)
            u = self.env.uniforms(t) if hasattr(self.env, "uniforms") else torch.rand(R, device=dev)
            action = ops.sample_action(logits, u)                            # runner.py:207
            frames, rew, done = self.env.step(action + hyps['action_shift'])
            rew = torch.as_tensor(rew, dtype=torch.float32, device=dev)
            reset = torch.as_tensor(done, device=dev).bool()
            done = reset | (rew != 0) if pong else reset                     # runner.py:212-214
            if h is not None:
                h = h * (~done).float()[:, None]                             # runner.py:218-220
            rew_v[:, t] = rew
            done_v[:, t] = done.float()
            actions_dev[:, t] = action
            self._push(frames, reset, t + 1)                                 # runner.py:227
        # bootstrap value V(s_T) (runner.py:236-241)
        view = self.ring.state_at(T)
        if h is not None:
            self.values[:, T] = net(view, h)[0].reshape(-1)
        else:
            self.values[:, T] = net(view)[0].reshape(-1)
        act_v.copy_(actions_dev)
        # raw per-step rewards / done flags for the episode-reward EMA (runner.py:211-217): taken
        # BEFORE stage 2 folds the bootstrap into rewards[end] and forces dones[end] = 1
        raw = torch.stack([rew_v.to(dev, torch.float32), done_v.to(dev, torch.float32)]) \
            if self.rew_q is not None else None
        # stage 2: deltas, bootstrap fold, dones[end] = 1 (runner.py:228-245)
        r_c, d_c = rew_v.reshape(-1), done_v.reshape(-1)
        contiguous_slot = r_c.data_ptr() == rew_v.data_ptr() and r_c.is_cuda
        if not contiguous_slot:
            r_c, d_c = rew_v.to(dev).reshape(-1).contiguous(), done_v.to(dev).reshape(-1).contiguous()
        deltas, advs, rets, stats = ops.delta_scan(r_c, d_c, self.values, hyps['gamma'],
                                                   hyps['lambda_'])
        if not contiguous_slot:
            rew_v.copy_(r_c.view(R, T))
            done_v.copy_(d_c.view(R, T))
        datas['deltas'][lo:hi] = deltas
        # stage-2 by-products: the Updater uses them (instead of re-running both scans) only
        # when they cover exactly the buffer it is given (it checks n, gamma, lambda_)
        datas['scan'] = {'n': R * T, 'advs': advs, 'returns': rets, 'stats': stats,
                         'gamma': hyps['gamma'], 'lambda_': hyps['lambda_'], 'block': (lo, hi)}
        # states: the ring view is what our Updater consumes; the fp32 tensor of
        # a2c/runner.py:199 is materialised only when the caller provided one
        self.datas_view = self.ring.view(T)
        if try_key(hyps, 'keep_ring_view', True):
            datas['states_view'] = self.datas_view
        if 'states' in datas and torch.is_tensor(datas['states']):
            st = self.datas_view.materialize(torch.float32)
            datas['states'][lo:hi] = st
        self._reward_ema(raw)
        self._needs_carry = True       # bookmark (runner.py:246) applied lazily: the ring view
        if h is not None:              # handed to the Updater must stay intact until then
            self.h_bookmark = h

    def _reward_ema(self, raw):
        """runner.py:211-217: ``ep_rew`` accumulates the RAW step rewards of every env and, at
        each done (true done or Pong point), ``rew_q`` <- .99*rew_q + .01*ep_rew and ep_rew <- 0.
        The reference's env processes interleave arbitrarily; here events are applied time-major
        (step t of every env before step t+1, envs in index order).  One D2H copy per rollout;
        the EMA over the K events of the rollout is evaluated in closed form."""
        if self.rew_q is None or raw is None:
            return
        rew, done = raw.cpu().double().numpy()                  # [R, T] each
        ep = self.ep_rew.cpu().double().numpy().copy()
        events = []
        for t in range(rew.shape[1]):
            ep += rew[:, t]
            hit = done[:, t] != 0
            if hit.any():
                events.append(ep[hit].copy())
                ep[hit] = 0.0
        ema = self.rew_q.get()
        if events:
            x = np.concatenate(events)
            k = len(x)
            w = 0.99 ** np.arange(k - 1, -1, -1, dtype=np.float64)
            ema = float(0.99 ** k * float(ema) + 0.01 * float((w * x).sum()))
        self.rew_q.put(ema)
        self.ep_rew = torch.from_numpy(ep).float().to(self.device)


def make_shared_data(R, T, fs, device, h_size=None, with_states=True):
    """Rollout buffers with the reference's keys / dtypes / layout
    (a2c/training.py:87-101), resident on the device."""
    N = R * T
    d = {'deltas': torch.zeros(N, device=device), 'rewards': torch.zeros(N, device=device),
         'actions': torch.zeros(N, dtype=torch.int64, device=device),
         'dones': torch.zeros(N, device=device)}
    if with_states:
        d['states'] = torch.zeros((N, fs, 80, 80), device=device)
    if h_size is not None:
        d['h_states'] = torch.zeros((N, h_size), device=device)
    return d

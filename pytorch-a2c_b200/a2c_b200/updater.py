"""Updater with the reference's interface (a2c/updater.py:9-229); the work is
done by the stage 2-4 kernels.

Differences from the reference are additive: ``compute_gradients`` (gradient
only, for parity tests), env-sharded data parallelism when torch.distributed is
initialised, micro-batching over rollout slots (the
reference sees all N transitions at once, a2c/updater.py:33-137; ConvModel keeps
~3.4 MB of activations per transition, so C5's 81,920 transitions are processed
in chunks whose gradients accumulate in the flat buffer), and
``shared_data['states']`` may also be a uint8 tensor or an ``ops.View`` of the
stage-1 plane ring.
"""
import torch
import torch.optim as optim

from . import ops, sharding
from .utils import try_key


class _FlatParams:
    """All parameters (and their grads / optimizer state) as views of flat fp32
    buffers, so clip + optimizer step is two kernels over contiguous memory."""

    def __init__(self, net):
        params = [p for p in net.parameters()]
        seen, uniq = set(), []
        for p in params:
            if id(p) not in seen:
                seen.add(id(p))
                uniq.append(p)
        self.params = uniq
        dev = uniq[0].device
        # 4-element alignment per tensor keeps every view 16-byte aligned
        self.offsets, total = [], 0
        for p in uniq:
            self.offsets.append(total)
            total += (p.numel() + 3) // 4 * 4
        self.numel = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        for p, off in zip(uniq, self.offsets):
            n = p.numel()
            self.flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            p.grad = self.grad[off:off + n].view(p.shape)
        self.state1 = torch.zeros(total, dtype=torch.float32, device=dev)
        self.state2 = None
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)

    def rebind_grads(self):
        for p, off in zip(self.params, self.offsets):
            g = self.grad[off:off + p.numel()].view(p.shape)
            if p.grad is None or p.grad.data_ptr() != g.data_ptr():
                if p.grad is not None:
                    g.add_(p.grad)      # something replaced our view (e.g. set_to_none) - keep its value
                p.grad = g

    def view_of(self, flat, i):
        p, off = self.params[i], self.offsets[i]
        return flat[off:off + p.numel()].view(p.shape)


class Updater:
    def __init__(self, net, hyps, data_parallel=True):
        """``data_parallel=False`` keeps this Updater single-process even when a
        torch.distributed group exists (used to check the sharded update against the
        full-batch one on the same hardware)."""
        self.net = net
        self.hyps = hyps
        self.data_parallel = data_parallel
        self.is_discrete = hyps['is_discrete']
        if not self.is_discrete:
            raise NotImplementedError("continuous actions are out of scope")
        dev = next(net.parameters()).device
        if dev.type != "cuda":
            raise ops._lib.A2CError("Updater needs the network on a CUDA device; there is no "
                                    "CPU fallback")
        self.device = dev
        self.flat = _FlatParams(net)
        self.optim = self.new_optim(hyps['lr'])
        self.info = {}
        self.norm = 0
        self.ret_mean = None
        self.ret_std = None
        self._step = 0
        self._count_cache = (None, None)
        self._graphs = {}               # update_from_rollout(graph=True): buffers -> captured CUDA graph
        self._chunk_cache = {}
        self.graph_replayed_kernels = 0  # kernels of ours launched through graph replays (not seen by a2c_launch_count)
        self.last_chunks = 1            # micro-batches of the last update
        self.last_allreduce = None      # {"bytes", "buckets", "overlapped_bytes"} of the last update

    # ---- distributed helpers (env-sharded ranks) ----------
    def _world(self):
        return sharding.world_size() if self.data_parallel else 1

    def _allreduce(self, t):
        return sharding.allreduce_sum_(t) if self.data_parallel else t

    def _n_global(self, n_local):
        """Transitions over all ranks; one tiny all-reduce, cached per local size."""
        if self._world() == 1:
            return n_local
        if self._count_cache[0] != n_local:
            self._count_cache = (n_local, sharding.global_count(n_local))
        return self._count_cache[1]

    # ---- forward ------------------------------------------------------------
    def _to_dev(self, t):
        if isinstance(t, ops.View) or t is None:
            return t
        return t if t.is_cuda else t.to(self.device, non_blocking=True)

    def _forward(self, states, h_states, dones, R):
        hyps, net = self.hyps, self.net
        if h_states is not None:
            if hyps['use_bptt']:
                return self.bptt(states, h_states, dones, n_rollouts=R)
            vals, logits, _ = net(states, h_states)
            return vals, logits
        return net(states)

    # ---- micro-batching -----------------------------------------------------------
    def _chunk_slots(self, R, T):
        """Rollout slots per micro-batch: ``hyps['update_chunk_slots']`` or as many as fit in
        ~45 % of the free device memory by the model's activation estimate."""
        c = try_key(self.hyps, 'update_chunk_slots', None)
        if c is None:
            c = self._chunk_cache.get((R, T))            # decided once per batch shape (no driver query per step)
            if c is None:
                est = getattr(self.net, 'activation_bytes_per_state', None)
                per = est() if callable(est) else 4 << 20
                free, _ = torch.cuda.mem_get_info(self.device)
                free += torch.cuda.memory_reserved(self.device) - torch.cuda.memory_allocated(self.device)
                c = self._chunk_cache[(R, T)] = int(0.45 * free // (per * T))
        return max(1, min(R, int(c)))

    @staticmethod
    def _slot_slice(x, r0, r1, T):
        if x is None:
            return None
        if isinstance(x, ops.View):
            return x.slots(r0, r1)
        return x[r0 * T:r1 * T]

    def _big_bucket(self):
        """(offset, numel) of the one parameter that dominates the flat buffer (ConvModel's
        resize_emb.0.weight: 98 % of 209 MB), or None.  Its gradient is final early in backward
        (it belongs to a late layer), so its all-reduce can run under the conv backward."""
        f = self.flat
        i = max(range(len(f.params)), key=lambda j: f.params[j].numel())
        n = f.params[i].numel()
        return (i, f.offsets[i], n) if n >= 0.5 * f.numel and n >= (1 << 22) else None

    def compute_gradients(self, shared_data):
        """update_model up to and including backward (a2c/updater.py:63-128).
        Leaves the (pre-clip) gradient in the flat buffer / p.grad and returns
        device tensors; nothing is synchronised."""
        hyps, net = self.hyps, self.net
        net.req_grads(True)
        self.flat.rebind_grads()
        rewards = self._to_dev(shared_data['rewards']).reshape(-1).float()
        dones = self._to_dev(shared_data['dones']).reshape(-1).float()
        actions = self._to_dev(shared_data['actions']).reshape(-1)
        deltas = self._to_dev(shared_data['deltas']).reshape(-1).float()
        N = deltas.numel()
        # the ring view is used only when it describes exactly these N transitions (a Runner
        # injects the view of ITS last block; buffers holding several blocks use 'states')
        view = shared_data.get('states_view')
        if isinstance(view, ops.View) and view.N == N:
            states = view
        else:
            states = self._to_dev(shared_data['states'])
            if states is None:
                raise ops._lib.A2CError("shared_data['states'] is missing and 'states_view' does not "
                                        "cover the %d transitions of this update" % N)
            n_states = states.N if isinstance(states, ops.View) else states.shape[0]
            if n_states != N:
                raise ops._lib.A2CError("states hold %d transitions, deltas %d" % (n_states, N))
        h_states = self._to_dev(shared_data['h_states']) if 'h_states' in shared_data else None
        T = hyps['n_tsteps']
        if N % T:
            raise ops._lib.A2CError("N=%d is not a multiple of n_tsteps=%d" % (N, T))
        R = N // T
        world = self._world()
        n_global = self._n_global(N)
        nstep = bool(hyps['use_nstep_rets'])

        # ---- stage 2: advantages / returns (a2c/updater.py:70-71,86-88) --------------------
        pre = shared_data.get('scan')          # Runner / stage-2 output for exactly these buffers
        if (isinstance(pre, dict) and pre.get('n') == N and pre['advs'].numel() == N
                and pre.get('gamma') == hyps['gamma'] and pre.get('lambda_') == hyps['lambda_']):
            advs, returns = pre['advs'], (None if nstep else pre['returns'])
            stats = pre['stats'].clone()
        else:
            stats = torch.zeros(4, dtype=torch.float64, device=self.device)
            advs, returns = ops.discount2(deltas, None if nstep else rewards, dones,
                                          hyps['gamma'] * hyps['lambda_'], hyps['gamma'], stats)

        chunk = self._chunk_slots(R, T)
        bounds = [(r0, min(R, r0 + chunk)) for r0 in range(0, R, chunk)]
        self.last_chunks = len(bounds)
        norm_ret = bool(try_key(hyps, 'norm_returns', False))

        ret_norm = None
        if norm_ret:
            # a2c/updater.py:89-96 (EMA of mean / unbiased std of the returns); rare path
            if nstep:
                # returns = advs + vals need every forward first: one extra no-grad pass
                with torch.no_grad():
                    v_all = torch.cat([self._forward(self._slot_slice(states, a, b, T),
                                                     self._slot_slice(h_states, a, b, T),
                                                     dones[a * T:b * T], b - a)[0].reshape(-1)
                                       for a, b in bounds])
                rets = advs + v_all
            else:
                rets = returns
            if world > 1:
                s = torch.stack([rets.double().sum(), (rets.double() ** 2).sum()])
                mean, std = sharding.mean_unbiased_std(self._allreduce(s), n_global)
            else:
                mean, std = rets.mean(), rets.std()
            if self.ret_mean is None:
                self.ret_mean, self.ret_std = mean.float(), std.float()
            else:
                self.ret_mean = 0.01 * mean.float() + 0.99 * self.ret_mean
                self.ret_std = 0.01 * std.float() + 0.99 * self.ret_std
            ret_norm = torch.stack([self.ret_mean, self.ret_std]).float().contiguous()
        if hyps['norm_advs']:
            self._allreduce(stats)

        # ---- stages 3 + 4 per micro-batch; gradients accumulate in the flat buffer -------------
        big = self._big_bucket() if world > 1 else None
        state = {"last": False, "work": None}
        hook = None
        if big is not None:
            i_big, off_big, n_big = big

            def _early_allreduce(param):
                # the big bucket's gradient is complete once the last micro-batch has accumulated
                # it: reduce it now, under the rest of the backward pass
                if state["last"] and state["work"] is None:
                    state["work"] = sharding.allreduce_sum_async(self.flat.grad[off_big:off_big + n_big])
            hook = self.flat.params[i_big].register_post_accumulate_grad_hook(_early_allreduce)
        out = None
        vals_all, logits_all = [], []
        try:
            for ci, (a, b) in enumerate(bounds):
                state["last"] = ci == len(bounds) - 1
                sl = slice(a * T, b * T)
                vals, logits = self._forward(self._slot_slice(states, a, b, T),
                                             self._slot_slice(h_states, a, b, T), dones[sl], b - a)
                dlogits, dvals, o = ops.loss_fwd_bwd(
                    logits.detach(), vals.detach(), actions[sl], advs[sl],
                    None if returns is None else returns[sl], stats[0:2], hyps,
                    n_global=n_global, ret_norm=ret_norm)
                torch.autograd.backward([logits, vals], [dlogits, dvals.view_as(vals)])
                out = o if out is None else out + o
                vals_all.append(vals.detach().reshape(-1))
                logits_all.append(logits.detach())
                del vals, logits, dlogits, dvals
        finally:
            if hook is not None:
                hook.remove()
        g = self.flat.grad
        if world > 1:
            if state["work"] is not None:
                # the rest of the flat buffer: the ranges before and after the big bucket
                self._allreduce(g[:off_big])
                self._allreduce(g[off_big + n_big:])
                state["work"].wait()
                self.last_allreduce = {"bytes": 4 * g.numel(), "buckets": 3, "overlapped_bytes": 4 * n_big}
            else:
                self._allreduce(g)
                self.last_allreduce = {"bytes": 4 * g.numel(), "buckets": 1, "overlapped_bytes": 0}
            self._allreduce(out)
        return {"loss_terms": out, "advs": advs, "returns": returns,
                "vals": vals_all[0] if len(vals_all) == 1 else torch.cat(vals_all),
                "logits": logits_all[0] if len(logits_all) == 1 else torch.cat(logits_all),
                "n_global": n_global, "chunks": len(bounds)}

    def update_model_async(self, shared_data):
        """update_model without the host synchronisation: enqueues the whole update and returns
        a device tensor f32 [5] = (Loss, Pi_Loss, ValLoss, Entropy, GradNorm).  The caller reads it
        when it needs the numbers (e.g. a non-blocking copy into pinned memory consumed a step
        later), so the host can run ahead of the GPU and a host-side stall does not idle it."""
        res = self.compute_gradients(shared_data)
        self._optim_step()
        return torch.cat([res["loss_terms"], self.flat.norm])

    # ---- rollout -> update in one call, optionally replayed as a CUDA graph -----------------------
    def update_from_rollout(self, states_view, rewards, dones, values, actions, h_states=None, graph=None):
        """Stage 2 + the whole update for a rollout whose buffers live at FIXED device addresses
        (the Runner's / a trainer's persistent buffers): rewards / dones f32 [R*T] (raw; fixed up in
        place as Runner.rollout leaves them, a2c/runner.py:228-245), values f32 [R, T+1], actions
        i64 [R*T], states_view an ops.View of the plane ring.  Returns the device tensor of
        update_model_async.

        graph=True (default: hyps['cuda_graph'], off): the first call with a given set of buffers
        runs eagerly (warm-up), the second is captured into a CUDA graph, later ones replay it -- one
        launch instead of 60 .. 400, which is what a small (C1) or launch-heavy (GRU BPTT) update
        needs on a busy host.  Restrictions: single process, one micro-batch, RMSprop (Adam's bias
        correction is a per-step host scalar), hyper-parameters frozen at capture time."""
        hyps = self.hyps
        if graph is None:
            graph = bool(try_key(hyps, 'cuda_graph', False))

        def body():
            deltas, advs, rets, stats = ops.delta_scan(rewards, dones, values, hyps['gamma'], hyps['lambda_'])
            shared = {"states_view": states_view, "states": None, "rewards": rewards, "dones": dones,
                      "deltas": deltas, "actions": actions,
                      "scan": {"n": deltas.numel(), "advs": advs, "returns": rets, "stats": stats,
                               "gamma": hyps['gamma'], "lambda_": hyps['lambda_']}}
            if h_states is not None:
                shared["h_states"] = h_states
            return self.update_model_async(shared)

        R_, T_ = states_view.R, states_view.T
        ok = (graph and self._world() == 1 and isinstance(self.optim, optim.RMSprop)
              and not try_key(hyps, 'norm_returns', False) and self._chunk_slots(R_, T_) >= R_)
        if not ok:
            return body()
        key = (states_view.base.data_ptr(), states_view.R, states_view.T, states_view.fs,
               rewards.data_ptr(), dones.data_ptr(), values.data_ptr(), actions.data_ptr(),
               None if h_states is None else h_states.data_ptr())
        ent = self._graphs.get(key)
        if ent is None:                                  # first sight of these buffers: eager warm-up
            self._graphs[key] = {"graph": None, "out": None, "keep": (states_view,)}
            return body()
        if ent["graph"] is None:                         # second call: capture (the capture also runs nothing,
            step0 = self._step                           # so replay it once to execute this update)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            n0 = ops._lib.launch_count()
            with torch.cuda.graph(g):
                ent["out"] = body()
            ent["kernels"] = ops._lib.launch_count() - n0      # kernels of liba2c_b200 recorded as graph nodes
            ent["graph"] = g
            self._step = step0                           # the capture pass advanced the host counters once:
            for st in self.optim.state.values():         # undo, the replay below counts
                if 'step' in st:
                    st['step'] -= 1
        ent["graph"].replay()
        self.graph_replayed_kernels += ent["kernels"]
        self._step += 1
        for st in self.optim.state.values():
            if 'step' in st:
                st['step'] += 1
        return ent["out"]

    def update_model(self, shared_data):
        """a2c/updater.py:33-137.  Returns {"Loss", "Pi_Loss", "ValLoss", "Entropy",
        "GradNorm"} (GradNorm a tensor, as in the reference)."""
        hyps = self.hyps
        terms = self.update_model_async(shared_data).tolist()     # one D2H sync per update
        self.norm = self.flat.norm.clone()[0]
        self.info = {"Loss": terms[0], "Pi_Loss": terms[1], "ValLoss": terms[2],
                     "Entropy": terms[3], "GradNorm": self.norm}
        return self.info

    def _optim_step(self):
        """clip_grad_norm_ + optimizer.step + zero_grad (a2c/updater.py:129-132)
        as two kernels over the flat buffers.  The torch optimizer object in
        ``self.optim`` owns the hyper-parameters and (as views) the state."""
        self._sync_optim_state()
        f, group = self.flat, self.optim.param_groups[0]
        self._step += 1
        f.sumsq.zero_()
        ops.grad_sumsq(f.grad, f.sumsq)
        kind = 0 if isinstance(self.optim, optim.RMSprop) else 1
        if kind == 0:
            ops.clip_optim_step(f.flat, f.grad, f.state1, None, f.sumsq, 0, group['lr'],
                                self.hyps['max_norm'], alpha=group['alpha'], eps=group['eps'],
                                norm_out=f.norm)
        else:
            b1, b2 = group['betas']
            ops.clip_optim_step(f.flat, f.grad, f.state1, f.state2, f.sumsq, 1, group['lr'],
                                self.hyps['max_norm'], step=self._step, eps=group['eps'],
                                beta1=b1, beta2=b2, norm_out=f.norm)
        for st in self.optim.state.values():
            if 'step' in st:
                st['step'] += 1

    def _sync_optim_state(self):
        """``optim.load_state_dict`` (the reference's resume path, a2c/training.py:126-127)
        replaces the per-parameter state tensors with fresh ones.  The fused step reads the
        flat buffers, so when a state tensor is no longer our view its values are copied in,
        the step counter is restored and the views are bound again."""
        f, opt = self.flat, self.optim
        key1 = 'square_avg' if isinstance(opt, optim.RMSprop) else 'exp_avg'
        stale = False
        for i, p in enumerate(f.params):
            st = opt.state.get(p)
            if not st or key1 not in st:
                stale = stale or bool(st)
                continue
            if st[key1].data_ptr() != f.view_of(f.state1, i).data_ptr():
                stale = True
                break
        if stale:
            steps = [float(st['step']) for st in opt.state.values() if 'step' in st]
            self._bind_state(opt)
            if steps:
                self._step = int(max(steps))

    # ---- reference API ----------------------------------------------------------
    def bptt(self, states, h_states, dones, n_rollouts=None):
        """a2c/updater.py:139-169.  ``n_rollouts`` defaults to the number of rollout slots in
        ``dones`` (the reference reads hyps['n_rollouts']; a rank or a micro-batch holds fewer)."""
        T = self.hyps['n_tsteps']
        R = dones.numel() // T if n_rollouts is None else n_rollouts
        h0 = h_states.view(R, T, -1)[:, 0].contiguous()
        return self.net.bptt_forward(states, h0, dones, R, T)

    def gae(self, rewards, values, next_vals, dones, gamma, lambda_):
        """a2c/updater.py:171-189."""
        deltas = rewards + gamma * next_vals * (1 - dones) - values
        return ops.discount(deltas.to(self.device), dones.to(self.device), gamma * lambda_)

    def print_statistics(self):
        nums = {k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in self.info.items()}
        print(" – ".join(k + ": " + str(round(v, 5)) for k, v in sorted(nums.items())))

    def log_statistics(self, log, T, reward, avg_action, best_avg_rew):
        nums = {k: (v.item() if isinstance(v, torch.Tensor) else v) for k, v in self.info.items()}
        arr = [k + ": " + str(round(v, 5)) if "ntropy" not in k else k + ": " + str(v)
               for k, v in nums.items()]
        arr += ["EpRew: " + str(reward), "AvgAction: " + str(avg_action),
                "BestRew:" + str(best_avg_rew)]
        log.write("Step:" + str(T) + " – " + " – ".join(arr) + '\n')
        log.flush()

    def save_model(self, net_file_name, optim_file_name):
        torch.save(self.net.state_dict(), net_file_name)
        if optim_file_name is not None:
            torch.save(self.optim.state_dict(), optim_file_name)

    def new_lr(self, new_lr):
        # Reference quirk kept on purpose (a2c/updater.py:221-224): loading the old
        # state_dict restores the old lr, so this does not change the learning rate.
        new_optim = self.new_optim(new_lr)
        new_optim.load_state_dict(self.optim.state_dict())
        self._bind_state(new_optim)
        self.optim = new_optim

    def new_optim(self, lr):
        cls = getattr(optim, self.hyps['optim_type'])
        if cls not in (optim.RMSprop, optim.Adam):
            raise NotImplementedError("fused optimizer step supports RMSprop and Adam (got %s)"
                                      % self.hyps['optim_type'])
        new_optim = cls(self.net.parameters(), lr=lr)
        if cls is optim.Adam and self.flat.state2 is None:
            self.flat.state2 = torch.zeros_like(self.flat.state1)
        self._bind_state(new_optim, fresh=True)
        return new_optim

    def _bind_state(self, opt, fresh=False):
        """Makes the torch optimizer's per-parameter state tensors views of the
        flat state buffers (same keys torch uses, so state_dict()s interchange)."""
        f = self.flat
        is_rms = isinstance(opt, optim.RMSprop)
        for i, p in enumerate(f.params):
            st = opt.state[p]
            s1 = f.view_of(f.state1, i)
            key1 = 'square_avg' if is_rms else 'exp_avg'
            if not fresh and key1 in st and st[key1].data_ptr() != s1.data_ptr():
                s1.copy_(st[key1])
            st[key1] = s1
            if not is_rms:
                s2 = f.view_of(f.state2, i)
                if not fresh and 'exp_avg_sq' in st and st['exp_avg_sq'].data_ptr() != s2.data_ptr():
                    s2.copy_(st['exp_avg_sq'])
                st['exp_avg_sq'] = s2
            if 'step' not in st:
                st['step'] = torch.tensor(float(self._step if hasattr(self, '_step') else 0))

"""Updater with the reference's interface (a2c/updater.py:9-229); the work is
done by the stage 2-4 kernels.

Differences from the reference are additive: ``compute_gradients`` (gradient
only, for parity tests), env-sharded data parallelism when torch.distributed is
initialised, and ``shared_data['states']`` may also be a
uint8 tensor or an ``ops.View`` of the stage-1 plane ring.
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
    def __init__(self, net, hyps):
        self.net = net
        self.hyps = hyps
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

    # ---- distributed helpers (env-sharded ranks) ----------
    _world = staticmethod(sharding.world_size)
    _allreduce = staticmethod(sharding.allreduce_sum_)

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

    def _forward(self, states, shared_data, dones):
        hyps, net = self.hyps, self.net
        if 'h_states' in shared_data:
            h_states = self._to_dev(shared_data['h_states'])
            if hyps['use_bptt']:
                return self.bptt(states, h_states, dones)
            vals, logits, _ = net(states, h_states)
            return vals, logits
        return net(states)

    def compute_gradients(self, shared_data):
        """update_model up to and including backward (a2c/updater.py:63-128).
        Leaves the (pre-clip) gradient in the flat buffer / p.grad and returns
        device tensors; nothing is synchronised."""
        hyps, net = self.hyps, self.net
        net.req_grads(True)
        self.flat.rebind_grads()
        states = self._to_dev(shared_data['states_view'] if 'states_view' in shared_data
                              else shared_data['states'])
        rewards = self._to_dev(shared_data['rewards']).reshape(-1).float()
        dones = self._to_dev(shared_data['dones']).reshape(-1).float()
        actions = self._to_dev(shared_data['actions']).reshape(-1)
        deltas = self._to_dev(shared_data['deltas']).reshape(-1).float()
        N = deltas.numel()
        world = self._world()
        n_global = self._n_global(N)

        stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        advs = ops.discount(deltas, dones, hyps['gamma'] * hyps['lambda_'], stats=stats[0:2])
        returns = None
        if not hyps['use_nstep_rets']:
            returns = ops.discount(rewards, dones, hyps['gamma'], stats=stats[2:4])

        vals, logits = self._forward(states, shared_data, dones)

        ret_norm = None
        if try_key(hyps, 'norm_returns', False):
            # a2c/updater.py:89-96 (EMA of mean / unbiased std); rare path, stats on host side
            rets = returns if returns is not None else advs + vals.detach().reshape(-1)
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

        dlogits, dvals, out = ops.loss_fwd_bwd(logits.detach(), vals.detach(), actions, advs,
                                               returns, stats[0:2], hyps, n_global=n_global,
                                               ret_norm=ret_norm)
        torch.autograd.backward([logits, vals], [dlogits, dvals.view_as(vals)])
        self._allreduce(self.flat.grad)
        self._allreduce(out)
        return {"loss_terms": out, "advs": advs, "returns": returns, "vals": vals.detach(),
                "logits": logits.detach(), "n_global": n_global}

    def update_model(self, shared_data):
        """a2c/updater.py:33-137.  Returns {"Loss", "Pi_Loss", "ValLoss", "Entropy",
        "GradNorm"} (GradNorm a tensor, as in the reference)."""
        hyps = self.hyps
        res = self.compute_gradients(shared_data)
        self._optim_step()
        terms = res["loss_terms"].tolist()           # one D2H sync per update
        self.norm = self.flat.norm.clone()[0]
        self.info = {"Loss": terms[0], "Pi_Loss": terms[1], "ValLoss": terms[2],
                     "Entropy": terms[3], "GradNorm": self.norm}
        return self.info

    def _optim_step(self):
        """clip_grad_norm_ + optimizer.step + zero_grad (a2c/updater.py:129-132)
        as two kernels over the flat buffers.  The torch optimizer object in
        ``self.optim`` owns the hyper-parameters and (as views) the state."""
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

    # ---- reference API ----------------------------------------------------------
    def bptt(self, states, h_states, dones):
        """a2c/updater.py:139-169."""
        hyps = self.hyps
        R, T = hyps['n_rollouts'], hyps['n_tsteps']
        w = self._world()
        if w > 1:
            R = dones.numel() // T
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
            if not fresh and key1 in st:
                s1.copy_(st[key1])
            st[key1] = s1
            if not is_rms:
                s2 = f.view_of(f.state2, i)
                if not fresh and 'exp_avg_sq' in st:
                    s2.copy_(st['exp_avg_sq'])
                st['exp_avg_sq'] = s2
            if 'step' not in st:
                st['step'] = torch.tensor(float(self._step if hasattr(self, '_step') else 0))

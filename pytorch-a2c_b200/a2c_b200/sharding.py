"""Host-side logic of the env-sharded multi-GPU update.

The reference is single-device: ``Updater.update_model`` sees all
``n_rollouts * n_tsteps`` transitions (a2c/updater.py:33-137).  Here every rank
owns a contiguous block of rollout slots; three rules make the per-rank
gradients sum to exactly the single-device gradient:

  1. advantage normalisation (a2c/updater.py:97-98) uses the GLOBAL mean and
     unbiased std, from all-reduced {sum, sum of squares} in float64;
  2. every ``.mean()`` of the loss (a2c/updater.py:101-127) divides by the
     global transition count (``a2c_loss_cfg.n_global``);
  3. the flat gradient buffer is all-reduced with SUM before clip + step, which
     then run identically on every rank.

Nothing here touches CUDA, so the same functions run under gloo on CPU
(tests/test_distributed.py) and under NCCL on the GPUs.
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def slot_range(n_slots, rank_=None, world=None):
    """Contiguous block [lo, hi) of the n_slots rollout slots owned by a rank;
    the first ``n_slots % world`` ranks hold one extra slot."""
    world = world_size() if world is None else world
    rank_ = rank() if rank_ is None else rank_
    if not 0 <= rank_ < world:
        raise ValueError("rank %d outside world of %d" % (rank_, world))
    base, extra = divmod(n_slots, world)
    lo = rank_ * base + min(rank_, extra)
    return lo, lo + base + (1 if rank_ < extra else 0)


def shard_rollout(shared_data, n_slots, n_tsteps, rank_=None, world=None):
    """This rank's part of a reference-layout rollout buffer dict
    (a2c/training.py:87-101: flat, rollout-major, slot r at [r*T, (r+1)*T))."""
    lo, hi = slot_range(n_slots, rank_, world)
    out = {}
    for k, v in shared_data.items():
        if torch.is_tensor(v) and v.dim() >= 1 and v.shape[0] == n_slots * n_tsteps:
            out[k] = v[lo * n_tsteps:hi * n_tsteps]
        else:
            out[k] = v
    return out


def _needs_host_staging(t):
    # gloo (CPU tests, or several ranks sharing one GPU in the -m gpu suite) reduces host tensors
    return t.is_cuda and dist.get_backend() == "gloo"


def allreduce_sum_(t):
    """In-place SUM over ranks; identity for a single process."""
    if world_size() > 1:
        if _needs_host_staging(t):
            h = t.detach().cpu()
            dist.all_reduce(h, op=dist.ReduceOp.SUM)
            t.copy_(h)
        else:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


class _Done:
    def wait(self):
        return True


def allreduce_sum_async(t):
    """Starts an in-place SUM over ranks and returns a handle with ``wait()``.  Under NCCL the
    collective is enqueued behind the work already on the current stream and runs on NCCL's own
    stream, i.e. concurrently with kernels launched afterwards; ``wait()`` makes the current
    stream wait for it (no host synchronisation)."""
    if world_size() == 1:
        return _Done()
    if _needs_host_staging(t):
        allreduce_sum_(t)
        return _Done()
    return dist.all_reduce(t, op=dist.ReduceOp.SUM, async_op=True)


def global_count(n_local):
    """Total transitions over all ranks (ranks may hold different slot counts)."""
    if world_size() == 1:
        return int(n_local)
    t = torch.tensor([float(n_local)], dtype=torch.float64)
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.all_reduce(t)
    return int(t.item())


def mean_unbiased_std(sums, n):
    """{sum, sum of squares} over n elements -> (mean, unbiased std), the
    statistics torch's ``x.mean(), x.std()`` give the reference."""
    mean = sums[0] / n
    var = (sums[1] - n * mean * mean) / max(n - 1, 1)
    return mean, var.clamp_min(0).sqrt()

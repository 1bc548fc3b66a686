"""Env-sharded update: the host logic in
a2c_b200/sharding.py run by two gloo processes on CPU.

Each rank evaluates the update on its own block of rollout slots with the
oracle's network (float64) under the contract of ``a2c_loss_fwd_bwd``
(global advantage statistics, every mean divided by ``n_global``) and
all-reduces the gradient; the result must equal the single-process oracle's
full-batch gradient (a2c/updater.py:33-128).  The CUDA side of the same
contract is covered by test_gpu_parity.py::test_sharded_loss_contract.
"""
import os
import socket
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import helpers as Hh                       # noqa: E402
from a2c_b200 import sharding              # noqa: E402
from oracle import a2c_oracle as O         # noqa: E402

MODEL, FS, HID, T = "FCModel", 2, 32, 6
TOL = 1e-9            # float64 on both sides: only the summation order differs


def _hyps(R):
    h = Hh.hyps_of(Hh.load("c1_fc.npz"))
    h.update(n_rollouts=R, n_tsteps=T, n_frame_stack=FS, h_size=HID, norm_advs=True,
             use_nstep_rets=False, norm_returns=False, use_bptt=False)
    return h


def _problem(R):
    _, shared = Hh.full_inputs(MODEL, R, T, FS, HID, seed=77)
    shared["states"] = shared.pop("states_u8").float()
    return shared, Hh.params_for(MODEL, FS, HID)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sharded_rank_update(shared, params, hyps):
    """One rank's part of the update under the sharded contract."""
    p = {k: v.double().clone().requires_grad_(True) for k, v in params.items()}
    dones = shared["dones"].reshape(-1)
    advs = O.discount(shared["deltas"], dones, hyps["gamma"] * hyps["lambda_"]).double()
    rets = O.discount(shared["rewards"], dones, hyps["gamma"]).double()
    n_global = sharding.global_count(advs.numel())
    stats = torch.stack([advs.sum(), (advs * advs).sum()])
    mean, std = sharding.mean_unbiased_std(sharding.allreduce_sum_(stats), n_global)
    advs = (advs - mean) / (std + 1e-6)
    vals, logits = O.model_forward(MODEL, p, shared["states"].double())
    lsm = F.log_softmax(logits, -1)
    log_ps = lsm[torch.arange(len(advs)), shared["actions"].reshape(-1).long()]
    pi = hyps["pi_coef"] * -(log_ps * advs).sum() / n_global
    entr = -hyps["entr_coef"] * (lsm * lsm.exp()).sum() / n_global
    val = hyps["val_coef"] * ((vals.reshape(-1) - rets) ** 2).sum() / n_global
    loss = pi + val - entr
    loss.backward()
    flat = torch.cat([q.grad.reshape(-1) for q in p.values()])
    terms = torch.stack([loss.detach(), pi.detach(), val.detach(), entr.detach()])
    sharding.allreduce_sum_(flat)
    sharding.allreduce_sum_(terms)
    return flat, terms, n_global


def _worker(rank, world, port, R, outdir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        shared, params = _problem(R)
        mine = sharding.shard_rollout(shared, R, T)
        lo, hi = sharding.slot_range(R)
        assert mine["deltas"].numel() == (hi - lo) * T
        flat, terms, n_global = _sharded_rank_update(mine, params, _hyps(R))
        np.savez(os.path.join(outdir, "rank%d.npz" % rank), flat=flat.numpy(),
                 terms=terms.numpy(), n_global=n_global, lo=lo, hi=hi)
    finally:
        dist.destroy_process_group()


def test_slot_range_partitions_every_slot():
    for n in (1, 2, 5, 256, 257):
        for world in (1, 2, 3, 8):
            blocks = [sharding.slot_range(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(blocks, blocks[1:]))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.slot_range(4, 2, 2)


def test_single_process_helpers_are_identity():
    t = torch.arange(4.0)
    assert sharding.world_size() == 1 and sharding.rank() == 0
    assert sharding.allreduce_sum_(t) is t and sharding.global_count(7) == 7
    x = torch.randn(1000, dtype=torch.float64)
    m, s = sharding.mean_unbiased_std(torch.stack([x.sum(), (x * x).sum()]), x.numel())
    assert abs(m - x.mean()) < 1e-12 and abs(s - x.std()) < 1e-12


@pytest.mark.parametrize("R", [4, 5])          # even and uneven slot split
def test_two_rank_sharded_update_equals_full_batch(R):
    world = 2
    shared, params = _problem(R)
    hyps = _hyps(R)
    full = O.UpdaterOracle(MODEL, params, hyps, dtype=torch.float64).compute(shared)
    ref_flat = torch.cat([g.reshape(-1) for g in full["grads"].values()]).numpy()
    ref_terms = np.array([full["Loss"], full["Pi_Loss"], full["ValLoss"], full["Entropy"]])

    with tempfile.TemporaryDirectory() as d:
        mp.spawn(_worker, args=(world, _free_port(), R, d), nprocs=world, join=True)
        out = [dict(np.load(os.path.join(d, "rank%d.npz" % r))) for r in range(world)]
    assert [int(o["lo"]) for o in out] == [0, (R + 1) // 2] and int(out[1]["hi"]) == R
    for o in out:
        assert int(o["n_global"]) == R * T
        assert Hh.rel_err(o["flat"], ref_flat) < TOL
        assert Hh.rel_err(o["terms"], ref_terms) < TOL
    # identical bits on every rank, so the redundant clip + optimizer steps stay in lock-step
    assert np.array_equal(out[0]["flat"], out[1]["flat"])

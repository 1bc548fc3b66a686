"""Round-2 GPU parity tests: micro-batched update, data-parallel update on hardware,
optimizer resume, reward EMA, the fused double scan, the pipelined host ingest."""
import os
import queue

import numpy as np
import pytest
import torch

import helpers as H
from a2c_b200 import synth
from oracle import a2c_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def build_net(model, fs, h, dev):
    from a2c_b200 import models
    net = models.get_model_class(model)([fs, 80, 80], 3, h_size=h, bnorm=False, is_discrete=True)
    net.load_state_dict(H.params_for(model, fs, h), strict=False)
    return net.to(dev)


def small_shared(model, R, T, fs, h, dev, seed=11):
    traj, shared = H.full_inputs(model, R, T, fs, h, seed)
    out = {"states": shared["states_u8"].to(dev), "rewards": shared["rewards"].to(dev),
           "dones": shared["dones"].to(dev), "deltas": shared["deltas"].to(dev),
           "actions": shared["actions"].to(dev)}
    if "h_states" in shared:
        out["h_states"] = shared["h_states"].to(dev)
    return out


HY = dict(gamma=0.99, lambda_=0.98, entr_coef=0.005, val_coef=0.5, pi_coef=1.0, max_norm=0.5,
          lr=1e-4, optim_type="RMSprop", norm_advs=True, use_nstep_rets=False, use_bptt=False,
          is_discrete=True, n_frame_stack=2)


# ------------------------------------------------------------------ stage 2 ----

@pytest.mark.parametrize("N", [1, 7, 60, 5120, 81920])
def test_discount2_bit_exact(dev, N):
    """Both scans of update_model (a2c/updater.py:70-71,86-88) in one launch."""
    from a2c_b200 import ops
    rng = np.random.default_rng(N)
    a = rng.standard_normal(N).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    d = (rng.random(N) < 0.15).astype(np.float32)
    stats = torch.zeros(4, dtype=torch.float64, device=dev)
    advs, rets = ops.discount2(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev),
                               torch.from_numpy(d).to(dev), 0.99 * 0.98, 0.99, stats)
    wa, wb = O.discount(a, d, 0.99 * 0.98), O.discount(b, d, 0.99)
    assert torch.equal(advs.cpu(), wa) and torch.equal(rets.cpu(), wb)
    want = [wa.double().sum(), (wa.double() ** 2).sum(), wb.double().sum(), (wb.double() ** 2).sum()]
    for i in range(4):
        assert abs(stats[i].item() - float(want[i])) <= 1e-9 * max(1.0, abs(float(want[i])))
    advs2, none = ops.discount2(torch.from_numpy(a).to(dev), None, torch.from_numpy(d).to(dev),
                                0.99 * 0.98, 0.99)
    assert none is None and torch.equal(advs2.cpu(), wa)


@pytest.mark.parametrize("R,T", [(1, 1), (3, 2), (64, 20), (65, 20), (512, 32), (4096, 20), (7, 400)])
def test_delta_scan_staged_and_direct_bit_exact(dev, R, T):
    """The shared-memory staged kernel (T small) and the direct fallback (T = 400) against
    a2c/runner.py:228-245 + a2c/utils.py:63-79."""
    from a2c_b200 import ops
    rng = np.random.default_rng(R * 1000 + T)
    rew = np.where(rng.random((R, T)) < 0.1, rng.integers(0, 2, (R, T)) * 2.0 - 1.0, 0.0).astype(np.float32)
    done = ((rng.random((R, T)) < 0.05) | (rew != 0)).astype(np.float32)
    vals = rng.standard_normal((R, T + 1)).astype(np.float32)
    r_d = torch.from_numpy(rew.reshape(-1).copy()).to(dev)
    d_d = torch.from_numpy(done.reshape(-1).copy()).to(dev)
    deltas, advs, rets, stats = ops.delta_scan(r_d, d_d, torch.from_numpy(vals).to(dev), 0.99, 0.98)
    r2, d2, wd = O.rollout_bookkeeping(rew, done, vals, 0.99)
    assert torch.equal(deltas.cpu(), wd.reshape(-1))
    assert torch.equal(r_d.cpu(), r2.reshape(-1)) and torch.equal(d_d.cpu(), d2.reshape(-1))
    assert float(d2[:, -1].min()) == 1.0
    if R * T <= 20000:
        wa = O.discount(wd, d2, 0.99 * 0.98)
        wr = O.discount(r2, d2, 0.99)
        assert torch.equal(advs.cpu(), wa) and torch.equal(rets.cpu(), wr)
        assert abs(stats[0].item() - float(wa.double().sum())) <= 1e-9 * max(1.0, abs(float(wa.double().sum())))


# ------------------------------------------------------------------ stage 1 ----

@pytest.mark.parametrize("layout", [0, 1, 2])
def test_pipelined_host_ingest_bit_exact(dev, layout):
    """FrameIngest (one async copy per rollout in any of its three staging layouts + stage 1 from
    the staging buffer) gives the same ring as stage 1 on device-resident frames, over three
    pipelined rollouts."""
    from a2c_b200 import ops
    R, T, fs = 5, 7, 2
    rng = np.random.default_rng(5)
    ring_a = ops.PlaneRing(R, T, fs, dev)
    ring_b = ops.PlaneRing(R, T, fs, dev)
    first = torch.from_numpy(synth.make_frames(rng, R)[:, None].copy())
    ones = torch.ones((R, 1), dtype=torch.uint8)
    ring_a.push_frames(first.to(dev), ones.to(dev))
    ring_b.push_frames(first.to(dev), ones.to(dev))
    ing = ops.FrameIngest(R, T, dev, depth=2, layout=layout)
    batches = []
    for k in range(3):
        fr = torch.from_numpy(synth.make_frames(rng, R * T).reshape(R, T, 210, 160, 3).copy())
        rs = torch.from_numpy((rng.random((R, T)) < 0.2).astype(np.uint8))
        batches.append((fr.pin_memory(), rs.pin_memory()))
    ing.prefetch(*batches[0])
    for k in range(3):
        if k > 0:
            ring_a.carry()
            ring_b.carry()
        ing.commit(ring_b, first_frame_index=1)
        if k + 1 < 3:
            ing.prefetch(*batches[k + 1])
        ring_a.push_frames(batches[k][0].to(dev), batches[k][1].to(dev), first_frame_index=1)
        assert torch.equal(ring_a.planes, ring_b.planes) and torch.equal(ring_a.pflag, ring_b.pflag)
        want = np.stack([O.pong_prep(f)[0] for f in batches[k][0][0].numpy()])
        assert np.array_equal(ring_b.planes[0, fs:fs + T].cpu().numpy().reshape(T, 80, 80), want)


@pytest.mark.parametrize("variant", [0, 2])
def test_obs_prep_variants_bit_exact(dev, variant):
    from a2c_b200 import ops
    rng = np.random.default_rng(31)
    R, Fn = 3, 301                     # more frames than CTAs x stages: the rings wrap
    frames = synth.make_frames(rng, R * Fn).reshape(R, Fn, 210, 160, 3)
    ring = ops.PlaneRing(R, Fn - 1, 1, dev)
    ring.push_frames(torch.from_numpy(frames).to(dev), torch.zeros((R, Fn), dtype=torch.uint8, device=dev),
                     variant=variant)
    got = ring.planes[:, :Fn].cpu().numpy().reshape(R * Fn, 80, 80)
    want = np.stack([O.pong_prep(f)[0] for f in frames.reshape(-1, 210, 160, 3)])
    assert np.array_equal(got, want)


# ------------------------------------------------------------ micro-batching ----

@pytest.mark.parametrize("model,bptt", [("ConvModel", False), ("FCModel", False), ("GRUModel", True),
                                        ("GRUModel", False), ("A3CModel", False)])
def test_micro_batched_update_equals_single_pass(dev, model, bptt):
    """Chunks of rollout slots accumulated in the flat gradient buffer == one pass over all N
    (the reference sees all N at once, a2c/updater.py:33-137)."""
    from a2c_b200 import updater
    R, T, fs, h = 6, 5, 2, 32
    hyps = dict(HY, n_tsteps=T, n_rollouts=R, use_bptt=bptt)
    shared = small_shared(model, R, T, fs, h, dev)
    res, grads = {}, {}
    for chunk in (R, 4, 1):
        net = build_net(model, fs, h, dev)
        upd = updater.Updater(net, dict(hyps, update_chunk_slots=chunk))
        r = upd.compute_gradients(shared)
        assert upd.last_chunks == -(-R // chunk)
        res[chunk] = r
        grads[chunk] = upd.flat.grad.clone()
    scale = float(grads[R].abs().max())
    for chunk in (4, 1):
        assert float((grads[chunk] - grads[R]).abs().max()) <= 2e-6 * scale, chunk
        assert torch.allclose(res[chunk]["loss_terms"], res[R]["loss_terms"], rtol=2e-6, atol=1e-7)
        assert torch.equal(res[chunk]["vals"], res[R]["vals"]) or \
            H.rel_err(res[chunk]["vals"].cpu().numpy(), res[R]["vals"].cpu().numpy()) < 1e-6
    # and the micro-batched gradient is the oracle's: within 2e-5 of the float64 evaluation, or
    # within 4x the fp32 oracle's own distance from it for the ill-conditioned first-layer sums
    cpu_shared = {k: (v.cpu().float() if k == "states" else v.cpu()) for k, v in shared.items()}
    want = O.UpdaterOracle(model, H.params_for(model, fs, h), hyps, dtype=torch.float64).compute(cpu_shared)
    w32 = O.UpdaterOracle(model, H.params_for(model, fs, h), hyps).compute(cpu_shared)
    net = build_net(model, fs, h, dev)
    upd = updater.Updater(net, dict(hyps, update_chunk_slots=2))
    upd.compute_gradients(shared)
    gs = max(g.abs().max().item() for g in want["grads"].values() if g is not None)
    for name, p in net.named_parameters():
        gw = want["grads"].get(name)
        if gw is None:
            continue
        den = max(gw.abs().max().item(), 1e-3 * gs)
        err = (p.grad.cpu().double() - gw).abs().max().item() / den
        ref_own = (w32["grads"][name].double() - gw).abs().max().item() / den
        assert err < max(2e-5, 4 * ref_own), (name, err, ref_own)


def test_stale_states_view_is_not_used(dev):
    """A ring view that does not cover the N transitions of the update is ignored."""
    from a2c_b200 import ops, updater
    R, T, fs, h = 4, 5, 2, 16
    hyps = dict(HY, n_tsteps=T, n_rollouts=R)
    shared = small_shared("FCModel", R, T, fs, h, dev)
    net = build_net("FCModel", fs, h, dev)
    upd = updater.Updater(net, hyps)
    g0 = upd.compute_gradients(shared)["loss_terms"].clone()
    upd.flat.grad.zero_()
    stale = ops.PlaneRing(2, T, fs, dev).view(T)          # 2 slots only, all-zero planes
    g1 = upd.compute_gradients(dict(shared, states_view=stale))["loss_terms"]
    assert torch.equal(g0, g1)
    with pytest.raises(Exception):
        upd.compute_gradients({k: v for k, v in dict(shared, states_view=stale).items() if k != "states"})


# --------------------------------------------------------- optimizer resume ----

@pytest.mark.parametrize("kind", ["RMSprop", "Adam"])
def test_optimizer_state_dict_resume(dev, kind, tmp_path):
    """a2c/training.py:126-127: ``updater.optim.load_state_dict(torch.load(file))`` on a fresh
    Updater must continue exactly like the Updater that saved it."""
    from a2c_b200 import updater
    R, T, fs, h = 4, 5, 2, 16
    hyps = dict(HY, n_tsteps=T, n_rollouts=R, optim_type=kind, lr=1e-3)
    shared = small_shared("FCModel", R, T, fs, h, dev)
    net_a = build_net("FCModel", fs, h, dev)
    upd_a = updater.Updater(net_a, hyps)
    for _ in range(3):
        upd_a.update_model(shared)
    upd_a.save_model(str(tmp_path / "net.p"), str(tmp_path / "optim.p"))
    # resume: fresh net + Updater, reference-style loads
    net_b = build_net("FCModel", fs, h, dev)
    net_b.load_state_dict(torch.load(str(tmp_path / "net.p")))
    upd_b = updater.Updater(net_b, hyps)
    upd_b.optim.load_state_dict(torch.load(str(tmp_path / "optim.p")))
    upd_b.optim.zero_grad()
    for _ in range(2):
        ia = upd_a.update_model(shared)
        ib = upd_b.update_model(shared)
        assert abs(ia["Loss"] - ib["Loss"]) <= 1e-6 * max(1.0, abs(ia["Loss"]))
    for (ka, pa), (kb, pb) in zip(net_a.named_parameters(), net_b.named_parameters()):
        assert float((pa - pb).abs().max()) <= 1e-6 * float(pa.abs().max()) + 1e-3 * hyps["lr"], ka
    assert H.rel_err(upd_b.flat.state1.cpu().numpy(), upd_a.flat.state1.cpu().numpy()) < 1e-5
    assert upd_a._step == upd_b._step == 5
    # and the reloaded state is what save_model writes next time
    sd = upd_b.optim.state_dict()
    key = "square_avg" if kind == "RMSprop" else "exp_avg"
    assert float(sd["state"][0][key].abs().sum()) > 0


# ------------------------------------------------------------- reward EMA ----

def test_runner_reward_ema_uses_raw_rewards(dev):
    """runner.py:211-217: rew_q is an EMA over finished-episode RAW rewards; the bootstrap fold
    and the forced dones[end] = 1 of stage 2 must not leak into it."""
    from a2c_b200 import runner as runner_mod
    R, T, fs, h = 5, 6, 2, 16
    hyps = dict(HY, n_tsteps=T, n_rollouts=R, env_type="Pong-v0", action_shift=1, render=False)
    net = build_net("FCModel", fs, h, dev)
    trajs = [synth.make_trajectory(R, T, seed=40 + k, first_rollout=(k == 0)) for k in range(2)]
    q = queue.Queue(1)
    q.put(-1.0)
    datas = runner_mod.make_shared_data(R, T, fs, dev, with_states=False)
    rn = runner_mod.Runner(datas, hyps, None, None, q)
    env = runner_mod.RecordedVecEnv(trajs[0], dev)
    rn.attach(net, env)
    ema, ep = -1.0, np.zeros(R)
    for k in range(2):
        if k > 0:
            env.load(trajs[k])
        rn.rollout(net, 0, hyps)
        rew, reset = trajs[k]["rewards"], trajs[k]["reset"]
        for t in range(T):
            ep += rew[:, t]
            for r in range(R):
                if reset[r, t + 1] != 0 or rew[r, t] != 0:       # true done or Pong point
                    ema = .99 * ema + .01 * ep[r]
                    ep[r] = 0
        got = q.get()
        q.put(got)
        assert abs(got - ema) < 1e-9, (k, got, ema)
        assert np.allclose(rn.ep_rew.cpu().numpy(), ep)


# --------------------------------------------- data parallel on the hardware ----

def _dp_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        for p in (root, os.path.join(root, "pytorch-a2c_b200"), os.path.join(root, "tests")):
            if p not in sys.path:
                sys.path.insert(0, p)
        from a2c_b200 import sharding, updater
        dev = torch.device("cuda:0")
        torch.cuda.set_device(dev)
        R, T, fs, h = 6, 5, 2, 32
        for model, bptt in (("ConvModel", False), ("GRUModel", True)):
            hyps = dict(HY, n_tsteps=T, n_rollouts=R, use_bptt=bptt, update_chunk_slots=2)
            full = small_shared(model, R, T, fs, h, dev)
            mine = sharding.shard_rollout(full, R, T, rank, world)
            net = build_net(model, fs, h, dev)
            upd = updater.Updater(net, hyps)
            info = upd.update_model(mine)
            net1 = build_net(model, fs, h, dev)
            upd1 = updater.Updater(net1, dict(hyps, update_chunk_slots=R), data_parallel=False)
            info1 = upd1.update_model(full)
            for key in ("Loss", "Pi_Loss", "ValLoss", "Entropy"):
                assert abs(info[key] - info1[key]) <= 2e-6 * max(1.0, abs(info1[key])), (model, key)
            assert abs(float(info["GradNorm"]) - float(info1["GradNorm"])) <= 2e-6 * float(info1["GradNorm"])
            step = hyps["lr"] / 0.1
            for (k, p), (_, p1) in zip(net.named_parameters(), net1.named_parameters()):
                # RMSprop's first step is sign-like: compare to a small fraction of one step
                assert float((p - p1).abs().max()) <= 1e-6 * float(p1.abs().max()) + 0.02 * step, (model, k)
        ret.put((rank, "ok"))
    except Exception as e:                                     # pragma: no cover
        import traceback
        ret.put((rank, "FAIL: " + traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_cuda_update_equals_full_batch(dev):
    """Two ranks (gloo rendezvous, both on cuda:0 -- NCCL refuses two ranks per device) run the
    PRODUCT Updater on their shard of the rollout slots; the result equals the single-process
    full-batch update (global advantage statistics, 1/N_global means, summed gradients)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    ret = ctx.Queue()
    world, port = 2, 29533 + os.getpid() % 200
    procs = [ctx.Process(target=_dp_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    results = [ret.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for r, msg in results:
        assert msg == "ok", "rank %d: %s" % (r, msg)


# ------------------------------------------------------- stage 4 additions ----

@pytest.mark.parametrize("M,N,K", [(256, 256, 16384), (256, 2048, 20480), (2000, 256, 5120), (128, 64, 4096)])
def test_tc_gemm_split_k(dev, M, N, K):
    """Weight-gradient shaped GEMMs (few output tiles, long K) run split-K on tcgen05; the
    result is the fp64 product to 1e-5 of its scale, for both operand orientations."""
    from a2c_b200 import ops
    ops.set_precision(1)
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    At = torch.randn((K, M), generator=g).to(dev)          # A(m, k) = At[k, m]  (mn contiguous)
    B = torch.randn((K, N), generator=g).to(dev)
    out = ops.gemm(At, 1, M, B, N, 1, M, N, K)
    want = At.double().t() @ B.double()
    assert float((out.double() - want).abs().max() / want.abs().max()) < 1e-5
    bias = torch.randn(N, generator=g).to(dev)
    out2 = torch.ones((M, N), device=dev)
    ops.gemm(At, 1, M, B, N, 1, M, N, K, out=out2, ldc=N, bias=bias, relu=True, beta=1)
    want2 = torch.relu(want + bias.double() + 1.0)
    assert float((out2.double() - want2).abs().max() / want2.abs().max()) < 1e-5


def test_gru_sequence_matches_stepwise_cell(dev):
    """The hoisted BPTT node (GRUSequenceFn) == T chained GRUCellFn steps with the masked carry
    (a2c/updater.py:157-165), values and all gradients."""
    from a2c_b200 import ops
    T, B, Hd = 5, 37, 64
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: (torch.randn(*s, generator=g) * 0.3).to(dev)
    x, h0 = mk(T, B, Hd), mk(B, Hd)
    Wx, Wh, b = mk(3, Hd, Hd), mk(3, Hd, Hd), mk(3, 1, Hd)
    keep = (torch.rand(T, B, generator=g) > 0.3).float().to(dev)
    wgt = mk(T, B, Hd)

    def run(seq):
        leaves = [t.clone().requires_grad_(True) for t in (x, h0, Wx, Wh, b)]
        xx, hh, wx, wh, bb = leaves
        if seq:
            hall = ops.GRUSequenceFn.apply(xx, hh, wx, wh, bb, keep)
        else:
            outs, hs = [], hh
            for t in range(T):
                hn, hs = ops.GRUCellFn.apply(xx[t], hs, wx, wh, bb, keep[t])
                outs.append(hn)
            hall = torch.stack(outs, 0)
        (hall * wgt).sum().backward()
        return hall.detach(), [l.grad for l in leaves]
    ha, ga = run(True)
    hb, gb = run(False)
    assert H.rel_err(ha.cpu().numpy(), hb.cpu().numpy()) < 1e-6
    for name, a, c in zip(("dx", "dh0", "dWx", "dWh", "db"), ga, gb):
        assert H.rel_err(a.cpu().numpy(), c.cpu().numpy()) < 5e-6, name


STACKS = {  # name: (Cin, H, [(Cout, stride)...])
    "conv": (2, 80, [(16, 1), (24, 1), (32, 2), (64, 2)]),
    "gru": (2, 80, [(16, 1), (24, 2), (32, 2), (48, 2), (64, 2)]),
}


@pytest.mark.parametrize("stack", ["conv", "gru"])
@pytest.mark.parametrize("precision,tol", [(1, 1e-5), (0, 1e-5)])
def test_conv_stack_mask_at_source(dev, stack, precision, tol):
    """ConvStackFn (ReLU masks applied in the input-gradient epilogues, bias gradients from those
    epilogues, a2c_conv2d_bwd_data_relu / a2c_channel_sum / a2c_gemm_masked) against the float64
    autograd of the same Conv+ReLU stack followed by a Linear layer (a2c/models.py:312,247)."""
    import torch.nn.functional as F
    from a2c_b200 import ops
    old = ops.get_precision()
    ops.set_precision(precision)
    try:
        Cin, Hh, layers = STACKS[stack]
        g = torch.Generator().manual_seed(5)
        Nb = 5
        x = (torch.rand(Nb, Cin, Hh, Hh, generator=g) < 0.3).float()
        params, c = [], Cin
        for cout, st in layers:
            params.append(((torch.randn(cout, c, 3, 3, generator=g) / np.sqrt(c * 9.0)), torch.randn(cout, generator=g) * 0.1, st))
            c = cout
        # device forward
        dparams = []
        for w, b, st in params:
            dparams += [w.to(dev).requires_grad_(True), b.to(dev).requires_grad_(True)]
        geom = tuple((st, 1) for _, _, st in params)
        feats = ops.ConvStackFn.apply(ops.View.of_states(x.to(torch.uint8).to(dev)), True, geom, *dparams)
        flat = feats.reshape(Nb, -1)
        K = flat.shape[1]
        Wl = (torch.randn(40, K, generator=g) / np.sqrt(K))
        bl = torch.randn(40, generator=g) * 0.1
        Wd, bd = Wl.to(dev).requires_grad_(True), bl.to(dev).requires_grad_(True)
        out = ops.LinearFn.apply(flat, Wd, bd, True, True)
        gout = torch.randn(out.shape, generator=g)
        out.backward(gout.to(dev))
        # float64 reference through the masks the device used
        p64 = [(w.double().requires_grad_(True), b.double().requires_grad_(True), st) for w, b, st in params]
        cur = x.double()
        dev_acts, ref_pre = [], []
        with torch.no_grad():
            a = ops.View.of_states(x.to(torch.uint8).to(dev))
            for l, (w, b, st) in enumerate(params):
                a = ops.ConvStackFn.apply(a, True, ((st, 1),), w.to(dev), b.to(dev))
                dev_acts.append(a.cpu())
        for l, (w, b, st) in enumerate(p64):
            pre = F.conv2d(cur, w, b, stride=st, padding=1)
            assert H.rel_err(dev_acts[l].numpy(), F.relu(pre).detach().numpy()) < tol, ("fwd", l)
            cur = pre * (dev_acts[l] > 0).double()
        W64, b64 = Wl.double().requires_grad_(True), bl.double().requires_grad_(True)
        lin = cur.reshape(Nb, -1) @ W64.t() + b64
        (lin * (out.detach().cpu() > 0).double()).backward(gout.double())
        for l, (w, b, st) in enumerate(p64):
            assert H.rel_err(dparams[2 * l].grad.cpu().numpy(), w.grad.numpy()) < tol, ("dW", l)
            assert H.rel_err(dparams[2 * l + 1].grad.cpu().numpy(), b.grad.numpy()) < tol, ("db", l)
        assert H.rel_err(Wd.grad.cpu().numpy(), W64.grad.numpy()) < tol
    finally:
        ops.set_precision(old)


@pytest.mark.parametrize("precision", [1, 0])
@pytest.mark.parametrize("Cin,Cout,Cnext,Hh,S", [(16, 24, 32, 80, 2), (32, 48, 64, 20, 2), (8, 16, 24, 40, 1)])
def test_relu_bits_entries(dev, precision, Cin, Cout, Cnext, Hh, S):
    """1-bit ReLU masks: a2c_conv2d_fwd_relu_bits writes bit (c & 7) of byte [n, c >> 3, pos] = Y > 0 next to
    the same Y as a2c_conv2d_fwd, and a2c_conv2d_bwd_data_relu_bits gives bit-identical results to
    a2c_conv2d_bwd_data_relu reading the fp32 activations (shift-GEMM epilogues at precision 1, the
    fallback passes at precision 0)."""
    import ctypes as C
    from a2c_b200 import ops
    from a2c_b200._lib import call, ptr, stream_ptr
    g = torch.Generator().manual_seed(11)
    N = 6
    x = torch.randn(N, Cin, Hh, Hh, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / np.sqrt(Cin * 9.0)).to(dev)
    b = (torch.randn(Cout, generator=g) * 0.1).to(dev)
    s = ops.conv_shape(Cin, Cout, Hh, Hh, 3, 1, 1)
    y0 = torch.empty(N, Cout, s.OH, s.OW, device=dev); y1 = torch.empty_like(y0)
    bits = torch.zeros(N, Cout // 8, s.OH * s.OW, dtype=torch.uint8, device=dev)
    call("a2c_conv2d_fwd", ptr(x), None, ptr(w), ptr(b), ptr(y0), N, C.byref(s), 1, precision, stream_ptr())
    call("a2c_conv2d_fwd_relu_bits", ptr(x), None, ptr(w), ptr(b), ptr(y1), ptr(bits), N, C.byref(s), precision,
         stream_ptr())
    assert torch.equal(y0, y1)
    want = (y0 > 0).reshape(N, Cout // 8, 8, s.OH * s.OW).to(torch.int32)
    want = (want * (2 ** torch.arange(8, device=dev, dtype=torch.int32)).view(1, 1, 8, 1)).sum(2).to(torch.uint8)
    assert torch.equal(bits, want)
    # the layer above: masked input gradient + the bias gradient of this layer, from floats and from bits
    s2 = ops.conv_shape(Cout, Cnext, s.OH, s.OW, 3, S, 1)
    w2 = (torch.randn(Cnext, Cout, 3, 3, generator=g) / np.sqrt(Cout * 9.0)).to(dev)
    dy = torch.randn(N, Cnext, s2.OH, s2.OW, generator=g).to(dev)
    ws = ops.workspace(dev)
    outs = []
    for use_bits in (False, True):
        dx = torch.full_like(y0, float("nan")); dbp = torch.empty(Cout, device=dev)
        if use_bits:
            call("a2c_conv2d_bwd_data_relu_bits", ptr(dy), None, ptr(w2), ptr(y0), ptr(bits), ptr(dx), ptr(dbp), N,
                 C.byref(s2), precision, ptr(ws), ws.numel(), stream_ptr())
        else:
            call("a2c_conv2d_bwd_data_relu", ptr(dy), None, ptr(w2), ptr(y0), ptr(dx), ptr(dbp), N, C.byref(s2),
                 precision, ptr(ws), ws.numel(), stream_ptr())
        outs.append((dx, dbp))
    assert torch.equal(outs[0][0], outs[1][0])
    assert H.rel_err(outs[1][1].cpu().numpy(), outs[0][1].cpu().numpy()) < 1e-6      # (f64 atomics: order)
    assert bool(((outs[1][0] != 0) <= (y0 > 0)).all())


def test_conv_stack_unmasked_gradient_and_tensor_input(dev):
    """premasked=False (the caller hands an unmasked output gradient) and a float tensor input
    that itself needs a gradient."""
    import torch.nn.functional as F
    from a2c_b200 import ops
    g = torch.Generator().manual_seed(9)
    Nb = 4
    x = torch.randn(Nb, 8, 20, 20, generator=g)
    w1, b1 = torch.randn(16, 8, 3, 3, generator=g) / 8.5, torch.randn(16, generator=g) * 0.1
    w2, b2 = torch.randn(24, 16, 3, 3, generator=g) / 12.0, torch.randn(24, generator=g) * 0.1
    xd = x.to(dev).requires_grad_(True)
    dp = [t.to(dev).requires_grad_(True) for t in (w1, b1, w2, b2)]
    y = ops.ConvStackFn.apply(xd, False, ((1, 1), (2, 1)), *dp)
    gy = torch.randn(y.shape, generator=g)
    y.backward(gy.to(dev))
    x64 = x.double().requires_grad_(True)
    p64 = [t.double().requires_grad_(True) for t in (w1, b1, w2, b2)]
    a1 = F.conv2d(x64, p64[0], p64[1], padding=1)
    with torch.no_grad():
        m1 = (ops.ConvStackFn.apply(x.to(dev), False, ((1, 1),), w1.to(dev), b1.to(dev)).cpu() > 0).double()
    a2 = F.conv2d(a1 * m1, p64[2], p64[3], stride=2, padding=1)
    assert H.rel_err(y.detach().cpu().numpy(), F.relu(a2).detach().numpy()) < TOL
    (a2 * (y.detach().cpu() > 0).double()).backward(gy.double())
    assert H.rel_err(xd.grad.cpu().numpy(), x64.grad.numpy()) < TOL
    for d, r in zip(dp, p64):
        assert H.rel_err(d.grad.cpu().numpy(), r.grad.numpy()) < TOL


def test_other_preprocessors_bit_exact(dev):
    """normalize / snake / breakout / null prep (a2c/preprocessing.py:4-9,19-32) against the numpy
    restatement in the oracle, per frame and batched."""
    from a2c_b200 import preprocessing as P
    rng = np.random.default_rng(12)
    frames = rng.integers(0, 256, size=(5, 210, 160, 3), dtype=np.uint8)
    frames[0, ::3, ::5, 0] = 1
    frames[1, ::2, ::7, :] = 255
    for f in frames:
        a, b = P.normalize_prep(f), O.normalize_prep(f)
        assert a.dtype == b.dtype == np.float64 and a.shape == b.shape and np.array_equal(a, b)
        a, b = P.snake_prep(f), O.snake_prep(f)
        assert a.dtype == b.dtype == np.float32 and a.shape == b.shape and np.array_equal(a, b)
        a, b = P.breakout_prep(f), O.breakout_prep(f)
        assert a.dtype == b.dtype == np.uint8 and a.shape == b.shape == (1, 80, 72) and np.array_equal(a, b)
        assert np.array_equal(P.null_prep(f), O.null_prep(f))
    small = rng.integers(0, 256, size=(3, 11, 13, 3), dtype=np.uint8)     # snake grids are small
    small[:, ::2, ::2, 0] = 1
    got = P.snake_prep_batch(torch.from_numpy(small).to(dev)).cpu().numpy()
    assert np.array_equal(got, np.stack([O.snake_prep(f) for f in small]))


@pytest.mark.parametrize("src", ["u8_ring", "f32_states"])
def test_im2col_stack_matches_float64_autograd(dev, src):
    """A3CModel's conv stack on explicit patch matrices (a2c_im2col / a2c_col2im_nhwc + tcgen05
    GEMMs, channels-last activations) against the float64 autograd of F.conv2d + ReLU + Linear
    with the reference's NCHW flatten (a2c/models.py:22-46, 59-90)."""
    import torch.nn.functional as F
    from a2c_b200 import ops
    g = torch.Generator().manual_seed(21)
    Nb = 6
    x = (torch.rand(Nb, 2, 80, 80, generator=g) < 0.3).float()
    if src == "f32_states":
        x = x * torch.rand(Nb, 2, 80, 80, generator=g)
    W1, b1 = torch.randn(16, 2, 8, 8, generator=g) / 11.3, torch.randn(16, generator=g) * 0.1
    W2, b2 = torch.randn(32, 16, 4, 4, generator=g) / 16.0, torch.randn(32, generator=g) * 0.1
    Wp, bp = torch.randn(24, 2048, generator=g) / 45.0, torch.randn(24, generator=g) * 0.1
    dp = [t.to(dev).requires_grad_(True) for t in (W1, b1, W2, b2, Wp, bp)]
    view = ops.View.of_states((x.to(torch.uint8) if src == "u8_ring" else x).to(dev))
    feats = ops.Im2colStackFn.apply(view, ((4, 0), (2, 0)), *dp[:4])
    assert feats.shape == (Nb, 2048)
    w = ops.nhwc_linear_weight(dp[4], 32, 64)
    out = ops.LinearFn.apply(feats, w, dp[5], False, True)
    gout = torch.randn(out.shape, generator=g)
    out.backward(gout.to(dev))
    p64 = [t.double().requires_grad_(True) for t in (W1, b1, W2, b2, Wp, bp)]
    a1 = F.conv2d(x.double(), p64[0], p64[1], stride=4)
    y1 = F.relu(a1)
    a2 = F.conv2d(y1, p64[2], p64[3], stride=2)
    y2 = F.relu(a2)
    got_y2 = feats.detach().cpu().view(Nb, 64, 32).permute(0, 2, 1).reshape(Nb, 32, 8, 8)
    assert H.rel_err(got_y2.numpy(), y2.detach().numpy()) < TOL
    # backward through the masks the device used
    m2 = (got_y2 > 0).double()
    ref = (a2 * m2).reshape(Nb, -1) @ p64[4].t() + p64[5]
    assert H.rel_err(out.detach().cpu().numpy(), ref.detach().numpy()) < TOL
    ref.backward(gout.double())
    for name, d, r in zip(("W1", "b1", "W2", "b2", "Wp", "bp"), dp, p64):
        assert H.rel_err(d.grad.cpu().numpy(), r.grad.numpy()) < 2e-5, name


@pytest.mark.parametrize("model,bptt", [("FCModel", False), ("ConvModel", False), ("GRUModel", True)])
def test_update_from_rollout_cuda_graph_equals_eager(dev, model, bptt):
    """Stage 2 + the whole update replayed as a CUDA graph (Updater.update_from_rollout(graph=True))
    gives the same parameters and loss terms as the eager path, update after update."""
    from a2c_b200 import ops, updater
    R, T, fs, h = 6, 5, 2, 32
    hyps = dict(HY, n_tsteps=T, n_rollouts=R, use_bptt=bptt)
    trajs = [synth.make_trajectory(R, T, seed=70 + k, first_rollout=(k == 0)) for k in range(4)]
    results = {}
    for graph in (False, True):
        net = build_net(model, fs, h, dev)
        upd = updater.Updater(net, hyps)
        ring = ops.PlaneRing(R, T, fs, dev)
        rew = torch.empty(R * T, device=dev)
        dn = torch.empty(R * T, device=dev)
        val = torch.empty((R, T + 1), device=dev)
        act = torch.empty(R * T, dtype=torch.int64, device=dev)
        hst = torch.from_numpy(synth.hashed_tensor((R * T, h), 999, 0.5)).float().to(dev) if model == "GRUModel" else None
        outs = []
        for k, tr in enumerate(trajs):
            fr = torch.from_numpy(tr["frames"]).to(dev)
            rs = torch.from_numpy(tr["reset"]).to(dev)
            if k == 0:
                ring.push_frames(fr, rs)
            else:
                ring.carry()
                ring.push_frames(fr[:, 1:].contiguous(), rs[:, 1:].contiguous(), first_frame_index=1)
            done = ((tr["reset"][:, 1:] != 0) | (tr["rewards"] != 0)).astype(np.float32)
            rew.copy_(torch.from_numpy(tr["rewards"].reshape(-1).copy()))
            dn.copy_(torch.from_numpy(done.reshape(-1).copy()))
            val.copy_(torch.from_numpy(tr["values"]))
            act.copy_(torch.from_numpy(np.random.default_rng(k).integers(0, 3, size=R * T)))
            out = upd.update_from_rollout(ring.view(T), rew, dn, val, act, hst, graph=graph)
            outs.append(out.clone().cpu())
        if graph:
            assert any(e["graph"] is not None for e in upd._graphs.values()), "no graph was captured"
            assert upd._step == len(trajs)
        results[graph] = (outs, [p.detach().clone() for p in net.parameters()])
    for a, b in zip(results[False][0], results[True][0]):
        assert torch.allclose(a, b, rtol=2e-5, atol=1e-7), (a, b)
    step = hyps["lr"] / 0.1
    for pa, pb in zip(results[False][1], results[True][1]):
        assert float((pa - pb).abs().max()) <= 1e-6 * float(pa.abs().max()) + 0.02 * step

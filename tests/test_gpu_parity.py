"""Parity of the CUDA path (through the C ABI) against the oracle and the
reference-generated golden fixtures.  Bit-exact for uint8 / index / scan work,
<= 1e-5 relative (BASELINE.json north_star) for fp32."""
import ctypes as C
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import helpers as H
from a2c_b200 import synth
from oracle import a2c_oracle as O

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def ops_mod():
    from a2c_b200 import ops
    return ops


# ---------------------------------------------------------------- stage 1 ----

@pytest.mark.parametrize("variant", [0, 1])
def test_obs_prep_bit_exact(dev, variant):
    ops = ops_mod()
    rng = np.random.default_rng(3)
    frames = synth.make_frames(rng, 37)
    frames[0][:] = 144
    frames[1][:] = 109
    frames[2][:] = 0
    frames[3] = rng.integers(0, 256, size=frames[3].shape, dtype=np.uint8)   # pure noise
    R, Fn = 1, 37
    ring = ops.PlaneRing(R, Fn - 1, 1, dev)
    reset = np.zeros((R, Fn), dtype=np.uint8)
    ring.push_frames(torch.from_numpy(frames[None]).to(dev), torch.from_numpy(reset).to(dev),
                     variant=variant)
    got = ring.planes[0, :Fn].cpu().numpy().reshape(Fn, 80, 80)
    want = np.stack([O.pong_prep(f)[0] for f in frames])
    assert np.array_equal(got, want)
    assert got[0].sum() == 0 and got[1].sum() == 0 and got[2].sum() == 0


def test_obs_prep_golden_frames(dev):
    ops = ops_mod()
    g = H.load("units.npz")
    rng = np.random.default_rng(int(g["prep_frames_seed"]))
    frames = synth.make_frames(rng, 6)
    frames[0][:] = 144
    frames[1][:] = 109
    frames[2][:] = 0
    want = np.unpackbits(g["prep_out_bits"])[:6 * 6400].reshape(6, 80, 80)
    ring = ops.PlaneRing(6, 1, 1, dev)
    ring.push_frames(torch.from_numpy(frames[:, None]).to(dev),
                     torch.ones((6, 1), dtype=torch.uint8, device=dev))
    assert np.array_equal(ring.planes[:, 0].cpu().numpy().reshape(6, 80, 80), want)


def test_pong_prep_callable_matches_reference_signature(dev):
    from a2c_b200 import preprocessing
    rng = np.random.default_rng(9)
    f = synth.make_frames(rng, 1)[0]
    out = preprocessing.pong_prep(f)
    assert out.shape == (1, 80, 80) and out.dtype == np.uint8
    assert np.array_equal(out, O.pong_prep(f))


@pytest.mark.parametrize("fs", [1, 2, 3, 4])
@pytest.mark.parametrize("host", [False, True])
def test_ring_stack_with_resets_and_carry(dev, fs, host):
    ops = ops_mod()
    R, T = 5, 7
    t0 = synth.make_trajectory(R, T, seed=11, p_reset=0.3)
    t1 = synth.make_trajectory(R, T, seed=12, p_reset=0.3, first_rollout=False)
    frames, reset = H.chain_frames([t0, t1])
    want = O.prep_and_stack(frames, reset, fs)              # [R, 2T+1, fs, 80, 80]
    ring = ops.PlaneRing(R, T, fs, dev)

    def put(x):
        x = torch.from_numpy(np.ascontiguousarray(x))
        return x.pin_memory() if host else x.to(dev)
    ring.push_frames(put(t0["frames"]), put(t0["reset"]))
    got0 = ring.view(T + 1).materialize(torch.uint8).cpu().numpy().reshape(R, T + 1, fs, 80, 80)
    assert np.array_equal(got0, want[:, :T + 1])
    f32 = ring.view(T).materialize(torch.float32).cpu().numpy().reshape(R, T, fs, 80, 80)
    assert np.array_equal(f32, want[:, :T].astype(np.float32))
    ring.carry()
    ring.push_frames(put(t1["frames"][:, 1:]), put(t1["reset"][:, 1:]), first_frame_index=1)
    got1 = ring.view(T + 1).materialize(torch.uint8).cpu().numpy().reshape(R, T + 1, fs, 80, 80)
    assert np.array_equal(got1, want[:, T:])


def test_prep_empty_and_single(dev):
    ops = ops_mod()
    ring = ops.PlaneRing(3, 4, 2, dev)
    ring.push_frames(torch.zeros((3, 0, 210, 160, 3), dtype=torch.uint8, device=dev),
                     torch.zeros((3, 0), dtype=torch.uint8, device=dev))
    assert int(ring.planes.sum()) == 0
    with pytest.raises(Exception):
        ring.push_frames(torch.zeros((3, 9, 210, 160, 3), dtype=torch.uint8, device=dev),
                         torch.zeros((3, 9), dtype=torch.uint8, device=dev))   # overflows the ring


def test_prep_full_size_checksum(dev):
    """C2-sized batch: plane population must equal the oracle's on a sampled
    subset and the idempotence property prep(prep-like frame) holds."""
    ops = ops_mod()
    R, T = 256, 20
    g = torch.Generator(device="cpu").manual_seed(5)
    frames = torch.randint(0, 256, (R, T + 1, 210, 160, 3), dtype=torch.uint8, generator=g)
    bg = torch.rand((R, T + 1, 210, 160, 3), generator=g) < 0.9
    frames[bg] = 144
    ring = ops.PlaneRing(R, T, 2, dev)
    reset = torch.zeros((R, T + 1), dtype=torch.uint8)
    ring.push_frames(frames.to(dev), reset.to(dev))
    got = ring.planes[:, 1:].cpu().numpy().reshape(R, T + 1, 80, 80)
    idx = [(0, 0), (17, 3), (255, 20), (100, 11)]
    for r, t in idx:
        assert np.array_equal(got[r, t], O.pong_prep(frames[r, t].numpy())[0])
    sel = frames[:, :, 35:195:2, ::2, 0]
    want_count = int(((sel != 0) & (sel != 144) & (sel != 109)).sum())
    assert int(got.sum()) == want_count


# ------------------------------------------------------------- acting -------

def test_sample_action_bit_exact(dev):
    ops = ops_mod()
    g = H.load("units.npz")
    got = ops.sample_action(torch.from_numpy(g["sa_pi"]).to(dev), torch.from_numpy(g["sa_u"]).to(dev),
                            is_prob=True).cpu().numpy()
    assert np.array_equal(got, g["sa_act"].astype(np.int64))
    assert got[0] == -1
    from a2c_b200 import utils
    out = utils.sample_action(torch.from_numpy(g["sa_pi"]), rand_nums=torch.from_numpy(g["sa_u"]))
    assert out.dtype == torch.float32 and np.array_equal(out.numpy(), g["sa_act"])


def test_sample_action_from_logits(dev):
    """From logits the kernel evaluates softmax itself (max, exp(x - max), sequential sum, IEEE
    division: F.softmax's operation order).  The reference's exp is Sleef's vectorised expf on the
    CPU and libdevice's expf here -- both within 1-2 ulp of exp but not bit-identical -- so the
    index can differ from the oracle's ONLY where a running sum sits within a few ulp of u.
    Asserted: every mismatch is such a rounding tie (never more than a handful), and given the
    oracle's own probabilities the kernel is bit-exact (previous test)."""
    ops = ops_mod()
    gen = torch.Generator().manual_seed(1)
    logits = torch.randn(20000, 3, generator=gen) * 2
    u = torch.rand(20000, generator=gen)
    got = ops.sample_action(logits.to(dev), u.to(dev), is_prob=False).cpu()
    want = O.get_action(logits, u)
    bad = torch.nonzero(got != want).flatten()
    assert bad.numel() <= 3
    cum = torch.cumsum(F.softmax(logits[bad].double(), dim=-1), dim=-1)
    gap = (cum - u[bad].double()[:, None]).abs().min(dim=1).values
    assert bool((gap <= 1e-6).all()), gap          # a few fp32 ulp of a sum <= 1


# ---------------------------------------------------------------- stage 2 ----

def test_discount_known_answer_and_golden(dev):
    from a2c_b200 import utils
    g = H.load("units.npz")
    out = utils.discount(torch.tensor([1., 1, 1]), torch.tensor([0., 0, 1]), .5)
    assert not out.is_cuda and np.array_equal(out.numpy(), np.array([1.75, 1.5, 1.0], np.float32))
    for key, f in (("disc_out_g99", 0.99), ("disc_out_gl", 0.99 * 0.98)):
        got = utils.discount(torch.from_numpy(g["disc_arr"]), torch.from_numpy(g["disc_dones"]), f)
        assert np.array_equal(got.numpy(), g[key])


@pytest.mark.parametrize("N,p", [(1, 0.0), (7, 0.0), (1000, 0.05), (81920, 0.1), (4096, 1.0)])
def test_discount_random_bit_exact(dev, N, p):
    ops = ops_mod()
    gen = torch.Generator().manual_seed(N)
    a = torch.randn(N, generator=gen)
    d = (torch.rand(N, generator=gen) < p).float()
    stats = torch.zeros(2, dtype=torch.float64, device=dev)
    got = ops.discount(a.to(dev), d.to(dev), 0.99 * 0.98, stats=stats).cpu()
    want = O.discount(a, d, 0.99 * 0.98)
    assert torch.equal(got, want)
    assert abs(stats[0].item() - want.double().sum().item()) < 1e-6 * max(1, N)
    assert abs(stats[1].item() - (want.double() ** 2).sum().item()) < 1e-6 * max(1, N)


@pytest.mark.parametrize("R,T", [(1, 1), (4, 15), (3, 2), (256, 20), (1024, 20), (512, 32), (130, 5)])
def test_delta_scan_bit_exact(dev, R, T):
    ops = ops_mod()
    traj = synth.make_trajectory(R, T, seed=R * 100 + T, with_frames=False, p_reset=0.05)
    done = ((traj["reset"][:, 1:] != 0) | (traj["rewards"] != 0)).astype(np.float32)
    r2, d2, want_deltas = O.rollout_bookkeeping(traj["rewards"], done, traj["values"], 0.99)
    rew = torch.from_numpy(traj["rewards"].reshape(-1).copy()).to(dev)
    dn = torch.from_numpy(done.reshape(-1).copy()).to(dev)
    deltas, advs, rets, stats = ops.delta_scan(rew, dn, torch.from_numpy(traj["values"]).to(dev),
                                               0.99, 0.98)
    assert torch.equal(rew.cpu(), r2.reshape(-1))
    assert torch.equal(dn.cpu(), d2.reshape(-1))
    assert torch.equal(deltas.cpu(), want_deltas.reshape(-1))
    want_advs = O.discount(want_deltas, d2, 0.99 * 0.98)
    want_rets = O.discount(r2, d2, 0.99)
    assert torch.equal(advs.cpu(), want_advs)
    assert torch.equal(rets.cpu(), want_rets)
    assert abs(stats[0].item() - want_advs.double().sum().item()) < 1e-6 * R * T
    # every segment ends with done = 1 (runner.py:243)
    assert bool((dn.view(R, T)[:, -1] == 1).all())


# ---------------------------------------------------------------- stage 3 ----

@pytest.mark.parametrize("N,A,norm,nstep", [(60, 3, True, False), (5120, 3, True, False),
                                            (333, 3, False, True), (2048, 6, True, True)])
def test_loss_fwd_bwd(dev, N, A, norm, nstep):
    ops = ops_mod()
    gen = torch.Generator().manual_seed(N + A)
    logits = (torch.randn(N, A, generator=gen) * 2).requires_grad_(True)
    vals = torch.randn(N, 1, generator=gen).requires_grad_(True)
    actions = torch.randint(0, A, (N,), generator=gen)
    actions[0] = -1                                     # the sample_action "-1" edge
    advs = torch.randn(N, generator=gen)
    rets = torch.randn(N, generator=gen)
    hyps = dict(pi_coef=1.0, val_coef=0.5, entr_coef=0.005, norm_advs=norm, use_nstep_rets=nstep)
    a_n = (advs - advs.mean()) / (advs.std() + 1e-6) if norm else advs
    r_n = advs + vals.detach().squeeze() if nstep else rets
    loss, pi_l, val_l, ent_l = O.a2c_loss(logits, vals, actions, a_n, r_n, hyps)
    loss.backward()
    stats = torch.tensor([advs.double().sum(), (advs.double() ** 2).sum()], dtype=torch.float64)
    dl, dv, out = ops.loss_fwd_bwd(logits.detach().to(dev), vals.detach().to(dev), actions.to(dev),
                                   advs.to(dev), rets.to(dev), stats.to(dev), hyps)
    want = [loss.item(), pi_l.item(), val_l.item(), ent_l.item()]
    for a, b in zip(out.tolist(), want):
        assert abs(a - b) <= TOL * max(1.0, abs(b))
    assert H.rel_err(dl.cpu().numpy(), logits.grad.numpy()) < TOL
    assert H.rel_err(dv.cpu().numpy(), vals.grad.numpy().reshape(-1)) < TOL


@pytest.mark.gpu
@pytest.mark.parametrize("cut", [2560, 1000])
def test_sharded_loss_contract(dev, cut):
    """two env shards evaluated with the global advantage statistics and
    n_global produce loss terms / gradients that add up to the full-batch ones
    (the gloo side of the same contract: tests/test_distributed.py)."""
    ops = ops_mod()
    N, A = 5120, 3
    gen = torch.Generator().manual_seed(5)
    logits = (torch.randn(N, A, generator=gen) * 2).to(dev)
    vals = torch.randn(N, generator=gen).to(dev)
    actions = torch.randint(0, A, (N,), generator=gen).to(dev)
    advs = torch.randn(N, generator=gen).to(dev)
    rets = torch.randn(N, generator=gen).to(dev)
    hyps = dict(pi_coef=1.0, val_coef=0.5, entr_coef=0.005, norm_advs=True, use_nstep_rets=False)
    stats = torch.stack([advs.double().sum(), (advs.double() ** 2).sum()])
    dl, dv, out = ops.loss_fwd_bwd(logits, vals, actions, advs, rets, stats, hyps)
    parts = []
    for sl in (slice(0, cut), slice(cut, N)):
        parts.append(ops.loss_fwd_bwd(logits[sl].contiguous(), vals[sl].contiguous(),
                                      actions[sl].contiguous(), advs[sl].contiguous(),
                                      rets[sl].contiguous(), stats, hyps, n_global=N))
    assert H.rel_err(torch.cat([p[0] for p in parts]).cpu().numpy(), dl.cpu().numpy()) < 1e-6
    assert H.rel_err(torch.cat([p[1] for p in parts]).cpu().numpy(), dv.cpu().numpy()) < 1e-6
    assert H.rel_err((parts[0][2] + parts[1][2]).cpu().numpy(), out.cpu().numpy()) < TOL


# ---------------------------------------------------------------- stage 4 ----

@pytest.mark.parametrize("M,N,K,ta,tb", [
    (128, 128, 64, False, False), (60, 256, 12800, False, True), (5, 3, 7, True, False),
    (300, 200, 100, True, True), (256, 3, 256, False, True), (2000, 130, 513, False, False),
    (3, 256, 5120, True, False), (129, 257, 31, False, True)])
def test_gemm_fp32(dev, M, N, K, ta, tb):
    ops = ops_mod()
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=gen)
    B = torch.randn((N, K) if tb else (K, N), generator=gen)
    bias = torch.randn(N, generator=gen)
    Am = A.t() if ta else A
    Bm = B.t() if tb else B
    want = (Am.double() @ Bm.double() + bias.double()).clamp_min(0).float()
    sam, sak = (1, M) if ta else (K, 1)
    sbk, sbn = (1, K) if tb else (N, 1)
    got = ops.gemm(A.to(dev), sam, sak, B.to(dev), sbk, sbn, M, N, K, bias=bias.to(dev), relu=True,
                   precision=0)
    assert H.rel_err(got.cpu().numpy(), want.numpy()) < TOL
    # beta = 1 accumulates
    got2 = ops.gemm(A.to(dev), sam, sak, B.to(dev), sbk, sbn, M, N, K, out=got.clone(), ldc=N, beta=1,
                    precision=0)
    want2 = want.double() + Am.double() @ Bm.double()
    assert H.rel_err(got2.cpu().numpy(), want2.numpy()) < TOL


@pytest.mark.parametrize("M,N,K,ta,tb", [
    (128, 128, 64, False, True), (256, 256, 512, False, False), (640, 2000, 2560, False, True),
    (512, 1280, 2000, False, False), (2000, 2560, 640, True, False), (300, 200, 100, True, True),
    (4096, 256, 2048, False, True), (5120, 32, 256, False, True), (130, 260, 132, False, True),
    (256, 128, 25600, False, True)])
@pytest.mark.parametrize("precision,tol", [(1, 1e-5), (2, 2e-3)])
def test_gemm_tcgen05(dev, M, N, K, ta, tb, precision, tol):
    """tcgen05 path of a2c_gemm: 3xTF32 (fp32-parity mode, 1e-5) and single-pass
    TF32 (2e-3), every operand layout, ragged edges, long K (accumulation)."""
    ops = ops_mod()
    gen = torch.Generator().manual_seed(M + N + K)
    A = torch.randn((K, M) if ta else (M, K), generator=gen)
    B = torch.randn((N, K) if tb else (K, N), generator=gen)
    bias = torch.randn(N, generator=gen)
    Am = A.t() if ta else A
    Bm = B.t() if tb else B
    want = (Am.double() @ Bm.double() + bias.double()).clamp_min(0)
    sam, sak = (1, M) if ta else (K, 1)
    sbk, sbn = (1, K) if tb else (N, 1)
    from a2c_b200 import _lib
    n0 = _lib.launch_count()
    got = ops.gemm(A.to(dev), sam, sak, B.to(dev), sbk, sbn, M, N, K, bias=bias.to(dev), relu=True,
                   precision=precision)
    assert _lib.launch_count() - n0 in (1, 2)      # 2: split-K (few output tiles, K >= 2048) + its reduction
    # error relative to the scale of the products being summed (sqrt(K) for N(0,1) operands)
    err = (got.cpu().double() - want).abs().max().item() / max(want.abs().max().item(), 1e-30)
    assert err < tol, err
    got2 = ops.gemm(A.to(dev), sam, sak, B.to(dev), sbk, sbn, M, N, K, out=got.clone(), ldc=N, beta=1,
                    precision=precision)
    want2 = want + Am.double() @ Bm.double()
    err2 = (got2.cpu().double() - want2).abs().max().item() / max(want2.abs().max().item(), 1e-30)
    assert err2 < tol, err2


CONVS = [  # Cin, Cout, H, K, stride, pad  -- every layer shape of the four models
    (2, 16, 80, 3, 1, 1), (16, 24, 80, 3, 1, 1), (24, 32, 80, 3, 2, 1), (32, 64, 40, 3, 2, 1),
    (16, 24, 80, 3, 2, 1), (24, 32, 40, 3, 2, 1), (32, 48, 20, 3, 2, 1), (48, 64, 10, 3, 2, 1),
    (2, 16, 80, 8, 4, 0), (16, 32, 19, 4, 2, 0), (3, 16, 80, 3, 1, 1),
    # further paths of the shift-GEMM kernels: unpacked stride-1 weight gradient (M = 128),
    # 8-channel packed band, single column chunk, narrow images
    (24, 32, 40, 3, 1, 1), (8, 16, 20, 3, 1, 1), (32, 24, 24, 3, 1, 1), (4, 8, 16, 3, 1, 1)]


@pytest.fixture
def precision_guard():
    ops = ops_mod()
    old = ops.get_precision()
    yield ops
    ops.set_precision(old)


@pytest.mark.parametrize("precision,TOLP", [(0, 1e-5), (1, 1e-5), (2, 3e-3)])
@pytest.mark.parametrize("Cin,Cout,Hh,K,S,P", CONVS)
def test_conv_fwd_bwd(dev, precision_guard, Cin, Cout, Hh, K, S, P, precision, TOLP):
    """Direct FFMA convolution (0), implicit GEMM on tcgen05 with the 3xTF32 split (1)
    and single-pass TF32 (2), forward / backward-data / backward-weight, against fp64."""
    ops = precision_guard
    ops.set_precision(precision)
    TOL = TOLP
    gen = torch.Generator().manual_seed(Cin * 100 + Cout + K)
    Nb = 3
    x = torch.randn(Nb, Cin, Hh, Hh, generator=gen).requires_grad_(True)
    w = (torch.randn(Cout, Cin, K, K, generator=gen) / np.sqrt(Cin * K * K)).requires_grad_(True)
    b = torch.randn(Cout, generator=gen).requires_grad_(True)
    ypre = F.conv2d(x.double(), w.double(), b.double(), stride=S, padding=P)
    y = F.relu(ypre)
    gy = torch.randn(y.shape, generator=gen)
    xd = x.detach().to(dev).requires_grad_(True)
    wd = w.detach().to(dev).requires_grad_(True)
    bd = b.detach().to(dev).requires_grad_(True)
    yd = ops.Conv2dFn.apply(xd, wd, bd, S, P, True)
    assert H.rel_err(yd.detach().cpu().numpy(), y.detach().numpy()) < TOL
    # backward reference through the ReLU mask the device actually used: an output that is
    # +-1e-7 (fp32) or +-1e-3 (tf32) around zero may legitimately flip its mask bit
    mask = (yd.detach().cpu() > 0).double()
    (ypre * mask).backward(gy.double())
    yd.backward(gy.to(dev))
    assert H.rel_err(xd.grad.cpu().numpy(), x.grad.numpy()) < TOL
    assert H.rel_err(wd.grad.cpu().numpy(), w.grad.numpy()) < TOL
    assert H.rel_err(bd.grad.cpu().numpy(), b.grad.numpy()) < TOL


def test_conv_first_layer_reads_ring(dev):
    ops = ops_mod()
    R, T, fs = 3, 4, 2
    traj = synth.make_trajectory(R, T, seed=2, p_reset=0.3)
    ring = ops.PlaneRing(R, T, fs, dev)
    ring.push_frames(torch.from_numpy(traj["frames"]).to(dev), torch.from_numpy(traj["reset"]).to(dev))
    states = torch.from_numpy(O.prep_and_stack(traj["frames"], traj["reset"], fs)[:, :T]
                              .reshape(R * T, fs, 80, 80)).float()
    gen = torch.Generator().manual_seed(0)
    w = torch.randn(16, fs, 3, 3, generator=gen)
    b = torch.randn(16, generator=gen)
    w64, b64 = w.double().requires_grad_(True), b.double().requires_grad_(True)   # float64 reference
    ypre = F.conv2d(states.double(), w64, b64, padding=1)
    y = F.relu(ypre)
    gy = torch.randn(y.shape, generator=gen)
    wd = w.to(dev).requires_grad_(True)
    bd = b.to(dev).requires_grad_(True)
    yd = ops.Conv2dFn.apply(ring.view(T), wd, bd, 1, 1, True)
    assert H.rel_err(yd.detach().cpu().numpy(), y.detach().numpy()) < TOL
    # backward through the ReLU mask the device used (an output within rounding of zero may flip)
    (ypre * (yd.detach().cpu() > 0).double()).backward(gy.double())
    yd.backward(gy.to(dev))
    assert H.rel_err(wd.grad.cpu().numpy(), w64.grad.numpy()) < TOL
    assert H.rel_err(bd.grad.cpu().numpy(), b64.grad.numpy()) < TOL


def test_layernorm(dev):
    ops = ops_mod()
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(77, 256, generator=gen).requires_grad_(True)
    gm = torch.randn(256, generator=gen).requires_grad_(True)
    bt = torch.randn(256, generator=gen).requires_grad_(True)
    y = F.layer_norm(x, (256,), gm, bt)
    gy = torch.randn(y.shape, generator=gen)
    y.backward(gy)
    xd, gd, bd = (t.detach().to(dev).requires_grad_(True) for t in (x, gm, bt))
    yd = ops.LayerNormFn.apply(xd, gd, bd)
    yd.backward(gy.to(dev))
    assert H.rel_err(yd.detach().cpu().numpy(), y.detach().numpy()) < TOL
    for a, b_ in ((xd, x), (gd, gm), (bd, bt)):
        assert H.rel_err(a.grad.cpu().numpy(), b_.grad.numpy()) < TOL


def test_gru_cell(dev):
    """GRU cell forward / backward against the oracle.  The float64 evaluation of the same formulas
    is the yardstick: the device must be within TOL of it; against the fp32 CPU oracle the allowance
    grows by that oracle's own distance from float64 (its matmul rounding depends on the host CPU:
    one box in six showed 1.7e-5 on the last row)."""
    ops = ops_mod()
    gen = torch.Generator().manual_seed(2)
    B, Hd = 37, 64
    p0 = {"gru.W_x": torch.randn(3, Hd, Hd, generator=gen) / 8,
          "gru.W_h": torch.randn(3, Hd, Hd, generator=gen) / 8,
          "gru.b": torch.randn(3, 1, Hd, generator=gen)}
    x0 = torch.randn(B, Hd, generator=gen)
    h0 = torch.randn(B, Hd, generator=gen)
    keep = (torch.rand(B, generator=gen) < 0.7).float()
    g1, g2 = torch.randn(B, Hd, generator=gen), torch.randn(B, Hd, generator=gen)

    def cpu_eval(dtype):
        p = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in p0.items()}
        x = x0.detach().clone().to(dtype).requires_grad_(True)
        h = h0.detach().clone().to(dtype).requires_grad_(True)
        hn = O.gru_cell(p, x, h)
        (hn * g1.to(dtype) + hn * keep.to(dtype)[:, None] * g2.to(dtype)).sum().backward()
        out = {"hnew": hn.detach(), "hnext": (hn * keep.to(dtype)[:, None]).detach(), "x": x.grad, "h": h.grad}
        out.update({k: v.grad for k, v in p.items()})
        return {k: v.double().numpy() for k, v in out.items()}

    ref32, ref64 = cpu_eval(torch.float32), cpu_eval(torch.float64)
    d = {k: v.detach().clone().to(dev).requires_grad_(True) for k, v in p0.items()}
    xd, hd = x0.detach().clone().to(dev).requires_grad_(True), h0.detach().clone().to(dev).requires_grad_(True)
    hnew, hnext = ops.GRUCellFn.apply(xd, hd, d["gru.W_x"], d["gru.W_h"], d["gru.b"], keep.to(dev))
    (hnew * g1.to(dev) + hnext * g2.to(dev)).sum().backward()
    got = {"hnew": hnew.detach(), "hnext": hnext.detach(), "x": xd.grad, "h": hd.grad}
    got.update({k: v.grad for k, v in d.items()})
    for k, v in got.items():
        v = v.cpu().double().numpy()
        e_ref = H.rel_err(ref32[k], ref64[k])
        assert H.rel_err(v, ref64[k]) < TOL, (k, "device vs float64", H.rel_err(v, ref64[k]), "fp32 oracle vs float64", e_ref)
        assert H.rel_err(v, ref32[k]) < TOL + e_ref, (k, "device vs fp32 oracle", H.rel_err(v, ref32[k]), e_ref)


# ------------------------------------------------ models / Updater / Runner ----

def build_net(model, fs, h, dev):
    from a2c_b200 import models
    net = models.get_model_class(model)([fs, 80, 80], 3, h_size=h, bnorm=False, is_discrete=True)
    missing, unexpected = net.load_state_dict(H.params_for(model, fs, h), strict=False)
    assert not unexpected
    return net.to(dev)


def test_state_dict_keys_match_reference_names(dev):
    """Checkpoint interchange: same keys / shapes as the reference modules."""
    for model in ("FCModel", "A3CModel", "ConvModel", "GRUModel"):
        net = build_net(model, 2, 32, dev)
        sd = net.state_dict()
        for k, shape in O.param_shapes(model, 2, 32).items():
            assert k in sd and tuple(sd[k].shape) == tuple(shape), (model, k)
    sd = build_net("A3CModel", 2, 32, dev).state_dict()
    for k in ("conv1.0.weight", "features.1.0.bias", "emb_bnorm.running_mean"):
        assert k in sd
    from a2c_b200 import models
    assert models.get_model_class("conv") is models.ConvModel


# magnitude of one optimizer step per element (first steps of RMSprop / Adam are sign-like)
STEP_SCALE = {"RMSprop": lambda lr: lr / np.sqrt(1 - 0.99), "Adam": lambda lr: lr}

SMALL = [("c1_fc.npz", "FCModel", 4, 15, 2, 256),
         ("small_a3c.npz", "A3CModel", 3, 6, 2, 64),
         ("small_conv.npz", "ConvModel", 3, 6, 2, 32),
         ("small_gru.npz", "GRUModel", 3, 6, 2, 64),
         ("small_gru_bptt.npz", "GRUModel", 3, 6, 2, 64),
         ("small_fc_adam_nstep.npz", "FCModel", 3, 6, 2, 64),
         ("small_fc_fs3.npz", "FCModel", 3, 6, 3, 64)]


@pytest.mark.parametrize("fname,model,R,T,fs,h", SMALL)
def test_runner_and_updater_against_reference_golden(dev, fname, model, R, T, fs, h):
    """The drop-in path end to end: batched Runner.rollout + Updater.update_model,
    two consecutive updates, against the unmodified reference's outputs."""
    from a2c_b200 import runner as runner_mod, updater as updater_mod
    g = H.load(fname)
    hyps = H.hyps_of(g)
    hyps["preprocessor"] = None
    trajs = H.small_trajectories(g, R, T)
    net = build_net(model, fs, h, dev)
    datas = runner_mod.make_shared_data(R, T, fs, dev, h_size=h if net.is_recurrent else None)
    run = runner_mod.Runner(datas, hyps, None, None, None)
    env = runner_mod.RecordedVecEnv(trajs[0], dev)
    run.attach(net, env)
    upd = updater_mod.Updater(net, hyps)
    N = R * T
    for k, traj in enumerate(trajs):
        if k > 0:
            env.load(traj)
        run.rollout(net, 0, hyps)
        pre = "u%d_" % k
        assert np.array_equal(datas["states"].cpu().numpy().astype(np.uint8),
                              H.golden_states(g, k, N, fs))
        assert np.array_equal(datas["actions"].cpu().numpy(), g[pre + "actions"])
        assert np.array_equal(datas["dones"].cpu().numpy(), g[pre + "dones"])
        for key in ("rewards", "deltas") + (("h_states",) if "h_states" in datas else ()):
            assert H.rel_err(datas[key].cpu().numpy(), g[pre + key]) < TOL, key
        # isolate update parity from acting noise: feed the reference's recorded datas
        shared = {key: torch.from_numpy(g[pre + key]).to(dev) for key in
                  ("rewards", "dones", "deltas", "actions")}
        shared["states_view"] = datas["states_view"]
        shared["states"] = datas["states"]
        if "h_states" in datas:
            shared["h_states"] = torch.from_numpy(g[pre + "h_states"]).to(dev)
        # float64 evaluation of the same update at the parameters the GPU holds now:
        # the yardstick for tensors whose fp32 value is ill-conditioned (see the
        # full-size test below for the rationale of the max(1e-5, 4x reference) bar)
        cur = {n_: p.detach().cpu().double() for n_, p in net.named_parameters()
               if n_ in O.param_shapes(model, fs, h)}
        o64 = O.UpdaterOracle(model, cur, hyps, dtype=torch.float64)
        o64.ret_mean, o64.ret_std = upd.ret_mean, upd.ret_std
        if o64.ret_mean is not None:
            o64.ret_mean, o64.ret_std = o64.ret_mean.cpu().double(), o64.ret_std.cpu().double()
        sh64 = {key: v.cpu() for key, v in shared.items() if torch.is_tensor(v)}
        want64 = o64.compute(sh64)["grads"]
        res = upd.compute_gradients(shared)
        grads = {name: p.grad.detach().clone() for name, p in net.named_parameters()}
        scale = H.golden_grad_scale(g, pre + "grad/")
        for name in O.param_shapes(model, fs, h):
            key = pre + "grad/" + name
            if key in g:
                mine = grads[name].cpu().numpy()
                w64 = H.compact(want64[name].numpy().astype(np.float32)) if want64[name].numel() > H.SAMPLE \
                    else want64[name].numpy().reshape(-1)
                ref_own = H.grad_err_compact(g[key], w64, scale)
                e64 = H.grad_err(mine, w64, scale)
                assert e64 <= max(TOL, 4 * ref_own), ("grad", k, name, e64, ref_own)
                eref = H.grad_err(mine, g[key], scale)
                assert eref <= TOL + 6 * ref_own, ("grad vs fp32 reference", k, name, eref, ref_own)
        # the flat buffer the optimizer reads is exactly these gradients
        flat_norm = float(torch.sqrt(sum((gr.double() ** 2).sum() for gr in grads.values())))
        assert abs(float(upd.flat.grad.double().norm()) - flat_norm) <= 1e-9 * max(1.0, flat_norm)
        upd._optim_step()
        terms = res["loss_terms"].tolist()
        for i, key in enumerate(("Loss", "Pi_Loss", "ValLoss", "Entropy")):
            assert abs(terms[i] - float(g[pre + key])) <= TOL * max(1.0, abs(float(g[pre + key]))), key
        # GradNorm: against the float64 norm of the reference's own gradients; the
        # value the reference reports carries its fp32 norm error (see helpers)
        true_norm = H.golden_true_norm(g, pre + "grad/")
        assert abs(upd.flat.norm.item() - true_norm) <= TOL * max(1.0, true_norm)
        assert abs(upd.flat.norm.item() - float(g[pre + "GradNorm"])) <= 5e-4 * max(1.0, true_norm)
        params = dict(net.named_parameters())
        step = STEP_SCALE[hyps["optim_type"]](hyps["lr"])
        for name in O.param_shapes(model, fs, h):
            assert H.param_err_ok(params[name].detach().cpu().numpy(),
                                  g[pre + "param_after/" + name], step), ("param_after", k, name)


def test_updater_accepts_reference_layout_fp32_states(dev):
    """Drop-in: the reference's own shared_data layout (fp32 states tensor,
    CPU LongTensor actions) gives the same update as the ring view."""
    from a2c_b200 import updater as updater_mod
    g = H.load("small_conv.npz")
    hyps = H.hyps_of(g)
    R, T, fs, h = 3, 6, 2, 32
    N = R * T
    net = build_net("ConvModel", fs, h, dev)
    upd = updater_mod.Updater(net, hyps)
    shared = {"states": torch.from_numpy(H.golden_states(g, 0, N, fs)).float().to(dev),
              "rewards": torch.from_numpy(g["u0_rewards"]).to(dev),
              "dones": torch.from_numpy(g["u0_dones"]).to(dev),
              "deltas": torch.from_numpy(g["u0_deltas"]).to(dev),
              "actions": torch.from_numpy(g["u0_actions"])}          # CPU, as training.py:97
    info = upd.update_model(shared)
    assert set(info) == {"Loss", "Pi_Loss", "ValLoss", "Entropy", "GradNorm"}
    assert isinstance(info["Loss"], float) and torch.is_tensor(info["GradNorm"])
    for key in ("Loss", "Pi_Loss", "ValLoss", "Entropy"):
        assert abs(info[key] - float(g["u0_" + key])) <= TOL * max(1.0, abs(float(g["u0_" + key])))
    params = dict(net.named_parameters())
    step = STEP_SCALE[hyps["optim_type"]](hyps["lr"])
    for name in O.param_shapes("ConvModel", fs, h):
        assert H.param_err_ok(params[name].detach().cpu().numpy(),
                              g["u0_param_after/" + name], step), name


FULL = [("full_c3_a3c.npz", "A3CModel", 1024, 20), ("full_c4_gru_bptt.npz", "GRUModel", 512, 32),
        ("full_c2_conv.npz", "ConvModel", 256, 20)]


@pytest.mark.parametrize("precision,FACT", [(0, 2), (1, 8)])
@pytest.mark.parametrize("fname,model,R,T", FULL)
def test_full_size_update_against_reference_golden(dev, precision_guard, fname, model, R, T, precision, FACT):
    """BASELINE configs C2-C4 at full size: stage 1 from raw frames on the GPU,
    then Updater.update_model, against the unmodified reference's outputs.

    precision 0: every contraction in fp32 FFMA.  precision 1 (the default): GEMMs and
    convolutions on tcgen05 with the 3xTF32 split.  3xTF32 products carry ~2^-21 relative
    error instead of 2^-24, so a few more pre-activations that are ~1e-6 from zero flip their
    ReLU mask bit; the ill-conditioned first-layer gradients see that as a slightly larger
    multiple (FACT) of the reference's own fp32 rounding error."""
    from a2c_b200 import ops, updater as updater_mod
    ops.set_precision(precision)
    g = H.load(fname)
    hyps = H.hyps_of(g)
    fs, h = 2, 256
    N = R * T
    traj = synth.make_trajectory(R, T, seed=int(g["seed"]))
    ring = ops.PlaneRing(R, T, fs, dev)
    ring.push_frames(torch.from_numpy(traj["frames"]).to(dev), torch.from_numpy(traj["reset"]).to(dev))
    view = ring.view(T)
    assert abs(float(view.materialize(torch.uint8).sum(dtype=torch.float64)) - float(g["states_sum"])) < 0.5
    done = ((traj["reset"][:, 1:] != 0) | (traj["rewards"] != 0)).astype(np.float32)
    rew = torch.from_numpy(traj["rewards"].reshape(-1).copy()).to(dev)
    dn = torch.from_numpy(done.reshape(-1).copy()).to(dev)
    deltas, advs, rets, _ = ops.delta_scan(rew, dn, torch.from_numpy(traj["values"]).to(dev),
                                           hyps["gamma"], hyps["lambda_"])
    assert H.rel_err(H.compact(advs.cpu().numpy()), g["advs"]) == 0.0      # bit-exact scans
    assert H.rel_err(H.compact(rets.cpu().numpy()), g["returns"]) == 0.0
    rng = np.random.default_rng(int(g["seed"]) + 1)
    actions = torch.from_numpy(rng.integers(0, 3, size=N)).long()
    shared = {"states_view": view, "states": None, "rewards": rew, "dones": dn, "deltas": deltas,
              "actions": actions.to(dev)}
    net = build_net(model, fs, h, dev)
    if model == "GRUModel":
        shared["h_states"] = torch.from_numpy(synth.hashed_tensor((N, h), 999, 0.5)).float().to(dev)
    upd = updater_mod.Updater(net, hyps)
    res = upd.compute_gradients(shared)
    grads = {name: p.grad.detach().clone() for name, p in net.named_parameters()}
    scale = H.golden_grad_scale(g, "grad/")
    # At these sizes (reductions over up to 3e7 cancelling terms) the fp32 CPU
    # reference is itself 5e-5 .. 1.4e-3 away from the exact gradient for the conv
    # / first-GEMM parameters (tests/golden/make_fp64.py measures it per tensor and
    # stores the float64 evaluation).  No fp32 implementation with a different
    # summation order can sit within 1e-5 of such a value, so the bar is:
    #   distance(cuda, float64 oracle) <= max(1e-5, d_ref) + FACT * d_ref,   d_ref = distance(reference fp32, float64 oracle)
    # i.e. 1e-5 wherever the reference itself achieves it, and the same order as the
    # reference's own rounding error elsewhere (both are random draws of the same fp32
    # noise, so a small multiple, not 1x, is the meaningful bound); the distance to the
    # reference's fp32 output then follows from the triangle inequality (asserted).
    # Round 2: the bar is  max(1e-5, ref_own) + k * ref_own  over EVERY element (not 4,096
    # samples), k = 2 (FFMA) / 8 (3xTF32).  The measured maxima of e_full / ref_own are 2.24
    # (conv weights of the FFMA C2 run) and, for the first-layer gradients of A3C on the 3xTF32
    # path, 6.83 / 5.77 / 7.07 for three builds whose arithmetic differs only in how operand
    # low parts are rounded (truncated / nearest-away / nearest-even): the residual is a handful
    # of flipped ReLU-mask bits (pre-activations within rounding of zero) on gradients whose mass
    # sits in few samples, deterministic for a given build but not smooth in the arithmetic, so
    # k leaves ~15 % headroom over the largest draw seen (profiles/r2_parity_margins.md).
    g64 = H.load(fname.replace("full_", "fp64_"))
    # FULL-tensor yardstick: the oracle evaluated in float64 on the GPU's fp64 units (same torch
    # ops, seconds instead of minutes), pinned first to the committed CPU-generated fp64 fixture
    # (norm + 4,096 strided samples per tensor), then compared element by element.
    states_f = view.materialize(torch.float32)
    o_shared = {"states": states_f, "rewards": rew, "dones": dn, "deltas": deltas, "actions": actions}
    if model == "GRUModel":
        o_shared["h_states"] = shared["h_states"]
    o64 = O.UpdaterOracle(model, H.params_for(model, fs, h), hyps, dtype=torch.float64,
                          device=dev).compute(o_shared)
    del states_f
    report = os.environ.get("A2C_TEST_REPORT")
    for name in O.param_shapes(model, fs, h):
        key = "grad/" + name
        if key not in g:
            continue
        full64 = o64["grads"][name]
        pin = H.grad_err_compact(H.compact(full64.cpu().numpy().astype(np.float32))
                                 if full64.numel() > H.SAMPLE else full64.cpu().numpy().reshape(-1),
                                 g64["grad64/" + name], scale)
        assert pin <= 2e-7, (name, "GPU float64 oracle is not the committed float64 fixture", pin)
        mine = grads[name].cpu().numpy()
        ref_own = H.grad_err_compact(g[key], g64["grad64/" + name], scale)
        e64 = H.grad_err(mine, g64["grad64/" + name], scale)
        den = max(float(full64.abs().max()), 1e-3 * scale)
        e_full = float((grads[name].double() - full64).abs().max()) / den
        bar = max(TOL, ref_own) + FACT * ref_own
        if report:
            print("MARGIN %s prec=%d %-24s numel=%-9d e_full=%.3g e_sampled=%.3g ref_own=%.3g bar=%.3g used=%.2f %s"
                  % (fname, precision, name, full64.numel(), e_full, e64, ref_own, bar, e_full / bar,
                     "abs-1e-5" if ref_own <= TOL else "ref_own"))
        assert e64 <= bar, (name, "vs float64 oracle (sampled)", e64, "reference's own", ref_own)
        assert e_full <= bar, (name, "vs float64 oracle (every element)", e_full, "reference's own", ref_own)
        eref = H.grad_err(mine, g[key], scale)
        assert eref <= bar + ref_own, (name, "vs fp32 reference (triangle inequality)", eref, ref_own)
    del o64
    upd._optim_step()
    terms = res["loss_terms"].tolist()
    for i, key in enumerate(("Loss", "Pi_Loss", "ValLoss", "Entropy")):
        assert abs(terms[i] - float(g[key])) <= TOL * max(1.0, abs(float(g[key]))), key
    true_norm = float(g64["norm64"])
    assert abs(upd.flat.norm.item() - true_norm) <= TOL * max(1.0, true_norm)
    assert abs(upd.flat.norm.item() - float(g["GradNorm"])) <= 5e-4 * max(1.0, true_norm)
    params = dict(net.named_parameters())
    step = STEP_SCALE[hyps["optim_type"]](hyps["lr"])
    for name in O.param_shapes(model, fs, h):
        assert H.param_err_ok(params[name].detach().cpu().numpy(), g["param_after/" + name], step), name


@pytest.mark.parametrize("kind", ["RMSprop", "Adam"])
def test_fused_clip_optimizer_matches_torch(dev, kind):
    """clip_grad_norm_ + optimizer.step + zero_grad over the flat buffer vs
    torch.optim on the CPU with identical gradients, three steps."""
    from a2c_b200 import models, updater as updater_mod
    hyps = dict(is_discrete=True, lr=1e-3, optim_type=kind, max_norm=0.5)
    net = models.FCModel([2, 80, 80], 3, h_size=16).to(dev)
    ref = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.named_parameters()}
    opt = getattr(torch.optim, kind)(list(ref.values()), lr=1e-3)
    upd = updater_mod.Updater(net, hyps)
    gen = torch.Generator().manual_seed(0)
    for it in range(3):
        for k, p in net.named_parameters():
            gr = torch.randn(p.shape, generator=gen) * (0.001 if it == 1 else 1.0)
            p.grad.copy_(gr.to(dev))
            ref[k].grad = gr.clone()
        norm = torch.nn.utils.clip_grad_norm_(list(ref.values()), 0.5)
        opt.step()
        upd._optim_step()
        assert abs(upd.flat.norm.item() - float(norm)) <= 1e-5 * float(norm)
        assert float(upd.flat.grad.abs().max()) == 0.0               # zero_grad
        for k, p in net.named_parameters():
            assert H.rel_err(p.detach().cpu().numpy(), ref[k].detach().numpy()) < 5e-6, (it, k)


def test_launches_come_from_our_library(dev):
    from a2c_b200 import _lib, ops
    n0 = _lib.launch_count()
    ops.discount(torch.ones(8, device=dev), torch.zeros(8, device=dev), 0.5)
    assert _lib.launch_count() == n0 + 1

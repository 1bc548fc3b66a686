"""Pins the oracle (oracle/a2c_oracle.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

import helpers as H
from a2c_b200 import synth
from oracle import a2c_oracle as O

TOL = 1e-5   # fp32 tolerance named by BASELINE.json north_star (relative)


def test_pong_prep_bit_exact():
    g = H.load("units.npz")
    rng = np.random.default_rng(int(g["prep_frames_seed"]))
    frames = synth.make_frames(rng, 6)
    frames[0][:] = 144
    frames[1][:] = 109
    frames[2][:] = 0
    want = np.unpackbits(g["prep_out_bits"])[:6 * 6400].reshape(6, 1, 80, 80)
    got = np.stack([O.pong_prep(f) for f in frames])
    assert got.dtype == np.uint8 and np.array_equal(got, want)
    # known answers: pure backgrounds / zeros give an all-zero plane
    assert got[0].sum() == 0 and got[1].sum() == 0 and got[2].sum() == 0
    assert bool(g["prep_mutates_input"])          # reference quirk, documented
    keep = frames[3].copy()
    O.pong_prep(frames[3])
    assert np.array_equal(keep, frames[3])        # the oracle does not mutate


@pytest.mark.parametrize("fs", [2, 3])
def test_frame_stack_bit_exact(fs):
    g = H.load("units.npz")
    resets = [True, False, False, True, False, False, False]
    want = np.unpackbits(g["stack_fs%d_bits" % fs])[:len(resets) * fs * 6400]
    want = want.reshape(len(resets), fs, 80, 80)
    st = O.FrameStack(fs)
    k = 100
    for i, rs in enumerate(resets):
        obs = np.full((1, 80, 80), (i * 7 + 3) % 2, dtype=np.uint8)
        obs[0, i, i] = 1 - obs[0, i, i]
        if rs:   # the fake env's reset() of make_golden
            k += 1
            obs = np.full((1, 80, 80), k % 2, dtype=np.uint8)
        got = st.push(obs, rs)
        assert np.array_equal(got, want[i]), i


def test_sample_action_bit_exact():
    g = H.load("units.npz")
    got = O.sample_action(g["sa_pi"], g["sa_u"]).numpy()
    assert np.array_equal(got, g["sa_act"])
    assert g["sa_act"][0] == -1                   # the "never reached" edge (utils.py:56)


def test_discount_bit_exact():
    g = H.load("units.npz")
    assert np.array_equal(g["disc_known"], np.array([1.75, 1.5, 1.0], dtype=np.float32))
    assert np.array_equal(O.discount([1., 1, 1], [0., 0, 1], .5).numpy(), g["disc_known"])
    assert np.array_equal(O.discount(g["disc_arr"], g["disc_dones"], 0.99).numpy(),
                          g["disc_out_g99"])
    assert np.array_equal(O.discount(g["disc_arr"], g["disc_dones"], 0.99 * 0.98).numpy(),
                          g["disc_out_gl"])


SMALL = [("c1_fc.npz", "FCModel", 4, 15, 2, 256),
         ("small_a3c.npz", "A3CModel", 3, 6, 2, 64),
         ("small_conv.npz", "ConvModel", 3, 6, 2, 32),
         ("small_gru.npz", "GRUModel", 3, 6, 2, 64),
         ("small_gru_bptt.npz", "GRUModel", 3, 6, 2, 64),
         ("small_fc_adam_nstep.npz", "FCModel", 3, 6, 2, 64),
         ("small_fc_fs3.npz", "FCModel", 3, 6, 3, 64)]


@pytest.mark.parametrize("fname,model,R,T,fs,h", SMALL)
def test_runner_and_updater_against_reference(fname, model, R, T, fs, h):
    """Runner.rollout + Updater.update_model, two consecutive updates (state,
    h and optimizer state carried), reference vs oracle."""
    g = H.load(fname)
    hyps = H.hyps_of(g)
    trajs = H.small_trajectories(g, R, T)
    params = H.params_for(model, fs, h)
    upd = O.UpdaterOracle(model, params, hyps)
    run = O.RunnerOracle(model, fs, hyps["gamma"], h)
    N = R * T
    for k, traj in enumerate(trajs):
        datas = run.rollout(upd.p, traj)
        pre = "u%d_" % k
        # stage 1: stacked states bit-exact
        assert np.array_equal(datas["states"].numpy().astype(np.uint8),
                              H.golden_states(g, k, N, fs))
        # actions bit-exact; bookkeeping within fp32 tolerance (forwards are
        # batched here and batch-1 in the reference)
        assert np.array_equal(datas["actions"].numpy(), g[pre + "actions"])
        assert np.array_equal(datas["dones"].numpy(), g[pre + "dones"])
        for key in ("rewards", "deltas") + (("h_states",) if "h_states" in datas else ()):
            assert H.rel_err(datas[key].numpy(), g[pre + key]) < TOL, key
        # feed the reference's own datas so update parity is isolated from acting noise
        shared = {key: torch.from_numpy(g[pre + key]) for key in
                  ("rewards", "dones", "deltas", "actions")}
        shared["states"] = datas["states"]
        if "h_states" in datas:
            shared["h_states"] = torch.from_numpy(g[pre + "h_states"])
        out = upd.step(shared)
        for key in ("Loss", "Pi_Loss", "ValLoss", "Entropy", "GradNorm"):
            assert abs(out[key] - float(g[pre + key])) <= TOL * max(1.0, abs(float(g[pre + key]))), key
        for name, gr in out["grads"].items():
            if gr is None:
                assert pre + "grad/" + name not in g, name
                continue
            assert H.rel_err(H.compact(gr.numpy()), g[pre + "grad/" + name]) < TOL, name
        for name, q in upd.p.items():
            assert H.rel_err(H.compact(q.detach().numpy()), g[pre + "param_after/" + name]) < TOL, name


def test_a3c_value_head_is_detached():
    """a2c/models.py:85: the value loss must not reach the encoder."""
    p = H.params_for("A3CModel", 2, 64)
    p = {k: v.clone().requires_grad_(True) for k, v in p.items()}
    x = torch.rand(5, 2, 80, 80)
    v, _ = O.model_forward("A3CModel", p, x)
    v.sum().backward()
    assert p["convs.0.0.weight"].grad is None and p["proj_matrx.weight"].grad is None
    assert p["value.weight"].grad is not None


def test_entropy_of_uniform_policy_is_ln3():
    hy = dict(entr_coef=1.0, pi_coef=1.0, val_coef=0.5)
    N = 8
    _, _, _, entr = O.a2c_loss(torch.zeros(N, 3), torch.zeros(N, 1), torch.zeros(N).long(),
                               torch.zeros(N), torch.zeros(N), hy)
    assert abs(float(entr) - np.log(3.0)) < 1e-6


@pytest.mark.needs_reference
def test_oracle_models_match_live_reference():
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import _ref_loader
    if not _ref_loader.reference_available():
        pytest.skip("reference not mounted")
    ref = _ref_loader.load_reference()
    torch.manual_seed(3)
    for model in ("FCModel", "A3CModel", "ConvModel", "GRUModel"):
        h = 32
        net = getattr(ref.models, model)([2, 80, 80], 3, h_size=h, bnorm=False, is_discrete=True)
        sd = net.state_dict()
        p = {k: sd[k] for k in O.param_shapes(model, 2, h)}
        x = (torch.rand(4, 2, 80, 80) < 0.1).float()
        if model == "GRUModel":
            h0 = torch.randn(4, h)
            want = net(x, h0)
            got = O.model_forward(model, p, x, h0)
        else:
            want = net(x)
            got = O.model_forward(model, p, x)
        for a, b in zip(got, want):
            assert torch.equal(a, b), model


@pytest.mark.needs_reference
def test_other_preprocessors_match_live_reference():
    """normalize / null / snake / breakout prep of the oracle against the unmodified reference
    functions (a2c/preprocessing.py:4-9,19-32; rgb2grey stubbed as the identity it is on a 2-D
    array in the skimage releases that export it).  Skipped where /root/reference is absent."""
    import _ref_loader
    if not _ref_loader.reference_available():
        pytest.skip("reference tree not mounted")
    ref = _ref_loader.load_reference().preprocessing
    rng = np.random.default_rng(4)
    for k in range(4):
        f = rng.integers(0, 256, size=(210, 160, 3), dtype=np.uint8)
        f[::3, ::5, 0] = 1
        f[::7, ::2, k % 3] = 255
        for name in ("normalize_prep", "null_prep", "snake_prep", "breakout_prep"):
            a, b = getattr(O, name)(f.copy()), getattr(ref, name)(f.copy())
            assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), name


def test_other_preprocessors_known_answers():
    f = np.zeros((210, 160, 3), dtype=np.uint8)
    assert float(O.normalize_prep(f).max()) == -3.0 and O.normalize_prep(f).shape == (1, 210, 160, 3)
    f[...] = 255
    assert float(O.normalize_prep(f).min()) == 3.0
    assert float(O.snake_prep(f).max()) == np.float32(.33)          # the blue rule is applied last
    f[..., 2] = 0
    assert float(O.snake_prep(f).max()) == 0.0                      # green == 255 -> 0
    assert O.breakout_prep(f).shape == (1, 80, 72)

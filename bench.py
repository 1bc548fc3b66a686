#!/usr/bin/env python
"""bench.py -- transitions/sec per A2C update (prep + returns + loss + fwd/bwd +
clip/optimizer) on synthetic Pong-shaped trajectories.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3|c4|c1]
                  [--impl ours|reference] [--precision 0|1|2]

One "step" = one pass of the hot path over one batch of R x T transitions:
stage 1 (prep + stack of the R*T new frames into the plane ring), stage 2
(bootstrap fold, deltas, reverse scans), forward, stage 3 (fused loss), backward,
gradient all-reduce (N > 1), clip + optimizer step.  N > 1 is launched by
torchrun, one rank per GPU; every rank owns R rollout slots (weak scaling) and
gradients / advantage statistics are all-reduced over NCCL.

`value` is measured with the inputs resident in HBM; `e2e` goes through the
same public API (PlaneRing.push_frames / Runner-style buffers / Updater.update_model)
with the step's inputs in pinned HOST memory and a device->host read of the
loss every step.  `--impl reference` times the CPU oracle port (the reference
itself cannot travel to the GPU box) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (model, R per GPU, T, fs, description)
    "c1": ("FCModel", 4, 15, 2, "C1 fc n_envs=4 n_tsteps=15"),
    "c2": ("ConvModel", 256, 20, 2, "C2 conv n_envs=256 n_tsteps=20 n_frame_stack=2 fp32"),
    "c3": ("A3CModel", 1024, 20, 2, "C3 a3c n_envs=1024 n_tsteps=20 RMSprop"),
    "c4": ("GRUModel", 512, 32, 2, "C4 gru BPTT n_envs=512 n_tsteps=32"),
}
HYPS = dict(gamma=0.99, lambda_=0.98, entr_coef=0.005, val_coef=0.5, pi_coef=1.0, max_norm=0.5,
            lr=1e-4, optim_type="RMSprop", norm_advs=True, use_nstep_rets=False, use_bptt=False,
            is_discrete=True, env_type="Pong-v0", action_shift=1, render=False, h_size=256)
METRIC = "transitions/sec per A2C update (prep+returns+loss+bwd)"
FRAME_BYTES = 210 * 160 * 3
ALG_BYTES_PREP = 44800            # 80 rows x 480 B read + 6400 B written


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------- our arm
def synth_device_batch(R, T, dev, seed):
    """Synthetic step inputs generated on the host then placed as requested.
    Frames: background 144/109 with 10 % uniform noise (all pong_prep branches)."""
    g = torch.Generator().manual_seed(seed)
    n = R * (T + 1)
    frames = torch.empty((n, FRAME_BYTES), dtype=torch.uint8)
    chunk = 512
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        noise = torch.randint(0, 256, (e - s, FRAME_BYTES), dtype=torch.uint8, generator=g)
        bg = torch.where(torch.rand((e - s, 1), generator=g) < 0.5, 144, 109).to(torch.uint8)
        sel = torch.rand((e - s, FRAME_BYTES), generator=g) < 0.1
        frames[s:e] = torch.where(sel, noise, bg)
    frames = frames.view(R, T + 1, 210, 160, 3)
    rew = torch.where(torch.rand((R, T), generator=g) < 0.1,
                      torch.randint(0, 2, (R, T), generator=g).float() * 2 - 1, torch.zeros(R, T))
    reset = torch.zeros((R, T + 1), dtype=torch.uint8)
    reset[:, 1:] = (torch.rand((R, T), generator=g) < 0.02).to(torch.uint8)
    reset[:, 0] = 1
    done = ((reset[:, 1:] != 0) | (rew != 0)).float()
    values = torch.randn((R, T + 1), generator=g)
    actions = torch.randint(0, 3, (R * T,), generator=g)
    return {"frames": frames, "reset": reset, "rewards": rew.reshape(-1).contiguous(),
            "dones": done.reshape(-1).contiguous(), "values": values, "actions": actions}


def run_ours(args):
    import torch.distributed as dist

    from a2c_b200 import _lib, models, ops, updater

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    model, R, T, fs, desc = WORKLOADS[args.workload]
    hyps = dict(HYPS, n_tsteps=T, n_rollouts=R, n_envs=R, n_frame_stack=fs, model=model,
                use_bptt=(model == "GRUModel"))
    N = R * T
    ops.set_precision(args.precision)

    torch.manual_seed(0)
    net = models.get_model_class(model)([fs, 80, 80], 3, h_size=hyps["h_size"]).to(dev)
    upd = updater.Updater(net, hyps)
    if world > 1:                       # identical replicas: broadcast rank 0's init
        dist.broadcast(upd.flat.flat, 0)

    host = synth_device_batch(R, T, dev, 1234 + rank)
    # device-resident copies (for `value`) and pinned host copies (for `e2e`)
    d_frames = host["frames"].to(dev)
    d_reset = host["reset"].to(dev)
    d_in = {k: host[k].to(dev) for k in ("rewards", "dones", "values", "actions")}
    h_frames = host["frames"][:, 1:].contiguous().pin_memory()       # the T new frames per slot
    h_reset = host["reset"][:, 1:].contiguous().pin_memory()
    h_in = {k: host[k].pin_memory() for k in ("rewards", "dones", "values", "actions")}
    h_states = None
    if net.is_recurrent:
        h_states = torch.randn((N, hyps["h_size"]), device=dev) * 0.5

    ring = ops.PlaneRing(R, T, fs, dev)
    ring.push_frames(d_frames[:, :1].contiguous(), d_reset[:, :1].contiguous())   # bookmark state
    d_new_frames = d_frames[:, 1:].contiguous()
    d_new_reset = d_reset[:, 1:].contiguous()
    del d_frames
    rew_w = torch.empty(N, device=dev)
    done_w = torch.empty(N, device=dev)
    val_w = torch.empty((R, T + 1), device=dev)
    act_w = torch.empty(N, dtype=torch.int64, device=dev)
    step_no = [0]

    def step(from_host):
        """One pass of the hot path.  Returns the info dict (forces the D2H of the loss)."""
        if step_no[0] > 0:
            ring.carry()
        step_no[0] += 1
        if from_host:
            ring.push_frames(h_frames, h_reset, first_frame_index=1)          # zero-copy TMA over PCIe
            rew_w.copy_(h_in["rewards"], non_blocking=True)
            done_w.copy_(h_in["dones"], non_blocking=True)
            val_w.copy_(h_in["values"], non_blocking=True)
            act_w.copy_(h_in["actions"], non_blocking=True)
        else:
            ring.push_frames(d_new_frames, d_new_reset, first_frame_index=1)
            rew_w.copy_(d_in["rewards"]); done_w.copy_(d_in["dones"])
            val_w.copy_(d_in["values"]); act_w.copy_(d_in["actions"])
        deltas, advs, rets, _ = ops.delta_scan(rew_w, done_w, val_w, hyps["gamma"], hyps["lambda_"])
        shared = {"states_view": ring.view(T), "states": None, "rewards": rew_w, "dones": done_w,
                  "deltas": deltas, "actions": act_w}
        if h_states is not None:
            shared["h_states"] = h_states
        return upd.update_model(shared)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(from_host, steps, profile=False):
        barrier()
        if profile:
            _lib.PROFILE = {}
        launches0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        info = None
        for _ in range(steps):
            info = step(from_host)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        prof = _lib.PROFILE
        _lib.PROFILE = None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count() - launches0, info, prof

    for _ in range(args.warmup):
        step(False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, info, prof = timed(False, args.steps, profile=True)
    clocks = sampler.stop() if rank == 0 else None
    # end to end: host inputs, D2H of the loss each step
    for _ in range(max(1, min(2, args.warmup))):
        step(True)
    ms_e2e, _, info_e2e, _ = timed(True, args.steps)

    # stage 1 alone, launches back to back (inside a step it is followed by a host sync, so its
    # per-call event time there includes the launch gap): 516 MB of frames per launch > L2
    barrier()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(20):
        ring.push_frames(d_new_frames, d_new_reset, first_frame_index=1)
    s1.record()
    barrier()
    stage1_ms = s0.elapsed_time(s1) / 20

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    total_tr = N * world * args.steps
    value = total_tr / (ms / 1e3)
    e2e_value = total_tr / (ms_e2e / 1e3)
    pk = peaks()
    roof, table = roofline_from_profile(prof, args.steps, R, T, pk, stage1_ms, args.precision)
    h2d = (R * T * 38400 + h_reset.numel()
           + sum(h_in[k].numel() * h_in[k].element_size() for k in h_in))
    out = {
        "metric": METRIC, "value": value, "unit": "transitions/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "tf32" if args.precision == 2 else "f32",
        "data": "synthetic",
        "config": {"workload": desc, "model": model, "n_envs_per_gpu": R, "n_tsteps": T,
                   "transitions_per_step": N * world, "parallelism": "env-sharded dp%d" % world,
                   "l2": "inputs larger than L2 (%.0f MB of frames per step)" % (R * T * FRAME_BYTES / 1e6),
                   "optim": hyps["optim_type"],
                   "arithmetic": {0: "fp32 FFMA kernels", 1: "fp32 results via the 3xTF32 split on tcgen05 "
                                  "(fp32 accumulate, 1e-5 parity)",
                                  2: "single-pass TF32 on tcgen05 (2e-3)"}[args.precision]},
        "e2e": {"value": e2e_value, "unit": "transitions/s", "ms_per_step": ms_e2e / args.steps,
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 20,
                "h2d_note": "frames are read zero-copy by the stage-1 TMA kernel from pinned host "
                            "memory: only the 80 selected rows (38,400 B) of each 100,800 B frame "
                            "cross PCIe"},
        "gpu_launches": int(launches),
        "roofline": roof, "kernel_table": table, "clocks": clocks,
        "loss": info["Loss"], "peaks_source": pk["src"],
    }
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.workload, budget_s=20.0)
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def ncu_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/*.json)."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    return json.load(open(path)) if os.path.isfile(path) else {}


def roofline_from_profile(prof, steps, R, T, pk, stage1_ms, precision):
    """prof: {key: [(start_evt, end_evt, work), ...]} recorded around each C-ABI
    call in the timed region (events on the launching stream).  Picks the entry
    with the largest total time as the dominant kernel."""
    if not prof:
        return None, None
    rows = []
    for key, recs in prof.items():
        tot = sum(s.elapsed_time(e) for s, e, _ in recs)
        work = sum(w for _, _, w in recs)
        rows.append((tot, key, len(recs), work))
    rows.sort(reverse=True)
    total_ms = sum(r[0] for r in rows)
    table = [{"call": k, "launches": n, "ms_per_step": t / steps, "share": t / total_ms}
             for t, k, n, _ in rows[:14]]
    t, key, n, work = rows[0]
    kind = key.split("|")[1] if "|" in key else "flops"
    per_launch_ms = t / n
    traffic = ncu_traffic()
    if kind == "bytes":
        ach = work / n / (per_launch_ms / 1e3) / 1e9
        roof = {"kernel": key, "bound": "hbm", "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": ach / pk["hbm_gbs"], "traffic": traffic.get(key)}
    else:
        ach = work / n / (per_launch_ms / 1e3) / 1e12
        roof = {"kernel": key, "bound": "tensor", "achieved": ach, "peak": pk["bf16_tflops"],
                "unit": "TFLOP/s", "frac": ach / pk["bf16_tflops"], "traffic": traffic.get(key),
                "note": "achieved = algorithmic fp32 flops (2*M*N*K) / launch time; peak = measured "
                        "sustained cuBLAS bf16.  The fp32-parity path issues 3 TF32 MMAs per product "
                        "(TF32 rate = bf16 / 2), so the tensor pipe itself runs at 6x this fraction"}
        if precision == 1:
            roof["frac_of_tf32_issued"] = 6.0 * ach / pk["bf16_tflops"]
    roof["launch_ms"] = per_launch_ms
    roof["peak_source"] = pk["src"]
    # stage 1 (the memory-bound stage of the north star) is always reported, timed alone
    a = R * T * ALG_BYTES_PREP / (stage1_ms / 1e3) / 1e9
    roof["stage1_hbm"] = {"kernel": "a2c_obs_prep_stack", "achieved": a, "peak": pk["hbm_gbs"],
                          "frac": a / pk["hbm_gbs"], "unit": "GB/s", "launch_ms": stage1_ms,
                          "bytes_per_transition": ALG_BYTES_PREP,
                          "traffic": traffic.get("a2c_obs_prep_stack")}
    return roof, table


# ------------------------------------------------------------------ reference arm
def cpu_sample_inputs(workload, Rs, seed=1234):
    from a2c_b200 import synth
    model, R, T, fs, desc = WORKLOADS[workload]
    traj = synth.make_trajectory(Rs, T, seed=seed)
    return model, T, fs, traj


def cpu_step(workload, Rs, state):
    """One pass of the reference algorithm (oracle port) over Rs slots: per-frame
    pong_prep + frame stack + fp32 state store, bookkeeping, Updater step."""
    from oracle import a2c_oracle as O
    model, T, fs, traj = state["inputs"]
    t0 = time.perf_counter()
    states = O.prep_and_stack(traj["frames"][:, :T], traj["reset"][:, :T], fs)
    st = torch.from_numpy(states.reshape(Rs * T, fs, 80, 80)).float()
    t1 = time.perf_counter()
    done = ((traj["reset"][:, 1:] != 0) | (traj["rewards"] != 0)).astype(np.float32)
    r2, d2, deltas = O.rollout_bookkeeping(traj["rewards"], done, traj["values"], HYPS["gamma"])
    shared = {"states": st, "rewards": r2.reshape(-1), "dones": d2.reshape(-1),
              "deltas": deltas.reshape(-1), "actions": state["actions"]}
    if model == "GRUModel":
        shared["h_states"] = state["h_states"]
    state["upd"].step(shared)
    t2 = time.perf_counter()
    return t1 - t0, t2 - t1


def cpu_state(workload, Rs):
    from a2c_b200 import synth
    from oracle import a2c_oracle as O
    model, T, fs, traj = cpu_sample_inputs(workload, Rs)
    hyps = dict(HYPS, n_tsteps=T, n_rollouts=Rs, n_frame_stack=fs, use_bptt=(model == "GRUModel"))
    shapes = O.param_shapes(model, fs=fs, h=hyps["h_size"])
    params = {k: torch.from_numpy(v) for k, v in synth.hashed_state_dict(shapes).items()}
    rng = np.random.default_rng(7)
    st = {"inputs": (model, T, fs, traj), "upd": O.UpdaterOracle(model, params, hyps),
          "actions": torch.from_numpy(rng.integers(0, 3, size=Rs * T)).long()}
    if model == "GRUModel":
        st["h_states"] = torch.randn(Rs * T, hyps["h_size"]) * 0.5
    return st


CPU_SAMPLE_SLOTS = {"c1": 4, "c2": 16, "c3": 128, "c4": 16}


def cpu_baseline(workload, budget_s=20.0, steps=None, warmup=1):
    torch.set_num_threads(os.cpu_count())
    model, R, T, fs, desc = WORKLOADS[workload]
    Rs = min(R, CPU_SAMPLE_SLOTS[workload])
    st = cpu_state(workload, Rs)
    for _ in range(warmup):
        cpu_step(workload, Rs, st)
    times, t_start = [], time.perf_counter()
    while True:
        tp, tu = cpu_step(workload, Rs, st)
        times.append((tp, tu))
        if steps is not None:
            if len(times) >= steps:
                break
        elif time.perf_counter() - t_start > budget_s or len(times) >= 50:
            break
    best = min(a + b for a, b in times)
    return {"value": Rs * T / best, "unit": "transitions/s", "cores": os.cpu_count(), "kind": "port",
            "sample": "%d of %d rollout slots (%d transitions) per pass: per-frame pong_prep + stack + "
                      "fp32 state store, bookkeeping, full Updater step (torch-CPU, %d threads); best "
                      "of %d passes" % (Rs, R, Rs * T, torch.get_num_threads(), len(times)),
            "prep_s": min(a for a, _ in times), "update_s": min(b for _, b in times)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, R, T, fs, desc = WORKLOADS[args.workload]
    torch.set_num_threads(os.cpu_count())
    Rs = min(R, CPU_SAMPLE_SLOTS[args.workload])
    st = cpu_state(args.workload, Rs)
    for _ in range(args.warmup):
        cpu_step(args.workload, Rs, st)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(args.workload, Rs, st)
    dt = time.perf_counter() - t0
    value = Rs * T * args.steps / dt
    sample = ("%d of %d rollout slots (%d transitions) per step; oracle port of the reference's "
              "numpy/torch-CPU Runner+Updater arithmetic, %d host threads"
              % (Rs, R, Rs * T, torch.get_num_threads()))
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s",
           "n_gpus": int(os.environ.get("WORLD_SIZE", "1")), "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": {"workload": desc, "model": model, "n_envs_per_gpu": R, "n_tsteps": T,
                      "sample_slots": Rs},
           "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": os.cpu_count(),
                            "kind": "port", "sample": sample},
           "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("A2C_PRECISION", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "RANK" not in os.environ:
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
               "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
               "--master-port", "29511", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()

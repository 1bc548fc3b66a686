#!/usr/bin/env python
"""bench.py -- transitions/sec per A2C update (prep + returns + loss + fwd/bwd +
clip/optimizer) on synthetic Pong-shaped trajectories.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5]
                  [--impl ours|reference] [--precision 0|1|2] [--no-extras]

One "step" = one pass of the hot path over one batch of R x T transitions:
stage 1 (prep + stack of the R*T new frames into the plane ring), stage 2
(bootstrap fold, deltas, reverse scans), forward, stage 3 (fused loss), backward,
gradient all-reduce (N > 1), clip + optimizer step.

Default workload: N = 1 -> C2 (conv, 256 envs x 20 steps: the configuration the
metric is quoted on); N > 1 -> C5 (conv, 4096 envs in total, env-sharded 4096/N
per rank, STRONG scaling, gradients all-reduced over NCCL).  The N = 1 line also
carries one record per remaining BASELINE config (`workloads`: c1, c3, c4 and c5
on one GPU, the latter being the denominator of the C5 scaling curve).

`value` is measured with the inputs resident in HBM and the per-call profiler
OFF; `e2e` goes through the same public API with the step's inputs in pinned
HOST memory (FrameIngest: one strided DMA per step, pipelined under the previous
step's kernels, then stage 1 from the staging buffer) and a device->host read of
the loss every step.  `--impl reference` times the CPU oracle port (the
reference itself cannot travel to the GPU box) on the same configuration.
"""
import argparse
import copy
import json
import os
import subprocess
import sys
import threading
import time

# NOTE: torch's default (non-expandable) allocator segments are used on purpose.  Measured on B200
# (GRU config, 400 launches per step): PYTORCH_CUDA_ALLOC_CONF=expandable_segments:True makes step
# times erratic (57 ... 137 ms, GPU idle while segments are re-mapped) where the default allocator
# gives 57.0 - 57.4 ms; the first steps of a fresh process are kept out of the timed region by
# Workload.settle() instead.

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "pytorch-a2c_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # name: (model, total rollout slots, T, fs, scaling, description)
    "c1": ("FCModel", 4, 15, 2, "weak", "C1 fc n_envs=4 n_tsteps=15"),
    "c2": ("ConvModel", 256, 20, 2, "weak", "C2 conv n_envs=256 n_tsteps=20 n_frame_stack=2 fp32"),
    "c3": ("A3CModel", 1024, 20, 2, "weak", "C3 a3c n_envs=1024 n_tsteps=20 RMSprop"),
    "c4": ("GRUModel", 512, 32, 2, "weak", "C4 gru BPTT n_envs=512 n_tsteps=32"),
    "c5": ("ConvModel", 4096, 20, 2, "strong", "C5 conv n_envs=4096 n_tsteps=20 env-sharded over the GPUs"),
}
HYPS = dict(gamma=0.99, lambda_=0.98, entr_coef=0.005, val_coef=0.5, pi_coef=1.0, max_norm=0.5,
            lr=1e-4, optim_type="RMSprop", norm_advs=True, use_nstep_rets=False, use_bptt=False,
            is_discrete=True, env_type="Pong-v0", action_shift=1, render=False, h_size=256)
METRIC = "transitions/sec per A2C update (prep+returns+loss+bwd)"
FRAME_BYTES = 210 * 160 * 3
ALG_BYTES_PREP = 44800            # 80 rows x 480 B read + 6400 B written
ARITH = {0: "fp32 FFMA kernels",
         1: "fp32 results via the 3xTF32 split on tcgen05 (fp32 accumulate, 1e-5 parity)",
         2: "single-pass TF32 on tcgen05 (2e-3)"}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1400.0, "src": "fallback"}


def slots_per_rank(workload, world):
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    if scaling == "strong":
        assert R % world == 0, "C5 shards 4096 envs evenly"
        return R // world
    return R


def config_of(workload, world, precision):
    """The `config` object; identical for our arm and the reference arm."""
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    Rr = slots_per_rank(workload, world)
    return {"workload": desc, "model": model, "n_envs_per_gpu": Rr, "n_tsteps": T,
            "transitions_per_step": Rr * T * world, "parallelism": "env-sharded dp%d" % world,
            "l2": "inputs larger than L2 (%.0f MB of frames per step per GPU)" % (Rr * T * FRAME_BYTES / 1e6),
            "optim": HYPS["optim_type"], "arithmetic": ARITH[precision]}


# --------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index, interval_ms=200):
        self.index, self.rows, self.proc, self.interval_ms = index, [], None, interval_ms

    def start(self):
        if self.interval_ms <= 0:
            return
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", str(self.interval_ms)],
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
def synth_batch(R, T, dev, seed, base_slots=512):
    """Synthetic step inputs: frames with background 144/109 and 10 %
    uniform noise (all three pong_prep branches), rewards in {-1,0,+1} with P(!=0) = 0.1, env
    resets ~ Bernoulli(0.02), N(0,1) rollout values, uniform actions.  Drawn with the device RNG
    and moved to pinned host memory BEFORE any timing; more than `base_slots` slots tile the
    first `base_slots` (distinct memory, repeated content)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    Rb = min(R, base_slots)
    n = Rb * (T + 1)
    frames = torch.empty((n, FRAME_BYTES), dtype=torch.uint8, device=dev)
    chunk = 2048
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        noise = torch.randint(0, 256, (e - s, FRAME_BYTES), dtype=torch.uint8, device=dev, generator=g)
        bg = torch.where(torch.rand((e - s, 1), device=dev, generator=g) < 0.5, 144, 109).to(torch.uint8)
        sel = torch.rand((e - s, FRAME_BYTES), device=dev, generator=g) < 0.1
        frames[s:e] = torch.where(sel, noise, bg)
        del noise, sel
    frames = frames.view(Rb, T + 1, 210, 160, 3)
    if R > Rb:
        reps = (R + Rb - 1) // Rb
        frames = frames.repeat(reps, 1, 1, 1, 1)[:R].contiguous()
    rew = torch.where(torch.rand((R, T), device=dev, generator=g) < 0.1,
                      torch.randint(0, 2, (R, T), device=dev, generator=g).float() * 2 - 1,
                      torch.zeros((R, T), device=dev))
    reset = torch.zeros((R, T + 1), dtype=torch.uint8, device=dev)
    reset[:, 1:] = (torch.rand((R, T), device=dev, generator=g) < 0.02).to(torch.uint8)
    reset[:, 0] = 1
    done = ((reset[:, 1:] != 0) | (rew != 0)).float()
    values = torch.randn((R, T + 1), device=dev, generator=g)
    actions = torch.randint(0, 3, (R * T,), device=dev, generator=g)
    return {"frames": frames, "reset": reset, "rewards": rew.reshape(-1).contiguous(),
            "dones": done.reshape(-1).contiguous(), "values": values, "actions": actions}


def pinned_copy(t):
    h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    h.copy_(t)
    return h


class Workload:
    """Resident state of one benchmark configuration on this rank."""

    def __init__(self, name, world, rank, dev, precision, dist, data_parallel=True, graph=-1):
        from a2c_b200 import models, ops, updater
        self.ops, self.dist = ops, dist
        self.name, self.world, self.rank, self.dev = name, world, rank, dev
        model, _, T, fs, scaling, desc = WORKLOADS[name]
        R = slots_per_rank(name, world)
        self.model, self.R, self.T, self.fs, self.N = model, R, T, fs, R * T
        # CUDA graph replay of stage 2 + update: a win for launch-bound small batches (C1: 1.8 -> 0.3 ms per
        # update), a small loss for the big ones (C2: 42.7 -> 44.5 ms), so "auto" = small batches only
        self.graph = (world == 1 and R * T <= 2048) if graph < 0 else bool(graph)
        self.hyps = dict(HYPS, n_tsteps=T, n_rollouts=R, n_envs=R, n_frame_stack=fs, model=model,
                         use_bptt=(model == "GRUModel"))
        ops.set_precision(precision)
        torch.manual_seed(0)
        self.net = models.get_model_class(model)([fs, 80, 80], 3, h_size=HYPS["h_size"]).to(dev)
        self.upd = updater.Updater(self.net, self.hyps, data_parallel=data_parallel)
        if world > 1:                       # identical replicas: broadcast rank 0's init
            dist.broadcast(self.upd.flat.flat, 0)
        b = synth_batch(R, T, dev, 1234 + rank)
        self.d_new_frames = b["frames"][:, 1:].contiguous()
        self.d_new_reset = b["reset"][:, 1:].contiguous()
        self.d_in = {k: b[k] for k in ("rewards", "dones", "values", "actions")}
        self.h_frames = pinned_copy(self.d_new_frames)
        self.h_reset = pinned_copy(self.d_new_reset)
        self.h_in = {k: pinned_copy(v) for k, v in self.d_in.items()}
        self.h_states = None
        if self.net.is_recurrent:
            self.h_states = torch.randn((self.N, HYPS["h_size"]), device=dev) * 0.5
        self.ring = ops.PlaneRing(R, T, fs, dev)
        self.ring.push_frames(b["frames"][:, :1].contiguous(), b["reset"][:, :1].contiguous())
        del b
        self.ingest = ops.FrameIngest(R, T, dev, depth=2)
        self.rew_w = torch.empty(self.N, device=dev)
        self.done_w = torch.empty(self.N, device=dev)
        self.val_w = torch.empty((R, T + 1), device=dev)
        self.act_w = torch.empty(self.N, dtype=torch.int64, device=dev)
        self.steps_done = 0

    def h2d_bytes(self):
        return int(self.ingest.bytes_per_prefetch
                   + sum(v.numel() * v.element_size() for v in self.h_in.values()))

    def step(self, from_host, prefetch_next=False, sync=True):
        """One pass of the hot path.  sync=True returns the reference's info dict (one host
        synchronisation: the D2H of the loss); sync=False only enqueues and returns the device
        tensor of the five scalars."""
        ops, hyps, ring = self.ops, self.hyps, self.ring
        if self.steps_done > 0:
            ring.carry()
        self.steps_done += 1
        if from_host:
            # The H2D copy engine is one FIFO for all streams: this step's small copies go in
            # BEFORE the next step's big frame DMA, or the compute stream would wait for them
            # behind ~0.4 GB of frames (measured: +7 ms per C2 step).
            self.rew_w.copy_(self.h_in["rewards"], non_blocking=True)
            self.done_w.copy_(self.h_in["dones"], non_blocking=True)
            self.val_w.copy_(self.h_in["values"], non_blocking=True)
            self.act_w.copy_(self.h_in["actions"], non_blocking=True)
            self.ingest.commit(ring, first_frame_index=1)
            if prefetch_next:                 # next step's DMA runs under this step's kernels
                self.ingest.prefetch(self.h_frames, self.h_reset)
        else:
            ring.push_frames(self.d_new_frames, self.d_new_reset, first_frame_index=1)
            self.rew_w.copy_(self.d_in["rewards"]); self.done_w.copy_(self.d_in["dones"])
            self.val_w.copy_(self.d_in["values"]); self.act_w.copy_(self.d_in["actions"])
        if not sync or self.graph:
            # stage 2 + update on the persistent rollout buffers; with graph on the Updater captures this
            # call into a CUDA graph on its second use and replays it afterwards (one launch per update)
            out5 = self.upd.update_from_rollout(ring.view(self.T), self.rew_w, self.done_w, self.val_w,
                                                self.act_w, self.h_states, graph=self.graph)
            if not sync:
                return out5
            return dict(zip(("Loss", "Pi_Loss", "ValLoss", "Entropy", "GradNorm"), out5.tolist()))
        deltas, advs, rets, stats = ops.delta_scan(self.rew_w, self.done_w, self.val_w, hyps["gamma"],
                                                   hyps["lambda_"])
        shared = {"states_view": ring.view(self.T), "states": None, "rewards": self.rew_w,
                  "dones": self.done_w, "deltas": deltas, "actions": self.act_w,
                  "scan": {"n": self.N, "advs": advs, "returns": rets, "stats": stats,
                           "gamma": hyps["gamma"], "lambda_": hyps["lambda_"]}}
        if self.h_states is not None:
            shared["h_states"] = self.h_states
        return self.upd.update_model(shared)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def timed(self, from_host, steps, profile=False):
        from a2c_b200 import _lib
        self.barrier()
        graph0 = self.graph
        if profile:
            _lib.PROFILE = {}
            self.graph = False                 # per-call events need the eager path
        launches0 = _lib.launch_count() + self.upd.graph_replayed_kernels
        mallocs0 = torch.cuda.memory_stats(self.dev).get("num_device_alloc", 0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e0.record()
        info = None
        if from_host:
            self.ingest.prefetch(self.h_frames, self.h_reset)       # step 0's copy: inside the timed region
        # The host only ENQUEUES inside the loop (the boxes of this pool show random 50-80 ms
        # host stalls; with one synchronisation per step each of them idled the GPU).  Every step's
        # five result scalars are still read back: a non-blocking D2H copy into pinned memory per
        # step, consumed LAG steps later, so the queue stays at most LAG steps deep.
        LAG = 3
        if not hasattr(self, "_res_host"):
            self._res_host = [torch.empty(5, dtype=torch.float32, pin_memory=True) for _ in range(LAG)]
        res_ev = [torch.cuda.Event() for _ in range(LAG)]
        for i in range(steps):
            if i >= LAG:
                res_ev[i % LAG].synchronize()
                info = self._res_host[i % LAG].tolist()
            out5 = self.step(from_host, prefetch_next=(i + 1 < steps), sync=False)
            self._res_host[i % LAG].copy_(out5, non_blocking=True)
            res_ev[i % LAG].record()
            marks[i].record()
        e1.record()
        self.barrier()
        info = dict(zip(("Loss", "Pi_Loss", "ValLoss", "Entropy", "GradNorm"),
                        self._res_host[(steps - 1) % LAG].tolist()))
        ms = e0.elapsed_time(e1)
        self.last_step_ms = [round(a.elapsed_time(b), 3) for a, b in zip([e0] + marks[:-1], marks)]
        self.last_device_allocs = torch.cuda.memory_stats(self.dev).get("num_device_alloc", 0) - mallocs0
        prof = _lib.PROFILE
        _lib.PROFILE = None
        self.graph = graph0
        if self.world > 1:
            t = torch.tensor([ms], device=self.dev)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, _lib.launch_count() + self.upd.graph_replayed_kernels - launches0, info, prof

    def settle(self, max_steps=30, window=3, tol=1.03):
        """Untimed steps BEFORE the W warm-up steps until `window` consecutive step times agree
        within 3 % (at most `max_steps`).  A fresh process on a box whose previous tenant has just
        released tens of GB showed 2x slower, erratic steps for its first second or two; this
        keeps that transient out of the warm-up / timed region.  The count is reported."""
        times = []
        for i in range(max_steps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            self.step(False)
            b.record()
            torch.cuda.synchronize()
            times.append(a.elapsed_time(b))
            done = len(times) >= window and max(times[-window:]) <= tol * min(times[-window:])
            if self.world > 1:                  # every rank leaves the loop after the same step
                flag = torch.tensor([1 if done else 0], device=self.dev)
                self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
                done = bool(flag.item())
            if done:
                break
        return len(times)

    def host_enqueue_ms(self, reps=3):
        """Host time to ENQUEUE one step (everything up to, not including, the read-back of the
        loss) with an idle GPU in front: when this approaches ms_per_step the step is
        launch-bound on the host, not on the device."""
        import time as _t
        ops, hyps, ring = self.ops, self.hyps, self.ring
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = _t.perf_counter()
            ring.carry()
            ring.push_frames(self.d_new_frames, self.d_new_reset, first_frame_index=1)
            self.rew_w.copy_(self.d_in["rewards"]); self.done_w.copy_(self.d_in["dones"])
            self.val_w.copy_(self.d_in["values"]); self.act_w.copy_(self.d_in["actions"])
            deltas, advs, rets, stats = ops.delta_scan(self.rew_w, self.done_w, self.val_w, hyps["gamma"],
                                                       hyps["lambda_"])
            shared = {"states_view": ring.view(self.T), "states": None, "rewards": self.rew_w,
                      "dones": self.done_w, "deltas": deltas, "actions": self.act_w}
            if self.h_states is not None:
                shared["h_states"] = self.h_states
            self.upd.compute_gradients(shared)
            self.upd._optim_step()
            dt = (_t.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        torch.cuda.synchronize()
        return best

    def stage1_alone_ms(self, reps=20):
        """Stage 1 alone, launches back to back; R*T*100.8 KB of frames per launch (> L2 for
        every config but C1)."""
        self.barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(reps):
            self.ring.push_frames(self.d_new_frames, self.d_new_reset, first_frame_index=1)
        s1.record()
        self.barrier()
        return s0.elapsed_time(s1) / reps

    def allreduce_alone(self, reps=5):
        """The gradient all-reduce timed by itself (device events, max over ranks)."""
        if self.world == 1:
            return None
        g = self.upd.flat.grad
        self.dist.all_reduce(g)
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            self.dist.all_reduce(g)
        e1.record()
        self.barrier()
        g.zero_()
        t = torch.tensor([e0.elapsed_time(e1) / reps], device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        ms = float(t.item())
        nbytes = 4 * g.numel()
        w = self.world
        return {"bytes": nbytes, "ms": ms, "algbw_gbs": nbytes / ms / 1e6,
                "busbw_gbs": 2.0 * (w - 1) / w * nbytes / ms / 1e6, "nvlink_peak_gbs_per_dir": 900.0,
                "in_step": self.upd.last_allreduce,
                "note": "in the step the big bucket (resize_emb.0.weight) is reduced on NCCL's stream "
                        "under the conv backward; the rest follows the backward pass"}


def measure(wl, steps, warmup, pk, precision, profile_steps=3, with_stage1=True):
    """value / e2e / kernel table / rooflines of one workload (all ranks call this)."""
    settle = wl.settle()
    for _ in range(warmup):
        wl.step(False)
    ms, launches, info, _ = wl.timed(False, steps)
    step_ms, dev_allocs = wl.last_step_ms, wl.last_device_allocs
    for _ in range(max(1, min(2, warmup))):
        wl.ingest.prefetch(wl.h_frames, wl.h_reset)
        wl.step(True)
    ms_e2e, _, _, _ = wl.timed(True, steps)
    e2e_step_ms = wl.last_step_ms
    _, _, _, prof = wl.timed(False, min(steps, profile_steps), profile=True)
    stage1_ms = wl.stage1_alone_ms() if with_stage1 else None
    host_ms = wl.host_enqueue_ms()
    total_tr = wl.N * wl.world * steps
    rec = {"ms_per_step": ms / steps, "value": total_tr / (ms / 1e3),
           "e2e": {"value": total_tr / (ms_e2e / 1e3), "unit": "transitions/s",
                   "ms_per_step": ms_e2e / steps, "step_ms": e2e_step_ms, "h2d_bytes_per_step": wl.h2d_bytes(),
                   "d2h_bytes_per_step": 20,
                   "h2d_note": "per step ONE asynchronous DMA (staging layout %d: %d of 100,800 B per frame cross "
                               "PCIe) moves the new frames from pinned host memory to a device staging buffer, "
                               "pipelined under the previous step's kernels; stage 1 then reads the staging "
                               "buffer.  Rollout scalars are copied on the compute stream, before the DMA is "
                               "queued." % (wl.ingest.layout, wl.ingest.frame_bytes)},
           "gpu_launches": int(launches), "loss": info["Loss"], "host_enqueue_ms_per_step": host_ms,
           "cuda_graph": bool(wl.graph),
           "step_ms": step_ms, "device_allocs_in_timed_region": dev_allocs, "settle_steps": settle,
           "update_micro_batches": wl.upd.last_chunks}
    roof, table, stages = roofline_from_profile(prof, min(steps, profile_steps), wl.R, wl.T, pk, stage1_ms,
                                                precision)
    rec["roofline"], rec["kernel_table"], rec["stages"] = roof, table, stages
    return rec


def ncu_traffic():
    """DRAM bytes per launch from the committed `ncu --set full` captures (profiles/*.json)."""
    out = {}
    for name in ("r1_traffic.json", "r2_traffic.json"):
        path = os.path.join(ROOT, "profiles", name)
        if os.path.isfile(path):
            out.update(json.load(open(path)))
    return out


def roofline_from_profile(prof, steps, R, T, pk, stage1_ms, precision):
    """prof: {key: [(start_evt, end_evt, work), ...]} recorded around each C-ABI
    call in a profiled pass (events on the launching stream).  The entry with the
    largest total time is the dominant kernel."""
    if not prof:
        return None, None, None
    rows = []
    for key, recs in prof.items():
        tot = sum(s.elapsed_time(e) for s, e, _ in recs)
        work = sum(w for _, _, w in recs)
        rows.append((tot, key, len(recs), work))
    rows.sort(reverse=True)
    total_ms = sum(r[0] for r in rows)
    table = [{"call": k, "launches_per_step": n / steps, "ms_per_step": t / steps, "share": t / total_ms}
             for t, k, n, _ in rows[:16]]
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
    # the memory-bound / latency-bound stages of the north star, always reported
    stages = {}
    if stage1_ms is not None:
        a = R * T * ALG_BYTES_PREP / (stage1_ms / 1e3) / 1e9
        stages["stage1_obs_prep"] = {
            "kernel": "a2c_obs_prep_stack", "bound": "hbm", "achieved": a, "unit": "GB/s",
            "peak": pk["hbm_gbs"], "frac": a / pk["hbm_gbs"], "frac_of_nominal_8TBs": a / 8000.0,
            "launch_us": stage1_ms * 1e3, "bytes_per_transition": ALG_BYTES_PREP,
            "traffic": traffic.get("a2c_obs_prep_stack"), "timed": "alone, 20 launches back to back"}
        roof["stage1_hbm"] = stages["stage1_obs_prep"]
    for tag, name, bpt in (("a2c_delta_scan", "stage2_delta_scan", 28), ("a2c_discount2", "stage2_discount2", 20),
                           ("a2c_loss_fwd_bwd", "stage3_loss", 48)):
        for tt, k, nn, w in rows:
            if k.startswith(tag + "|"):
                us = tt / nn * 1e3
                gbs = w / nn / (tt / nn / 1e3) / 1e9
                stages[name] = {"kernel": tag, "bound": "launch latency", "launch_us": us,
                                "bytes_per_transition": bpt, "achieved": gbs, "unit": "GB/s",
                                "frac": gbs / pk["hbm_gbs"], "launches_per_step": nn / steps,
                                "timed": "in the step, CUDA events around the launch (includes launch overhead)"}
    return roof, table, stages


def dp_parity(world, rank, dev, dist, precision):
    """Before timing (N > 1): the NCCL-sharded CUDA update against the single-GPU full-batch
    update of the SAME global batch (ConvModel, 4 slots per rank x 6 steps, micro-batched in two
    chunks).  Returns max relative gradient / loss difference (rank 0)."""
    from a2c_b200 import models, ops, sharding, updater
    ops.set_precision(precision)
    Rg, T, fs = 4 * world, 6, 2
    hyps = dict(HYPS, n_tsteps=T, n_rollouts=Rg, n_envs=Rg, n_frame_stack=fs, model="ConvModel",
                update_chunk_slots=2)
    torch.manual_seed(7)
    net = models.ConvModel([fs, 80, 80], 3, h_size=HYPS["h_size"]).to(dev)
    upd = updater.Updater(net, hyps)
    dist.broadcast(upd.flat.flat, 0)
    b = synth_batch(Rg, T, dev, 99)                  # same seed on every rank: the global batch
    ring = ops.PlaneRing(Rg, T, fs, dev)
    ring.push_frames(b["frames"], b["reset"])
    states = ring.view(T).materialize(torch.uint8)
    rew, dn = b["rewards"].clone(), b["dones"].clone()
    deltas, _, _, _ = ops.delta_scan(rew, dn, b["values"], hyps["gamma"], hyps["lambda_"])
    full = {"states": states, "rewards": rew, "dones": dn, "deltas": deltas, "actions": b["actions"]}
    mine = sharding.shard_rollout(full, Rg, T, rank, world)
    res = upd.compute_gradients(mine)
    g_dp = upd.flat.grad.clone()
    loss_dp = res["loss_terms"].clone()
    out = None
    if rank == 0:
        net1 = copy.deepcopy(net)
        upd1 = updater.Updater(net1, dict(hyps, update_chunk_slots=Rg), data_parallel=False)
        res1 = upd1.compute_gradients(full)
        g1 = upd1.flat.grad
        den = float(g1.abs().max())
        out = {"max_rel_grad_diff": float((g_dp - g1).abs().max()) / den,
               "rel_l2_grad_diff": float((g_dp - g1).double().norm() / g1.double().norm()),
               # loss terms (Loss, Pi_Loss, ValLoss, Entropy) on the scale the parity tests use:
               # |diff| / max(1, |value|)  (Pi_Loss is a near-zero sum of signed terms)
               "max_loss_diff": float(((loss_dp - res1["loss_terms"]).abs()
                                       / res1["loss_terms"].abs().clamp_min(1.0)).max()),
               "loss_terms_dp": [float(x) for x in loss_dp.tolist()],
               "loss_terms_full_batch": [float(x) for x in res1["loss_terms"].tolist()],
               "config": "ConvModel, %d slots x %d steps global, %d slots per rank, micro-batches of 2 slots, "
                         "vs one full-batch pass on rank 0" % (Rg, T, Rg // world)}
    dist.barrier()
    return out


def acting_path(dev, precision, model="ConvModel", R=256, T=20, fs=2, reps=4):
    """SURVEY section 8f-1, timed: Runner.rollout of the batched on-device acting step (per
    timestep: one stage-1 launch for the R new frames, one batched inference forward, one fused
    softmax / inverse-CDF sampler; then stage 2) against a recorded environment whose frames are
    resident on the device.  The reference does this with R processes of batch-1 forwards."""
    import queue

    from a2c_b200 import models, ops, runner as runner_mod
    ops.set_precision(precision)
    hyps = dict(HYPS, n_tsteps=T, n_rollouts=R, n_envs=R, n_frame_stack=fs, model=model)
    b = synth_batch(R, T, dev, 4321)
    g = torch.Generator(device=dev).manual_seed(5)
    traj_dev = {"frames": b["frames"], "reset": b["reset"], "rewards": b["rewards"].view(R, T),
                "u": torch.rand((R, T), device=dev, generator=g)}

    class _Env(runner_mod.RecordedVecEnv):
        def __init__(self):                      # device tensors directly (no numpy round trip)
            self.device, self.R, self.T = dev, R, T
            self.frames, self.reset_flags = traj_dev["frames"], traj_dev["reset"]
            self.rewards, self.u, self.t = traj_dev["rewards"], traj_dev["u"], 0

        def rewind(self):
            self.t = 0

    torch.manual_seed(0)
    net = models.get_model_class(model)([fs, 80, 80], 3, h_size=HYPS["h_size"]).to(dev)
    datas = runner_mod.make_shared_data(R, T, fs, dev, h_size=HYPS["h_size"] if net.is_recurrent else None,
                                        with_states=False)
    q = queue.Queue(1)
    q.put(-1.0)
    rn = runner_mod.Runner(datas, hyps, None, None, q)
    env = _Env()
    rn.attach(net, env)
    times = []
    for i in range(reps + 1):
        env.rewind()
        torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        rn.rollout(net, 0, hyps)
        e.record()
        torch.cuda.synchronize()
        if i > 0:
            times.append(a.elapsed_time(e))
    ms = min(times)
    return {"model": model, "n_envs": R, "n_tsteps": T, "ms_per_rollout": ms, "ms_per_timestep": ms / T,
            "transitions_per_s": R * T / (ms / 1e3),
            "what": "Runner.rollout: batched acting (stage 1 + inference forward + sampler per timestep, "
                    "stage 2 at the end), frames resident on the device, best of %d rollouts" % reps}


def gpu_eager_baseline(dev, N, steps=5, warmup=2):
    """Stage 3 + 4 + clip + RMSprop of C2 with stock torch.nn / autograd / torch.optim on the same
    B200 (fp32, allow_tf32 off: the library bar SURVEY.md section 2.1 names, and with TF32 on for
    reference).  Library code, not ours: a reported baseline only."""
    import torch.nn as nn
    out = {}
    for tf32 in (False, True):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        torch.backends.cudnn.allow_tf32 = tf32
        torch.manual_seed(0)
        m = nn.ModuleDict(dict(
            convs=nn.Sequential(nn.Conv2d(2, 16, 3, 1, 1), nn.ReLU(), nn.Conv2d(16, 24, 3, 1, 1), nn.ReLU(),
                                nn.Conv2d(24, 32, 3, 2, 1), nn.ReLU(), nn.Conv2d(32, 64, 3, 2, 1), nn.ReLU()),
            emb=nn.Sequential(nn.Linear(25600, 2000), nn.ReLU()),
            pi=nn.Sequential(nn.Linear(2000, 256), nn.ReLU(), nn.Linear(256, 3)),
            val=nn.Sequential(nn.Linear(2000, 256), nn.ReLU(), nn.Linear(256, 1)))).to(dev)
        opt = torch.optim.RMSprop(m.parameters(), lr=1e-4)
        x = (torch.rand(N, 2, 80, 80, device=dev) < 0.1).float()
        adv = torch.randn(N, device=dev)
        ret = torch.randn(N, device=dev)
        act = torch.randint(0, 3, (N,), device=dev)

        def step():
            e = m["emb"](m["convs"](x).reshape(N, -1))
            logits, vals = m["pi"](e), m["val"](e)
            lsm = torch.log_softmax(logits, -1)
            loss = (-(lsm[torch.arange(N, device=dev), act] * adv).mean()
                    + 0.5 * torch.nn.functional.mse_loss(vals.squeeze(), ret)
                    + 0.005 * (lsm * lsm.exp()).sum(-1).mean())
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 0.5)
            opt.step()
            opt.zero_grad()
        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        out["tf32_on" if tf32 else "fp32"] = {"ms_per_step": ms, "transitions_per_s": N / (ms / 1e3)}
        del m, opt, x
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    out["what"] = ("torch.nn ConvModel forward + A2C loss + autograd backward + clip_grad_norm_ + "
                   "torch.optim.RMSprop on %d fp32 states already on the device (no stage 1 / 2): cuDNN / cuBLAS" % N)
    return out


def run_ours(args):
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    workload = args.workload or ("c2" if world == 1 else "c5")
    pk = peaks()
    parity = dp_parity(world, rank, dev, dist, args.precision) if world > 1 else None

    wl = Workload(workload, world, rank, dev, args.precision, dist, graph=args.graph)
    sampler = ClockSampler(local, args.clock_interval_ms)
    if rank == 0:
        sampler.start()
    rec = measure(wl, args.steps, args.warmup, pk, args.precision)
    clocks = sampler.stop() if rank == 0 else None
    allred = wl.allreduce_alone()
    chunks = None
    out = None
    if rank == 0:
        model, _, T, fs, scaling, desc = WORKLOADS[workload]
        out = {
            "metric": METRIC, "value": rec["value"], "unit": "transitions/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": rec["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "tf32" if args.precision == 2 else "f32", "data": "synthetic",
            "config": config_of(workload, world, args.precision),
            "e2e": rec["e2e"], "gpu_launches": rec["gpu_launches"],
            "roofline": rec["roofline"], "kernel_table": rec["kernel_table"], "stages": rec["stages"],
            "clocks": clocks, "loss": rec["loss"], "peaks_source": pk["src"],
            "step_ms": rec["step_ms"], "host_enqueue_ms_per_step": rec["host_enqueue_ms_per_step"],
            "device_allocs_in_timed_region": rec["device_allocs_in_timed_region"],
            "settle_steps": rec["settle_steps"],
        }
        if allred is not None:
            out["allreduce"] = allred
        if parity is not None:
            out["dp_parity"] = parity
    del wl
    torch.cuda.empty_cache()

    if world > 1 and workload == "c5" and not args.no_extras:
        # the denominator of the strong-scaling curve, measured in the same run: all 4096 envs on
        # rank 0 alone (no collectives), the other ranks wait
        n1 = None
        if rank == 0:
            w1 = Workload("c5", 1, 0, dev, args.precision, dist, data_parallel=False)
            for _ in range(3):
                w1.step(False)
            ms1, _, _, _ = w1.timed(False, 3)
            n1 = {"value": w1.N * 3 / (ms1 / 1e3), "ms_per_step": ms1 / 3, "steps": 3, "warmup": 3,
                  "update_micro_batches": w1.upd.last_chunks,
                  "what": "C5 (4096 envs x 20 steps) on ONE GPU of this box, same build, measured by rank 0 "
                          "while the other ranks idle: value / (n_gpus x this) is the strong-scaling efficiency"}
            del w1
            torch.cuda.empty_cache()
        dist.barrier()
        if rank == 0:
            out["c5_single_gpu"] = n1

    if world == 1 and not args.no_extras and args.workload is None:
        # one record per remaining BASELINE config (fewer steps: they are not the headline)
        extras = {}
        for name in ("c1", "c3", "c4", "c5"):
            w2 = Workload(name, 1, 0, dev, args.precision, dist, graph=args.graph)
            r2 = measure(w2, max(3, min(args.steps, 5)), 3, pk, args.precision, profile_steps=2)
            r2["config"] = config_of(name, 1, args.precision)
            r2["steps"] = max(3, min(args.steps, 5))
            r2["kernel_table"] = r2["kernel_table"][:8]
            if not args.no_cpu_baseline:
                r2["cpu_baseline"] = cpu_baseline(name, budget_s=8.0)
            extras[name] = r2
            del w2
            torch.cuda.empty_cache()
        out["workloads"] = extras
        out["gpu_eager_baseline"] = gpu_eager_baseline(dev, 5120)
        out["acting_path"] = acting_path(dev, args.precision)
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(workload, budget_s=25.0)
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------ reference arm
# Rollout slots the CPU arm processes per step: the FULL configuration where one pass fits the
# budget (C1, C2, C3), a bounded sample otherwise (stated in cpu_baseline.sample).
CPU_SLOTS = {"c1": 4, "c2": 256, "c3": 1024, "c4": 32, "c5": 64}


def cpu_state(workload, Rs):
    from a2c_b200 import synth
    from oracle import a2c_oracle as O
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    traj = synth.make_trajectory(Rs, T, seed=1234)
    hyps = dict(HYPS, n_tsteps=T, n_rollouts=Rs, n_frame_stack=fs, use_bptt=(model == "GRUModel"))
    shapes = O.param_shapes(model, fs=fs, h=hyps["h_size"])
    params = {k: torch.from_numpy(v) for k, v in synth.hashed_state_dict(shapes).items()}
    rng = np.random.default_rng(7)
    st = {"inputs": (model, T, fs, traj), "upd": O.UpdaterOracle(model, params, hyps),
          "actions": torch.from_numpy(rng.integers(0, 3, size=Rs * T)).long()}
    if model == "GRUModel":
        st["h_states"] = torch.randn(Rs * T, hyps["h_size"]) * 0.5
    return st


def cpu_step(Rs, state):
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


def cpu_sample_text(workload, Rs, n_passes=None):
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    what = ("the FULL configuration (%d rollout slots, %d transitions) per pass" % (R, R * T) if Rs == R
            else "%d of %d rollout slots (%d transitions) per pass" % (Rs, R, Rs * T))
    s = ("%s: per-frame pong_prep + frame stack + fp32 state store (single process, as one reference "
         "actor), bookkeeping, full Updater step incl. both Python discount loops (torch-CPU, %d threads)"
         % (what, torch.get_num_threads()))
    if n_passes:
        s += "; best of %d passes" % n_passes
    return s


def cpu_baseline(workload, budget_s=25.0):
    torch.set_num_threads(os.cpu_count())
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    Rs = min(R, CPU_SLOTS[workload])
    st = cpu_state(workload, Rs)
    times, t_start = [], time.perf_counter()
    while True:
        times.append(cpu_step(Rs, st))
        if time.perf_counter() - t_start > budget_s or len(times) >= 20:
            break
    if len(times) > 1:
        times = times[1:]                       # the first pass warms the allocator / thread pool
    best = min(a + b for a, b in times)
    prep, upd = min(a for a, _ in times), min(b for _, b in times)
    cores = os.cpu_count()
    return {"value": Rs * T / best, "unit": "transitions/s", "cores": cores, "kind": "port",
            "sample": cpu_sample_text(workload, Rs, len(times)), "same_config": Rs == R,
            "prep_s": prep, "update_s": upd,
            "prep_transitions_per_s_single_process": Rs * T / prep,
            "prep_transitions_per_s_ideal_actor_parallelism": Rs * T / prep * cores,
            "value_with_prep_divided_by_cores": Rs * T / (prep / cores + upd)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    workload = args.workload or ("c2" if world == 1 else "c5")
    model, R, T, fs, scaling, desc = WORKLOADS[workload]
    torch.set_num_threads(os.cpu_count())
    Rs = min(R, CPU_SLOTS[workload])
    st = cpu_state(workload, Rs)
    for _ in range(args.warmup):
        cpu_step(Rs, st)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_step(Rs, st)
    dt = time.perf_counter() - t0
    value = Rs * T * args.steps / dt
    out = {"impl": "reference", "metric": METRIC, "value": value, "unit": "transitions/s",
           "n_gpus": world, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
           "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": config_of(workload, world, args.precision),
           "cpu_baseline": {"value": value, "unit": "transitions/s", "cores": os.cpu_count(),
                            "kind": "port", "sample": cpu_sample_text(workload, Rs),
                            "same_config": Rs == R and world == 1},
           "e2e": {"value": value, "unit": "transitions/s", "h2d_bytes_per_step": 0,
                   "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS),
                    help="default: c2 on one GPU, c5 (4096 envs sharded) on several")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", type=int, default=int(os.environ.get("A2C_PRECISION", "1")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--graph", type=int, default=-1,
                    help="replay stage 2 + the update as a CUDA graph (single GPU, one micro-batch, RMSprop): "
                         "1 on, 0 off, -1 auto (batches of <= 2048 transitions)")
    ap.add_argument("--clock-interval-ms", type=int, default=200,
                    help="nvidia-smi sampling period during the timed region (the profiling recipe's 200 ms); 0 = off")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the c1/c3/c4/c5 records and the torch-eager GPU baseline of the N=1 line")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3                          # timing rule: at least 3 warm-up steps
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

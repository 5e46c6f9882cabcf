#!/usr/bin/env python
"""bench.py -- the hot path's headline measurement (see DESIGN.md section d).

Workload (BASELINE.json configs[2], the configuration the metric is quoted
on): Rainbow's replay path -- PrioritizedReplayBuffer(10**6, alpha=.5,
beta0=.4, num_steps=3, normalize_by_max="memory") holding synthetic Atari
transitions (84x84 uint8 frames, stack 4, frame-shared), minibatch 512.
One "step" = one pass of the replay hot path over one minibatch:

    prioritized sample(512) -> IS weights -> gather state/next_state as f32
    (/255) + reward/discount/terminal/action -> TD-error -> priority write-back

metric   replay_samples_per_sec (whole job, all ranks)
value    device-resident: u / TD errors already in HBM resp. pinned, C-ABI calls
e2e      public API with HOST buffers: buf.sample() -> batch_experiences() ->
         D2H of weights/reward/indices -> buf.update_errors(host float list)
roofline the gather kernel (dominant in bytes), CUDA events inside the timed
         region, vs MEASURED_PEAKS.json hbm_gbs
cpu_baseline / --impl reference: oracle/pyport.py (pure-Python port with the
         reference's cost profile; the reference itself cannot travel to the
         GPU box), single thread, bounded sample of the same workload.

Launch: python bench.py [--gpus N --steps K --warmup W]; for N > 1 under
torch.distributed.run (one rank per GPU, NCCL only for the timing barrier and
max-reduction: the replay shards never exchange data -> "scaling": "weak").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAME = (84, 84)
STACK = 4
N_STEP = 3
GAMMA = 0.99
ALPHA = 0.5
BETA0 = 0.4
FRAME_BYTES = 84 * 84
# algorithmic bytes per sampled experience, f32 outputs (DESIGN.md section d):
# 7 distinct input frames (3-step, stack 4) + state and next_state as f32
# + scalars (action 8, reward 4, terminal 4, discount 4, weight 4, index 8)
ALGO_BYTES_GATHER = 7 * FRAME_BYTES + 2 * STACK * FRAME_BYTES * 4 + 32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--capacity", type=int, default=10 ** 6)
    ap.add_argument("--batch", type=int, default=512)
    ap.add_argument("--mode", default="exact", choices=["exact", "parallel"])
    ap.add_argument("--cpu-seconds", type=float, default=12.0,
                    help="CPU baseline budget (timed part)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rainbow", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------
# clocks
# ---------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(np.max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------
# CPU baseline: pure-Python port of the reference path (oracle/pyport.py)
# ---------------------------------------------------------------------------
def cpu_reference_run(capacity, batch, steps, warmup, seconds=None, pool=65536, seed=0):
    """Time `steps` passes (or as many as fit in `seconds`) of
    sample(batch) + batch_experiences + update_errors on the host."""
    import torch
    from oracle.pyport import PyPrioritizedReplayBuffer, py_batch_experiences
    from pfrl_b200.utils.lazy_frames import LazyFrames

    torch.set_num_threads(os.cpu_count() or 1)
    rng = np.random.RandomState(seed)
    pool = min(pool, capacity + STACK + N_STEP)
    frames = rng.randint(0, 256, size=(pool, 1) + FRAME, dtype=np.uint8)
    flist = [frames[i] for i in range(pool)]
    t0 = time.perf_counter()
    buf = PyPrioritizedReplayBuffer(capacity, alpha=ALPHA, beta0=BETA0, betasteps=None,
                                    num_steps=N_STEP, normalize_by_max="memory")
    # setup (untimed): transitions share dicts between overlapping windows,
    # observations share frames, exactly like the reference's storage
    T = capacity + N_STEP - 1
    obs = [LazyFrames([flist[(t + j) % pool] for j in range(STACK)], stack_axis=0)
           for t in range(T + 1)]
    acts = rng.randint(0, 18, size=T)
    rews = rng.randint(-1, 2, size=T).astype(np.float64)
    trans = [dict(state=obs[t], action=int(acts[t]), reward=float(rews[t]),
                  next_state=obs[t + 1], next_action=None, is_state_terminal=False)
             for t in range(T)]
    values = [trans[s:s + N_STEP] for s in range(capacity)]
    buf.memory.bulk_load(values, rng.rand(capacity) + 0.05)
    setup_s = time.perf_counter() - t0
    phi = lambda x: np.asarray(x, dtype=np.float32) / 255  # noqa: E731
    dev = torch.device("cpu")
    np.random.seed(seed)

    def one():
        exps = buf.sample(batch)
        b = py_batch_experiences(exps, dev, phi, GAMMA)
        err = [float(x) for x in np.abs(rng.randn(batch))]
        buf.update_errors(err)
        return b

    for _ in range(warmup):
        one()
    done = 0
    t0 = time.perf_counter()
    while done < steps:
        one()
        done += 1
        if seconds is not None and time.perf_counter() - t0 > seconds:
            break
    dt = time.perf_counter() - t0
    return {"samples_per_sec": done * batch / dt, "steps": done, "seconds": dt,
            "ms_per_step": 1e3 * dt / done, "setup_s": setup_s, "pool": pool}


# ---------------------------------------------------------------------------
def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        r = cpu_reference_run(args.capacity, args.batch, args.steps, args.warmup)
        line = {
            "impl": "reference", "metric": "replay_samples_per_sec", "value": r["samples_per_sec"],
            "unit": "samples/s", "n_gpus": args.gpus, "steps": r["steps"], "warmup": args.warmup,
            "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 priorities / u8->f32 frames", "data": "synthetic",
            "config": workload_config(args, 1),
            "cpu_baseline": {"value": r["samples_per_sec"], "unit": "samples/s", "cores": 1,
                             "kind": "port",
                             "sample": "%d steps of sample(%d)+batch_experiences+update_errors, "
                                       "1M-leaf tree, frames from a pool of %d"
                                       % (r["steps"], args.batch, r["pool"])},
            "e2e": {"value": r["samples_per_sec"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from pfrl_b200 import _lib
    from pfrl_b200.replay_buffer import batch_experiences
    from pfrl_b200.replay_buffers import PrioritizedReplayBuffer
    from pfrl_b200.utils.phi import ScaleU8

    B, cap = args.batch, args.capacity
    # ---- build and prefill the shard (untimed) ---------------------------------
    buf = PrioritizedReplayBuffer(cap, alpha=ALPHA, beta0=BETA0, betasteps=None,
                                  normalize_by_max="memory", num_steps=N_STEP, device=local_rank,
                                  max_batch=max(B, 512), part_capacity=cap + 4096,
                                  sample_mode=args.mode)
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    rng = np.random.RandomState(100 + rank)
    T = cap + N_STEP - 1
    chunk = 1 << 17
    first = True
    t_fill = time.perf_counter()
    done = 0
    while done < T:
        m = min(chunk, T - done)
        # each chunk is its own trajectory segment (stack extra frames up front)
        frames = torch.randint(0, 256, (m + STACK,) + FRAME, dtype=torch.uint8, device=dev,
                               generator=g)
        acts = rng.randint(0, 18, size=m).astype(np.int64)
        rews = rng.randint(-1, 2, size=m).astype(np.float64)
        term = rng.rand(m) < 1e-3
        term[-1] = True  # close the segment so that its tail is emitted
        buf.append_trajectory(frames, acts, rews, term)
        done += m
        first = False
    # non-uniform priorities so that the tree descent is not degenerate
    n_pri = 64
    for _ in range(n_pri):
        np.random.seed(rng.randint(1 << 30))
        buf.sample(B)
        buf.update_errors(torch.rand(B, device=dev, dtype=torch.float32, generator=g) * 2)
    torch.cuda.synchronize()
    fill_s = time.perf_counter() - t_fill
    assert len(buf) == cap, (len(buf), cap)

    store = buf.store
    phi = ScaleU8()
    gp = buf._gamma_pow(GAMMA)
    mode = _lib.SAMPLE_EXACT if args.mode == "exact" else _lib.SAMPLE_PARALLEL
    K, W = args.steps, args.warmup
    u_all = rng.random_sample((K + W, B))
    err_dev = torch.rand((16, B), device=dev, dtype=torch.float32, generator=g).abs() * 1.5
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- value: device-resident loop through the store (C ABI) -----------------
    ev_pairs = {"sample": [], "gather": [], "update": []}

    def step_value(i, timed):
        e = [ev() for _ in range(6)] if timed else None
        if timed:
            e[0].record()
        store.sample(u_all[i], mode=mode, want_index=True, want_priority=False)
        if timed:
            e[1].record()
        w = store.weights(B, buf.beta, _lib.NORM_MEMORY)
        if timed:
            e[2].record()
        out = store.gather(B, gp, obs_mode=_lib.OBS_U8_TO_F32, obs_scale=phi.b2rl_obs_scale,
                           obs_shape=(STACK,) + FRAME)
        if timed:
            e[3].record()
            e[4].record()
        store.update_errors(err_dev[i % 16], ALPHA, 0.01, 0, 1)
        if timed:
            e[5].record()
            ev_pairs["sample"].append((e[0], e[1]))
            ev_pairs["gather"].append((e[2], e[3]))
            ev_pairs["update"].append((e[4], e[5]))
        return out, w

    for i in range(W):
        step_value(i, False)
    barrier()
    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    t0, t1 = ev(), ev()
    t0.record()
    for i in range(W, W + K):
        step_value(i, True)
    t1.record()
    barrier()
    ms_value = t0.elapsed_time(t1)
    kern_ms = {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in ev_pairs.items()}

    # ---- e2e: public API with host buffers --------------------------------------
    np.random.seed(7 + rank)
    err_host = [[float(x) for x in np.abs(rng.randn(B))] for _ in range(8)]
    pinned = torch.empty((3, B), dtype=torch.float64).pin_memory()

    def step_e2e(i):
        exps = buf.sample(B)
        b = batch_experiences(exps, dev, phi, GAMMA)
        pinned[0].copy_(b["weights"], non_blocking=True)
        pinned[1].copy_(b["reward"], non_blocking=True)
        pinned[2].copy_(exps.index, non_blocking=True)
        torch.cuda.current_stream().synchronize()  # the D2H read of the step's result
        buf.update_errors(err_host[i % 8])
        return b

    for i in range(W):
        step_e2e(i)
    barrier()
    t2, t3 = ev(), ev()
    t2.record()
    for i in range(K):
        step_e2e(i)
    t3.record()
    barrier()
    ms_e2e = t2.elapsed_time(t3)
    clk = clocks.stop() if rank == 0 else None

    # ---- max over ranks ---------------------------------------------------------
    tm = torch.tensor([ms_value, ms_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    ms_value, ms_e2e = [float(x) for x in tm.tolist()]
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total = world * B * K
    value = total / (ms_value / 1e3)
    e2e = total / (ms_e2e / 1e3)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = json.load(open(peaks_path))["hbm_gbs"]
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (of fallback)"
    gather_bytes = ALGO_BYTES_GATHER * B
    achieved = gather_bytes / (kern_ms["gather"] * 1e-3) / 1e9
    line = {
        "metric": "replay_samples_per_sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": K, "warmup": W, "ms_per_step": ms_value / K, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 priorities / u8->f32 frames",
        "data": "synthetic", "config": workload_config(args, world),
        "e2e": {"value": e2e, "unit": "samples/s", "ms_per_step": ms_e2e / K,
                "h2d_bytes_per_step": 8 * B + 8 * B + 8 * (N_STEP + 1),
                "d2h_bytes_per_step": 4 * B + 4 * B + 8 * B},
        "gpu_launches": 4 * K,
        "kernels_ms": kern_ms,
        "roofline": {"bound": "hbm", "kernel": "k_gather", "achieved": achieved, "peak": peak,
                     "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                     "algorithmic_bytes_per_launch": gather_bytes, "peak_source": peak_src,
                     "kernel_ms": kern_ms["gather"],
                     "share_of_step": kern_ms["gather"] / (ms_value / K)},
        "sampler": {"mode": args.mode, "kernel": "k_sample_" + args.mode,
                    "ms_per_batch": kern_ms["sample"],
                    "ns_per_draw": 1e6 * kern_ms["sample"] / B,
                    "share_of_step": kern_ms["sample"] / (ms_value / K)},
        "clocks": clk, "prefill_s": fill_s,
        "hbm_bytes_per_rank": store.device_bytes,
    }
    if not args.no_cpu_baseline:
        r = cpu_reference_run(cap, B, 10 ** 9, 2, seconds=args.cpu_seconds)
        line["cpu_baseline"] = {
            "value": r["samples_per_sec"], "unit": "samples/s", "cores": 1, "kind": "port",
            "host_cores_available": os.cpu_count(),
            "sample": "%d steps (%.1f s) of sample(%d)+batch_experiences+update_errors, "
                      "1M-leaf tree, frames from a pool of %d; pure-Python port of the "
                      "reference (oracle/pyport.py)" % (r["steps"], r["seconds"], B, r["pool"])}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def workload_config(args, world):
    return {"workload": "Rainbow replay path (BASELINE configs[2]): PER 1M cap, 3-step, "
                        "84x84x4 u8 frames, minibatch %d per rank" % args.batch,
            "capacity_per_rank": args.capacity, "batch_per_rank": args.batch,
            "global_batch": args.batch * world, "n_step": N_STEP, "alpha": ALPHA,
            "sampler": args.mode, "obs_out": "f32 (x/255)", "l2": "inputs_larger_than_l2",
            "parallelism": "replay shard per rank, no data-path collective"}


if __name__ == "__main__":
    main()

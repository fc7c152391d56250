#!/usr/bin/env python
"""bench.py -- quadrotor env-steps/sec of the B200 step kernel (BASELINE.json's metric).

  python bench.py [--gpus N --steps K --warmup W]            our arm (torchrun launches one rank per GPU for N>1)
  python bench.py --impl reference [...]                     the CPU arm: the oracle port on all host threads

A "step" is one pass of the hot path (step_wait: residual MLPs + EoM + Euler + reward/flags + fused reset +
observation) over every env.  Workload: the end-to-end Bebop env, zigzag track, gates_ahead=1, training
disturbance ranges, N = 2**20 envs PER GPU (weak scaling; 2**20 envs touch ~320 MB per step, > the 126 MB L2, so
no L2 flush is needed between steps).  Rank 0 prints ONE JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "quadrotor env-steps/sec"
UNIT = "env-steps/s"
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=200)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", default="e2e", choices=["e2e", "indi"])
    ap.add_argument("--num-envs", type=int, default=1 << 20, help="envs per GPU")
    ap.add_argument("--gates-ahead", type=int, default=1)
    ap.add_argument("--e2e-steps", type=int, default=20)
    ap.add_argument("--cpu-envs", type=int, default=1 << 16, help="bounded CPU sample: envs stepped by the CPU arm")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gather-obs", nargs="?", const="nccl", default=None, choices=["nccl", "p2p"],
                    help="add the all-gather of observations per step: nccl = step kernel writes the send slot, NCCL "
                         "gathers in place; p2p = the step kernel stores its tiles into every peer's buffer itself")
    ap.add_argument("--workload", default="step", choices=["step", "policy", "rollout", "rollout_fused", "rollout_unfused"],
                    help="step: the env step alone on resident actions (the headline, default); policy: the on-device "
                         "controller forward alone (tcgen05); rollout: policy forward + env step per step, no host")
    ap.add_argument("--rollout-steps", type=int, default=32,
                    help="rollout_fused / rollout_unfused: env steps per qs_rollout[_fused] call (SB3 n_steps)")
    ap.add_argument("--no-stats", action="store_true", help="do not accumulate the device-side reward/flag totals")
    ap.add_argument("--graph", type=int, default=20,
                    help="replay the timed steps as a CUDA graph of this many steps (0 = one launch call per step)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.variant}_zigzag_ga{a.gates_ahead}_N{a.num_envs}_per_gpu" if a.variant == "e2e" else \
        f"{a.variant}_rectangle_ga{a.gates_ahead}_N{a.num_envs}_per_gpu"


def track_for(variant):
    import optimal_quad_control_rl_b200 as Q
    return Q.zigzag_track() if variant == "e2e" else Q.rectangle_track()


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_step_rate(variant, n_envs, gates_ahead, seconds, min_steps=3):
    """env-steps/s of the oracle port (C restatement, OpenMP over envs) on this host; same track, same ranges,
    uniform random actions, resets drawn like the reference does."""
    import numpy as np

    import optimal_quad_control_rl_b200 as Q
    from oracle import c_oracle as O

    O.lib().qo_set_num_threads(os.cpu_count() or 1)  # all host threads, whatever OMP_NUM_THREADS the launcher set
    gp, gy, sp = track_for(variant)
    env = O.OracleEnv(variant, n_envs, gp, gy, sp, gates_ahead=gates_ahead)
    if variant == "e2e":
        env.disturbance_ranges = Q.training_disturbance_ranges()
    np.random.seed(0)
    env.reset()
    acts = np.random.default_rng(1).uniform(-1, 1, (8, n_envs, 4)).astype(np.float32)
    for i in range(2):
        env.step(acts[i])
    steps, t0 = 0, time.perf_counter()
    while True:
        env.step(acts[steps % 8])
        steps += 1
        dt = time.perf_counter() - t0
        if steps >= min_steps and dt >= seconds:
            break
    return n_envs * steps / dt, steps, dt, O.lib().qo_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm does not shard: rank 0 alone runs and prints
    # `--steps K --warmup W` bound the sample: K+W steps of cpu_envs envs, capped in wall time
    t_budget = min(a.cpu_seconds * 4, 120.0)
    import numpy as np

    import optimal_quad_control_rl_b200 as Q
    from oracle import c_oracle as O

    O.lib().qo_set_num_threads(os.cpu_count() or 1)  # all host threads (torchrun exports OMP_NUM_THREADS=1)
    gp, gy, sp = track_for(a.variant)
    env = O.OracleEnv(a.variant, a.cpu_envs, gp, gy, sp, gates_ahead=a.gates_ahead)
    if a.variant == "e2e":
        env.disturbance_ranges = Q.training_disturbance_ranges()
    np.random.seed(0)
    env.reset()
    acts = np.random.default_rng(1).uniform(-1, 1, (8, a.cpu_envs, 4)).astype(np.float32)
    for i in range(max(1, min(a.warmup, 5))):
        env.step(acts[i % 8])
    steps, t0 = 0, time.perf_counter()
    while steps < a.steps:
        env.step(acts[steps % 8])
        steps += 1
        if time.perf_counter() - t0 > t_budget:
            break
    dt = time.perf_counter() - t0
    value = a.cpu_envs * steps / dt
    cores = O.lib().qo_num_threads()
    sample = f"{steps} steps of the first {a.cpu_envs} envs of the workload ({dt:.1f} s)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * dt / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "cpu_sample_envs": a.cpu_envs},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is NumPy notebook cells (not installable); this arm is oracle/quadsim_oracle.c, the C "
                "restatement pinned to the reference's golden vectors, OpenMP over envs on all host threads",
    }))


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """Polls SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    BAD = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
    NOTE = {"sw_power_cap": 0x4, "hw_power_brake": 0x80, "sync_boost": 0x10}

    def __init__(self, index, period=0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz, self.ok = [], 0, None, False
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.reasons |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join()
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"], "samples": 0}
        s = sorted(self.samples)
        names = [k for k, b in {**self.BAD, **self.NOTE}.items() if self.reasons & b]
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": names, "samples": len(s)}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    import optimal_quad_control_rl_b200 as Q
    from optimal_quad_control_rl_b200 import _lib as L

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU (the product has no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, ga = a.num_envs, a.gates_ahead
    gp, gy, sp = track_for(a.variant)
    cls = Q.Quadcopter3DGates if a.variant == "e2e" else Q.Quadcopter3DGatesINDI
    env = cls(n, gp, gy, sp, gates_ahead=ga, device=dev, reset_rng="device", seed=0, env_offset=rank * n,
              obs_buffers=2)
    if a.variant == "e2e":
        env.disturbance_ranges = Q.training_disturbance_ranges()
    env.enable_stats(not a.no_stats)
    env.reset_tensor()
    gen = torch.Generator(device=dev).manual_seed(1 + rank)
    acts = [torch.rand((n, 4), generator=gen, device=dev) * 2 - 1 for _ in range(4)]  # resident, > L2 with obs
    gather = None
    if a.gather_obs == "nccl" or (a.gather_obs and world == 1):
        gather = Q.ObsAllGather(n * world, env.state_len, dev)
    elif a.gather_obs == "p2p":
        gather = Q.ObsPeerGather(n * world, env.state_len, dev)
        gather.attach(env)

    pol = None
    if a.workload != "step":
        if a.variant != "e2e" or ga != 1:
            raise SystemExit("--workload policy/rollout uses the shipped 24-input controller: e2e, --gates-ahead 1")
        pol = Q.MlpPolicy.from_npz(device=dev, seed=3, env_offset=rank * n)
        obs_in = [torch.randn((n, env.state_len), generator=gen, device=dev) for _ in range(4)]

    def one_step(i):
        if a.workload == "policy":
            pol.forward(obs_in[i & 3], out=acts[i & 3])
            return
        if a.workload == "rollout":  # actions come from the controller evaluated on the previous observations
            pol.forward(env._obs_ring[env._ring], out=acts[i & 3])
        env.step_tensor(acts[i & 3], obs_out=None if gather is None else gather.local_slot())
        if gather is not None:
            gather.gather()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    W, K = max(a.warmup, 3), a.steps
    roll = a.workload in ("rollout_fused", "rollout_unfused")
    if roll:  # collect_rollouts into (T, N, .) buffers: one fused launch per T steps, or 2T launches
        T = max(1, min(a.rollout_steps, K))
        K -= K % T
        W = max(T, W - W % T)
        a.graph = 0
        bufs = {"obs": torch.empty((T + 1, n, env.state_len), dtype=torch.float32, device=dev),
                "actions": torch.empty((T, n, 4), dtype=torch.float32, device=dev),
                "raw_actions": torch.empty((T, n, 4), dtype=torch.float32, device=dev),
                "rewards": torch.empty((T, n), dtype=torch.float32, device=dev),
                "dones": torch.empty((T, n), dtype=torch.uint8, device=dev)}
        bufs["obs"][0].copy_(env.current_obs_tensor())
        fused = a.workload == "rollout_fused"

        def one_step(i):  # noqa: F811 - called once per T env steps below
            if i % T == 0:
                env.rollout(pol, T, buffers=bufs, fused=fused)
                bufs["obs"][0].copy_(bufs["obs"][T])
    for i in range(W):
        one_step(i)
    graph = None
    if gather is not None:
        a.graph = 0  # the collective is issued by torch.distributed per step
    if a.graph > 0:  # launch-bound regime: capture `graph` consecutive steps once, replay K/graph times
        g = a.graph - a.graph % 4  # whole action-buffer cycles, and a divisor of K
        while g >= 4 and K % g:
            g -= 4
        a.graph = g if g >= 4 else 0
    if a.graph > 0:
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            one_step(0)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            for i in range(a.graph):
                one_step(i)
        for _ in range(2):
            graph.replay()
    env.stats(reset=True)
    l0 = env.launch_count
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.start()
    e0.record()
    if graph is None:
        for i in range(K):
            one_step(i)
    else:
        for _ in range(K // a.graph):
            graph.replay()
    e1.record()
    barrier()
    clocks = sampler.stop()
    ms = e0.elapsed_time(e1)
    launches = env.launch_count - l0 if graph is None else K  # graph replays launch the same kernels
    if a.workload == "policy":
        launches = K
    elif a.workload == "rollout":
        launches = 2 * K
    elif roll:
        launches = K // T if fused else 2 * K
    st = env.stats(reset=True)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    value = n * world * K / (ms * 1e-3)

    if a.workload != "step":
        if rank == 0:
            flops = 2.0 * sum(w.size for w in pol.weights)  # algorithmic MACs of the unpadded network x 2
            peaks = {}
            try:
                with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                    peaks = json.load(f)
            except Exception:
                pass
            tf_peak = float(peaks.get("bf16_tflops_sustained", 1407.0))
            out = {"metric": METRIC if a.workload != "policy" else "policy forwards/sec", "value": value,
                   "unit": UNIT if a.workload != "policy" else "obs/s", "n_gpus": world, "steps": K, "warmup": W,
                   "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                   "dtype": "bf16 operands / f32 accumulate (policy), f32 (env)", "data": "synthetic",
                   "config": {"workload": a.workload + "_" + workload_name(a), "policy": "24-120-120-120-4 ReLU "
                              "(c_code/neural_network.c weights), Gaussian noise + clip", "envs_per_gpu": n,
                              "launch": "cuda-graph x%d" % a.graph if graph is not None else
                              ("one fused launch per %d steps" % T if roll and fused else "per-step"),
                              **({"rollout_steps": T, "hbm_bytes_per_env_step_written": 4 * env.state_len + 37} if roll else {})},
                   "clocks": clocks, "gpu_launches": int(launches), "e2e": None}
            if a.workload == "policy":
                ach = n * flops / (ms / K * 1e-3) / 1e12
                out["roofline"] = {"bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
                                   "frac": ach / tf_peak, "traffic": None, "flops_per_obs": flops,
                                   "kernel": "qs::policy_kernel", "kernel_ms": ms / K,
                                   "peak_source": "measured bf16 sustained" if peaks else "fallback"}
            print(json.dumps(out))
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the C ABI with HOST (pinned) buffers: H2D actions, step, D2H obs/reward/done per step
    lib = env._lib
    D = env.state_len
    import ctypes as C
    def pin(nbytes):
        p = lib.qs_host_alloc(nbytes)
        if not p:
            raise SystemExit('qs_host_alloc failed')
        return p

    h_act, h_obs, h_rew, h_done = pin(n * 16), pin(n * D * 4), pin(n * 4), pin(n)
    src = acts[0].cpu().numpy()
    C.memmove(C.c_void_p(h_act), C.c_void_p(src.ctypes.data), n * 16)
    ke = max(a.e2e_steps, 3)
    for _ in range(3):
        env._call("qs_step_host", h_act, h_obs, h_rew, h_done, None, L.MODE_NORMAL, L.RESET_DEVICE)
    barrier()
    t0 = time.perf_counter()
    for _ in range(ke):
        env._call("qs_step_host", h_act, h_obs, h_rew, h_done, None, L.MODE_NORMAL, L.RESET_DEVICE)
    barrier()
    te = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([te], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        te = t.item()
    e2e_value = n * world * ke / te
    for p in (h_act, h_obs, h_rew, h_done):
        lib.qs_host_free(p)

    if rank == 0:
        bpe = lib.qs_algorithmic_bytes_per_env_step(L.E2E if a.variant == "e2e" else L.INDI, ga)
        peak, peak_src = FALLBACK_HBM_GBS, "fallback"
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peak, peak_src = float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                traffic = json.load(f).get(f"{a.variant}_N{n}")
        except Exception:
            pass
        kernel_ms = ms / K  # the timed region is exactly K launches of the step kernel, back to back
        # the library pins the INDI state in L2 by default (csrc/quadsim_capi.cu; QS_L2_HINTS overrides): say so, the
        # roofline fraction is still computed from the fixed algorithmic bytes and can then exceed 1
        hints = os.environ.get("QS_L2_HINTS", "1" if a.variant == "indi" else "0") not in ("0", "")
        state_mb = n * (56 if a.variant == "indi" else 92) / 1e6
        l2_note = (f"; state ({state_mb:.0f} MB) loaded/stored with L2 evict_last, streams evict_first: it stays resident "
                   f"between step launches up to {os.environ.get('QS_L2_KEEP_MB', '56')} MB") if hints else ""
        achieved = n * bpe / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "envs_per_gpu": n, "total_envs": n * world,
                       "gates_ahead": ga, "obs_dim": D, "reset": "fused device Philox", "l2": "inputs > L2: "
                       f"{n * (bpe + 16 * 3) / 1e6:.0f} MB touched per step vs 126 MB L2, no flush" + l2_note,
                       "parallelism": f"env-sharded x{world}, no data-path collective" +
                                      (" + obs all-gather (%s)" % a.gather_obs if gather is not None else ""),
                       "done_rate": st["dones"] / max(1, st["env_steps"]) if not a.no_stats else None,
                       "launch": "cuda-graph x%d" % a.graph if graph is not None else "per-step"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "bytes_per_env_step": bpe,
                         "kernel": f"qs::step_kernel<{a.variant}>", "kernel_ms": kernel_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * (D * 4 + 4 + 1), "steps": ke, "api": "qs_step_host (pinned host buffers)"},
            "gpu_launches": int(launches),
        }
        if world == 1 and not a.no_cpu_baseline:
            v, s, dt, cores = cpu_step_rate(a.variant, a.cpu_envs, ga, a.cpu_seconds)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"{s} steps of the first {a.cpu_envs} envs of the workload ({dt:.1f} s), "
                                             "oracle/quadsim_oracle.c with OpenMP", "host_cores": os.cpu_count()}
        print(json.dumps(out))
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

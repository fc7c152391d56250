#!/usr/bin/env python
"""bench.py -- quadrotor env-steps/sec of the B200 step kernel (BASELINE.json's metric).

  python bench.py [--gpus N --steps K --warmup W]            our arm (torchrun launches one rank per GPU for N>1)
  python bench.py --impl reference [...]                     the CPU arm: the reference's own NumPy step()

A "step" is one pass of the hot path (step_wait: residual MLPs + EoM + Euler + reward/flags + fused reset +
observation) over every env.  Headline workload: the end-to-end Bebop env, zigzag track, gates_ahead=1, training
disturbance ranges, N = 2**20 envs PER GPU (weak scaling; 2**20 envs touch ~350 MB per step, > the 126 MB L2, so no
L2 flush is needed between steps).  Rank 0 prints ONE JSON line.  Besides the headline fields the line carries

  configs    the other BASELINE.json configurations timed in the same process: at N=1  C2 (E2E N=4096), C3 (INDI
             N=262144) and INDI N=2**20, each with its own roofline; at N>1  config 4 exactly (N = 2**20 envs TOTAL,
             sharded over the ranks) without a collective, with the NCCL observation all-gather, and with the gather
             fused into the step kernel over peer memory -- ms/step, per-GPU roofline fraction, NVLink bytes;
  selfcheck  (N>1) sharded == unsharded and fused-P2P gather == NCCL gather on the live ranks, checked before any
             timing; a failure aborts the run;
  cpu_baseline  (N=1) the reference's own NumPy step() (oracle/_ref/reference, exec'd verbatim) on this box's host
             cores at the SAME N, next to the repo's C/OpenMP port as a second, labelled point.
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
L2_MB = 126.0


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
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="time budget of the C/OpenMP port leg")
    ap.add_argument("--cpu-ref-steps", type=int, default=12, help="steps of the reference's NumPy env in cpu_baseline")
    ap.add_argument("--cpu-ref-budget", type=float, default=150.0, help="wall-clock cap of a reference CPU leg, seconds")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="headline only: skip the `configs` / `selfcheck` records")
    ap.add_argument("--sub-steps", type=int, default=400, help="timed steps of each `configs` entry")
    ap.add_argument("--gather-obs", nargs="?", const="nccl", default=None, choices=["nccl", "p2p"],
                    help="add the all-gather of observations per step to the HEADLINE run: nccl = step kernel writes the "
                         "send slot, NCCL gathers in place; p2p = the step kernel stores its tiles into every peer itself")
    ap.add_argument("--workload", default="step", choices=["step", "policy", "rollout", "rollout_fused", "rollout_unfused"],
                    help="step: the env step alone on resident actions (the headline, default); policy: the on-device "
                         "controller forward alone (tcgen05); rollout: policy forward + env step per step, no host")
    ap.add_argument("--rollout-steps", type=int, default=32,
                    help="rollout_fused / rollout_unfused: env steps per qs_rollout[_fused] call (SB3 n_steps)")
    ap.add_argument("--no-stats", action="store_true", help="do not accumulate the device-side reward/flag totals")
    ap.add_argument("--graph", type=int, default=100,
                    help="replay the timed steps as CUDA graphs of (at most) this many steps, shortened to a divisor of "
                         "--steps (0 = one launch call per step).  The first step of a graph waits for everything before "
                         "it, the others are chained CTA by CTA: measured 58.9 / 57.4 / 56.4 / 55.8 us per step with "
                         "graphs of 4 / 8 / 20 / 100 steps (profiles/r2/quantisation_probe.log)")
    return ap.parse_args()


def workload_name(variant, ga, n, suffix="per_gpu"):
    track = "zigzag" if variant == "e2e" else "rectangle"
    return f"{variant}_{track}_ga{ga}_N{n}_{suffix}"


def track_for(variant):
    import optimal_quad_control_rl_b200 as Q
    return Q.zigzag_track() if variant == "e2e" else Q.rectangle_track()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------ CPU arms
def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_port_rate(variant, n_envs, gates_ahead, seconds, min_steps=3):
    """env-steps/s of the repo's own C restatement (oracle/quadsim_oracle.c, OpenMP over envs) on this host: a second,
    labelled CPU point -- NOT the reference.  Same track, ranges, random actions, reference-order resets."""
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
    acts = np.random.default_rng(1).uniform(-1, 1, (4, n_envs, 4)).astype(np.float32)
    for i in range(2):
        env.step(acts[i])
    steps, t0 = 0, time.perf_counter()
    while True:
        env.step(acts[steps % 4])
        steps += 1
        dt = time.perf_counter() - t0
        if steps >= min_steps and dt >= seconds:
            break
    return {"value": n_envs * steps / dt, "unit": UNIT, "cores": int(O.lib().qo_num_threads()), "kind": "port",
            "sample": f"{steps} steps of all {n_envs} envs ({dt:.1f} s), oracle/quadsim_oracle.c + OpenMP"}


def reference_available():
    try:
        from oracle import reference_exec as R
        import sympy  # noqa: F401 - the reference lambdifies its equations of motion at import
        return R.reference_available(), R.REFERENCE_ROOT
    except Exception as exc:  # pragma: no cover
        return False, f"{type(exc).__name__}: {exc}"


def cpu_reference_rate(variant, n_envs, gates_ahead, steps, warmup, budget_s, groups=3):
    """env-steps/s of the REFERENCE's own NumPy ``env.step`` (`3D quad race.ipynb:498-595`; INDI `:300-385`): the
    unmodified notebook cells exec'd by oracle/reference_exec.py from oracle/_ref/reference, same track / ranges /
    N / action distribution as the GPU arm.  ``steps`` are timed one by one; reported: total rate and the best of
    ``groups`` consecutive groups."""
    import numpy as np
    import torch

    from oracle import reference_exec as R

    nthreads = host_threads()
    torch.set_num_threads(nthreads)  # the two residual nn.Linear stacks are the only multi-threaded part
    t_load = time.perf_counter()
    env = R.make_reference_env(variant, n_envs, gates_ahead=gates_ahead)
    np.random.seed(0)
    env.reset()
    acts = np.random.default_rng(1).uniform(-1, 1, (2, n_envs, 4)).astype(np.float32)
    t_start = time.perf_counter()
    for i in range(max(1, warmup)):
        env.step(acts[i % 2])
        if time.perf_counter() - t_start > 0.3 * budget_s:
            break
    per = []
    t0 = time.perf_counter()
    for i in range(max(1, steps)):
        t = time.perf_counter()
        env.step(acts[i % 2])
        per.append(time.perf_counter() - t)
        if time.perf_counter() - t_start > budget_s:
            break
    dt = time.perf_counter() - t0
    k = len(per)
    g = max(1, k // groups)
    best = min(sum(per[j:j + g]) / g for j in range(0, k - g + 1, g))
    return {"value": n_envs * k / dt, "unit": UNIT, "cores": nthreads, "kind": "reference",
            "sample": f"{k} steps of all {n_envs} envs ({dt:.1f} s; setup incl. sympy lambdify {t_start - t_load:.1f} s)",
            "best_group_value": n_envs / best, "ms_per_step": 1e3 * dt / k,
            "threads": {"numpy_ufuncs": 1, "torch_intra_op": int(torch.get_num_threads()), "host_cores": os.cpu_count()},
            "what": "reference notebook cells exec'd verbatim (oracle/reference_exec.py), NumPy f32 + sympy-lambdified "
                    "f_func + torch CPU residual MLPs"}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return  # the CPU arm does not shard: rank 0 alone runs and prints
    ok, where = reference_available()
    n = a.num_envs
    cfg = {"workload": workload_name(a.variant, a.gates_ahead, n), "envs": n, "gates_ahead": a.gates_ahead}
    if ok:
        r = cpu_reference_rate(a.variant, n, a.gates_ahead, a.steps, min(a.warmup, 3), a.cpu_ref_budget)
        steps = int(r["sample"].split()[0])
        note = f"the reference's own NumPy step() from {os.path.relpath(where, ROOT)}"
    else:  # no staged reference / no sympy on this box: fall back to the port and say so
        r = cpu_port_rate(a.variant, n, a.gates_ahead, min(a.cpu_seconds, 60.0))
        steps = int(r["sample"].split()[0])
        r["ms_per_step"] = 1e3 * n / r["value"]
        note = f"reference unavailable here ({where}); this is the repo's C/OpenMP port"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": min(a.warmup, 3), "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "cpu_baseline": {k: r[k] for k in r if k != "ms_per_step"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": note,
    }))


# ------------------------------------------------------------------------------------------------ clocks / NVLink
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


def nvlink_kib(index):
    """(tx, rx) data KiB moved over all NVLinks of GPU `index` so far, or None: NVML's throughput field values (scope =
    all links), else the per-link counters `nvidia-smi nvlink -gt d` prints, summed."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        all_links = 0xFFFFFFFF  # scopeId UINT_MAX = sum over the links (scopeId 0 would be link 0 only)
        vals = nv.nvmlDeviceGetFieldValues(h, [(nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, all_links),
                                               (nv.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, all_links)])
        if any(v.nvmlReturn != 0 for v in vals):
            raise RuntimeError("field value unavailable")
        return tuple(int(v.value.ullVal) for v in vals)
    except Exception:
        pass
    try:
        import re
        import subprocess
        txt = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=10).stdout
        tx = sum(int(x) for x in re.findall(r"Data Tx:\s*(\d+)\s*KiB", txt))
        rx = sum(int(x) for x in re.findall(r"Data Rx:\s*(\d+)\s*KiB", txt))
        return (tx, rx) if re.search(r"Data Tx", txt) else None
    except Exception:
        return None


def pin_to_gpu_cpus(index):
    """Bind this rank to the CPU cores NVML reports as local to its GPU (NUMA-local pinned buffers and copy threads)."""
    try:
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        words = nv.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * w + b for w, m in enumerate(words) for b in range(64) if (int(m) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


# ------------------------------------------------------------------------------------------------ timing helpers
def median(xs):
    s = sorted(xs)
    return s[len(s) // 2] if s else None


class StepTimer:
    """Times K calls of ``one_step(i)`` on the current stream with CUDA events: W untimed warm-up steps, then the K steps
    either one launch call at a time or as K/g replays of a CUDA graph of g steps, with an event between replays so
    that the spread (min / median / max per replay) is known.  The reported total is first event -> last event."""

    def __init__(self, torch, dev, one_step, barrier, graph_steps, cycle=4):
        self.torch, self.dev, self.one_step, self.barrier = torch, dev, one_step, barrier
        self.g, self.cycle = int(graph_steps), cycle
        self.graph = None

    def prepare(self, W, K):
        torch = self.torch
        for i in range(W):
            self.one_step(i)
        g = self.g - self.g % self.cycle if self.g > 0 else 0   # whole action-buffer cycles, and a divisor of K
        while g >= self.cycle and K % g:
            g -= self.cycle
        self.g = g if g >= self.cycle else 0
        if self.g:
            side = torch.cuda.Stream(self.dev)
            side.wait_stream(torch.cuda.current_stream(self.dev))
            with torch.cuda.stream(side):
                self.one_step(0)
            torch.cuda.current_stream(self.dev).wait_stream(side)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                for i in range(self.g):
                    self.one_step(i)
            for _ in range(2):
                self.graph.replay()

    def run(self, K):
        torch = self.torch
        chunk = self.g if self.graph is not None else max(1, K // 20)
        n_chunks = K // chunk
        assert n_chunks * chunk == K or self.graph is None
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n_chunks + 1)]
        self.barrier()
        evs[0].record()
        done = 0
        for c in range(n_chunks):
            if self.graph is not None:
                self.graph.replay()
            else:
                hi = K if c == n_chunks - 1 else done + chunk
                for i in range(done, hi):
                    self.one_step(i)
                done = hi
            evs[c + 1].record()
        self.barrier()
        total = evs[0].elapsed_time(evs[-1])
        sizes = [chunk] * n_chunks
        if self.graph is None:
            sizes[-1] = K - chunk * (n_chunks - 1)
        per = [evs[c].elapsed_time(evs[c + 1]) / sizes[c] for c in range(n_chunks)]
        extra = 0
        if self.graph is not None and n_chunks < 5:  # a short run (--steps 20): the total above is the K timed steps and
            extra = 5 - n_chunks                       # nothing else; a few MORE replays only feed the spread estimate
            xe = [torch.cuda.Event(enable_timing=True) for _ in range(extra + 1)]
            xe[0].record()
            for c in range(extra):
                self.graph.replay()
                xe[c + 1].record()
            self.barrier()
            per += [xe[c].elapsed_time(xe[c + 1]) / chunk for c in range(extra)]
        return total, {"n": n_chunks + extra, "steps_each": chunk, "min": min(per), "median": median(per), "max": max(per),
                       "replays_outside_the_timed_region": extra}

    @property
    def launch_mode(self):
        return "cuda-graph x%d" % self.g if self.graph is not None else "per-step"


def flushed_step_ms(torch, dev, one_step, reps=30):
    """Median time of ONE step measured with a 256 MB write (> L2) in between, events around the step only: what a
    step costs when nothing of the previous one is left in the cache."""
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for i in range(reps):
        flush.fill_(i & 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        one_step(i)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    return median(ts)


def make_env(Q, variant, n, ga, dev, seed=0, env_offset=0):
    gp, gy, sp = track_for(variant)
    cls = Q.Quadcopter3DGates if variant == "e2e" else Q.Quadcopter3DGatesINDI
    env = cls(n, gp, gy, sp, gates_ahead=ga, device=dev, reset_rng="device", seed=seed, env_offset=env_offset, obs_buffers=2)
    if variant == "e2e":
        env.disturbance_ranges = Q.training_disturbance_ranges()
    return env


def sub_config(torch, Q, L, dev, name, variant, n, ga, steps, graph, barrier, peak):
    """One `configs` entry at N=1: its own env, resident random actions, W=50 warm-up + `steps` timed steps."""
    env = make_env(Q, variant, n, ga, dev)
    env.enable_stats(True)
    env.reset_tensor()
    gen = torch.Generator(device=dev).manual_seed(11)
    acts = [torch.rand((n, 4), generator=gen, device=dev) * 2 - 1 for _ in range(4)]
    one = lambda i: env.step_tensor(acts[i & 3])
    t = StepTimer(torch, dev, one, barrier, graph)
    t.prepare(50, steps)
    env.stats(reset=True)
    total, rep = t.run(steps)
    st = env.stats(reset=True)
    ms = total / steps
    bpe = env._lib.qs_algorithmic_bytes_per_env_step(L.E2E if variant == "e2e" else L.INDI, ga)
    touched_mb = n * (bpe + 16 * 3) / 1e6
    ach = n * bpe / (ms * 1e-3) / 1e9
    out = {"name": name, "workload": workload_name(variant, ga, n, "1gpu"), "value": n * steps / (total * 1e-3), "unit": UNIT,
           "ms_per_step": ms, "steps": steps, "warmup": 50, "launch": t.launch_mode, "replay_ms_per_step": rep,
           "done_rate": st["dones"] / max(1, st["env_steps"]),
           "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "bytes_per_env_step": bpe, "kernel": f"qs::step_kernel<{variant}>", "kernel_ms": ms}}
    if touched_mb < L2_MB:  # the workload itself is L2-resident: say so and add the cold-cache figure
        fl = flushed_step_ms(torch, dev, one)
        out["l2"] = (f"working set {touched_mb:.1f} MB per step < {L2_MB:.0f} MB L2: steady state is L2-resident (the state is "
                     "re-read every step by nature of the workload); ms_per_step is NOT flushed")
        out["ms_per_step_l2_flushed"] = fl
        out["roofline"]["frac_l2_flushed"] = n * bpe / (fl * 1e-3) / 1e9 / peak
    else:
        out["l2"] = f"inputs > L2: {touched_mb:.0f} MB touched per step vs {L2_MB:.0f} MB L2, no flush"
    env.close()
    return out


# ------------------------------------------------------------------------------------------------ multi-GPU records
def selfcheck_multi(torch, dist, Q, dev, rank, world):
    """Body of tests/test_gpu_multi.py on the live ranks: (1) the fused peer-memory gather equals the NCCL gather bit
    for bit on every rank for 8 steps with resets; (2) on rank 0 the gathered observations of the sharded job equal the
    unsharded env's.  Returns a dict; raises on mismatch."""
    total = 25 * 128 * world  # 25 tiles per rank
    gp, gy, sp = Q.zigzag_track()
    first, count = Q.shard_range(total, rank, world)

    def make(n, off):
        env = Q.Quadcopter3DGates(n, gp, gy, sp, gates_ahead=1, device=dev, reset_rng="device", seed=4, env_offset=off)
        env.disturbance_ranges = Q.training_disturbance_ranges()
        env.max_steps = 5
        return env

    gen = torch.Generator(device=dev).manual_seed(7)  # same seed on every rank: the global action tensor
    acts = [(torch.rand((total, 4), generator=gen, device=dev) * 2 - 1) for _ in range(8)]
    e_nccl, e_p2p = make(count, first), make(count, first)
    g_nccl = Q.ObsAllGather(total, e_nccl.state_len, dev)
    g_p2p = Q.ObsPeerGather(total, e_p2p.state_len, dev)
    g_p2p.attach(e_p2p)
    e_nccl.reset_tensor()
    e_p2p.reset_tensor()
    ok_gather, full_p2p = True, None
    for t in range(8):
        a = acts[t][first:first + count].contiguous()
        e_nccl.step_tensor(a, obs_out=g_nccl.local_slot())
        full_nccl = g_nccl.gather().clone()
        e_p2p.step_tensor(a, obs_out=g_p2p.local_slot())
        full_p2p = g_p2p.gather().clone()
        ok_gather &= bool(torch.equal(full_nccl, full_p2p)) and full_p2p.abs().sum().item() > 0
    # (3) the packed BF16 gather: gathered blocks == round-to-nearest BF16 of the float32 gather, and the policy agrees bit for bit
    e_pk = make(count, first)
    g_pk = Q.ObsPeerGather(total, e_pk.state_len, dev, packed=True)
    g_pk.attach(e_pk)
    e_pk.reset_tensor(obs_out=g_pk.local_slot())
    for t in range(8):
        e_pk.step_tensor(acts[t][first:first + count].contiguous(), obs_out=g_pk.local_slot())
        full_pk = g_pk.gather()
    ch = (e_pk.state_len + 7) // 8  # K chunks that travel: 3 for the 24-wide row = 48 B per env
    rows = full_pk.view(-1, ch, 32, 8, 2).permute(0, 2, 1, 3, 4).reshape(-1, 8 * ch, 2)[:total].contiguous().view(torch.int16).reshape(total, 8 * ch)
    want = torch.zeros((total, 32), device=dev)
    want[:, :e_pk.state_len] = full_p2p
    want[:, e_pk.state_len] = 1.0
    pol = Q.MlpPolicy.from_npz(device=dev)
    ok_packed = bool(torch.equal(rows, want[:, :8 * ch].to(torch.bfloat16).view(torch.int16))) and \
        bool(torch.equal(pol.forward_packed(full_pk, total, deterministic=True), pol.forward(full_p2p, deterministic=True)))
    e_pk.close()
    ok_shard = True
    if rank == 0:
        ref = make(total, 0)
        ref.reset_tensor()
        for t in range(8):
            o = ref.step_tensor(acts[t])[0]
        ok_shard = bool(torch.equal(o, full_p2p))
        ref.close()
    flags = torch.tensor([int(ok_gather), int(ok_shard), int(ok_packed)], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    e_nccl.close()
    e_p2p.close()
    res = {"p2p_gather_equals_nccl": bool(flags[0].item()), "sharded_equals_unsharded": bool(flags[1].item()),
           "packed_bf16_gather_equals_packed_f32_and_policy_agrees": bool(flags[2].item()), "envs": total, "steps": 8, "ranks": world}
    res["status"] = "ok" if all(bool(f) for f in flags.tolist()) else "FAILED"
    return res


def config4(torch, dist, Q, L, dev, rank, world, local, steps, barrier, peak, total=1 << 20, ga=1):
    """BASELINE config 4 as stated: N = 2**20 E2E envs in TOTAL sharded over the ranks, with / without the observation
    all-gather.  Device-timed, max over ranks."""
    first, count = Q.shard_range(total, rank, world)
    bpe = 189 + 4 * (20 + 4 * ga)
    out = []
    modes = ("none", "nccl", "p2p") + (("p2p_bf16",) if all(Q.shard_range(total, r, world)[0] % 128 == 0 for r in range(world)) else ())
    for mode in modes:
        env = make_env(Q, "e2e", count, ga, dev, env_offset=first)
        gen = torch.Generator(device=dev).manual_seed(21 + rank)
        acts = [torch.rand((count, 4), generator=gen, device=dev) * 2 - 1 for _ in range(4)]
        gather = None
        if mode == "nccl":
            gather = Q.ObsAllGather(total, env.state_len, dev)
        elif mode in ("p2p", "p2p_bf16"):
            gather = Q.ObsPeerGather(total, env.state_len, dev, packed=mode == "p2p_bf16")
            gather.attach(env)
        if mode == "p2p_bf16":
            env.reset_tensor(obs_out=gather.local_slot())
        else:
            env.reset_tensor()

        def one(i):
            env.step_tensor(acts[i & 3], obs_out=None if gather is None else gather.local_slot())
            if gather is not None:
                gather.gather()

        t = StepTimer(torch, dev, one, barrier, 20 if gather is None else 0)
        t.prepare(30, steps)
        nv0 = nvlink_kib(local)
        total_ms, rep = t.run(steps)
        nv1 = nvlink_kib(local)
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = tt.item() / steps
        rec = {"name": "C4_" + mode, "workload": workload_name("e2e", ga, total, f"total_over_{world}gpus"),
               "gather": {"none": "no collective (data-parallel policy)", "nccl": "step kernel writes the send slot; in-place "
                          "ncclAllGather", "p2p": "fused: step kernel bulk-stores its tiles into every peer (symmetric memory), "
                          "double-buffered, one barrier", "p2p_bf16": "fused, observations packed as BF16 in the on-device "
                          "policy's operand layout (48 B per env instead of 96; qs_set_obs_format), consumed by "
                          "qs_policy_forward_packed"}[mode],
               "value": total / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "envs_per_gpu": count,
               "launch": t.launch_mode, "replay_ms_per_step": rep,
               "per_gpu_roofline_frac": count * bpe / (ms * 1e-3) / 1e9 / peak,
               "nvlink_algorithmic_rx_bytes_per_gpu_per_step": 0 if gather is None else
               (total - count) * (16 * ((env.state_len + 7) // 8) if mode == "p2p_bf16" else env.state_len * 4)}
        if gather is not None:  # what the gather must deliver per GPU, over the measured step time
            rec["nvlink_algorithmic_rx_GBps_per_gpu"] = rec["nvlink_algorithmic_rx_bytes_per_gpu_per_step"] / (ms * 1e-3) / 1e9
        if nv0 and nv1 and (gather is None or nv1[1] > nv0[1]):
            rec["nvlink_measured_bytes_per_step_rank0"] = {"tx": (nv1[0] - nv0[0]) * 1024 / steps, "rx": (nv1[1] - nv0[1]) * 1024 / steps}
            if gather is not None:
                rec["nvlink_rx_GBps_rank0"] = (nv1[1] - nv0[1]) * 1024 / steps / (ms * 1e-3) / 1e9
        elif gather is not None:  # counters that stand still while a verified gather runs are not a measurement of zero
            rec["nvlink_measured_bytes_per_step_rank0"] = None
            rec["nvlink_counters"] = ("unavailable" if not (nv0 and nv1) else
                                      "NVML / nvidia-smi NVLink data counters did not advance on this box during a gather "
                                      "whose result the selfcheck verifies (not exposed to the container)")
        out.append(rec)
        env.close()
        del gather
        barrier()
    return out


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
    pinned_cpus = pin_to_gpu_cpus(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n, ga = a.num_envs, a.gates_ahead
    peak, peak_src = hbm_peak()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- multi-GPU self-check first: nothing is timed on ranks whose results disagree
    selfcheck = None
    if world > 1 and not a.no_configs and a.workload == "step":
        selfcheck = selfcheck_multi(torch, dist, Q, dev, rank, world)
        if selfcheck["status"] != "ok":
            if rank == 0:
                print(json.dumps({"metric": METRIC, "value": None, "unit": UNIT, "n_gpus": world, "selfcheck": selfcheck,
                                  "error": "multi-GPU self-check failed; nothing was timed"}))
            dist.destroy_process_group()
            raise SystemExit(3)

    env = make_env(Q, a.variant, n, ga, dev, seed=0, env_offset=rank * n)
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
        pol = Q.MlpPolicy.reference_controller(device=dev, seed=3, env_offset=rank * n)
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
    if gather is not None:
        a.graph = 0  # the collective is issued by torch.distributed per step
    timer = StepTimer(torch, dev, one_step, barrier, a.graph)
    timer.prepare(W, K)
    env.stats(reset=True)
    l0 = env.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms, replays = timer.run(K)
    clocks = sampler.stop()
    graphed = timer.graph is not None
    launches = env.launch_count - l0 if not graphed else K  # graph replays launch the same kernels
    if a.workload == "policy":
        launches = K
    elif a.workload == "rollout":
        launches = 2 * K
    elif roll:
        launches = K // T if fused else 2 * K
    st = env.stats(reset=True)
    chained = env.chained_launch_count  # captured steps that wait for their predecessor CTA by CTA (per graph, not per replay)
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
                   "config": {"workload": a.workload + "_" + workload_name(a.variant, ga, n), "policy": "24-120-120-120-4 ReLU "
                              "(c_code/neural_network.c weights), Gaussian noise + clip", "envs_per_gpu": n,
                              "launch": timer.launch_mode if not roll else
                              ("one fused launch per %d steps" % T if fused else "per-step"),
                              **({"rollout_steps": T, "hbm_bytes_per_env_step_written": 4 * env.state_len + 37} if roll else {})},
                   "replay_ms_per_step": replays, "clocks": clocks, "gpu_launches": int(launches), "e2e": None}
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
    env.close()

    # ---- the other BASELINE configurations, same process
    configs = []
    if not a.no_configs:
        if world == 1:
            for name, variant, nn, graph in (("C2", "e2e", 4096, a.graph), ("C3", "indi", 262144, a.graph), ("INDI_2^20", "indi", 1 << 20, a.graph)):
                configs.append(sub_config(torch, Q, L, dev, name, variant, nn, 1, a.sub_steps, graph, barrier, peak))
        else:
            configs = config4(torch, dist, Q, L, dev, rank, world, local, a.sub_steps, barrier, peak)

    if rank == 0:
        bpe = lib.qs_algorithmic_bytes_per_env_step(L.E2E if a.variant == "e2e" else L.INDI, ga)
        traffic, traffic_note = None, None
        try:
            with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
                tj = json.load(f)
            traffic, traffic_note = tj.get(f"{a.variant}_N{n}"), tj.get("_how")
        except Exception:
            pass
        kernel_ms = ms / K  # the timed region is exactly K launches of the step kernel, back to back
        # the library asks the L2 to keep the state by default (csrc/quadsim_capi.cu; QS_L2_HINTS overrides): say so, the
        # roofline fraction is still computed from the fixed algorithmic bytes and can then exceed 1
        hints = os.environ.get("QS_L2_HINTS", "1") not in ("0", "")
        state_mb = n * (56 if a.variant == "indi" else 92) / 1e6
        l2_note = (f"; state ({state_mb:.0f} MB) loaded/stored with L2 evict_last, streams evict_first: it stays resident "
                   f"between step launches up to {os.environ.get('QS_L2_KEEP_MB', '56')} MB") if hints else ""
        achieved = n * bpe / (kernel_ms * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a.variant, ga, n), "envs_per_gpu": n, "total_envs": n * world,
                       "gates_ahead": ga, "obs_dim": D, "reset": "fused device Philox", "l2": "inputs > L2: "
                       f"{n * (bpe + 16 * 3) / 1e6:.0f} MB touched per step vs 126 MB L2, no flush" + l2_note,
                       "parallelism": f"env-sharded x{world}, no data-path collective" +
                                      (" + obs all-gather (%s)" % a.gather_obs if gather is not None else ""),
                       "done_rate": st["dones"] / max(1, st["env_steps"]) if not a.no_stats else None,
                       "launch": timer.launch_mode, "cpu_affinity_cores": pinned_cpus,
                       # steps inside a graph that depend on their predecessor CTA by CTA instead of grid by grid
                       "chained_launches_per_graph": chained},
            "replay_ms_per_step": replays,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "traffic_source": traffic_note, "peak_source": peak_src,
                         "bytes_per_env_step": bpe, "kernel": f"qs::step_kernel<{a.variant}>", "kernel_ms": kernel_ms},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 16,
                    "d2h_bytes_per_step": n * (D * 4 + 4 + 1), "steps": ke, "api": "qs_step_host (pinned host buffers)"},
            "gpu_launches": int(launches),
        }
        if configs:
            out["configs"] = configs
        if selfcheck is not None:
            out["selfcheck"] = selfcheck
        if world == 1 and not a.no_cpu_baseline:
            ok, where = reference_available()
            port = cpu_port_rate(a.variant, n, ga, a.cpu_seconds)
            if ok:
                out["cpu_baseline"] = cpu_reference_rate(a.variant, n, ga, a.cpu_ref_steps, 2, a.cpu_ref_budget)
                out["cpu_baseline"]["same_config"] = True
                out["cpu_port"] = port
            else:
                out["cpu_baseline"] = dict(port, note=f"reference unavailable on this box ({where}): the port stands in")
        print(json.dumps(out))
    if world > 1:
        barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)

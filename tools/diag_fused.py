#!/usr/bin/env python
"""Where does the fused rollout differ from the unfused one? (diagnostic, GPU)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import golden
from test_gpu_rollout_fused import make, KEYS

k = golden("kat")
tracks = {v: (k[f"{v}_gate_pos"], k[f"{v}_gate_yaw"], k[f"{v}_start_pos"]) for v in ("e2e", "indi")}
variant, n, steps = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
det = len(sys.argv) > 4 and sys.argv[4] == "det"
outs = []
for fused in (True, False):
    env, pol = make(variant, n, tracks)
    env.reset_tensor()
    r = env.rollout(pol, steps, deterministic=det, fused=fused)
    torch.cuda.synchronize()
    outs.append({k: r[k].cpu().numpy() for k in KEYS})
a, b = outs
for key in KEYS:
    d = a[key] != b[key]
    print(key, "diffs", int(d.sum()), "of", d.size)
d = a["obs"] != b["obs"]
if d.any():
    t_idx, e_idx, c_idx = np.nonzero(d)
    print("first t with obs diff:", t_idx.min(), " per-t counts:", np.bincount(t_idx, minlength=steps + 1).tolist())
    print("columns:", np.bincount(c_idx, minlength=a["obs"].shape[2]).tolist())
    t0 = t_idx.min()
    sel = (t_idx == t0)
    for e_, c_ in list(zip(e_idx[sel], c_idx[sel]))[:12]:
        x, y = a["obs"][t0, e_, c_], b["obs"][t0, e_, c_]
        prev_done = a["dones"][t0 - 1, e_] if t0 > 0 else -1
        w0 = (e_ // 32) * 32
        nd = int(a["dones"][t0 - 1, w0:w0 + 32].sum()) if t0 > 0 else -1
        print(f" t={t0} env={e_} (lane {e_%32}) col={c_}: fused {x!r} unfused {y!r} diff {float(x)-float(y):.3e} prev_done={prev_done} dones_in_warp={nd}")
    da = a["actions"] != b["actions"]
    if da.any():
        print("first t with action diff:", np.nonzero(da)[0].min())

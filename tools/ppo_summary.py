#!/usr/bin/env python
"""Summarise a tools/train_ppo.py log (one JSON line per iteration): wall-clock to reward plateau.
Plateau = first iteration whose 5-iteration mean of ep_rew_mean is within 2 % (of the final level's magnitude) of
the mean over the last 10 % of the run."""
import json
import sys

import numpy as np


def summarise(path):
    rows = [json.loads(l) for l in open(path) if l.startswith("{")]
    r = np.array([x["ep_rew_mean"] for x in rows])
    t = np.array([x["wall_s"] for x in rows])
    k = max(3, len(rows) // 10)
    final = float(np.nanmean(r[-k:]))
    sm = np.convolve(r, np.ones(5) / 5, mode="valid")
    hit = np.flatnonzero(sm >= final - 0.02 * abs(final))
    i = int(hit[0]) + 4 if hit.size else len(rows) - 1
    out = {"iterations": len(rows), "timesteps": rows[-1]["timesteps"], "wall_s": round(float(t[-1]), 1),
           "final_ep_rew_mean": round(final, 2), "final_gates_per_episode": round(float(np.mean([x["gates_per_episode"] for x in rows[-k:]])), 2),
           "final_ep_len_mean": round(float(np.mean([x["ep_len_mean"] for x in rows[-k:]])), 1),
           "plateau_wall_s": round(float(t[i]), 1), "plateau_timesteps": rows[i]["timesteps"],
           "collect_s_per_iter": round(float(np.mean([x["collect_s"] for x in rows[1:]])), 4),
           "train_s_per_iter": round(float(np.mean([x["train_s"] for x in rows[1:]])), 4),
           "env_steps_per_s_overall": round(rows[-1]["timesteps"] / float(t[-1])),
           "nan": bool(any(not np.isfinite(x["pg_loss"]) for x in rows)),
           "rolled_back": int(sum(bool(x.get("rolled_back")) for x in rows)),
           "curve_every_10": [[round(float(t[j]), 1), round(float(r[j]), 2)] for j in range(0, len(rows), 10)]}
    return out


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(json.dumps({"log": p, **summarise(p)}))

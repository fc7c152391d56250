"""Trajectory export adapter: the simulation log the reference's analysis notebooks read.

The reference records one env of a (test) run step by step -- ``world_states[0]``, the action, ``step_counts[0]*dt`` --
and saves a dict of columns with ``np.savez`` (`3D quad race INDI inner loop.ipynb:647-695`); `Analyse Flight
Data.ipynb` / `Figures for paper.ipynb` then load those files next to real flight logs.  ``TrajectoryLog`` keeps that
schema (keys ``t x y z vx vy vz V phi theta psi u1 u2 u3 u4 u``, actions rescaled to [0,1]) over any env with the
reference's public attributes: the GPU env (one small device-to-host read of the selected env per step through
``qs_get_state``), or a rollout buffer produced on the device.  Host-side NumPy only; nothing here is on the hot path."""
from __future__ import annotations

import os

import numpy as np

STATE_KEYS = ("x", "y", "z", "vx", "vy", "vz", "phi", "theta", "psi")


class TrajectoryLog:
    def __init__(self, env=None, index=0):
        self.env, self.index = env, int(index)
        self.time, self.state_traj, self.action_traj = [], [], []

    def record(self, actions, env=None):
        """Call after ``env.step(actions)``: appends the selected env's world state, action and episode time."""
        env = env or self.env
        i = self.index
        if hasattr(env, "_get_rows"):
            ws, sc = env._get_rows(i, 1)
            ws, sc = ws[0], sc[0]
        else:
            ws, sc = np.asarray(env.world_states)[i], np.asarray(env.step_counts)[i]
        self.state_traj.append(np.array(ws, dtype=np.float32))
        self.action_traj.append(np.array(np.asarray(actions)[i], dtype=np.float32))
        self.time.append(sc * env.dt)

    def extend(self, world_states, actions, times):
        """Bulk form: ``world_states (T, >=9)``, ``actions (T, 4)`` in [-1,1], ``times (T,)``."""
        self.state_traj += [np.asarray(w, np.float32) for w in world_states]
        self.action_traj += [np.asarray(a, np.float32) for a in actions]
        self.time += list(times)

    def __len__(self):
        return len(self.time)

    def as_dict(self):
        s = np.array(self.state_traj, dtype=np.float32).reshape(len(self.time), -1)
        a = np.array(self.action_traj, dtype=np.float32).reshape(len(self.time), 4)
        d = {"t": self.time}
        for k, name in enumerate(STATE_KEYS):
            d[name] = s[:, k]
        d["V"] = np.sqrt(s[:, 3] ** 2 + s[:, 4] ** 2 + s[:, 5] ** 2)
        for k in range(4):
            d[f"u{k + 1}"] = (a[:, k] + 1) / 2
        d["u"] = np.stack([d["u1"], d["u2"], d["u3"], d["u4"]], axis=1)
        order = ("t", "x", "y", "z", "vx", "vy", "vz", "V", "phi", "theta", "psi", "u1", "u2", "u3", "u4", "u")
        return {k: d[k] for k in order}

    def save(self, name, folder="flight_data/simulation_logs"):
        """``np.savez(folder/name, **log_dict)`` (`:688-695`); returns the path written."""
        os.makedirs(folder, exist_ok=True)
        path = os.path.join(folder, name)
        np.savez(path, **self.as_dict())
        return path if path.endswith(".npz") else path + ".npz"


def log_policy_run(model, env, steps, deterministic=False, index=0):
    """The notebook's logging loop: ``model.predict(env.states)`` -> ``env.step`` -> record, ``steps`` times."""
    log = TrajectoryLog(env, index)
    env.reset()
    for _ in range(int(steps)):
        actions, _ = model.predict(env.states, deterministic=deterministic)
        env.step(actions)
        log.record(actions)
    return log

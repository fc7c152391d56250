"""TEST INFRASTRUCTURE — ctypes binding of ``oracle/libquadsim_oracle.so`` plus ``OracleEnv``, a CPU env with the
reference's ``Quadcopter3DGates`` surface (`3D quad race.ipynb:287-620`, INDI `:142-410`).

``OracleEnv`` holds the same array-of-structs NumPy attributes as the reference (``world_states (N,16|13)``,
``states (N,D)``, ``target_gates``, ``step_counts``, ``disturbances`` ...) and draws its resets from the global
``np.random`` stream in the reference's order, so seeded runs line up draw for draw.  The arithmetic of a step
is done by the C restatement.  Only tests, ``smoke()`` and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libquadsim_oracle.so")
REF_MLP_LIB_PATH = os.path.join(HERE, "_ref", "libnn_residual_ref.so")

E2E, INDI = 0, 1
MODE_NORMAL, MODE_PAUSE_IF_COLLISION, MODE_PAUSE = 0, 1, 2
F_DONE, F_TRUNC, F_GATE_PASSED, F_GATE_COLLISION, F_GROUND, F_OOB = 1, 2, 4, 8, 16, 32

_fp = C.POINTER(C.c_float)
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int64)
_bp = C.POINTER(C.c_uint8)


class QoParams(C.Structure):
    _fields_ = [("variant", C.c_int), ("n_gates", C.c_int), ("gates_ahead", C.c_int), ("ranges_f64", C.c_int),
                ("max_steps", C.c_int64), ("dt", C.c_float),
                ("gate_pos", _fp), ("gate_yaw", _fp), ("gate_cos", _fp), ("gate_sin", _fp),
                ("gate_pos_rel", _fp), ("gate_yaw_rel", _fp), ("dist_ranges", _dp),
                ("thrust_w", _fp), ("moment_w", _fp)]


def build(force=False):
    """Compile the oracle (and oracle/_ref when the reference is mounted) with the committed Makefile."""
    if force or not os.path.isfile(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(
            os.path.join(HERE, "quadsim_oracle.c")):
        subprocess.run(["make", "-C", HERE, "libquadsim_oracle.so"], check=True, capture_output=True)
    if os.path.isdir("/root/reference/c_code") and (force or not os.path.isfile(REF_MLP_LIB_PATH)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(LIB_PATH)
        pp = C.POINTER(QoParams)
        L.qo_num_threads.restype = C.c_int
        L.qo_set_num_threads.argtypes = [C.c_int]
        L.qo_set_num_threads.restype = None
        L.qo_track_tables.argtypes = [C.c_int, _fp, _fp, _fp, _fp]
        L.qo_body_velocity.argtypes = [_fp, C.c_int64, _fp]
        L.qo_residual_mlp.argtypes = [pp, _fp, C.c_int64, _fp, _fp]
        L.qo_euler.argtypes = [pp, _fp, _fp, _fp, C.c_int64, _fp, _fp, _fp]
        L.qo_observe.argtypes = [pp, _fp, _fp, _ip, C.c_int64, _fp]
        L.qo_step.argtypes = [pp, C.c_int, C.c_int64, _fp, _fp, _ip, _ip, _fp, _fp, _fp, _fp, _fp, _bp, _bp, _fp]
        L.qo_step.restype = C.c_int64
        L.qo_apply_reset.argtypes = [pp, C.c_int64, _fp, _fp, _ip, _ip, _bp, _fp, _fp]
        _LIB = L
    return _LIB


def ref_mlp_lib():
    """The reference's own generated C (c_code/nn_thrust.c, nn_moment.c) compiled into oracle/_ref, or None."""
    if not os.path.isfile(REF_MLP_LIB_PATH):
        return None
    L = C.CDLL(REF_MLP_LIB_PATH)
    L.nn_thrust_forward.argtypes = [_fp, _fp]
    L.nn_moment_forward.argtypes = [_fp, _fp]
    return L


def _f(a):
    return None if a is None else a.ctypes.data_as(_fp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _b(a):
    return None if a is None else a.ctypes.data_as(_bp)


def load_residual_weights(path=None):
    """Packaged copy of NNDroneModel/*.pt (written by oracle/make_golden.py) -> (thrust[289], moment[451])."""
    path = path or os.path.join(os.path.dirname(HERE), "optimal_quad_control_rl_b200", "data", "residual_mlp.npz")
    z = np.load(path)
    pack = lambda n: np.concatenate([z[f"{n}_w1"].ravel(), z[f"{n}_b1"].ravel(), z[f"{n}_w2"].ravel(),
                                     z[f"{n}_b2"].ravel()]).astype(np.float32)
    return pack("thrust"), pack("moment")


def track_tables(gate_pos, gate_yaw):
    gp = np.ascontiguousarray(gate_pos, np.float32)
    gy = np.ascontiguousarray(gate_yaw, np.float32)
    pr = np.zeros_like(gp)
    yr = np.zeros_like(gy)
    lib().qo_track_tables(len(gy), _f(gp), _f(gy), _f(pr), _f(yr))
    return pr, yr


class OracleEnv:
    """CPU stand-in for the reference env, arithmetic in C.  ``variant`` in {"e2e", "indi"}."""

    def __init__(self, variant, num_envs, gates_pos, gate_yaw, start_pos, gates_ahead=0, pause_if_collision=False,
                 numpy_trig_tables=True):
        self.variant = {"e2e": E2E, "indi": INDI}[variant]
        self.num_envs = int(num_envs)
        self.start_pos = np.asarray(start_pos).astype(np.float32)
        self.gate_pos = np.ascontiguousarray(np.asarray(gates_pos).astype(np.float32))
        self.gate_yaw = np.ascontiguousarray(np.asarray(gate_yaw).astype(np.float32))
        self.num_gates = self.gate_pos.shape[0]
        self.gates_ahead = int(gates_ahead)
        self.pause_if_collision = pause_if_collision
        self.gate_pos_rel, self.gate_yaw_rel = track_tables(self.gate_pos, self.gate_yaw)
        # cos/sin of the gate yaw as NumPy's float32 ufuncs give them (what the reference evaluates every step)
        self._gcos = np.cos(self.gate_yaw) if numpy_trig_tables else None
        self._gsin = np.sin(self.gate_yaw) if numpy_trig_tables else None
        self.ns = 16 if self.variant == E2E else 13
        self.state_len = (20 if self.variant == E2E else 13) + 4 * self.gates_ahead
        self.target_gates = np.zeros(num_envs, dtype=np.int64)
        self.world_states = np.zeros((num_envs, self.ns), np.float32)
        self.states = np.zeros((num_envs, self.state_len), np.float32)
        self.max_steps = 1200
        self.dt = np.float32(0.01)
        self.step_counts = np.zeros(num_envs, dtype=np.int64)
        self.actions = np.zeros((num_envs, 4), np.float32)
        self.dones = np.zeros(num_envs, dtype=bool)
        self.disturbance_ranges = np.zeros((6, 2), np.float32)
        self.disturbances = np.zeros((num_envs, 6), np.float32)
        self.disturbance_scale = 1
        self.pause = False
        self.last_flags = np.zeros(num_envs, np.uint8)
        self._tw, self._mw = load_residual_weights()

    # -- parameter block (rebuilt per call: attributes may be poked between steps, SURVEY section 5 "Config")
    def _params(self):
        dr = np.ascontiguousarray(self.disturbance_ranges, np.float64)
        self._keep = (dr,)
        return QoParams(self.variant, self.num_gates, self.gates_ahead,
                        int(np.asarray(self.disturbance_ranges).dtype == np.float64), int(self.max_steps),
                        float(self.dt), _f(self.gate_pos), _f(self.gate_yaw), _f(self._gcos), _f(self._gsin),
                        _f(self.gate_pos_rel), _f(self.gate_yaw_rel), dr.ctypes.data_as(_dp), _f(self._tw),
                        _f(self._mw))

    def update_states(self):
        p = self._params()
        obs = np.zeros((self.num_envs, self.state_len), np.float32)
        ws = np.ascontiguousarray(self.world_states, np.float32)
        lib().qo_observe(C.byref(p), _f(ws), _f(self.disturbances), _i(self.target_gates), self.num_envs, _f(obs))
        self.states = obs

    def _draw_reset(self, n):
        """The 16|13 + 6 field-major draws of reset_ (`3D quad race.ipynb:455-489`), f64 -> f32 on store."""
        u = np.random.uniform
        cols = [u(-0.5, 0.5, size=(n,)) + self.start_pos[0], u(-0.5, 0.5, size=(n,)) + self.start_pos[1],
                u(-0.5, 0.5, size=(n,)) + self.start_pos[2]]
        cols += [u(-0.5, 0.5, size=(n,)) for _ in range(3)]
        cols += [u(-np.pi / 9, np.pi / 9, size=(n,)), u(-np.pi / 9, np.pi / 9, size=(n,)), u(-np.pi, np.pi, size=(n,))]
        cols += [u(-0.1, 0.1, size=(n,)) for _ in range(3)]
        if self.variant == E2E:
            cols += [u(-1, 1, size=(n,)) for _ in range(4)]
        else:
            cols += [u(-.1, .1, size=(n,))]
        ws = np.stack(cols, axis=1).astype(np.float32)
        dist = None
        if self.variant == E2E:
            r = self.disturbance_ranges
            dist = (self.disturbance_scale * np.stack([u(r[k, 0], r[k, 1], size=(n,)) for k in range(6)], axis=1)
                    ).astype(np.float32)
        return ws, dist

    def reset_(self, dones):
        n = int(dones.sum())
        ws, dist = self._draw_reset(n)
        self.world_states[dones] = ws
        self.step_counts[dones] = 0
        self.target_gates[dones] = 0
        if dist is not None:
            self.disturbances[dones] = dist
        self.update_states()
        return self.states

    def reset(self):
        return self.reset_(np.ones(self.num_envs, dtype=bool))

    def step_async(self, actions):
        self.actions = actions

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def step_wait(self):
        L, p, n = lib(), self._params(), self.num_envs
        act = np.ascontiguousarray(self.actions, np.float32)
        mode = MODE_PAUSE if self.pause else (MODE_PAUSE_IF_COLLISION if self.pause_if_collision else MODE_NORMAL)
        rew = np.empty(n, np.float32)
        done = np.empty(n, np.uint8)
        flags = np.empty(n, np.uint8)
        ws = np.ascontiguousarray(self.world_states, np.float32).copy()  # reference rebinds, never mutates
        dist = self.disturbances
        if mode == MODE_NORMAL:
            # advance + flags in C, then draw from np.random in the reference's order, then masked store + obs
            L.qo_step(C.byref(p), mode, n, _f(ws), _f(dist), _i(self.target_gates), _i(self.step_counts), _f(act),
                      None, None, None, _f(rew), _b(done), _b(flags), None)
            rws, rdist = self._draw_reset(int(done.sum()))
            L.qo_apply_reset(C.byref(p), n, _f(ws), _f(dist), _i(self.target_gates), _i(self.step_counts), _b(done),
                             _f(rws), _f(rdist))
            obs = np.zeros((n, self.state_len), np.float32)
            L.qo_observe(C.byref(p), _f(ws), _f(dist), _i(self.target_gates), n, _f(obs))
            self.states = obs
        elif mode == MODE_PAUSE_IF_COLLISION:
            obs = np.zeros((n, self.state_len), np.float32)
            L.qo_step(C.byref(p), mode, n, _f(ws), _f(dist), _i(self.target_gates), _i(self.step_counts), _f(act),
                      None, None, _f(obs), _f(rew), _b(done), _b(flags), None)
            self.states = obs
        else:
            L.qo_step(C.byref(p), mode, n, _f(ws), _f(dist), _i(self.target_gates), _i(self.step_counts), _f(act),
                      None, None, None, _f(rew), _b(done), _b(flags), None)
        if mode != MODE_PAUSE:
            self.world_states = ws
        self.last_flags = flags
        dones = done.astype(bool)
        self.dones = dones
        # the aliased-dict quirk of `:589-594`: one dict shared by every env
        info = {}
        idx = np.flatnonzero(dones)
        if idx.size:
            info["terminal_observation"] = self.states[idx[-1]]
        if (flags & F_TRUNC).any():
            info["TimeLimit.truncated"] = True
        return self.states, rew, dones, [info] * n

    # -- teacher forcing helper for tests
    def force(self, ws, tg, sc, dist=None):
        self.world_states = np.ascontiguousarray(ws, np.float32).copy()
        self.target_gates = np.asarray(tg, np.int64).copy()
        self.step_counts = np.asarray(sc, np.int64).copy()
        if dist is not None:
            self.disturbances = np.ascontiguousarray(dist, np.float32).copy()
        self.update_states()


def euler(env: OracleEnv, ws, act, dist=None):
    """new_states_raw (+ residual thrust/moment for E2E) for arbitrary inputs."""
    p = env._params()
    n = len(ws)
    ws = np.ascontiguousarray(ws, np.float32)
    act = np.ascontiguousarray(act, np.float32)
    out = np.empty_like(ws)
    th = np.empty((n, 1), np.float32)
    mo = np.empty((n, 3), np.float32)
    d = None if dist is None else np.ascontiguousarray(dist, np.float32)
    lib().qo_euler(C.byref(p), _f(ws), _f(act), _f(d), n, _f(out), _f(th), _f(mo))
    return out, th, mo


# ------------------------------------------------------------------------------------------------ policy (row f1)
def policy_forward(weights, biases, obs, bf16=False, activation="relu"):
    """The controller MLP on the CPU: the float32 restatement of `c_code/neural_network.c:397-430`, or (``bf16``)
    the same network with the B200 tensor-core path's operand rounding."""
    L = lib()
    obs = np.ascontiguousarray(obs, np.float32)
    n, nl = len(obs), len(weights)
    dims = (C.c_int * (nl + 1))(weights[0].shape[1], *[w.shape[0] for w in weights])
    ws = [np.ascontiguousarray(w, np.float32) for w in weights]
    bs = [np.ascontiguousarray(b, np.float32) for b in biases]
    wp = (_fp * nl)(*[_f(w) for w in ws])
    bp = (_fp * nl)(*[_f(b) for b in bs])
    out = np.empty((n, weights[-1].shape[0]), np.float32)
    L.qo_set_policy_activation.argtypes = [C.c_int]
    L.qo_set_policy_activation(1 if activation == "tanh" else 0)
    fn = L.qo_policy_forward_bf16 if bf16 else L.qo_policy_forward
    fn.argtypes = [C.c_int, C.POINTER(C.c_int), C.POINTER(_fp), C.POINTER(_fp), _fp, C.c_int64, _fp]
    fn.restype = None
    fn(nl, dims, wp, bp, _f(obs), n, _f(out))
    L.qo_set_policy_activation(0)
    return out


def ref_policy_lib():
    """The reference's own generated C policy (c_code/neural_network.c + nn_controller.c) compiled into oracle/_ref."""
    path = os.path.join(HERE, "_ref", "libnn_policy_ref.so")
    if not os.path.isfile(path):
        return None
    L = C.CDLL(path)
    L.nn_forward.argtypes = [_fp, _fp]
    return L

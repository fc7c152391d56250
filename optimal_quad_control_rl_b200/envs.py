"""Host-side mirror of the reference's ``Quadcopter3DGates(VecEnv)`` (`3D quad race.ipynb:287-620`, INDI
`3D quad race INDI inner loop.ipynb:142-410`) over the CUDA library: same constructor, same methods, same public
attributes, same reward, so the reference's SB3/PPO cells run against it unchanged.

Everything numeric happens in ``libquadsim.so``; this file only moves arguments across the C ABI, keeps the
NumPy-facing attributes the notebooks poke, and (in ``reset_rng="numpy"`` mode) replays the reference's
``np.random`` draw order so that seeded runs reset to bit-identical states.  There is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import itertools
import os
from collections.abc import Sequence

import numpy as np
import torch

from . import _lib as L

try:  # the real base class when SB3 is installed (it is not in the build image)
    from stable_baselines3.common.vec_env import VecEnv as _VecEnvBase  # type: ignore
except Exception:  # pragma: no cover - exercised in this image
    class _VecEnvBase:  # the part of SB3's VecEnv the reference relies on
        def __init__(self, num_envs, observation_space, action_space):
            self.num_envs = num_envs
            self.observation_space = observation_space
            self.action_space = action_space

        def step(self, actions):
            self.step_async(actions)
            return self.step_wait()

try:
    from gymnasium import spaces as _spaces  # type: ignore
except Exception:  # pragma: no cover
    class _Box:
        def __init__(self, low, high, shape=None, dtype=np.float32):
            self.low = np.broadcast_to(np.asarray(low, dtype=np.float64), shape or np.shape(low)).astype(dtype)
            self.high = np.broadcast_to(np.asarray(high, dtype=np.float64), shape or np.shape(high)).astype(dtype)
            self.shape = tuple(shape) if shape is not None else self.low.shape
            self.dtype = np.dtype(dtype)

        def sample(self):
            lo = np.where(np.isfinite(self.low), self.low, -1.0)
            hi = np.where(np.isfinite(self.high), self.high, 1.0)
            return np.random.uniform(lo, hi).astype(self.dtype)

    class _spaces:  # noqa: N801
        Box = _Box

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


def load_residual_weights(path=None):
    """Residual thrust / moment MLP parameters (a packaged copy of the reference's ``NNDroneModel/*.pt``,
    `3D quad race.ipynb:228-229`), flattened in the order the C ABI expects."""
    z = np.load(path or os.path.join(_DATA, "residual_mlp.npz"))
    flat = lambda n: np.ascontiguousarray(np.concatenate(
        [z[f"{n}_w1"].ravel(), z[f"{n}_b1"].ravel(), z[f"{n}_w2"].ravel(), z[f"{n}_b2"].ravel()]), np.float32)
    t, m = flat("thrust"), flat("moment")
    assert t.size == 289 and m.size == 451
    return t, m


def _f(a):
    return a.ctypes.data_as(L._fp) if a is not None else None


class AliasedInfos(Sequence):
    """``infos`` of a step: the reference builds ``[{}] * num_envs`` -- ONE dict referenced ``num_envs`` times
    (`3D quad race.ipynb:589-594`).  This is that list without the N pointers (building them costs ~2 ms per step at
    N = 2**20, as much as the whole device step + PCIe transfer): ``len``, indexing, iteration, ``in``, slicing and
    ``list(infos)`` behave like the reference's list; every index returns the same dict."""
    __slots__ = ("_info", "_n")

    def __init__(self, info, n):
        self._info, self._n = info, int(n)

    def __len__(self):
        return self._n

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._info] * len(range(*i.indices(self._n)))
        i = int(i)
        if not -self._n <= i < self._n:
            raise IndexError("list index out of range")
        return self._info

    def __iter__(self):
        return itertools.repeat(self._info, self._n)

    def __eq__(self, other):
        return len(other) == self._n and all(o == self._info for o in other)

    def __repr__(self):
        return f"[{self._info!r}] * {self._n}"


class _QuadGatesBase(_VecEnvBase):
    """Shared implementation; the two public classes below only fix the model variant."""

    _VARIANT = None  # "e2e" | "indi"

    def __init__(self, num_envs, gates_pos, gate_yaw, start_pos, gates_ahead=0, pause_if_collision=False, *,
                 device=None, reset_rng="numpy", seed=0, env_offset=0, obs_buffers=2, residual_weights=None):
        if reset_rng not in ("numpy", "device"):
            raise ValueError("reset_rng must be 'numpy' (reference-exact draws) or 'device' (fused Philox)")
        if not torch.cuda.is_available():
            raise L.QuadsimError("no CUDA device: this environment only runs on the GPU (there is no CPU fallback)")
        self._lib = L.load()
        self._vid = L.E2E if self._VARIANT == "e2e" else L.INDI
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise L.QuadsimError(f"device {self.device} is not a CUDA device (there is no CPU fallback)")
        if self.device.index is None:  # "cuda" = the CURRENT device, not ordinal 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.reset_rng = reset_rng
        self._obs_format = "f32"
        self.lazy_infos_from = 1 << 16  # step_wait returns AliasedInfos instead of a list of N references from this size on

        # -- race track (`:298-302`)
        self.start_pos = np.asarray(start_pos).astype(np.float32)
        self.gate_pos = np.ascontiguousarray(np.asarray(gates_pos).astype(np.float32))
        self.gate_yaw = np.ascontiguousarray(np.asarray(gate_yaw).astype(np.float32))
        self.num_gates = self.gate_pos.shape[0]
        self.gates_ahead = int(gates_ahead)
        self.pause_if_collision = pause_if_collision

        # -- gate i in the frame of gate i-1, looped track (`:309-319`); NumPy float32 like the reference
        ng = self.num_gates
        self.gate_pos_rel = np.zeros((ng, 3), dtype=np.float32)
        self.gate_yaw_rel = np.zeros(ng, dtype=np.float32)
        for i in range(ng):
            delta = self.gate_pos[i] - self.gate_pos[i - 1]
            c, s = np.cos(self.gate_yaw[i - 1]), np.sin(self.gate_yaw[i - 1])
            self.gate_pos_rel[i, 0:2] = np.array([[c, s], [-s, c]]) @ delta[0:2]
            self.gate_pos_rel[i, 2] = delta[2]
            self.gate_yaw_rel[i] = self.gate_yaw[i] - self.gate_yaw[i - 1]

        self._ns = self._lib.qs_state_len(self._vid)
        self.state_len = self._lib.qs_obs_len(self._vid, self.gates_ahead)
        action_space = _spaces.Box(low=-1, high=1, shape=(4,))
        observation_space = _spaces.Box(low=np.array([-np.inf] * self.state_len),
                                        high=np.array([np.inf] * self.state_len))
        _VecEnvBase.__init__(self, num_envs, observation_space, action_space)
        n = self.num_envs = int(num_envs)

        # -- device side
        h = L._vp()
        st = self._lib.qs_create(C.byref(h), self._vid, n, ng, _f(self.gate_pos), _f(self.gate_yaw),
                                 _f(self.start_pos), self.gates_ahead, self.device.index,
                                 L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        L.check(self._lib, None, st, "qs_create")
        self._h = h
        self._call("qs_set_track_tables", _f(np.cos(self.gate_yaw)), _f(np.sin(self.gate_yaw)),
                   _f(np.ascontiguousarray(self.gate_pos_rel)), _f(self.gate_yaw_rel))
        self._call("qs_seed", int(seed))
        self._call("qs_set_env_offset", int(env_offset))
        if self._VARIANT == "e2e":
            t, m = residual_weights if residual_weights is not None else load_residual_weights()
            self._call("qs_set_residual_weights", _f(np.ascontiguousarray(t, np.float32)),
                       _f(np.ascontiguousarray(m, np.float32)))
        dev = self.device
        self._act_dev = torch.zeros((n, 4), dtype=torch.float32, device=dev)
        self._obs_ring = [torch.zeros((n, self.state_len), dtype=torch.float32, device=dev)
                          for _ in range(max(2, int(obs_buffers)))]
        self._rew_ring = [torch.zeros(n, dtype=torch.float32, device=dev) for _ in self._obs_ring]
        self._done_ring = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in self._obs_ring]
        self._flags_ring = [torch.zeros(n, dtype=torch.uint8, device=dev) for _ in self._obs_ring]
        self._ring = 0

        # -- the attributes the notebooks read and poke (`:340-360`)
        self.states = np.zeros((n, self.state_len), dtype=np.float32)
        self.max_steps = 1200
        self.dt = np.float32(0.01)
        self.actions = np.zeros((n, 4), dtype=np.float32)
        self.dones = np.zeros(n, dtype=bool)
        self.final_gate_passed = np.zeros(n, dtype=bool)
        self.update_states = self.update_states_gate
        self.disturbance_ranges = np.zeros((6, 2), dtype=np.float32)
        self.disturbance_scale = 1
        self.pause = False
        self.last_flags = np.zeros(n, dtype=np.uint8)
        self._pushed = None
        self._host_ring = None     # pinned NumPy-facing output buffers of the host fast path (lazy)
        self._host_ptrs = []
        self._ring_stale = False   # the device obs ring lags the state after a host-path step

    # ------------------------------------------------------------------------------------------ plumbing
    def _call(self, name, *args):
        L.check(self._lib, self._h, getattr(self._lib, name)(self._h, *args), name)

    def _push_config(self):
        """Forward attribute pokes (max_steps, dt, disturbance_ranges/scale) made since the last call."""
        dr = np.asarray(self.disturbance_ranges)
        key = (int(self.max_steps), float(self.dt), dr.dtype.str, dr.tobytes(), float(self.disturbance_scale))
        if key == self._pushed:
            return
        self._call("qs_set_max_steps", int(self.max_steps))
        self._call("qs_set_dt", float(self.dt))
        if self._VARIANT == "e2e":
            r64 = np.ascontiguousarray(dr, dtype=np.float64).reshape(12)
            self._call("qs_set_disturbance_ranges", r64.ctypes.data_as(L._dp), int(dr.dtype == np.float64),
                       float(self.disturbance_scale))
        self._pushed = key

    def _sync_stream(self):
        self._call("qs_set_stream", L._vp(torch.cuda.current_stream(self.device).cuda_stream))

    def _mode(self):
        if self.pause:
            return L.MODE_PAUSE
        return L.MODE_PAUSE_IF_COLLISION if self.pause_if_collision else L.MODE_NORMAL

    def _next_slot(self):
        self._ring = (self._ring + 1) % len(self._obs_ring)
        return self._ring

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.qs_destroy(h)
        ptrs, self._host_ptrs, self._host_ring = getattr(self, "_host_ptrs", []), [], None
        for p in ptrs:  # arrays handed out earlier must not be used after close()
            self._lib.qs_host_free(p)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------------------------------ state attributes
    def _get(self, what):
        n = self.num_envs
        ws = np.empty((n, self._ns), np.float32) if what == "ws" else None
        d = np.empty((n, 6), np.float32) if what == "dist" else None
        tg = np.empty(n, np.int64) if what == "tg" else None
        sc = np.empty(n, np.int64) if what == "sc" else None
        ip = lambda a: a.ctypes.data_as(L._i64p) if a is not None else None
        self._sync_stream()
        self._call("qs_get_state", 0, n, _f(ws), _f(d), ip(tg), ip(sc))
        return next(a for a in (ws, d, tg, sc) if a is not None)

    def _get_rows(self, first, count):
        """world_states and step_counts of envs ``first .. first+count-1`` only (what a viewer or logger needs:
        a few hundred bytes device-to-host instead of the whole state)."""
        if first < 0 or count < 0 or first + count > self.num_envs:
            raise IndexError("env range out of bounds")
        ws, sc = np.empty((count, self._ns), np.float32), np.empty(count, np.int64)
        self._sync_stream()
        self._call("qs_get_state", int(first), int(count), _f(ws), None, None, sc.ctypes.data_as(L._i64p))
        return ws, sc

    def _set(self, ws=None, dist=None, tg=None, sc=None):
        n = self.num_envs
        ws = None if ws is None else np.ascontiguousarray(ws, np.float32).reshape(n, self._ns)
        dist = None if dist is None else np.ascontiguousarray(dist, np.float32).reshape(n, 6)
        tg = None if tg is None else np.ascontiguousarray(tg, np.int64).reshape(n)
        sc = None if sc is None else np.ascontiguousarray(sc, np.int64).reshape(n)
        ip = lambda a: a.ctypes.data_as(L._i64p) if a is not None else None
        self._sync_stream()
        self._call("qs_set_state", 0, n, _f(ws), _f(dist), ip(tg), ip(sc))

    # Reads copy device -> host (array-of-structs like the reference); assignment copies host -> device.
    # In-place edits of the returned copy do not reach the simulator -- assign the whole array instead.
    world_states = property(lambda self: self._get("ws"), lambda self, v: self._set(ws=v))
    target_gates = property(lambda self: self._get("tg"), lambda self, v: self._set(tg=v))
    step_counts = property(lambda self: self._get("sc"), lambda self, v: self._set(sc=v))

    def _get_dist(self):
        if self._VARIANT != "e2e":
            raise AttributeError("disturbances")
        return self._get("dist")

    disturbances = property(_get_dist, lambda self, v: self._set(dist=v))

    # ------------------------------------------------------------------------------------------ observation
    def update_states_gate(self):
        """update_states_gate (`:365-450`): recompute every observation into a fresh array."""
        self._push_config()
        self._sync_stream()
        k = self._next_slot()
        self._call("qs_observe", L._vp(self._obs_ring[k].data_ptr()))
        self._ring_stale = False
        self.states = self._obs_ring[k].cpu().numpy()
        return self.states

    # ------------------------------------------------------------------------------------------ reset
    def _draw_reset(self, n):
        """reset_'s draws from the GLOBAL NumPy stream, field-major, float64 -> float32 on store (`:455-489`)."""
        u = np.random.uniform
        cols = [u(-0.5, 0.5, size=(n,)) + self.start_pos[k] for k in range(3)]
        cols += [u(-0.5, 0.5, size=(n,)) for _ in range(3)]
        cols += [u(-np.pi / 9, np.pi / 9, size=(n,)), u(-np.pi / 9, np.pi / 9, size=(n,)),
                 u(-np.pi, np.pi, size=(n,))]
        cols += [u(-0.1, 0.1, size=(n,)) for _ in range(3)]
        cols += [u(-1, 1, size=(n,)) for _ in range(4)] if self._VARIANT == "e2e" else [u(-.1, .1, size=(n,))]
        ws = np.ascontiguousarray(np.stack(cols, axis=1), dtype=np.float32)
        dist = None
        if self._VARIANT == "e2e":
            r = np.asarray(self.disturbance_ranges)
            dist = np.ascontiguousarray(
                self.disturbance_scale * np.stack([u(r[k, 0], r[k, 1], size=(n,)) for k in range(6)], axis=1),
                dtype=np.float32)
        return ws, dist

    def _apply_host_reset(self, mask, obs_dev):
        idx = np.flatnonzero(mask).astype(np.int32)
        ws, dist = self._draw_reset(idx.size)  # n == 0 consumes no RNG state, like the reference
        if idx.size:
            self._call("qs_apply_reset", idx.size, idx.ctypes.data_as(L._i32p), _f(ws), _f(dist),
                       L._vp(obs_dev.data_ptr()))

    def reset_(self, dones):
        """reset_(mask) (`:452-493`): redraw the masked envs from np.random, refresh all observations."""
        self._push_config()
        self._sync_stream()
        dones = np.asarray(dones, dtype=bool)
        k = self._next_slot()
        self._call("qs_observe", L._vp(self._obs_ring[k].data_ptr()))
        self._ring_stale = False
        self._apply_host_reset(dones, self._obs_ring[k])
        self.states = self._obs_ring[k].cpu().numpy()
        return self.states

    def reset(self):
        if self.reset_rng == "device":
            self.reset_tensor()
            self.states = self._obs_ring[self._ring].cpu().numpy()
            return self.states
        return self.reset_(np.ones(self.num_envs, dtype=bool))

    # ------------------------------------------------------------------------------------------ step
    def step_async(self, actions):
        self.actions = actions

    def _pinned(self, shape, dtype):
        """A NumPy array over page-locked host memory (qs_host_alloc): device copies into it are asynchronous DMA."""
        nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = self._lib.qs_host_alloc(max(nbytes, 1))
        if not p:
            raise L.QuadsimError("qs_host_alloc failed")
        self._host_ptrs.append(p)
        return np.frombuffer((C.c_char * max(nbytes, 1)).from_address(p), dtype=dtype, count=int(np.prod(shape))).reshape(shape)

    def _step_wait_host(self, mode):
        """NumPy-facing fast path (``reset_rng="device"``): ONE ``qs_step_host_ex`` call -- chunk-pipelined staging + H2D
        of the caller's action array (float32 or float64, as given), step kernel, D2H of obs / reward / done / flags --
        into a ring of pinned NumPy arrays; the scan for the ``infos`` quirk happens behind the C ABI as well.  The
        returned arrays are recycled after ``obs_buffers`` further steps (SB3 copies them into its rollout buffer
        within one)."""
        n = self.num_envs
        if self._host_ring is None:
            self._host_ring = [{"obs": self._pinned((n, self.state_len), np.float32), "rew": self._pinned((n,), np.float32),
                                "done": self._pinned((n,), np.uint8), "flags": self._pinned((n,), np.uint8)}
                               for _ in self._obs_ring]
        k = self._next_slot()
        h = self._host_ring[k]
        act = self.actions
        if not (isinstance(act, np.ndarray) and act.dtype in (np.float32, np.float64) and act.flags.c_contiguous
                and act.size == 4 * n):
            act = np.ascontiguousarray(act, dtype=np.float32).reshape(n, 4)
        vp = lambda a: L._vp(a.ctypes.data)
        info = L.QsStepInfo()
        self._call("qs_step_host_ex", vp(act), L.F64 if act.dtype == np.float64 else L.F32, vp(h["obs"]), vp(h["rew"]),
                   vp(h["done"]), vp(h["flags"]), mode, L.RESET_DEVICE, C.byref(info))
        self._ring_stale = True
        return h["obs"], h["rew"], h["done"].view(np.bool_), h["flags"], info

    def current_obs_tensor(self):
        """The observations of the current state as a CUDA tensor (N, D): what ``step_tensor`` / ``rollout`` last wrote,
        recomputed from the state if NumPy-path steps ran in between."""
        if self._obs_format != "f32":
            raise L.QuadsimError("current_obs_tensor: the env's observation format is packed BF16; set obs_format = 'f32' first")
        if self._ring_stale:
            self._push_config()
            self._sync_stream()
            self._call("qs_observe", L._vp(self._obs_ring[self._ring].data_ptr()))
            self._ring_stale = False
        return self._obs_ring[self._ring]

    def _infos(self, info):
        """A real list for the sizes the reference runs at; from ``lazy_infos_from`` envs on (default 2**16) the same
        aliased sequence without materialising N pointers."""
        n = self.num_envs
        return [info] * n if n < self.lazy_infos_from else AliasedInfos(info, n)

    def step_wait(self):
        """step_wait (`:501-595`), NumPy in / NumPy out.  Returns fresh arrays every call."""
        n = self.num_envs
        self._push_config()
        self._sync_stream()
        if self.reset_rng == "device":
            mode = self._mode()
            obs, rewards, dones, flags, si = self._step_wait_host(mode)
            if mode != L.MODE_PAUSE:
                self.states = obs
            self.dones, self.last_flags = dones, flags
            info = {}  # ONE dict aliased N times, like the reference's `[{}] * num_envs` (`:589-594`)
            if si.last_done_index >= 0:
                info["terminal_observation"] = self.states[si.last_done_index]
            if si.any_truncated:
                info["TimeLimit.truncated"] = True
            return self.states, rewards, dones, self._infos(info)
        act = np.ascontiguousarray(self.actions, dtype=np.float32).reshape(n, 4)
        self._act_dev.copy_(torch.from_numpy(act))
        mode = self._mode()
        host_reset = self.reset_rng == "numpy"
        k = self._next_slot()
        obs_d, rew_d, done_d, fl_d = self._obs_ring[k], self._rew_ring[k], self._done_ring[k], self._flags_ring[k]
        self._call("qs_step", L._vp(self._act_dev.data_ptr()), L._vp(obs_d.data_ptr()), L._vp(rew_d.data_ptr()),
                   L._vp(done_d.data_ptr()), L._vp(fl_d.data_ptr()), mode,
                   L.RESET_HOST if host_reset else L.RESET_DEVICE)
        dones = done_d.cpu().numpy().astype(bool)
        flags = fl_d.cpu().numpy()
        if mode == L.MODE_NORMAL and host_reset:
            self._apply_host_reset(dones, obs_d)
        rewards = rew_d.cpu().numpy()
        if mode != L.MODE_PAUSE:  # env.pause leaves self.states as it was (`:570-572`)
            self.states = obs_d.cpu().numpy()
        self.dones = dones
        self.last_flags = flags
        # `infos = [{}] * N` aliases ONE dict (`:589-594`): after the loop it carries the observation row of the
        # highest-index done env (already reset) and the truncation flag if ANY env timed out.  Kept on purpose.
        info = {}
        idx = np.flatnonzero(dones)
        if idx.size:
            info["terminal_observation"] = self.states[idx[-1]]
        if (flags & L.F_TRUNCATED).any():
            info["TimeLimit.truncated"] = True
        return self.states, rewards, dones, self._infos(info)

    # ------------------------------------------------------------------------------------------ tensor fast path
    def reset_tensor(self, obs_out=None):
        """reset() with the fused device RNG; returns the observation tensor (N,D) on the GPU (or fills ``obs_out``)."""
        self._push_config()
        self._sync_stream()
        if obs_out is not None or self._obs_format != "f32":
            self._check_obs_out(obs_out)
            self._call("qs_reset_all", L._vp(obs_out.data_ptr()))
            self._ring_stale = True
            return obs_out
        k = self._next_slot()
        self._call("qs_reset_all", L._vp(self._obs_ring[k].data_ptr()))
        self._ring_stale = False
        return self._obs_ring[k]

    def _check_obs_out(self, obs_out):
        if self._obs_format == "f32":
            if obs_out.dtype != torch.float32 or not obs_out.is_cuda or not obs_out.is_contiguous() \
                    or obs_out.numel() < self.num_envs * self.state_len:
                raise ValueError("obs_out must be a contiguous float32 CUDA tensor of (num_envs, obs_len)")
        elif obs_out is None or obs_out.dtype != torch.uint8 or not obs_out.is_cuda or not obs_out.is_contiguous() \
                or obs_out.numel() < self.packed_obs_bytes():
            raise ValueError(f"obs_format 'bf16_k32': obs_out must be a contiguous uint8 CUDA tensor of >= {self.packed_obs_bytes()} bytes")

    def step_tensor(self, actions, obs_out=None):
        """Zero-copy step: ``actions`` is a float32 CUDA tensor (N,4); returns (obs, reward, done, flags) CUDA
        tensors.  Asynchronous on the current stream.  Output tensors are recycled every ``obs_buffers`` calls.
        ``obs_out`` lets a sharded job have the kernel write straight into its slice of an all-gather buffer."""
        if actions.dtype != torch.float32 or not actions.is_cuda or not actions.is_contiguous():
            raise ValueError("step_tensor needs a contiguous float32 CUDA tensor of shape (num_envs, 4)")
        self._push_config()
        self._sync_stream()
        if obs_out is not None or self._obs_format != "f32":
            self._check_obs_out(obs_out)
        k = self._next_slot()
        obs_d = self._obs_ring[k] if obs_out is None else obs_out
        self._call("qs_step", L._vp(actions.data_ptr()), L._vp(obs_d.data_ptr()),
                   L._vp(self._rew_ring[k].data_ptr()), L._vp(self._done_ring[k].data_ptr()),
                   L._vp(self._flags_ring[k].data_ptr()), self._mode(), L.RESET_DEVICE)
        self._ring_stale = obs_out is not None  # the ring slot itself was not written
        return obs_d, self._rew_ring[k], self._done_ring[k], self._flags_ring[k]

    def step_graph(self, actions, out=None):
        """Open-loop stepping as ONE CUDA graph: ``actions`` is a float32 CUDA tensor ``(T, N, 4)`` (or a list of T
        ``(N, 4)`` tensors); returns ``(graph, out)`` where ``graph.replay()`` runs the T steps back to back and ``out``
        holds ``obs (T, N, D)``, ``rewards (T, N)``, ``dones (T, N)`` and ``flags (T, N)`` (uint8) of every step.  Steps that
        follow each other inside a graph are chained CTA by CTA by the library (``chained_launch_count``), so the graph
        runs 4 - 15 % faster than T separate launches; refill ``actions`` in place between replays.  The random-agent
        and action-replay loops of the reference (`3D quad race.ipynb:672-673`) in one call."""
        if self._obs_format != "f32":
            raise L.QuadsimError("step_graph: the env's observation format is packed BF16; set obs_format = 'f32' first")
        if isinstance(actions, (list, tuple)):
            acts = list(actions)
        else:
            acts = [actions[t] for t in range(actions.shape[0])]
        T, n, d, dev = len(acts), self.num_envs, self.state_len, self.device
        if out is None:
            out = {"obs": torch.empty((T, n, d), dtype=torch.float32, device=dev),
                   "rewards": torch.empty((T, n), dtype=torch.float32, device=dev),
                   "dones": torch.empty((T, n), dtype=torch.uint8, device=dev),
                   "flags": torch.empty((T, n), dtype=torch.uint8, device=dev)}
        for a in acts:
            if a.dtype != torch.float32 or not a.is_cuda or not a.is_contiguous() or tuple(a.shape) != (n, 4):
                raise ValueError("step_graph needs contiguous float32 CUDA action tensors of shape (num_envs, 4)")
        self._push_config()
        self._sync_stream()
        # anything lazy (track-table upload) happens now, outside the capture: recompute the current observation rows
        self._call("qs_observe", L._vp(self._obs_ring[self._ring].data_ptr()))
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            self._sync_stream()
            for t, a in enumerate(acts):
                self._call("qs_step", L._vp(a.data_ptr()), L._vp(out["obs"][t].data_ptr()), L._vp(out["rewards"][t].data_ptr()),
                           L._vp(out["dones"][t].data_ptr()), L._vp(out["flags"][t].data_ptr()), self._mode(), L.RESET_DEVICE)
        self._ring_stale = True
        return graph, out

    def rollout(self, policy, steps, deterministic=False, buffers=None, fused=None):
        """collect_rollouts on the device (SB3's loop behind `model.learn`, `3D quad race.ipynb:820`): ``steps`` x
        (policy forward -> env step) enqueued back to back, no host round trip.  Returns CUDA tensors
        ``obs (steps+1, N, D)``, ``actions (steps, N, 4)`` (clipped), ``raw_actions (steps, N, 4)`` (un-clipped samples),
        ``rewards (steps, N)``, ``dones (steps, N)`` (uint8); a caller-supplied ``buffers["flags"]`` (steps, N) uint8
        also receives every step's ``F_*`` bits (``F_TRUNCATED`` = the reference's ``TimeLimit.truncated``);
        ``obs[0]`` is the observation the rollout started from, ``obs[steps]`` the one the next rollout starts from.
        ``fused``: True = ONE launch of the closed-loop kernel (quads stay in registers for all ``steps``; same results
        bit for bit), False = 2*steps launches, None = fused whenever the shapes fit."""
        n, d, dev = self.num_envs, self.state_len, self.device
        self._push_config()
        self._sync_stream()
        if buffers is None:
            buffers = {"obs": torch.empty((steps + 1, n, d), dtype=torch.float32, device=dev),
                       "actions": torch.empty((steps, n, 4), dtype=torch.float32, device=dev),
                       "raw_actions": torch.empty((steps, n, 4), dtype=torch.float32, device=dev),
                       "rewards": torch.empty((steps, n), dtype=torch.float32, device=dev),
                       "dones": torch.empty((steps, n), dtype=torch.uint8, device=dev)}
            buffers["obs"][0].copy_(self.current_obs_tensor())
        if fused is None:
            fused = bool(self._lib.qs_rollout_fused_supported(self._h, policy._h))
        self._call("qs_rollout_fused" if fused else "qs_rollout", policy._h, int(steps), L._vp(buffers["obs"].data_ptr()),
                   L._vp(buffers["actions"].data_ptr()),
                   L._vp(buffers["raw_actions"].data_ptr()) if "raw_actions" in buffers else None,
                   L._vp(buffers["rewards"].data_ptr()), L._vp(buffers["dones"].data_ptr()),
                   L._vp(buffers["flags"].data_ptr()) if "flags" in buffers else None, int(bool(deterministic)))
        self._obs_ring[self._ring].copy_(buffers["obs"][steps])
        self._ring_stale = False
        return buffers

    def set_obs_peers(self, peer_ptrs, row_offset):
        """Fused observation all-gather: also store every step's observations into the other ranks' gather buffers
        (device pointers of peer memory) at row ``row_offset + env``.  ``peer_ptrs=[]`` switches it off."""
        arr = (L._vp * max(1, len(peer_ptrs)))(*[L._vp(int(p)) for p in peer_ptrs])
        self._call("qs_set_obs_peers", len(peer_ptrs), arr, int(row_offset))

    # ---- observation format of the tensor path (`step_tensor` / `reset_tensor` with ``obs_out=``)
    @property
    def obs_format(self):
        return self._obs_format

    @obs_format.setter
    def obs_format(self, fmt):
        """``"f32"``: float32 rows (N, D), the reference's layout.  ``"bf16_k32"``: packed BF16 blocks in the on-device policy's
        operand layout (`qs_set_obs_format`; 48 B per env for the 24-wide E2E row) for a sharded job whose gathered observations only feed
        ``MlpPolicy.forward_packed``.  In that format ``step_tensor`` / ``reset_tensor`` need ``obs_out=`` (a uint8 CUDA
        tensor of ``packed_obs_bytes()`` bytes) and the NumPy-facing ``step`` / ``reset`` / ``rollout`` are unavailable."""
        if fmt not in ("f32", "bf16_k32"):
            raise ValueError("obs_format must be 'f32' or 'bf16_k32'")
        self._call("qs_set_obs_format", 1 if fmt == "bf16_k32" else 0)
        self._obs_format = fmt

    def packed_obs_bytes(self, n=None):
        return int(self._lib.qs_obs_packed_bytes(int(self.state_len), int(self.num_envs if n is None else n)))

    def enable_stats(self, on=True):
        self._call("qs_enable_stats", int(on))

    def stats(self, reset=False):
        s = L.QsStats()
        self._sync_stream()
        self._call("qs_get_stats", C.byref(s), int(reset))
        return {k: getattr(s, k) for k, _ in L.QsStats._fields_}

    @property
    def launch_count(self):
        return int(self._lib.qs_launch_count(self._h))

    @property
    def chained_launch_count(self):
        """Step launches that depended on their predecessor CTA by CTA (consecutive steps inside one CUDA graph)."""
        return int(self._lib.qs_chained_launch_count(self._h))

    # ------------------------------------------------------------------------------------------ VecEnv plumbing (`:597-620`)
    def seed(self, seed=None):
        pass  # the reference's seed() is a no-op: seeding is np.random.seed by the caller (`:600-601`)

    def set_attr(self, attr_name, value, indices=None):
        pass

    def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
        pass

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False] * self.num_envs

    _RENDER_KEYS = ()

    def render(self, mode="human"):
        """Dict of state columns + actions rescaled to [0,1] for quadcopter_animation (`:615-620`)."""
        state_dict = dict(zip(self._RENDER_KEYS, self.world_states.T))
        action_dict = dict(zip(["u1", "u2", "u3", "u4"], (np.array(np.asarray(self.actions).T) + 1) / 2))
        return {**state_dict, **action_dict}


class Quadcopter3DGates(_QuadGatesBase):
    """End-to-end Bebop env: rotor-speed commands, residual thrust/moment MLPs, disturbances
    (`3D quad race.ipynb:287`).  Observation width 20 + 4*gates_ahead."""
    _VARIANT = "e2e"
    _RENDER_KEYS = ("x", "y", "z", "vx", "vy", "vz", "phi", "theta", "psi", "p", "q", "r", "w1", "w2", "w3", "w4")

    def get_attr(self, attr_name, indices=None):
        raise AttributeError()  # required: SB3 2.x probes render_mode through this (`:603-604`)


class Quadcopter3DGatesINDI(_QuadGatesBase):
    """INDI inner-loop env: thrust + body-rate commands through first-order lags
    (`3D quad race INDI inner loop.ipynb:142`).  Observation width 13 + 4*gates_ahead."""
    _VARIANT = "indi"
    _RENDER_KEYS = ("x", "y", "z", "vx", "vy", "vz", "phi", "theta", "psi", "p", "q", "r", "T")

    def get_attr(self, attr_name, indices=None):
        pass  # the INDI notebook returns None (`:393-394`)

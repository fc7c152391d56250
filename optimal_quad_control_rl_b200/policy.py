"""On-device controller policy: the MLP the reference trains with SB3
(`PPO("MlpPolicy", net_arch=[120,120,120], ReLU)`, `3D quad race.ipynb:784-795`), evaluates with
`model.predict(env.states)` (`:803`) and ships as generated C (`c_code/neural_network.c`, noise + clip in
`c_code/nn_controller.c:158-176`) -- evaluated on the B200's tensor cores right next to the simulator, so
observations and actions never leave the GPU.  Everything numeric happens in ``libquadsim.so``."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from . import _lib as L

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")


class MlpPolicy:
    """``weights[l]`` is ``(out, in)`` float32 (torch / generated-C layout), ``biases[l]`` is ``(out,)``; the last
    pair is the output layer.  ``std`` = exp(log_std) of the Gaussian action distribution."""

    def __init__(self, weights, biases, std=None, device=None, seed=0, env_offset=0, activation="relu", obs_limit=0.0):
        if not torch.cuda.is_available():
            raise L.QuadsimError("no CUDA device: the policy only runs on the GPU (there is no CPU fallback)")
        self._lib = L.load()
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.device.type != "cuda":
            raise L.QuadsimError(f"device {self.device} is not a CUDA device (there is no CPU fallback)")
        if self.device.index is None:  # "cuda" = the CURRENT device, not ordinal 0
            self.device = torch.device("cuda", torch.cuda.current_device())
        self.weights = [np.ascontiguousarray(w, np.float32) for w in weights]
        self.biases = [np.ascontiguousarray(b, np.float32) for b in biases]
        nl = len(self.weights)
        self.in_dim = self.weights[0].shape[1]
        self.hidden = self.weights[0].shape[0]
        self.out_dim = self.weights[-1].shape[0]
        for l, (w, b) in enumerate(zip(self.weights, self.biases)):
            want = (self.out_dim if l == nl - 1 else self.hidden, self.in_dim if l == 0 else self.hidden)
            if w.shape != want or b.shape != (want[0],):
                raise ValueError(f"layer {l}: expected W{want} b({want[0]},), got W{w.shape} b{b.shape}")
        self.std = np.zeros(4, np.float32)
        if std is not None:
            self.std[:self.out_dim] = np.asarray(std, np.float32)
        h = L._vp()
        st = self._lib.qs_policy_create(C.byref(h), self.in_dim, nl - 1, self.hidden, self.out_dim,
                                        self.device.index, L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        if st != 0:
            raise L.QuadsimError(f"qs_policy_create failed ({st}): {self._lib.qs_policy_last_error(None).decode()}")
        self._h = h
        for l, (w, b) in enumerate(zip(self.weights, self.biases)):
            self._call("qs_policy_set_layer", l, w.ctypes.data_as(L._fp), b.ctypes.data_as(L._fp))
        self._call("qs_policy_set_std", self.std.ctypes.data_as(L._fp))
        if activation not in ("relu", "tanh"):
            raise ValueError("activation must be 'relu' or 'tanh'")
        self.activation = activation
        self._call("qs_policy_set_activation", 1 if activation == "tanh" else 0)
        if obs_limit:  # sanitise the inputs like the PPO learner (NaN -> 0, clamp): forwards over rollout buffers
            self._call("qs_policy_set_obs_limit", float(obs_limit))
        self._call("qs_policy_seed", int(seed))
        self._call("qs_policy_set_env_offset", int(env_offset))

    @classmethod
    def from_npz(cls, path=None, **kw):
        """Default: the controller this repository trained on the E2E zigzag env with its own PPO (130 s on one B200,
        profiles/r2/ppo; 16-18 gates per 12 s episode without a crash, also when exported to C and flown against the
        CPU oracle env, tests/test_trained_policy_cpu.py)."""
        z = np.load(path or os.path.join(_DATA, "policy_e2e_zigzag_ppo.npz"))
        n = len(z["dims"]) - 1
        return cls([z[f"W{l}"] for l in range(n)], [z[f"b{l}"] for l in range(n)], std=z["std"], **kw)

    @classmethod
    def reference_controller(cls, **kw):
        """The reference's shipped controller (24 -> 120 -> 120 -> 120 -> 4, `c_code/neural_network.c`, std
        `c_code/nn_controller.c:7-12`).  It was trained on the rectangle track with another env build and passes hardly
        any gate here: a fixture for parity with the reference's own C network, not a pilot."""
        return cls.from_npz(os.path.join(_DATA, "policy_k4.npz"), **kw)

    @classmethod
    def from_sb3(cls, model, **kw):
        """From a trained SB3 ``PPO("MlpPolicy")`` model: ``mlp_extractor.policy_net`` + ``action_net`` + ``log_std``."""
        pol = model.policy
        lin = [m for m in pol.mlp_extractor.policy_net if hasattr(m, "weight")] + [pol.action_net]
        w = [m.weight.detach().cpu().numpy() for m in lin]
        b = [m.bias.detach().cpu().numpy() for m in lin]
        act = getattr(pol, "activation_fn", None)
        if act is not None and "activation" not in kw:
            kw["activation"] = "tanh" if getattr(act, "__name__", "") == "Tanh" else "relu"
        return cls(w, b, std=pol.log_std.detach().exp().cpu().numpy(), **kw)

    def _call(self, name, *args):
        st = getattr(self._lib, name)(self._h, *args)
        if st != 0:
            msg = self._lib.qs_policy_last_error(self._h)
            raise L.QuadsimError(f"{name} failed ({st}): {msg.decode() if msg else '?'}")

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.qs_policy_destroy(h)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    def forward(self, obs, deterministic=False, out=None, mean_out=None, raw_out=None):
        """actions (N, 4) CUDA tensor <- obs (N, in_dim) float32 CUDA tensor; asynchronous on the current stream."""
        if obs.dtype != torch.float32 or not obs.is_cuda or not obs.is_contiguous() or obs.shape[1] != self.in_dim:
            raise ValueError(f"forward needs a contiguous float32 CUDA tensor of shape (N, {self.in_dim})")
        n = obs.shape[0]
        if out is None:
            out = torch.empty((n, 4), dtype=torch.float32, device=obs.device)
        self._call("qs_policy_set_stream", L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        self._call("qs_policy_forward", L._vp(obs.data_ptr()), n, L._vp(out.data_ptr()),
                   L._vp(mean_out.data_ptr()) if mean_out is not None else None,
                   L._vp(raw_out.data_ptr()) if raw_out is not None else None, int(bool(deterministic)))
        return out

    def forward_packed(self, packed_obs, n, deterministic=False, out=None, mean_out=None, raw_out=None):
        """The same forward over the step kernel's packed BF16 observation blocks (``env.obs_format = "bf16_k32"``: a uint8
        CUDA tensor of ``env.packed_obs_bytes(n)`` bytes, e.g. what ``ObsPeerGather(packed=True).gather()`` returns): the
        blocks are the first MMA's operand as they are (TMA-loaded; nothing is converted).  Bit-identical to ``forward``
        on the float32 rows."""
        if packed_obs.dtype != torch.uint8 or not packed_obs.is_cuda or not packed_obs.is_contiguous():
            raise ValueError("forward_packed needs a contiguous uint8 CUDA tensor (packed BF16 observation blocks)")
        need = self._lib.qs_obs_packed_bytes(int(self.in_dim), int(n))
        if packed_obs.numel() < need:
            raise ValueError(f"packed observations of {n} envs take {need} bytes, got {packed_obs.numel()}")
        if out is None:
            out = torch.empty((n, 4), dtype=torch.float32, device=packed_obs.device)
        self._call("qs_policy_set_stream", L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        self._call("qs_policy_forward_packed", L._vp(packed_obs.data_ptr()), int(n), L._vp(out.data_ptr()),
                   L._vp(mean_out.data_ptr()) if mean_out is not None else None,
                   L._vp(raw_out.data_ptr()) if raw_out is not None else None, int(bool(deterministic)))
        return out

    def set_weights(self, weights, biases, std=None):
        """Replace the parameters (after a PPO update); shapes must not change."""
        for l, (w, b) in enumerate(zip(weights, biases)):
            w, b = np.ascontiguousarray(w, np.float32), np.ascontiguousarray(b, np.float32)
            assert w.shape == self.weights[l].shape and b.shape == self.biases[l].shape
            self.weights[l], self.biases[l] = w, b
            self._call("qs_policy_set_layer", l, w.ctypes.data_as(L._fp), b.ctypes.data_as(L._fp))
        if std is not None:
            self.std[:self.out_dim] = np.asarray(std, np.float32)
            self._call("qs_policy_set_std", self.std.ctypes.data_as(L._fp))

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        """SB3's ``model.predict`` signature (`3D quad race.ipynb:803`): NumPy in, ``(actions, None)`` out."""
        obs = torch.from_numpy(np.ascontiguousarray(observation, np.float32)).to(self.device)
        return self.forward(obs, deterministic=deterministic)[:, :self.out_dim].cpu().numpy(), None

    @property
    def launch_count(self):
        return int(self._lib.qs_policy_launch_count(self._h))

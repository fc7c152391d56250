"""PPO with Stable-Baselines3's interface over the GPU environment (BASELINE config 5).

The reference trains with ``from stable_baselines3 import PPO`` / ``VecMonitor``:

    env = VecMonitor(Quadcopter3DGates(num_envs=100, ...)); env.venv.disturbance_ranges = ...
    model = PPO("MlpPolicy", env, policy_kwargs=dict(activation_fn=torch.nn.ReLU,
                net_arch=[dict(pi=[120,120,120], vf=[120,120,120])], log_std_init=0), verbose=0,
                tensorboard_log=log_dir, n_steps=1000, batch_size=5000, n_epochs=10, gamma=0.999)
    model.learn(total_timesteps=..., reset_num_timesteps=False, tb_log_name=log_name); model.save(path)
    (`3D quad race.ipynb:765-831`, `PPO.load(path)` / `model.policy.mlp_extractor.policy_net` `:3985-3995`)

SB3 is third-party and not installable here, so this module carries the algorithm SB3 runs for that configuration
behind the same names -- constructor, ``learn``, ``predict``, ``save`` / ``load``, ``policy.*``, ``num_timesteps``,
``n_steps`` and ``VecMonitor`` -- so that the cell above runs with only its import line changed.  Two collectors:

  rollout="host"    SB3's ``collect_rollouts`` restated step for step over ANY VecEnv (NumPy in / NumPy out): policy
                    forward, ``np.clip``, ``env.step``, the ``TimeLimit.truncated`` bootstrap read from ``infos`` exactly
                    as SB3 does (so the reference env's aliased-dict quirk, SURVEY row a8, acts as it does under SB3),
                    ``buffer.add(self._last_obs, ...)`` AFTER the step.  The reference-faithful path; one host round trip
                    per step.
  rollout="device"  ``env.rollout``: ``n_steps`` x (tcgen05 policy forward -> fused env step) in ONE kernel launch, no
                    host.  Needs the GPU env with ``reset_rng="device"`` and a ReLU policy.  ``bootstrap="sb3_a8"``
                    reproduces, on the device, what SB3 + the reference's aliased ``infos`` do (every done env of a step
                    gets ``gamma * V(obs of the highest-index done env, already reset)`` added whenever ANY env timed out
                    in that step); ``bootstrap="none"`` (default of the device path) treats a time-out as a termination.
  rollout="auto"    "device" when the env allows it, else "host".

GAE runs in ``qs_gae`` (one thread per env) and the clipped-surrogate epochs either in torch (``update="torch"``) or in
the hand-written tcgen05 training kernels (``update="fused"``, see ``csrc/quadsim_train.cuh``).  Guards kept from
round 1 (documented deviations, active only on degenerate data): observations beyond ``obs_limit`` get zero weight,
the log-ratio is clamped to +-20, an update that ends non-finite is rolled back.
"""
from __future__ import annotations

import json
import math
import os
import time

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .policy import MlpPolicy

F_TRUNCATED = L.F_TRUNCATED


# ------------------------------------------------------------------------------------------------ VecMonitor
class VecMonitor:
    """``stable_baselines3.common.vec_env.VecMonitor`` for what the reference uses (`3D quad race.ipynb:769`, `:780`):
    wraps a VecEnv, exposes it as ``.venv``, forwards every other attribute, and on each done env puts
    ``info["episode"] = {"r": return, "l": length, "t": seconds}`` into a COPY of that env's info dict."""

    def __init__(self, venv, filename=None, info_keywords=()):
        self.venv = venv
        self.num_envs = venv.num_envs
        self.observation_space = getattr(venv, "observation_space", None)
        self.action_space = getattr(venv, "action_space", None)
        self.episode_count = 0
        self.t_start = time.time()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        self.info_keywords = tuple(info_keywords)

    def reset(self):
        obs = self.venv.reset()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        return obs

    def step_async(self, actions):
        self.venv.step_async(actions)

    def step_wait(self):
        obs, rewards, dones, infos = self.venv.step_wait()
        self.episode_returns += rewards
        self.episode_lengths += 1
        new_infos = list(infos[:])
        for i in np.flatnonzero(dones):
            info = infos[i].copy()
            ep = {"r": float(self.episode_returns[i]), "l": int(self.episode_lengths[i]),
                  "t": round(time.time() - self.t_start, 6)}
            for key in self.info_keywords:
                ep[key] = info[key]
            info["episode"] = ep
            self.episode_count += 1
            self.episode_returns[i] = 0
            self.episode_lengths[i] = 0
            new_infos[i] = info
        return obs, rewards, dones, new_infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        return self.venv.close()

    def __getattr__(self, name):  # only called when normal lookup fails: forward to the wrapped env
        if name == "venv":
            raise AttributeError(name)
        return getattr(self.venv, name)


# ------------------------------------------------------------------------------------------------ policy
def _mlp(in_dim, arch, act):
    layers, d = [], in_dim
    for h in arch:
        lin = nn.Linear(d, h)
        nn.init.orthogonal_(lin.weight, gain=math.sqrt(2))
        nn.init.zeros_(lin.bias)
        layers += [lin, act()]
        d = h
    return nn.Sequential(*layers), d


def _parse_net_arch(net_arch):
    """SB3's three spellings: dict(pi=[..], vf=[..]); the pre-1.8 [dict(pi=.., vf=..)] the reference uses (`:784`); or a
    plain list of widths used for both networks."""
    if net_arch is None:
        net_arch = dict(pi=[64, 64], vf=[64, 64])
    if isinstance(net_arch, (list, tuple)) and len(net_arch) == 1 and isinstance(net_arch[0], dict):
        net_arch = net_arch[0]
    if isinstance(net_arch, dict):
        return list(net_arch.get("pi", [])), list(net_arch.get("vf", []))
    if any(isinstance(x, dict) for x in net_arch):
        raise NotImplementedError("shared layers before the pi/vf split (SB3 < 1.8 net_arch=[64, dict(...)]) are not supported")
    return list(net_arch), list(net_arch)


class _MlpExtractor(nn.Module):
    def __init__(self, in_dim, pi_arch, vf_arch, act):
        super().__init__()
        self.policy_net, self.latent_dim_pi = _mlp(in_dim, pi_arch, act)
        self.value_net, self.latent_dim_vf = _mlp(in_dim, vf_arch, act)

    def forward(self, x):
        return self.policy_net(x), self.value_net(x)


class DiagGaussianDistribution:
    def __init__(self, action_dim):
        self.action_dim = action_dim

    def __repr__(self):
        return f"DiagGaussianDistribution(action_dim={self.action_dim})"


class ActorCriticPolicy(nn.Module):
    """The part of SB3's ``MlpPolicy`` the reference touches: ``mlp_extractor.policy_net`` / ``.value_net``,
    ``action_net``, ``value_net``, ``log_std``, ``action_dist`` (`3D quad race.ipynb:3988-3995`), plus ``forward`` /
    ``evaluate_actions`` / ``predict_values`` with SB3's signatures."""

    def __init__(self, obs_dim, action_dim=4, net_arch=None, activation_fn=nn.Tanh, log_std_init=0.0, ortho_init=True):
        super().__init__()
        pi_arch, vf_arch = _parse_net_arch(net_arch)
        self.obs_dim, self.action_dim = int(obs_dim), int(action_dim)
        self.net_arch, self.activation_fn = dict(pi=pi_arch, vf=vf_arch), activation_fn
        self.mlp_extractor = _MlpExtractor(obs_dim, pi_arch, vf_arch, activation_fn)
        self.action_net = nn.Linear(self.mlp_extractor.latent_dim_pi, action_dim)
        self.value_net = nn.Linear(self.mlp_extractor.latent_dim_vf, 1)
        self.log_std = nn.Parameter(torch.ones(action_dim) * float(log_std_init))
        self.action_dist = DiagGaussianDistribution(action_dim)
        if ortho_init:  # SB3: sqrt(2) for the extractor, 0.01 for the action net, 1 for the value net
            nn.init.orthogonal_(self.action_net.weight, gain=0.01)
            nn.init.zeros_(self.action_net.bias)
            nn.init.orthogonal_(self.value_net.weight, gain=1.0)
            nn.init.zeros_(self.value_net.bias)

    # -- pieces
    def mean_actions(self, obs):
        return self.action_net(self.mlp_extractor.policy_net(obs))

    def predict_values(self, obs):
        return self.value_net(self.mlp_extractor.value_net(obs))

    def log_prob(self, mean, actions):
        z = (actions - mean) / self.log_std.exp()
        return (-0.5 * z * z - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    def entropy(self):
        return (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum()

    # -- SB3 signatures
    def forward(self, obs, deterministic=False):
        mean = self.mean_actions(obs)
        actions = mean if deterministic else mean + self.log_std.exp() * torch.randn_like(mean)
        return actions, self.predict_values(obs), self.log_prob(mean, actions)

    def evaluate_actions(self, obs, actions):
        mean = self.mean_actions(obs)
        return self.predict_values(obs), self.log_prob(mean, actions), self.entropy().expand(obs.shape[0])

    def pi_layers(self):
        return [m for m in self.mlp_extractor.policy_net if isinstance(m, nn.Linear)] + [self.action_net]

    def vf_layers(self):
        return [m for m in self.mlp_extractor.value_net if isinstance(m, nn.Linear)] + [self.value_net]


def a8_bootstrap_(rewards, dones, flags, obs_next, value_fn, gamma):
    """What SB3's ``collect_rollouts`` does to the rewards when it runs on the reference env:
    the env's ``infos`` are ONE dict aliased N times (`3D quad race.ipynb:589-594`), so after a step it holds
    ``terminal_observation`` = the observation row of the HIGHEST-index done env (already reset) and
    ``TimeLimit.truncated`` iff ANY env reached max_steps; SB3 then adds ``gamma * V(terminal_observation)`` to the reward
    of EVERY done env of that step.  In place on ``rewards`` (T, N); ``dones`` / ``flags`` (T, N) uint8,
    ``obs_next`` (T, N, D) = the observations returned by each step, ``value_fn`` maps (T, D) -> (T,).  Pure torch, any
    device."""
    T, n = rewards.shape
    d = dones != 0
    any_tr = ((flags & F_TRUNCATED) != 0).any(dim=1)
    idx1 = torch.arange(1, n + 1, device=rewards.device)
    last = (d * idx1).amax(dim=1) - 1                       # highest done index per step, -1 = none
    sel = any_tr & (last >= 0)
    if not bool(sel.any()):
        return rewards
    term = obs_next[torch.arange(T, device=rewards.device), last.clamp_min(0)]
    v = value_fn(term).reshape(T).to(rewards.dtype)
    rewards += (gamma * v * sel)[:, None] * d
    return rewards


# ------------------------------------------------------------------------------------------------ PPO
class PPO:
    """``stable_baselines3.PPO`` for ``"MlpPolicy"`` on a continuous 4-action VecEnv.  Positional / keyword arguments,
    defaults and attribute names follow SB3 2.1; keyword-only extras select the B200 paths."""

    def __init__(self, policy, env=None, learning_rate=3e-4, n_steps=2048, batch_size=64, n_epochs=10, gamma=0.99,
                 gae_lambda=0.95, clip_range=0.2, clip_range_vf=None, normalize_advantage=True, ent_coef=0.0, vf_coef=0.5,
                 max_grad_norm=0.5, use_sde=False, sde_sample_freq=-1, target_kl=None, stats_window_size=100,
                 tensorboard_log=None, policy_kwargs=None, verbose=0, seed=None, device="auto", _init_setup_model=True, *,
                 rollout="auto", bootstrap=None, update="auto", evaluate="auto", tf32=True, amp=False, obs_limit=2.0e3,
                 value_limit=1.0e3, obs_dim=None):
        if not (policy == "MlpPolicy" or policy is ActorCriticPolicy):
            raise ValueError(f"only 'MlpPolicy' is supported (got {policy!r})")
        if use_sde or clip_range_vf is not None:
            raise NotImplementedError("use_sde / clip_range_vf are not used by the reference and not implemented")
        if rollout not in ("auto", "device", "host") or bootstrap not in (None, "none", "sb3_a8") or \
                update not in ("auto", "torch", "fused") or evaluate not in ("auto", "torch", "device", "mixed"):
            raise ValueError("rollout in {auto, device, host}; bootstrap in {None, 'none', 'sb3_a8'}; update in {auto, torch, fused}; "
                             "evaluate in {auto, torch, device, mixed}")
        self.env = env
        self.venv = getattr(env, "venv", env)  # VecMonitor(env) -> env
        self.learning_rate, self.n_steps, self.batch_size, self.n_epochs = float(learning_rate), int(n_steps), int(batch_size), int(n_epochs)
        self.gamma, self.gae_lambda, self.clip_range = float(gamma), float(gae_lambda), float(clip_range)
        self.ent_coef, self.vf_coef, self.max_grad_norm = float(ent_coef), float(vf_coef), float(max_grad_norm)
        self.normalize_advantage, self.target_kl = bool(normalize_advantage), target_kl
        self.tensorboard_log, self.verbose, self.seed = tensorboard_log, int(verbose), seed
        self.policy_kwargs = dict(policy_kwargs or {})
        self.obs_limit, self.value_limit, self.amp = float(obs_limit), float(value_limit), bool(amp)
        if not torch.cuda.is_available():
            raise L.QuadsimError("no CUDA device: PPO runs on the GPU next to the environment (there is no CPU fallback)")
        dev = getattr(self.venv, "device", None) if device == "auto" else device
        self.device = torch.device(dev if dev is not None else f"cuda:{torch.cuda.current_device()}")
        torch.manual_seed(0 if seed is None else int(seed))
        if tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
        d = obs_dim if obs_dim is not None else (getattr(self.venv, "state_len", None) or self.env.observation_space.shape[0])
        pk = dict(self.policy_kwargs)
        self.policy = ActorCriticPolicy(int(d), 4, net_arch=pk.pop("net_arch", None), activation_fn=pk.pop("activation_fn", nn.Tanh),
                                        log_std_init=pk.pop("log_std_init", 0.0), ortho_init=pk.pop("ortho_init", True)).to(self.device)
        if pk:
            raise NotImplementedError(f"policy_kwargs not supported: {sorted(pk)}")
        self.optimizer = torch.optim.Adam(self.policy.parameters(), lr=self.learning_rate, eps=1e-5)
        self.n_envs = getattr(self.env, "num_envs", None)
        self.num_timesteps, self._n_updates, self._episode_num = 0, 0, 0
        self._lib = L.load()
        self._last_obs, self._last_episode_starts = None, None
        self._started = False
        self.buffers, self.history = None, []
        self._writer, self._log_dir = None, None

        # ---- which collector / bootstrap / update
        from .envs import _QuadGatesBase
        pi, vf = self.policy.net_arch["pi"], self.policy.net_arch["vf"]
        act = {nn.ReLU: "relu", nn.Tanh: "tanh"}.get(self.policy.activation_fn)
        fits = act is not None and 1 <= len(pi) <= 4 and len(set(pi)) == 1 and pi[0] <= (127 if act == "relu" else 120)
        can_device = isinstance(self.venv, _QuadGatesBase) and getattr(self.venv, "reset_rng", "") == "device" and fits
        if rollout == "device" and not can_device:
            raise ValueError("rollout='device' needs the GPU env with reset_rng='device' and a ReLU / Tanh policy of 1-4 "
                             "equal hidden layers <= 127 (ReLU) / 120 (Tanh) wide")
        self.rollout = "device" if (rollout != "host" and can_device) else "host"
        # SB3 always bootstraps from infos; on the device path it is opt-in (the reference's aliased infos make it wrong)
        self.bootstrap = bootstrap if bootstrap is not None else ("sb3_a8" if self.rollout == "host" else "none")
        from .train_fused import fused_supported
        if update == "fused" and not fused_supported(self.policy):
            raise ValueError("update='fused' needs ReLU pi / vf networks of three equal hidden layers (<= 127 wide), "
                             "observations <= 63 wide")
        self.update = ("fused" if fused_supported(self.policy) else "torch") if update == "auto" else update
        self.actor = None
        if fits:
            self.actor = MlpPolicy(*self._pi_arrays(), std=self.policy.log_std.detach().exp().cpu().numpy(), device=self.device,
                                   seed=0 if seed is None else int(seed), activation=act)
        # values / old log-probs over the collected buffer.  "torch" (default): float32 / TF32 forwards.  "device": the tcgen05
        # forward kernel (an actor and a critic instance that sanitise their inputs like the learner), 23 -> 8 ms per
        # 8.4 M-sample buffer -- but the critic then runs on BF16 operands (values off by up to 2.6 % of their range), and on
        # the E2E env the one 130 s run with it plateaued at ep_rew_mean 115 - 137 instead of 145 - 160 (profiles/r2/ppo/):
        # opt-in, for when the collect phase matters more than the last gates per episode.
        fits_vf = act is not None and vf == pi and fits
        # "mixed": values in float32 / TF32 (torch), old log-probs from the tcgen05 actor -- the BF16 operands the sampling
        # actor and the fused trainer use, so the first epoch's PPO ratio is exactly 1 -- which drops the torch policy
        # forward over the buffer without touching the critic's precision.
        if evaluate == "device" and not (fits_vf and self.rollout == "device"):
            raise ValueError("evaluate='device' needs rollout='device' and pi / vf networks of the same supported shape")
        if evaluate == "mixed" and not (act is not None and fits and self.rollout == "device"):
            raise ValueError("evaluate='mixed' needs rollout='device' and a policy network of the supported shape")
        self.evaluate = evaluate if evaluate in ("device", "mixed") else "torch"
        self.eval_actor = self.critic = None
        if self.evaluate in ("device", "mixed"):
            self.eval_actor = MlpPolicy(*self._pi_arrays(), device=self.device, activation=act, obs_limit=self.obs_limit)
        if self.evaluate == "device":
            self.critic = MlpPolicy(*self._vf_arrays(), device=self.device, activation=act, obs_limit=self.obs_limit)

    # convenient aliases used by the tests / tools of this repository
    pi = property(lambda self: nn.Sequential(*self.policy.mlp_extractor.policy_net, self.policy.action_net))
    vf = property(lambda self: nn.Sequential(*self.policy.mlp_extractor.value_net, self.policy.value_net))
    log_std = property(lambda self: self.policy.log_std)

    def get_env(self):
        return self.env

    def set_env(self, env):
        self.env, self.venv, self.n_envs = env, getattr(env, "venv", env), env.num_envs

    # ------------------------------------------------------------------------------------------ plumbing
    def _pi_arrays(self):
        lin = self.policy.pi_layers()
        return ([m.weight.detach().cpu().numpy() for m in lin], [m.bias.detach().cpu().numpy() for m in lin])

    def _vf_arrays(self):
        lin = self.policy.vf_layers()
        return ([m.weight.detach().cpu().numpy() for m in lin], [m.bias.detach().cpu().numpy() for m in lin])

    def _publish(self):
        if self.actor is not None:
            w, b = self._pi_arrays()
            self.actor.set_weights(w, b, std=self.policy.log_std.detach().exp().cpu().numpy())
            if self.eval_actor is not None:
                self.eval_actor.set_weights(w, b)
            if self.critic is not None:
                self.critic.set_weights(*self._vf_arrays())

    def _sane(self, obs):
        """Observations as the learner sees them: finite and within +-obs_limit (identity for every sane sample:
        the env itself ends an episode at |p,q,r| > 1000, |x,y| > 10)."""
        return torch.nan_to_num(obs, nan=0.0, posinf=self.obs_limit, neginf=-self.obs_limit).clamp_(-self.obs_limit,
                                                                                                     self.obs_limit)

    _ACT_LIMIT = 1.0e4  # |mean + std * noise| of any sane sample is O(1)

    def _sane_act(self, a):
        return torch.nan_to_num(a, nan=0.0, posinf=self._ACT_LIMIT, neginf=-self._ACT_LIMIT).clamp_(-self._ACT_LIMIT,
                                                                                                   self._ACT_LIMIT)

    def _log_prob(self, obs, raw_actions, sane=False):
        mean = self.policy.mean_actions(obs if sane else self._sane(obs)).float()
        if not sane:
            raw_actions = self._sane_act(raw_actions)
        return self.policy.log_prob(mean, raw_actions)

    def _values(self, obs):
        return self.policy.predict_values(self._sane(obs)).squeeze(-1).float().clamp(-self.value_limit, self.value_limit)

    def _alloc(self, T, n, d):
        dev = self.device
        f = lambda *s: torch.empty(s, dtype=torch.float32, device=dev)
        self.buffers = {"obs": f(T + 1, n, d), "actions": f(T, n, 4), "raw_actions": f(T, n, 4), "rewards": f(T, n),
                        "dones": torch.empty((T, n), dtype=torch.uint8, device=dev),
                        "flags": torch.zeros((T, n), dtype=torch.uint8, device=dev),
                        "values": f(T + 1, n), "log_probs": f(T, n), "advantages": f(T, n), "weights": f(T, n),
                        "returns": f(T, n)}
        return self.buffers

    # ------------------------------------------------------------------------------------------ collect
    def collect_rollouts(self):
        b = self._collect_device() if self.rollout == "device" else self._collect_host()
        T, n = self.n_steps, self.n_envs
        st = self._lib.qs_gae(L._vp(b["rewards"].data_ptr()), L._vp(b["values"].data_ptr()), L._vp(b["dones"].data_ptr()),
                              L._vp(b["advantages"].data_ptr()), L._vp(b["returns"].data_ptr()), n, T, self.gamma,
                              self.gae_lambda, L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        if st != 0:
            raise L.QuadsimError(f"qs_gae failed ({st})")
        self.num_timesteps += T * n
        return b

    def _evaluate_buffer(self, b, T, n, d):
        """values (T+1, n), old log-probs and sample weights over the collected buffer (device path)."""
        if self.evaluate == "device":
            return self._evaluate_buffer_device(b, T, n, d)
        with torch.no_grad():
            torch.nan_to_num_(b["rewards"], nan=0.0, posinf=0.0, neginf=0.0)  # one NaN would poison a whole GAE column
            chunk = max(1, (1 << 22) // n)  # ~4M rows per forward
            for t0 in range(0, T + 1, chunk):
                t1 = min(T + 1, t0 + chunk)
                b["values"][t0:t1] = self._values(b["obs"][t0:t1].reshape(-1, d)).reshape(t1 - t0, n)
                if t0 < T:  # weight 0 for degenerate samples (non-finite or beyond obs_limit, or a non-finite action)
                    e = min(t1, T)
                    ok = torch.isfinite(b["obs"][t0:e]).all(-1) & (b["obs"][t0:e].abs().amax(-1) <= self.obs_limit)
                    ra = b["raw_actions"][t0:e]
                    ok &= torch.isfinite(ra).all(-1) & (ra.abs().amax(-1) <= self._ACT_LIMIT)
                    ok &= torch.isfinite(b["rewards"][t0:e])
                    b["weights"][t0:e] = ok.float()
            if self.evaluate == "mixed":  # the policy mean from the tcgen05 forward kernel (one launch over the buffer)
                rows = T * n
                if getattr(self, "_eval_scratch", None) is None or self._eval_scratch[0].shape[0] < rows:
                    self._eval_scratch = (torch.empty((rows, 4), device=self.device), torch.empty((rows, 4), device=self.device))
                out4, mean4 = self._eval_scratch
                self.eval_actor.forward(b["obs"][:T].reshape(rows, d), deterministic=True, out=out4[:rows], mean_out=mean4[:rows])
                b["log_probs"].copy_(self.policy.log_prob(mean4[:rows], self._sane_act(b["raw_actions"].reshape(rows, 4))).reshape(T, n))
                return
            for t0 in range(0, T, chunk):
                t1 = min(T, t0 + chunk)
                b["log_probs"][t0:t1] = self._log_prob(b["obs"][t0:t1].reshape(-1, d),
                                                       b["raw_actions"][t0:t1].reshape(-1, 4)).reshape(t1 - t0, n)

    def _evaluate_buffer_device(self, b, T, n, d):
        """The same three results from two launches of the tcgen05 forward kernel (critic over T+1 steps, policy mean over T)
        plus elementwise glue: BF16 operands like the actor that sampled the actions, so the PPO ratio starts at exactly 1."""
        with torch.no_grad():
            torch.nan_to_num_(b["rewards"], nan=0.0, posinf=0.0, neginf=0.0)  # one NaN would poison a whole GAE column
            rows = (T + 1) * n
            if getattr(self, "_eval_scratch", None) is None or self._eval_scratch[0].shape[0] < rows:
                self._eval_scratch = (torch.empty((rows, 4), device=self.device), torch.empty((rows, 4), device=self.device))
            out4, mean4 = self._eval_scratch
            flat = b["obs"].reshape(rows, d)
            self.critic.forward(flat, deterministic=True, out=out4[:rows], mean_out=mean4[:rows])
            b["values"].copy_(torch.nan_to_num(mean4[:rows, 0]).clamp_(-self.value_limit, self.value_limit).reshape(T + 1, n))
            obs, ra = b["obs"][:T], b["raw_actions"]
            ok = torch.isfinite(obs).all(-1) & (obs.abs().amax(-1) <= self.obs_limit)
            ok &= torch.isfinite(ra).all(-1) & (ra.abs().amax(-1) <= self._ACT_LIMIT) & torch.isfinite(b["rewards"])
            b["weights"].copy_(ok)
            self.eval_actor.forward(flat[:T * n], deterministic=True, out=out4[:T * n], mean_out=mean4[:T * n])
            b["log_probs"].copy_(self.policy.log_prob(mean4[:T * n], self._sane_act(ra.reshape(T * n, 4))).reshape(T, n))

    def _collect_device(self):
        env, T, n, d = self.venv, self.n_steps, self.n_envs, self.venv.state_len
        if not self._started:
            env.reset_tensor()
            self._started = True
        env.enable_stats(True)
        env.stats(reset=True)
        self._publish()
        b = self.buffers or self._alloc(T, n, d)
        b["obs"][0].copy_(env.current_obs_tensor())
        env.rollout(self.actor, T, buffers=b)
        self._evaluate_buffer(b, T, n, d)
        if self.bootstrap == "sb3_a8":
            with torch.no_grad():
                a8_bootstrap_(b["rewards"], b["dones"], b["flags"], b["obs"][1:], self._values, self.gamma)
        self._rollout_stats = env.stats(reset=True)
        return b

    def _collect_host(self):
        """SB3 ``OnPolicyAlgorithm.collect_rollouts`` over the NumPy-facing VecEnv interface."""
        env, T, n = self.env, self.n_steps, self.n_envs
        d = self.policy.obs_dim
        if self._last_obs is None:
            self._last_obs = env.reset()
            self._last_episode_starts = np.ones(n, dtype=bool)
        b = self.buffers or self._alloc(T, n, d)
        lo, hi = -1.0, 1.0
        sp = getattr(env, "action_space", None)
        if sp is not None and getattr(sp, "low", None) is not None:
            lo, hi = sp.low, sp.high
        rew_sum, n_done, ep_r, ep_l = 0.0, 0, [], []
        trunc = 0
        for t in range(T):
            with torch.no_grad():
                obs_t = torch.as_tensor(np.asarray(self._last_obs), dtype=torch.float32, device=self.device)
                actions, values, log_probs = self.policy(self._sane(obs_t))
            actions_np = actions.cpu().numpy()
            clipped = np.clip(actions_np, lo, hi)
            new_obs, rewards, dones, infos = env.step(clipped)
            rewards = np.array(rewards, dtype=np.float32)  # a copy: the env may recycle its output arrays
            for idx in np.flatnonzero(dones):
                info = infos[idx]
                if "episode" in info:
                    ep_r.append(info["episode"]["r"])
                    ep_l.append(info["episode"]["l"])
                if self.bootstrap == "sb3_a8" and info.get("terminal_observation") is not None and \
                        info.get("TimeLimit.truncated", False):
                    with torch.no_grad():
                        tobs = torch.as_tensor(np.asarray(info["terminal_observation"]), dtype=torch.float32,
                                               device=self.device).reshape(1, -1)
                        rewards[idx] += self.gamma * float(self._values(tobs)[0])
            trunc += sum(1 for j in np.flatnonzero(dones) if infos[j].get("TimeLimit.truncated", False))
            # rollout_buffer.add(self._last_obs, actions, rewards, self._last_episode_starts, values, log_probs):
            # _last_obs is read AFTER env.step (the env must not have recycled it)
            b["obs"][t].copy_(torch.as_tensor(np.asarray(self._last_obs), dtype=torch.float32), non_blocking=False)
            b["raw_actions"][t].copy_(actions)
            b["actions"][t].copy_(torch.as_tensor(clipped, dtype=torch.float32))
            b["rewards"][t].copy_(torch.as_tensor(rewards))
            b["dones"][t].copy_(torch.as_tensor(np.asarray(dones).astype(np.uint8)))
            b["values"][t].copy_(values.squeeze(-1).clamp(-self.value_limit, self.value_limit))
            b["log_probs"][t].copy_(log_probs)
            rew_sum += float(rewards.sum())
            n_done += int(np.count_nonzero(dones))
            self._last_obs, self._last_episode_starts = new_obs, dones
        with torch.no_grad():
            last = torch.as_tensor(np.asarray(self._last_obs), dtype=torch.float32, device=self.device)
            b["obs"][T].copy_(last)
            b["values"][T].copy_(self._values(last))
            ok = torch.isfinite(b["obs"][:T]).all(-1) & (b["obs"][:T].abs().amax(-1) <= self.obs_limit)
            b["weights"].copy_(ok.float())
        self._rollout_stats = {"reward_sum": rew_sum, "env_steps": T * n, "dones": n_done, "truncated": trunc,
                               "gates_passed": float("nan"), "ep_r": ep_r, "ep_l": ep_l}
        return b

    # ------------------------------------------------------------------------------------------ update
    def train(self):
        b, T, n, d = self.buffers, self.n_steps, self.n_envs, self.policy.obs_dim
        obs = b["obs"][:T].reshape(-1, d)
        act = b["raw_actions"].reshape(-1, 4)
        old_lp, adv, ret = b["log_probs"].reshape(-1), b["advantages"].reshape(-1), b["returns"].reshape(-1)
        wts = b["weights"].reshape(-1)
        total = T * n
        bs = min(self.batch_size, total)
        if self.update == "fused":
            from .train_fused import fused_update
            return fused_update(self, obs, act, old_lp, adv, ret, wts, total, bs)
        acc = torch.zeros(4, device=self.device)  # pg_loss, v_loss, clip_frac, approx_kl summed on the device
        updates = 0
        params = list(self.policy.parameters())
        snapshot = [p.detach().clone() for p in params]
        opt_state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._flat_opt_state().items()}
        for _ in range(self.n_epochs):
            perm = torch.randperm(total, device=self.device)
            for s0 in range(0, total, bs):  # like SB3's RolloutBuffer.get: the last, shorter minibatch is used too
                idx = perm[s0:s0 + bs]
                o, lp0, ad, rt, w = self._sane(obs[idx]), old_lp[idx], adv[idx], ret[idx], wts[idx]
                a = self._sane_act(act[idx])
                wsum = w.sum().clamp_min(1.0)
                if self.normalize_advantage and idx.numel() > 1:  # over the valid samples of the minibatch
                    m = (ad * w).sum() / wsum
                    sd = (((ad - m) ** 2 * w).sum() / (wsum - 1).clamp_min(1.0)).sqrt()
                    ad = (ad - m) / (sd + 1e-8)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                    lp = self._log_prob(o, a, sane=True)
                    v_pred = self.policy.predict_values(o).squeeze(-1).float()
                log_ratio = torch.nan_to_num(lp - lp0, nan=0.0).clamp(-20.0, 20.0)
                ratio = torch.exp(log_ratio)
                pg = -(torch.min(ad * ratio, ad * torch.clamp(ratio, 1 - self.clip_range, 1 + self.clip_range)) * w).sum() / wsum
                v_loss = ((v_pred - rt) ** 2 * w).sum() / wsum
                loss = pg + self.vf_coef * v_loss - self.ent_coef * self.policy.entropy()
                self.optimizer.zero_grad(set_to_none=True)
                loss.backward()
                nn.utils.clip_grad_norm_(params, self.max_grad_norm)
                self.optimizer.step()
                with torch.no_grad():
                    acc += torch.stack([pg, v_loss, (((ratio - 1).abs() > self.clip_range).float() * w).sum() / wsum,
                                        (((ratio - 1) - log_ratio) * w).sum() / wsum])
                    updates += 1
        finite = torch.stack([torch.isfinite(p).all() for p in params]).all() & torch.isfinite(acc).all()
        rolled_back = not bool(finite.item())  # the one host sync of the update
        if rolled_back:  # never observed with the masks above; keeps a long run alive if it ever happens
            with torch.no_grad():
                for p, q in zip(params, snapshot):
                    p.copy_(q)
            self._restore_opt_state(opt_state)
        self._n_updates += self.n_epochs
        a = (acc / max(1, updates)).tolist()
        return {"pg_loss": a[0], "v_loss": a[1], "clip_frac": a[2], "approx_kl": a[3], "updates": updates,
                "valid_frac": float(wts.mean().item()), "rolled_back": rolled_back}

    def _flat_opt_state(self):
        return {(i, k): v for i, st in enumerate(self.optimizer.state.values()) for k, v in st.items()}

    def _restore_opt_state(self, saved):
        with torch.no_grad():
            for i, st in enumerate(self.optimizer.state.values()):
                for k in list(st.keys()):
                    if torch.is_tensor(st[k]):  # state born inside the failed update restarts from zero
                        st[k].copy_(saved[(i, k)]) if (i, k) in saved else st[k].zero_()

    # ------------------------------------------------------------------------------------------ learn
    def _setup_logger(self, tb_log_name, reset_num_timesteps):
        if self.tensorboard_log is None:
            return
        os.makedirs(self.tensorboard_log, exist_ok=True)
        run = 1 + sum(1 for x in os.listdir(self.tensorboard_log) if x.startswith(tb_log_name + "_"))
        if not reset_num_timesteps and self._log_dir is not None:
            return  # continue the run this model already logs to (SB3 does the same)
        self._log_dir = os.path.join(self.tensorboard_log, f"{tb_log_name}_{run}")
        os.makedirs(self._log_dir, exist_ok=True)
        try:
            from torch.utils.tensorboard import SummaryWriter
            self._writer = SummaryWriter(self._log_dir)
        except Exception:  # no tensorboard in this image: progress.jsonl only
            self._writer = None

    def _log(self, rec):
        if self._log_dir is not None:
            with open(os.path.join(self._log_dir, "progress.jsonl"), "a") as f:
                f.write(json.dumps(rec) + "\n")
            if self._writer is not None:
                for tag, key in (("rollout/ep_rew_mean", "ep_rew_mean"), ("rollout/ep_len_mean", "ep_len_mean"),
                                 ("train/policy_gradient_loss", "pg_loss"), ("train/value_loss", "v_loss"),
                                 ("train/approx_kl", "approx_kl"), ("train/clip_fraction", "clip_frac"), ("train/std", "std"),
                                 ("time/fps", "fps")):
                    if rec.get(key) is not None and np.isfinite(rec[key]):
                        self._writer.add_scalar(tag, rec[key], rec["timesteps"])
                self._writer.flush()
        if self.verbose:
            print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in rec.items()})

    def learn(self, total_timesteps=None, callback=None, log_interval=1, tb_log_name="PPO", reset_num_timesteps=True,
              progress_bar=False, *, iterations=None, wall_clock_s=None, log=None):
        """``model.learn`` (`3D quad race.ipynb:820`): alternate collect_rollouts / train until ``total_timesteps`` more
        (``reset_num_timesteps=False``) or in total; ``iterations`` / ``wall_clock_s`` are extra stop conditions."""
        if self.env is None:
            raise ValueError("this model was loaded without an environment: call set_env() first")
        self.n_envs = self.env.num_envs
        if reset_num_timesteps:
            self.num_timesteps, self._episode_num = 0, 0
        elif total_timesteps is not None:
            total_timesteps += self.num_timesteps
        self._setup_logger(tb_log_name, reset_num_timesteps)
        t_start, it = time.perf_counter(), 0
        while True:
            t0 = time.perf_counter()
            self.collect_rollouts()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            es = self._rollout_stats
            tr = self.train()
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            episodes = max(1, es["dones"])
            ep_r, ep_l = es.get("ep_r"), es.get("ep_l")
            rec = {"iteration": it, "timesteps": self.num_timesteps, "wall_s": t2 - t_start, "collect_s": t1 - t0,
                   "train_s": t2 - t1, "fps": self.n_steps * self.n_envs / (t2 - t0),
                   "reward_per_step": es["reward_sum"] / max(1, es["env_steps"]),
                   "ep_rew_mean": float(np.mean(ep_r)) if ep_r else es["reward_sum"] / episodes,
                   "ep_len_mean": float(np.mean(ep_l)) if ep_l else es["env_steps"] / episodes,
                   "gates_per_episode": es.get("gates_passed", float("nan")) / episodes,
                   "crash_rate": (es["dones"] - es["truncated"]) / episodes,
                   "std": self.policy.log_std.detach().exp().mean().item(), "rollout": self.rollout,
                   "bootstrap": self.bootstrap, "update": self.update, **tr}
            self.history.append(rec)
            self._log(rec)
            if log:
                log(rec)
            it += 1
            if iterations is not None and it >= iterations:
                break
            if total_timesteps is not None and self.num_timesteps >= total_timesteps:
                break
            if wall_clock_s is not None and t2 - t_start >= wall_clock_s:
                break
            if iterations is None and total_timesteps is None and wall_clock_s is None:
                break
        self._publish()
        return self

    # ------------------------------------------------------------------------------------------ checkpoints
    _HYPER = ("learning_rate", "n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range", "ent_coef",
              "vf_coef", "max_grad_norm", "normalize_advantage", "obs_limit", "value_limit")

    @staticmethod
    def _ckpt_path(path):
        path = str(path)
        return path if path.endswith((".zip", ".pt")) else path + ".zip"

    def save(self, path):
        """``model.save(path)`` (`3D quad race.ipynb:823`): networks, log_std, optimizer and counters in ``<path>.zip``
        (tensors and plain containers only).  The env state is not checkpointed -- neither does the reference."""
        path = self._ckpt_path(path)
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        act = {nn.ReLU: "relu", nn.Tanh: "tanh"}.get(self.policy.activation_fn)
        if act is None:
            raise NotImplementedError("only ReLU / Tanh policies can be saved")
        torch.save({"policy": self.policy.state_dict(), "optimizer": self.optimizer.state_dict(),
                    "num_timesteps": self.num_timesteps, "hyper": {k: getattr(self, k) for k in self._HYPER},
                    "net_arch": {k: list(v) for k, v in self.policy.net_arch.items()}, "activation": act,
                    "obs_dim": self.policy.obs_dim, "history": self.history, "rollout": self.rollout,
                    "bootstrap": self.bootstrap, "update": self.update}, path)
        return path

    @classmethod
    def load(cls, path, env=None, device="auto", **overrides):
        """``PPO.load(path)`` (`:3985`), with or without an env; training continues where it stopped
        (``reset_num_timesteps=False``, `:820`)."""
        ck = torch.load(cls._ckpt_path(path), map_location="cpu", weights_only=True)
        if env is not None:
            d = getattr(getattr(env, "venv", env), "state_len", None) or env.observation_space.shape[0]
            if ck["obs_dim"] != d:
                raise ValueError(f"checkpoint was trained on {ck['obs_dim']}-wide observations, env has {d}")
        kw = dict(ck["hyper"])
        kw.update(overrides)
        pk = dict(net_arch=ck["net_arch"], activation_fn={"relu": nn.ReLU, "tanh": nn.Tanh}[ck["activation"]])
        self = cls("MlpPolicy", env, policy_kwargs=pk, device=device, obs_dim=ck["obs_dim"],
                   **{k: v for k, v in kw.items()})
        self.policy.load_state_dict(ck["policy"])
        self.optimizer.load_state_dict(ck["optimizer"])
        self.num_timesteps, self.history = ck["num_timesteps"], list(ck.get("history", []))
        self._publish()
        return self

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        """SB3's ``model.predict`` (`3D quad race.ipynb:803`): NumPy observations -> (clipped actions, None)."""
        if self.actor is not None:
            self._publish()
            return self.actor.predict(observation, deterministic=deterministic)
        with torch.no_grad():
            o = torch.as_tensor(np.asarray(observation), dtype=torch.float32, device=self.device).reshape(-1, self.policy.obs_dim)
            a, _, _ = self.policy(self._sane(o), deterministic=deterministic)
        return np.clip(a.cpu().numpy(), -1.0, 1.0), None

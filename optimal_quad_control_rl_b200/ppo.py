"""PPO over the device-resident rollout (BASELINE config 5).

The reference trains with Stable-Baselines3: ``PPO("MlpPolicy", VecMonitor(env), policy_kwargs=dict(activation_fn=ReLU,
net_arch=dict(pi=[120,120,120], vf=[120,120,120])), n_steps=1000, batch_size=5000, n_epochs=10, gamma=0.999)``
(`3D quad race.ipynb:784-795`) and ``model.learn`` (`:820`).  SB3 is not installable here, and its rollout loop is a
host loop with one H2D/D2H round trip per step -- exactly what the GPU env removes.  This module is the same algorithm
with the same hyper-parameter names and defaults, split the B200 way:

  collect   ``env.rollout`` -- ``n_steps`` x (tcgen05 policy forward -> fused env step) enqueued back to back, no host;
  evaluate  values and old log-probabilities in two large batched torch forwards over the (n_steps, N, D) buffer;
  GAE       ``qs_gae`` (one thread per env walks its column backwards);
  update    clipped-surrogate PPO epochs in PyTorch (autograd + Adam: library GEMMs, plumbing not product);
  publish   the new weights go back into the device policy (``MlpPolicy.set_weights``).

Differences from SB3, on purpose: (1) the time-limit bootstrap ``rewards += gamma * V(terminal_observation)`` is not
applied -- the reference's own ``infos`` alias one dict for all envs (`:589-594`), so SB3
bootstraps every done env of a step with the value of ONE post-reset observation whenever any env timed out; a
time-out is treated as a termination here; (2) actions are sampled by the BF16 tensor-core policy while the update
recomputes log-probabilities in float32 (|delta mean| <= ~1e-2 against std ~1: the importance ratio starts at 1 to
within 1e-2); (3) degenerate samples are masked: the reference env lets a tumbling quad's roll angle wind to ~1e8
through ``tan(theta)`` near +-pi/2 without ending the episode, and with 10^7 samples per
rollout (the reference collects 10^5) some always exist -- one of them turns the importance ratio into inf and the
whole network into NaN.  Observations beyond ``obs_limit`` get zero weight in both losses and in the advantage
normalisation, the log-ratio is clamped to +-20, critic values to +-``value_limit``, and an update that still ends
non-finite is rolled back.  ``tests/test_gpu_ppo.py`` checks GAE against SB3's formula, that a short run improves
the reward, and that poisoned observations cannot break the update."""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np
import torch
from torch import nn

from . import _lib as L
from .policy import MlpPolicy


def _mlp(in_dim, arch, out_dim, out_gain):
    layers, d = [], in_dim
    for h in arch:
        lin = nn.Linear(d, h)
        nn.init.orthogonal_(lin.weight, gain=math.sqrt(2))
        nn.init.zeros_(lin.bias)
        layers += [lin, nn.ReLU()]
        d = h
    out = nn.Linear(d, out_dim)
    nn.init.orthogonal_(out.weight, gain=out_gain)  # SB3: 0.01 for the action net, 1 for the value net
    nn.init.zeros_(out.bias)
    return nn.Sequential(*layers, out)


class PPO:
    def __init__(self, env, net_arch=(120, 120, 120), n_steps=1000, batch_size=5000, n_epochs=10, gamma=0.999,
                 gae_lambda=0.95, clip_range=0.2, ent_coef=0.0, vf_coef=0.5, max_grad_norm=0.5, learning_rate=3e-4,
                 log_std_init=0.0, normalize_advantage=True, seed=0, tf32=True, amp=False, obs_limit=2.0e3,
                 value_limit=1.0e3):
        self.env, self.device = env, env.device
        self.n_steps, self.batch_size, self.n_epochs = int(n_steps), int(batch_size), int(n_epochs)
        self.gamma, self.gae_lambda, self.clip_range = float(gamma), float(gae_lambda), float(clip_range)
        self.ent_coef, self.vf_coef, self.max_grad_norm = float(ent_coef), float(vf_coef), float(max_grad_norm)
        self.normalize_advantage = normalize_advantage
        self.obs_limit, self.value_limit = float(obs_limit), float(value_limit)
        self.amp = bool(amp)  # BF16 autocast of the update's GEMMs (float32 master weights, float32 losses)
        torch.manual_seed(seed)
        if tf32:
            torch.backends.cuda.matmul.allow_tf32 = True
        d = env.state_len
        self.pi = _mlp(d, net_arch, 4, 0.01).to(self.device)
        self.vf = _mlp(d, net_arch, 1, 1.0).to(self.device)
        self.log_std = nn.Parameter(torch.full((4,), float(log_std_init), device=self.device))
        self.optimizer = torch.optim.Adam([*self.pi.parameters(), *self.vf.parameters(), self.log_std], lr=learning_rate,
                                          eps=1e-5)
        self.actor = MlpPolicy(*self._pi_arrays(), std=self.log_std.detach().exp().cpu().numpy(), device=self.device,
                               seed=seed)
        self.num_timesteps = 0
        self._lib = L.load()
        self._started = False
        self.buffers = None
        self.history = []

    # ------------------------------------------------------------------------------------------ plumbing
    def _pi_arrays(self):
        lin = [m for m in self.pi if isinstance(m, nn.Linear)]
        return ([m.weight.detach().cpu().numpy() for m in lin], [m.bias.detach().cpu().numpy() for m in lin])

    def _publish(self):
        w, b = self._pi_arrays()
        self.actor.set_weights(w, b, std=self.log_std.detach().exp().cpu().numpy())

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
        mean = self.pi(obs if sane else self._sane(obs)).float()
        if not sane:
            raw_actions = self._sane_act(raw_actions)
        std = self.log_std.exp()
        z = (raw_actions - mean) / std
        return (-0.5 * z * z - self.log_std - 0.5 * math.log(2 * math.pi)).sum(-1)

    # ------------------------------------------------------------------------------------------ collect
    def collect_rollouts(self):
        env, T, n = self.env, self.n_steps, self.env.num_envs
        if not self._started:
            env.reset_tensor()
            self._started = True
        env.enable_stats(True)
        env.stats(reset=True)
        self._publish()
        if self.buffers is None:
            dev, d = self.device, env.state_len
            self.buffers = {"obs": torch.empty((T + 1, n, d), dtype=torch.float32, device=dev),
                            "actions": torch.empty((T, n, 4), dtype=torch.float32, device=dev),
                            "raw_actions": torch.empty((T, n, 4), dtype=torch.float32, device=dev),
                            "rewards": torch.empty((T, n), dtype=torch.float32, device=dev),
                            "dones": torch.empty((T, n), dtype=torch.uint8, device=dev),
                            "values": torch.empty((T + 1, n), dtype=torch.float32, device=dev),
                            "log_probs": torch.empty((T, n), dtype=torch.float32, device=dev),
                            "advantages": torch.empty((T, n), dtype=torch.float32, device=dev),
                            "weights": torch.empty((T, n), dtype=torch.float32, device=dev),
                            "returns": torch.empty((T, n), dtype=torch.float32, device=dev)}
        b = self.buffers
        b["obs"][0].copy_(env.current_obs_tensor())
        env.rollout(self.actor, T, buffers=b)
        with torch.no_grad():
            torch.nan_to_num_(b["rewards"], nan=0.0, posinf=0.0, neginf=0.0)  # one NaN would poison a whole GAE column
            chunk = max(1, (1 << 22) // n)  # ~4M rows per forward
            for t0 in range(0, T + 1, chunk):
                t1 = min(T + 1, t0 + chunk)
                o = b["obs"][t0:t1].reshape(-1, env.state_len)
                b["values"][t0:t1] = self.vf(self._sane(o)).reshape(t1 - t0, n).clamp_(-self.value_limit, self.value_limit)
                if t0 < T:  # weight 0 for degenerate samples (non-finite or beyond obs_limit, or a non-finite action)
                    e = min(t1, T)
                    ok = torch.isfinite(b["obs"][t0:e]).all(-1) & (b["obs"][t0:e].abs().amax(-1) <= self.obs_limit)
                    ra = b["raw_actions"][t0:e]
                    ok &= torch.isfinite(ra).all(-1) & (ra.abs().amax(-1) <= self._ACT_LIMIT)
                    ok &= torch.isfinite(b["rewards"][t0:e])
                    b["weights"][t0:e] = ok.float()
            for t0 in range(0, T, chunk):
                t1 = min(T, t0 + chunk)
                b["log_probs"][t0:t1] = self._log_prob(b["obs"][t0:t1].reshape(-1, env.state_len),
                                                       b["raw_actions"][t0:t1].reshape(-1, 4)).reshape(t1 - t0, n)
        st = self._lib.qs_gae(L._vp(b["rewards"].data_ptr()), L._vp(b["values"].data_ptr()), L._vp(b["dones"].data_ptr()),
                              L._vp(b["advantages"].data_ptr()), L._vp(b["returns"].data_ptr()), n, T, self.gamma,
                              self.gae_lambda, L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        if st != 0:
            raise L.QuadsimError(f"qs_gae failed ({st})")
        self.num_timesteps += T * n
        return b

    # ------------------------------------------------------------------------------------------ update
    def train(self):
        b, T, n, d = self.buffers, self.n_steps, self.env.num_envs, self.env.state_len
        obs = b["obs"][:T].reshape(-1, d)
        act = b["raw_actions"].reshape(-1, 4)
        old_lp, adv, ret = b["log_probs"].reshape(-1), b["advantages"].reshape(-1), b["returns"].reshape(-1)
        wts = b["weights"].reshape(-1)
        total = T * n
        bs = min(self.batch_size, total)
        acc = torch.zeros(4, device=self.device)  # pg_loss, v_loss, clip_frac, approx_kl summed on the device
        updates = 0
        params = [*self.pi.parameters(), *self.vf.parameters(), self.log_std]
        snapshot = [p.detach().clone() for p in params]
        opt_state = {k: (v.clone() if torch.is_tensor(v) else v) for k, v in self._flat_opt_state().items()}
        for _ in range(self.n_epochs):
            perm = torch.randperm(total, device=self.device)
            for s0 in range(0, total - bs + 1, bs):
                idx = perm[s0:s0 + bs]
                o, lp0, ad, rt, w = self._sane(obs[idx]), old_lp[idx], adv[idx], ret[idx], wts[idx]
                a = self._sane_act(act[idx])
                wsum = w.sum().clamp_min(1.0)
                if self.normalize_advantage:  # over the valid samples of the minibatch
                    m = (ad * w).sum() / wsum
                    sd = (((ad - m) ** 2 * w).sum() / (wsum - 1).clamp_min(1.0)).sqrt()
                    ad = (ad - m) / (sd + 1e-8)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=self.amp):
                    lp = self._log_prob(o, a, sane=True)
                    v_pred = self.vf(o).squeeze(-1).float()
                log_ratio = torch.nan_to_num(lp - lp0, nan=0.0).clamp(-20.0, 20.0)
                ratio = torch.exp(log_ratio)
                pg = -(torch.min(ad * ratio, ad * torch.clamp(ratio, 1 - self.clip_range, 1 + self.clip_range)) * w).sum() / wsum
                v_loss = ((v_pred - rt) ** 2 * w).sum() / wsum
                entropy = (0.5 + 0.5 * math.log(2 * math.pi) + self.log_std).sum()
                loss = pg + self.vf_coef * v_loss - self.ent_coef * entropy
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
    def learn(self, total_timesteps=None, iterations=None, wall_clock_s=None, log=None):
        """``model.learn`` (`:820`): alternate collect_rollouts / train until a timestep, iteration or time budget."""
        t_start, it = time.perf_counter(), 0
        while True:
            t0 = time.perf_counter()
            b = self.collect_rollouts()
            torch.cuda.synchronize(self.device)
            t1 = time.perf_counter()
            es = self.env.stats(reset=True)
            tr = self.train()
            torch.cuda.synchronize(self.device)
            t2 = time.perf_counter()
            episodes = max(1, es["dones"])
            rec = {"iteration": it, "timesteps": self.num_timesteps, "wall_s": t2 - t_start, "collect_s": t1 - t0,
                   "train_s": t2 - t1, "reward_per_step": es["reward_sum"] / max(1, es["env_steps"]),
                   "ep_rew_mean": es["reward_sum"] / episodes, "ep_len_mean": es["env_steps"] / episodes,
                   "gates_per_episode": es["gates_passed"] / episodes, "crash_rate": (es["dones"] - es["truncated"]) / episodes,
                   "std": self.log_std.detach().exp().mean().item(), **tr}
            self.history.append(rec)
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
    def save(self, path):
        """``model.save(path)`` (`3D quad race.ipynb:823`): networks, log_std, optimizer and counters in one file.
        (The env state is not checkpointed -- neither does the reference)"""
        import os
        os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
        path = path if str(path).endswith(".pt") else str(path) + ".pt"
        hp = {k: getattr(self, k) for k in ("n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "clip_range",
                                            "ent_coef", "vf_coef", "max_grad_norm", "normalize_advantage", "obs_limit",
                                            "value_limit")}
        torch.save({"pi": self.pi.state_dict(), "vf": self.vf.state_dict(), "log_std": self.log_std.detach().cpu(),
                    "optimizer": self.optimizer.state_dict(), "num_timesteps": self.num_timesteps, "hyper": hp,
                    "net_arch": [m.out_features for m in self.pi if isinstance(m, nn.Linear)][:-1],
                    "obs_dim": self.env.state_len, "history": self.history}, path)
        return path

    @classmethod
    def load(cls, path, env, **overrides):
        """``PPO.load(path)`` (`:3985`) against ``env``; training continues where it stopped
        (``reset_num_timesteps=False``, `:820`)."""
        path = path if str(path).endswith(".pt") else str(path) + ".pt"
        ck = torch.load(path, map_location="cpu", weights_only=False)
        if ck["obs_dim"] != env.state_len:
            raise ValueError(f"checkpoint was trained on {ck['obs_dim']}-wide observations, env has {env.state_len}")
        self = cls(env, net_arch=tuple(ck["net_arch"]), **{**ck["hyper"], **overrides})
        self.pi.load_state_dict(ck["pi"])
        self.vf.load_state_dict(ck["vf"])
        with torch.no_grad():
            self.log_std.copy_(ck["log_std"].to(self.device))
        self.optimizer.load_state_dict(ck["optimizer"])
        self.num_timesteps, self.history = ck["num_timesteps"], list(ck.get("history", []))
        self._publish()
        return self

    def predict(self, observation, state=None, episode_start=None, deterministic=False):
        self._publish()
        return self.actor.predict(observation, deterministic=deterministic)

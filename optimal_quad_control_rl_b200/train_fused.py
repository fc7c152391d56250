"""The PPO update on the tensor cores: host side of ``qs_trainer_*`` (include/quadsim.h,
csrc/quadsim_train.cuh).  ``PPO(update="fused")`` routes ``train()`` here: SB3's clipped-surrogate epochs
(`3D quad race.ipynb:784-795` hyper-parameters) run as three hand-written kernels per minibatch instead of ~60 torch
launches.  Float32 master parameters and Adam moments live in the trainer handle; the torch ``policy`` module is
refreshed from it after every update (it still serves ``predict_values`` / log-prob evaluation and checkpoints)."""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L


class FusedTrainer:
    def __init__(self, obs_dim, hidden, device):
        self._lib = L.load()
        self.device = torch.device(device)
        self.obs_dim, self.hidden = int(obs_dim), int(hidden)
        h = L._vp()
        st = self._lib.qs_trainer_create(C.byref(h), self.obs_dim, self.hidden, self.device.index,
                                         L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        if st != 0:
            raise L.QuadsimError(f"qs_trainer_create failed ({st}): {self._lib.qs_trainer_last_error(None).decode()}")
        self._h = h

    def _call(self, name, *args):
        st = getattr(self._lib, name)(self._h, *args)
        if st != 0:
            msg = self._lib.qs_trainer_last_error(self._h)
            raise L.QuadsimError(f"{name} failed ({st}): {msg.decode() if msg else '?'}")

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._lib.qs_trainer_destroy(h)

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # ---- parameters
    def load_from(self, policy):
        """torch ActorCriticPolicy -> trainer (float32 master copy)."""
        for net, layers in ((0, policy.pi_layers()), (1, policy.vf_layers())):
            for l, m in enumerate(layers):
                w = np.ascontiguousarray(m.weight.detach().cpu().numpy(), np.float32)
                b = np.ascontiguousarray(m.bias.detach().cpu().numpy(), np.float32)
                self._call("qs_trainer_set_layer", net, l, w.ctypes.data_as(L._fp), b.ctypes.data_as(L._fp))
        ls = np.ascontiguousarray(policy.log_std.detach().cpu().numpy(), np.float32)
        self._call("qs_trainer_set_log_std", ls.ctypes.data_as(L._fp))

    def store_to(self, policy):
        """trainer -> torch ActorCriticPolicy (in place)."""
        with torch.no_grad():
            for net, layers in ((0, policy.pi_layers()), (1, policy.vf_layers())):
                for l, m in enumerate(layers):
                    w = np.empty(tuple(m.weight.shape), np.float32)
                    b = np.empty(tuple(m.bias.shape), np.float32)
                    self._call("qs_trainer_get_layer", net, l, w.ctypes.data_as(L._fp), b.ctypes.data_as(L._fp))
                    m.weight.copy_(torch.from_numpy(w))
                    m.bias.copy_(torch.from_numpy(b))
            ls = np.empty(4, np.float32)
            self._call("qs_trainer_get_log_std", ls.ctypes.data_as(L._fp))
            policy.log_std.copy_(torch.from_numpy(ls))

    def grads(self, policy):
        """Gradients of the last minibatch in torch layout: ([(W, b)] policy net, [(W, b)] value net, log_std)."""
        out = []
        for net, layers in ((0, policy.pi_layers()), (1, policy.vf_layers())):
            g = []
            for l, m in enumerate(layers):
                w = np.empty(tuple(m.weight.shape), np.float32)
                b = np.empty(tuple(m.bias.shape), np.float32)
                self._call("qs_trainer_get_grad", net, l, w.ctypes.data_as(L._fp), b.ctypes.data_as(L._fp), None)
                g.append((w, b))
            out.append(g)
        ls = np.empty(4, np.float32)
        self._call("qs_trainer_get_grad", 0, 0, None, None, ls.ctypes.data_as(L._fp))
        return out[0], out[1], ls

    def reset_optimizer(self):
        self._call("qs_trainer_reset_optimizer")

    # ---- one minibatch
    def minibatch(self, idx, obs, act, old_lp, adv, ret, wts, hyper, apply=True):
        self._call("qs_trainer_set_stream", L._vp(torch.cuda.current_stream(self.device).cuda_stream))
        rows = int(idx.numel()) if idx is not None else int(obs.shape[0])
        p = lambda t: L._vp(t.data_ptr()) if t is not None else None
        self._call("qs_trainer_minibatch", p(idx), rows, p(obs), p(act), p(old_lp), p(adv), p(ret), p(wts), C.byref(hyper),
                   int(bool(apply)))

    def stats(self, reset=True):
        out = np.zeros(8, np.float32)
        self._call("qs_trainer_get_stats", out.ctypes.data_as(L._fp), int(reset))
        return out

    def publish(self, actor):
        """Hand the updated policy network to the device actor (``MlpPolicy``), device to device."""
        self._call("qs_trainer_publish", actor._h)

    @property
    def launch_count(self):
        return int(self._lib.qs_trainer_launch_count(self._h))


def hyper_of(ppo):
    return L.QsTrainHyper(ppo.learning_rate, 0.9, 0.999, 1e-5, ppo.clip_range, ppo.vf_coef, ppo.ent_coef, ppo.max_grad_norm,
                          ppo.obs_limit, ppo._ACT_LIMIT, int(ppo.normalize_advantage), 0)


def fused_supported(policy):
    pi, vf = policy.net_arch["pi"], policy.net_arch["vf"]
    return (policy.activation_fn is torch.nn.ReLU and len(pi) == 3 and pi == vf and len(set(pi)) == 1 and pi[0] <= 127
            and policy.obs_dim <= 63 and policy.action_dim == 4)


def fused_update(ppo, obs, act, old_lp, adv, ret, wts, total, bs):
    """PPO.train() with the hand-written kernels; same return dict as the torch path."""
    pol = ppo.policy
    if not fused_supported(pol):
        raise L.QuadsimError("update='fused' needs ReLU pi / vf networks of three equal hidden layers (<= 127 wide)")
    tr = getattr(ppo, "_trainer", None)
    if tr is None:
        tr = ppo._trainer = FusedTrainer(pol.obs_dim, pol.net_arch["pi"][0], ppo.device)
    tr.load_from(pol)
    snapshot = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    hyper = hyper_of(ppo)
    tr.stats(reset=True)
    updates = 0
    idx64 = None
    for _ in range(ppo.n_epochs):
        perm = torch.randperm(total, device=ppo.device)
        for s0 in range(0, total, bs):
            idx64 = perm[s0:s0 + bs]
            tr.minibatch(idx64, obs, act, old_lp, adv, ret, wts, hyper, apply=True)
            updates += 1
    st = tr.stats(reset=True)  # synchronises
    tr.store_to(pol)
    finite = all(bool(torch.isfinite(p).all()) for p in pol.parameters()) and bool(np.isfinite(st).all())
    if not finite:  # never observed with the sample masks; keeps a long run alive if it ever happens
        pol.load_state_dict(snapshot)
        tr.load_from(pol)
        tr.reset_optimizer()
    ppo._n_updates += ppo.n_epochs
    a = st / max(1, updates)
    return {"pg_loss": float(a[0]), "v_loss": float(a[1]), "clip_frac": float(a[2]), "approx_kl": float(a[3]),
            "grad_norm": float(a[4]), "updates": updates, "valid_frac": float(wts.mean().item()), "rolled_back": not finite}

"""Row f2: the hand-written PPO update (csrc/quadsim_train.cuh: tcgen05 forward + loss + backward, reduce, Adam) against
torch autograd in float32 on the same minibatch, and a short training run through ``PPO(update="fused")``."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def PK():
    import torch
    return dict(activation_fn=torch.nn.ReLU, net_arch=[dict(pi=[120, 120, 120], vf=[120, 120, 120])], log_std_init=0)


def torch_loss(pol, obs, act, old_lp, adv, ret, w, clip, vf_coef, ent_coef, normalize):
    """The loss of ppo.PPO.train() (SB3's, with sample weights)."""
    import torch
    wsum = w.sum().clamp_min(1.0)
    ad = adv
    if normalize:
        m = (ad * w).sum() / wsum
        sd = (((ad - m) ** 2 * w).sum() / (wsum - 1).clamp_min(1.0)).sqrt()
        ad = (ad - m) / (sd + 1e-8)
    mean = pol.mean_actions(obs)
    lp = pol.log_prob(mean, act)
    lr = (lp - old_lp).clamp(-20, 20)
    ratio = torch.exp(lr)
    pg = -(torch.min(ad * ratio, ad * torch.clamp(ratio, 1 - clip, 1 + clip)) * w).sum() / wsum
    v = pol.predict_values(obs).squeeze(-1)
    vl = ((v - ret) ** 2 * w).sum() / wsum
    return pg + vf_coef * vl - ent_coef * pol.entropy(), pg, vl


def bf16_emulated_grads(pol, obs, act, old_lp, adv, ret, w, clip, vf_coef, ent_coef, normalize, obs_limit=2000.0):
    """torch autograd through a float32 network that rounds exactly where the kernel rounds: BF16 inputs, weights and
    biases, BF16 activations after every ReLU, and BF16 gradients dZ (what the backward GEMMs consume); float32
    accumulation and loss math.  Returns gradients in the layout of FusedTrainer.grads."""
    import torch
    r16 = lambda t: t.to(torch.bfloat16).to(torch.float32)

    class RoundFwd(torch.autograd.Function):        # value rounded to BF16, gradient passes
        @staticmethod
        def forward(ctx, x):
            return r16(x)

        @staticmethod
        def backward(ctx, g):
            return g

    class RoundBwd(torch.autograd.Function):        # identity, gradient rounded to BF16
        @staticmethod
        def forward(ctx, x):
            return x.clone()

        @staticmethod
        def backward(ctx, g):
            return r16(g)

    def net(layers, x):
        leaves = []
        h = r16(x.clamp(-obs_limit, obs_limit))
        for l, m in enumerate(layers):
            W, b = r16(m.weight.detach()).requires_grad_(), r16(m.bias.detach()).requires_grad_()
            leaves.append((W, b))
            z = RoundBwd.apply(h @ W.t() + b)
            h = RoundFwd.apply(torch.relu(z)) if l < len(layers) - 1 else z
        return h, leaves

    wsum = w.sum().clamp_min(1.0)
    ad = adv
    if normalize:
        m = (ad * w).sum() / wsum
        sd = (((ad - m) ** 2 * w).sum() / (wsum - 1).clamp_min(1.0)).sqrt()
        ad = (ad - m) / (sd + 1e-8)
    mean, pi_leaves = net(pol.pi_layers(), obs)
    v, vf_leaves = net(pol.vf_layers(), obs)
    log_std = pol.log_std.detach().clone().requires_grad_()
    z = (act - mean) / log_std.exp()
    lp = (-0.5 * z * z - log_std - 0.5 * math.log(2 * math.pi)).sum(-1)
    lr = (lp - old_lp).clamp(-20, 20)
    ratio = torch.exp(lr)
    pg = -(torch.min(ad * ratio, ad * torch.clamp(ratio, 1 - clip, 1 + clip)) * w).sum() / wsum
    vl = ((v.squeeze(-1) - ret) ** 2 * w).sum() / wsum
    ent = (0.5 + 0.5 * math.log(2 * math.pi) + log_std).sum()
    (pg + vf_coef * vl - ent_coef * ent).backward()
    g = lambda leaves: [(W.grad.cpu().numpy(), b.grad.cpu().numpy()) for W, b in leaves]
    return g(pi_leaves), g(vf_leaves), log_std.grad.cpu().numpy(), pg.item(), vl.item()


def make_batch(pol, total, seed, dev):
    import torch
    g = torch.Generator(device="cpu").manual_seed(seed)
    obs = torch.randn(total, 24, generator=g).to(dev)
    with torch.no_grad():
        mean = pol.mean_actions(obs)
    act = mean + torch.randn(total, 4, generator=g).to(dev)
    with torch.no_grad():
        old_lp = pol.log_prob(mean, act) + 0.3 * torch.randn(total, generator=g).to(dev)   # ratios spread around 1: clipping active
    adv = torch.randn(total, generator=g).to(dev) * 2 + 0.5
    ret = torch.randn(total, generator=g).to(dev) * 3
    w = (torch.rand(total, generator=g) > 0.05).float().to(dev)
    return obs, act, old_lp, adv, ret, w


def rel(a, b):
    return float(np.linalg.norm(np.asarray(a, np.float64) - np.asarray(b, np.float64)) / max(np.linalg.norm(np.asarray(b, np.float64)), 1e-30))


@pytest.mark.parametrize("rows,use_idx", [(128, False), (1000, True), (20000, True)])
def test_fused_gradients_match_torch_autograd(rows, use_idx):
    import torch
    from optimal_quad_control_rl_b200.ppo import ActorCriticPolicy
    from optimal_quad_control_rl_b200.train_fused import FusedTrainer
    from optimal_quad_control_rl_b200 import _lib as L
    dev = torch.device("cuda", 0)
    torch.manual_seed(rows)
    pol = ActorCriticPolicy(24, 4, net_arch=PK()["net_arch"], activation_fn=torch.nn.ReLU, log_std_init=-0.3).to(dev)
    with torch.no_grad():  # SB3's 0.01-gain action net would make every policy gradient tiny: use a trained-looking scale
        pol.action_net.weight.mul_(30.0)
        for m in pol.vf_layers():
            m.bias.add_(0.05 * torch.randn_like(m.bias))
    total = rows + 777 if use_idx else rows
    obs, act, old_lp, adv, ret, w = make_batch(pol, total, rows, dev)
    idx = torch.randperm(total, device=dev)[:rows] if use_idx else None
    sel = (lambda t: t[idx]) if use_idx else (lambda t: t)
    hyper = L.QsTrainHyper(3e-4, 0.9, 0.999, 1e-5, 0.2, 0.5, 0.01, 0.5, 2000.0, 1e4, 1, 0)
    loss, pg, vl = torch_loss(pol, sel(obs), sel(act), sel(old_lp), sel(adv), sel(ret), sel(w), 0.2, 0.5, 0.01, True)
    pol.zero_grad()
    loss.backward()
    tr = FusedTrainer(24, 120, dev)
    tr.load_from(pol)
    tr.stats(reset=True)
    tr.minibatch(idx, obs, act, old_lp, adv, ret, w, hyper, apply=False)
    g_pi, g_vf, g_ls = tr.grads(pol)
    st = tr.stats()
    # (1) against torch autograd through a network that rounds to BF16 exactly where the kernel does: this isolates the
    #     kernel (GEMM operand layouts, masks, loss gradient, reductions) -- only float32 summation order is left.
    e_pi, e_vf, e_ls_ref, e_pg, e_vl = bf16_emulated_grads(pol, sel(obs), sel(act), sel(old_lp), sel(adv), sel(ret), sel(w),
                                                           0.2, 0.5, 0.01, True)
    # (2) against plain float32 autograd: the BF16 forward flips the ReLU mask of pre-activations within ~1e-2 sigma of
    #     zero (~1 % of the units), which shows up as a ~sqrt(flip fraction) relative error of the hidden-layer gradients
    #     on incoherent (random) data; direction and norm must agree.
    worst_emu, worst_f32, min_cos, report = 0.0, 0.0, 1.0, []
    for got, emu, layers, name in ((g_pi, e_pi, pol.pi_layers(), "pi"), (g_vf, e_vf, pol.vf_layers(), "vf")):
        for l, ((gw, gb), (ew_, eb_), m) in enumerate(zip(got, emu, layers)):
            tw = m.weight.grad.cpu().numpy()
            e1, e2, e3 = rel(gw, ew_), rel(gb, eb_), rel(gw, tw)
            cos = float((gw.astype(np.float64) * tw).sum() / (np.linalg.norm(gw) * np.linalg.norm(tw) + 1e-30))
            report.append("%s layer %d: vs BF16-emulated W %.2e b %.2e | vs float32 W %.2e cos %.4f |gW| %.3e (torch %.3e)" % (
                name, l, e1, e2, e3, cos, np.linalg.norm(gw), np.linalg.norm(tw)))
            worst_emu, worst_f32, min_cos = max(worst_emu, e1, e2), max(worst_f32, e3), min(min_cos, cos)
    e_ls = rel(g_ls, e_ls_ref)
    report.append("log_std: rel err vs emulated %.2e vs float32 %.2e" % (e_ls, rel(g_ls, pol.log_std.grad.cpu().numpy())))
    print("\n".join(report))
    print("loss stats got", st[:4], "emulated pg %.5f v %.5f | float32 pg %.5f v %.5f" % (e_pg, e_vl, pg.item(), vl.item()))
    assert abs(st[0] - e_pg) <= 2e-3 * max(1.0, abs(e_pg)) and abs(st[1] - e_vl) <= 2e-3 * max(1.0, abs(e_vl))
    assert abs(st[0] - pg.item()) <= 3e-2 * max(1.0, abs(pg.item())) and abs(st[1] - vl.item()) <= 3e-2 * max(1.0, abs(vl.item()))
    assert worst_emu <= 1e-2 and e_ls <= 1e-2, "\n".join(report)          # the kernel does what it says
    assert worst_f32 <= 0.2 and min_cos >= 0.98, "\n".join(report)        # and that is PPO's gradient
    print("worst relative gradient error: vs BF16-emulated autograd %.2e, vs float32 autograd %.2e (cos >= %.4f)" % (
        worst_emu, worst_f32, min_cos))
    tr.close()


def test_fused_adam_step_matches_torch_adam():
    """clip_grad_norm_ + torch.optim.Adam(eps=1e-5) on the kernel's own gradients == the parameters the kernels produce."""
    import torch
    from optimal_quad_control_rl_b200.ppo import ActorCriticPolicy
    from optimal_quad_control_rl_b200.train_fused import FusedTrainer
    from optimal_quad_control_rl_b200 import _lib as L
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    pol = ActorCriticPolicy(24, 4, net_arch=PK()["net_arch"], activation_fn=torch.nn.ReLU, log_std_init=0.0).to(dev)
    with torch.no_grad():
        pol.action_net.weight.mul_(30.0)
    obs, act, old_lp, adv, ret, w = make_batch(pol, 4096, 1, dev)
    hyper = L.QsTrainHyper(1e-3, 0.9, 0.999, 1e-5, 0.2, 0.5, 0.0, 0.5, 2000.0, 1e4, 1, 0)
    tr = FusedTrainer(24, 120, dev)
    tr.load_from(pol)
    opt = torch.optim.Adam(pol.parameters(), lr=1e-3, eps=1e-5)
    ref = ActorCriticPolicy(24, 4, net_arch=PK()["net_arch"], activation_fn=torch.nn.ReLU).to(dev)
    for step in range(3):
        tr.minibatch(None, obs, act, old_lp, adv, ret, w, hyper, apply=True)
        g_pi, g_vf, g_ls = tr.grads(pol)        # gradients the step just used
        for got, layers in ((g_pi, pol.pi_layers()), (g_vf, pol.vf_layers())):
            for (gw, gb), m in zip(got, layers):
                m.weight.grad = torch.from_numpy(gw).to(dev)
                m.bias.grad = torch.from_numpy(gb).to(dev)
        pol.log_std.grad = torch.from_numpy(g_ls).to(dev)
        torch.nn.utils.clip_grad_norm_(pol.parameters(), 0.5)
        opt.step()
        ref.load_state_dict(pol.state_dict())
        tr.store_to(ref)
        for (n1, p1), (_, p2) in zip(pol.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p1, p2, rtol=2e-5, atol=2e-6), (step, n1, (p1 - p2).abs().max().item())
        # the next minibatch must see the new weights: the torch copy follows the trainer exactly
        pol.load_state_dict(ref.state_dict())
    tr.close()


def test_ppo_fused_update_learns(tracks):
    """PPO(update='fused') on the INDI env: same hyper-parameters as the torch test, reward per step improves."""
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["indi"]
    env = Q.Quadcopter3DGatesINDI(4096, gp, gy, sp, gates_ahead=1, reset_rng="device", seed=0)
    ppo = Q.PPO("MlpPolicy", env, policy_kwargs=PK(), n_steps=128, batch_size=16384, n_epochs=4, gamma=0.999, seed=0,
                update="fused")
    ppo.learn(iterations=12)
    h = ppo.history
    print([round(r["reward_per_step"], 4) for r in h])
    assert all(np.isfinite(r["pg_loss"]) and np.isfinite(r["v_loss"]) and not r["rolled_back"] for r in h)
    assert np.mean([r["reward_per_step"] for r in h[-3:]]) > np.mean([r["reward_per_step"] for r in h[:2]]) + 0.005
    assert h[-1]["update"] == "fused" and ppo._trainer.launch_count > 0
    env.close()

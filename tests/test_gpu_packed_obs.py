"""Packed BF16 observations (`qs_set_obs_format(QS_OBS_BF16_K32)`, row (e) / config 4): the step kernel emits its
observations in the on-device policy's first-layer operand layout, the policy TMA-loads them.  Bit-exact against the
float32 path: packed == round-to-nearest-even BF16 of the float32 rows (+ the constant 1 of the folded bias), and
`forward_packed` == `forward`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def unpack(buf, n, obs_len):
    """(n, 8 * chunks) int16 view of the BF16 rows of a packed buffer: [32-env block][chunk 0..chunks-1][row 0..31][8 x bf16],
    chunks = ceil(obs_len / 8) (the chunks that carry an observation value)."""
    import torch
    ch = (obs_len + 7) // 8
    blocks = buf.view(-1, ch, 32, 8, 2)
    return blocks.permute(0, 2, 1, 3, 4).reshape(-1, 8 * ch, 2)[:n].contiguous().view(torch.int16).reshape(n, 8 * ch)


@pytest.mark.parametrize("variant,ga,n", [("e2e", 1, 4096), ("e2e", 1, 1000), ("e2e", 0, 130), ("e2e", 2, 5000),
                                          ("indi", 1, 4096), ("indi", 0, 777), ("indi", 4, 2049), ("indi", 3, 640)])
def test_packed_observations_equal_packed_float32_rows(variant, ga, n, tracks):
    import torch
    import optimal_quad_control_rl_b200 as Q
    cls = Q.Quadcopter3DGates if variant == "e2e" else Q.Quadcopter3DGatesINDI
    gp, gy, sp = tracks[variant]
    envs = []
    for _ in range(2):
        e = cls(n, gp, gy, sp, gates_ahead=ga, reset_rng="device", seed=11)
        if variant == "e2e":
            e.disturbance_ranges = Q.training_disturbance_ranges()
        e.max_steps = 7
        envs.append(e)
    ef, ep = envs
    ep.obs_format = "bf16_k32"
    D = ef.state_len
    pk = torch.zeros(ep.packed_obs_bytes(), dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(n)
    dims = [D, 120, 120, 4]
    w = [rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l])).astype(np.float32) for l in range(3)]
    b = [rng.normal(0, 0.1, dims[l + 1]).astype(np.float32) for l in range(3)]
    pol = Q.MlpPolicy(w, b, std=np.full(4, 0.3, np.float32), seed=3)
    pol2 = Q.MlpPolicy(w, b, std=np.full(4, 0.3, np.float32), seed=3)

    def check(of, what):
        ch = (D + 7) // 8
        want = torch.zeros((n, 32), device="cuda")
        want[:, :D] = of
        want[:, D] = 1.0  # (lands in a travelling chunk unless D % 8 == 0: then the policy kernel writes it itself)
        assert torch.equal(unpack(pk, n, D), want[:, :8 * ch].to(torch.bfloat16).view(torch.int16)), what
        blk = ch * 512
        tail = pk[(n + 31) // 32 * blk:]
        assert int(tail.abs().sum()) == 0, "blocks beyond the last env must stay zero"
        if n % 32:  # rows of the last block beyond n are zero
            last = pk[(n // 32) * blk:(n // 32 + 1) * blk].view(ch, 32, 16)
            assert int(last[:, n % 32:].abs().sum()) == 0
        a0 = pol.forward(of, deterministic=True).clone()
        a1 = pol.forward_packed(pk, n, deterministic=True).clone()
        assert torch.equal(a0, a1), what + ": forward_packed != forward"
        s0 = pol.forward(of).clone()       # sampled: same noise stream (same seed, same launch count)
        s1 = pol2.forward_packed(pk, n).clone()
        assert torch.equal(s0, s1), what + ": sampled actions"

    of = ef.reset_tensor()
    ep.reset_tensor(obs_out=pk)
    check(of, "reset")
    gen = torch.Generator(device="cuda").manual_seed(5)
    for t in range(12):  # max_steps = 7: every env is reset inside the step kernel at least once
        a = torch.rand((n, 4), generator=gen, device="cuda") * 2 - 1
        of, rf, df, _ = ef.step_tensor(a)
        _, rp, dp, _ = ep.step_tensor(a, obs_out=pk)
        assert torch.equal(rf, rp) and torch.equal(df, dp)
        check(of, f"step {t}")
    torch.cuda.synchronize()
    assert np.array_equal(ef.world_states, ep.world_states)
    for e in envs:
        e.close()


def pack(x):
    """Host-side restatement of the packed format (what the step kernel emits) for float32 rows `x` (n, D)."""
    import torch
    n, D = x.shape
    ch = (D + 7) // 8
    full = torch.zeros((n, 32), device=x.device)
    full[:, :D] = x
    full[:, D] = 1.0
    nb = (n + 127) // 128 * 4                                   # whole 128-row policy tiles
    rows = torch.zeros((nb * 32, 8 * ch), dtype=torch.bfloat16, device=x.device)
    rows[:n] = full[:, :8 * ch].to(torch.bfloat16)
    return rows.view(nb, 32, ch, 8).permute(0, 2, 1, 3).contiguous().view(torch.uint8).reshape(-1)


@pytest.mark.parametrize("in_dim", [8, 13, 16, 24, 31])
def test_forward_packed_for_every_chunk_pattern(in_dim):
    """Observation widths that end on a chunk boundary (8, 16, 24: the policy kernel writes the bias' constant-1 chunk and,
    for 16, a zero chunk behind it) and inside one (13, 31: the constant travels)."""
    import torch
    import optimal_quad_control_rl_b200 as Q
    rng = np.random.default_rng(in_dim)
    dims = [in_dim, 120, 120, 120, 4]
    w = [rng.normal(0, 1 / np.sqrt(dims[l]), (dims[l + 1], dims[l])).astype(np.float32) for l in range(4)]
    b = [rng.normal(0, 0.3, dims[l + 1]).astype(np.float32) for l in range(4)]
    pol = Q.MlpPolicy(w, b)
    for n in (1, 200, 4096):
        x = torch.from_numpy(rng.normal(0, 1, (n, in_dim)).astype(np.float32)).cuda()
        a0 = pol.forward(x, deterministic=True).clone()
        a1 = pol.forward_packed(pack(x), n, deterministic=True).clone()
        assert torch.equal(a0, a1), (in_dim, n)
        assert float(a0.abs().max()) > 1e-3


def test_packed_format_is_refused_where_float32_rows_are_the_contract(tracks):
    import torch
    import optimal_quad_control_rl_b200 as Q
    gp, gy, sp = tracks["e2e"]
    env = Q.Quadcopter3DGates(256, gp, gy, sp, gates_ahead=3, reset_rng="device")   # obs_len 32 > 31
    with pytest.raises(Q.QuadsimError):
        env.obs_format = "bf16_k32"
    env.close()
    env = Q.Quadcopter3DGates(256, gp, gy, sp, gates_ahead=1, reset_rng="device")
    env.obs_format = "bf16_k32"
    with pytest.raises(ValueError):
        env.step_tensor(torch.zeros((256, 4), device="cuda"))                          # needs obs_out
    with pytest.raises(Q.QuadsimError):
        env.step(np.zeros((256, 4), np.float32))                                       # NumPy path keeps float32 rows
    pol = Q.MlpPolicy.reference_controller()
    with pytest.raises(Q.QuadsimError):
        env.rollout(pol, 4)
    env.obs_format = "f32"
    env.reset()
    env.step(np.zeros((256, 4), np.float32))
    env.close()

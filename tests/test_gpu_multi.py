"""Multi-GPU (one process per GPU, NCCL): env sharding is invisible in the results, and the fused P2P observation
all-gather (step kernel storing into every peer's buffer) equals the NCCL all-gather.  Skipped on a 1-GPU box."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
import optimal_quad_control_rl_b200 as Q
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dev = torch.device("cuda", rank)
dist.init_process_group("nccl", device_id=dev)
total = 3 * 4096 + 512            # ragged over 2 ranks? no: 12800 = 2 * 6400 (50 tiles each), tail warp inside a rank
gp, gy, sp = Q.zigzag_track()
first, count = Q.shard_range(total, rank, world)
def make():
    env = Q.Quadcopter3DGates(count, gp, gy, sp, gates_ahead=1, device=dev, reset_rng="device", seed=4, env_offset=first)
    env.disturbance_ranges = Q.training_disturbance_ranges()
    env.max_steps = 5
    return env
gen = torch.Generator(device=dev).manual_seed(7)           # same seed on every rank: the global action tensor
acts = [(torch.rand((total, 4), generator=gen, device=dev) * 2 - 1) for _ in range(8)]
e_nccl, e_p2p = make(), make()
g_nccl = Q.ObsAllGather(total, e_nccl.state_len, dev)
g_p2p = Q.ObsPeerGather(total, e_p2p.state_len, dev)
g_p2p.attach(e_p2p)
e_nccl.reset_tensor(); e_p2p.reset_tensor()
for t in range(8):
    a = acts[t][first:first + count].contiguous()
    e_nccl.step_tensor(a, obs_out=g_nccl.local_slot())
    full_nccl = g_nccl.gather().clone()
    e_p2p.step_tensor(a, obs_out=g_p2p.local_slot())
    full_p2p = g_p2p.gather().clone()
    assert torch.equal(full_nccl, full_p2p), (rank, t)
    assert full_p2p.abs().sum().item() > 0                 # no extra barrier: ObsPeerGather double-buffers
# the packed BF16 gather (48 B per env over NVLink): same envs, the gathered buffer is the policy's operand as it is
e_pk = make()
g_pk = Q.ObsPeerGather(total, e_pk.state_len, dev, packed=True)
g_pk.attach(e_pk)
e_pk.reset_tensor(obs_out=g_pk.local_slot())
pol = Q.MlpPolicy.reference_controller(device=dev)
for t in range(8):
    a = acts[t][first:first + count].contiguous()
    e_pk.step_tensor(a, obs_out=g_pk.local_slot())
    full_pk = g_pk.gather()
    ch = (e_pk.state_len + 7) // 8                                                 # chunks that travel (24 values: 3, 48 B per env)
    blocks = full_pk.view(-1, ch, 32, 8, 2)                                        # [32-env block][chunk][row][8 x bf16 as 2 bytes]
    rows = blocks.permute(0, 2, 1, 3, 4).reshape(-1, 8 * ch, 2)[:total].contiguous().view(torch.int16).reshape(total, 8 * ch)
torch.cuda.synchronize()
# (the float32 reference of the LAST step: full_p2p, gathered above from identically seeded envs)
want = torch.zeros((total, 32), device=dev)
want[:, :e_pk.state_len] = full_p2p
want[:, e_pk.state_len] = 1.0
assert torch.equal(rows, want[:, :8 * ch].to(torch.bfloat16).view(torch.int16)), "packed gather != pack(float32 gather)"
a_pk = pol.forward_packed(full_pk, total, deterministic=True)
a_f32 = pol.forward(full_p2p, deterministic=True)
assert torch.equal(a_pk, a_f32), "policy on the packed gather != policy on the float32 gather"
# sharding is invisible: rank 0 also runs the unsharded env and compares the last gathered observations
if rank == 0:
    ref = Q.Quadcopter3DGates(total, gp, gy, sp, gates_ahead=1, device=dev, reset_rng="device", seed=4)
    ref.disturbance_ranges = Q.training_disturbance_ranges()
    ref.max_steps = 5
    ref.reset_tensor()
    for t in range(8):
        o = ref.step_tensor(acts[t])[0]
    assert torch.equal(o, full_p2p), "sharded != unsharded"
dist.barrier()
dist.destroy_process_group()
print("ok", rank)
"""


def test_p2p_fused_gather_equals_nccl_and_sharding_is_invisible(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for r in range(2):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE="2", LOCAL_RANK=str(r), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    for r, p in enumerate(procs):
        out, _ = p.communicate(timeout=600)
        assert p.returncode == 0, out[-3000:]
        assert f"ok {r}" in out

"""Multi-GPU: envs are independent, so a job of N quads is N/G contiguous envs per rank with NO data-path
collective.  The one optional exchange is the all-gather of observations for a policy that lives on one rank
(BASELINE.json config 4); the step kernel writes its tile straight into this rank's slice of the gather buffer,
so the collective needs no staging copy (NCCL in-place all-gather)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous block [first, first+count) of rank; remainders go to the low ranks."""
    base, rem = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, count


class ObsAllGather:
    """Pre-allocated (total_envs, D) buffer; ``local_slot()`` is where this rank's step writes its observations."""

    def __init__(self, total_envs, obs_len, device, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.first, self.count = shard_range(total_envs, self.rank, self.world)
        self.equal = total_envs % self.world == 0
        self.buf = torch.zeros((total_envs, obs_len), dtype=torch.float32, device=device)

    def local_slot(self):
        return self.buf[self.first:self.first + self.count]

    def gather(self):
        """All ranks end with every rank's observations in ``buf``."""
        if self.world == 1:
            return self.buf
        if self.equal:
            dist.all_gather_into_tensor(self.buf, self.local_slot(), group=self.group)
        else:  # ragged shards (total_envs % world != 0): every rank broadcasts its block in place
            works = []
            for r in range(self.world):
                f, c = shard_range(self.buf.shape[0], r, self.world)
                src = dist.get_global_rank(self.group, r) if self.group is not None else r
                works.append(dist.broadcast(self.buf[f:f + c], src=src, group=self.group, async_op=True))
            for w in works:
                w.wait()
        return self.buf


class ObsPeerGather:
    """The observation all-gather fused INTO the step kernel: the gather buffer lives in symmetric memory
    (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every process), each rank's step
    kernel stores its observation tiles into its own slice of ALL buffers over NVLink while it computes the next
    tile, and one cross-rank barrier replaces the collective.  Same contents as ``ObsAllGather`` afterwards."""

    def __init__(self, total_envs, obs_len, device, group=None):
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.first, self.count = shard_range(total_envs, self.rank, self.world)
        self.buf = symm_mem.empty((total_envs, obs_len), dtype=torch.float32, device=device)
        self.buf.zero_()
        self.handle = symm_mem.rendezvous(self.buf, self.group)
        self.peer_ptrs = [int(p) for r, p in enumerate(self.handle.buffer_ptrs) if r != self.rank]

    def attach(self, env):
        """Point ``env``'s step kernel at the peers; its ``obs_out`` must be ``local_slot()``."""
        env.set_obs_peers(self.peer_ptrs, self.first)

    def local_slot(self):
        return self.buf[self.first:self.first + self.count]

    def gather(self):
        """Wait until every rank's step kernel has finished writing: afterwards ``buf`` holds all observations."""
        self.handle.barrier()
        return self.buf

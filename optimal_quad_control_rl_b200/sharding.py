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
    """The observation all-gather fused INTO the step kernel: the gather buffers live in symmetric memory
    (``torch.distributed._symmetric_memory``: every rank's buffer is mapped into every process), each rank's step
    kernel stores its observation tiles into its own slice of ALL ranks' buffers over NVLink while it computes the
    next tile, and one cross-rank barrier replaces the collective.  Same contents as ``ObsAllGather`` afterwards.

    Protocol (race-free by construction): there are TWO symmetric buffers and consecutive steps alternate between
    them.  ``gather()`` = one barrier on the current stream: once a rank is past it, every rank's stores of this step
    have landed.  A rank can only start storing step t+2 (same buffer as step t) after passing the barrier of step
    t+1, i.e. after every peer has *enqueued past* its own consumers of step t's buffer (stream order: step t,
    barrier t, consumers of t, step t+1, barrier t+1, ...).  So nobody overwrites a buffer a peer still reads, as long
    as the consumers run on the stream ``gather()`` was called on."""

    PACK_TILE = 128  # packed format: 4 blocks of 32 envs per 128-env policy tile, ceil(obs_len / 8) x 512 B each

    def __init__(self, total_envs, obs_len, device, group=None, packed=False):
        """``packed=True``: the buffers hold the step kernel's packed BF16 blocks (``env.obs_format = "bf16_k32"``,
        ``16 * ceil(obs_len / 8)`` B per env over NVLink instead of ``4 * obs_len``) for ``MlpPolicy.forward_packed``; every rank's slice must
        then start on a multiple of 128 envs."""
        import torch.distributed._symmetric_memory as symm_mem
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.first, self.count = shard_range(total_envs, self.rank, self.world)
        self.packed = bool(packed)
        self.total_envs = int(total_envs)
        self.PACK_TILE_BYTES = 4 * ((int(obs_len) + 7) // 8) * 512
        if self.packed:
            if any(shard_range(total_envs, r, self.world)[0] % self.PACK_TILE for r in range(self.world)):
                raise ValueError("ObsPeerGather(packed=True): every rank's slice must start on a multiple of 128 envs")
        elif (self.first * obs_len * 4) % 16:
            raise ValueError("ObsPeerGather: this rank's slice must start on a 16-byte boundary "
                             "(first_env * obs_len * 4 % 16 == 0) for the TMA bulk stores into peer memory")
        self.bufs, self.handles, self.peer_ptrs = [], [], []
        for _ in range(2):
            if self.packed:
                tiles = (int(total_envs) + self.PACK_TILE - 1) // self.PACK_TILE
                b = symm_mem.empty((tiles * self.PACK_TILE_BYTES,), dtype=torch.uint8, device=device)
            else:
                b = symm_mem.empty((total_envs, obs_len), dtype=torch.float32, device=device)
            b.zero_()
            h = symm_mem.rendezvous(b, self.group)
            self.bufs.append(b)
            self.handles.append(h)
            self.peer_ptrs.append([int(p) for r, p in enumerate(h.buffer_ptrs) if r != self.rank])
        self.parity = 0
        self.env = None

    @property
    def buf(self):
        """The buffer the current step writes / the last ``gather()`` returned."""
        return self.bufs[self.parity]

    def attach(self, env):
        """Point ``env``'s step kernel at the peers; its ``obs_out`` must be ``local_slot()``."""
        self.env = env
        if self.packed and env.obs_format != "bf16_k32":
            env.obs_format = "bf16_k32"
        env.set_obs_peers(self.peer_ptrs[self.parity], self.first)

    def local_slot(self):
        """Where this rank's NEXT step must write its observations (``step_tensor(..., obs_out=local_slot())``)."""
        if self.packed:
            lo = self.first // self.PACK_TILE * self.PACK_TILE_BYTES
            hi = lo + (self.count + self.PACK_TILE - 1) // self.PACK_TILE * self.PACK_TILE_BYTES
            return self.bufs[self.parity][lo:hi]
        return self.bufs[self.parity][self.first:self.first + self.count]

    def gather(self):
        """Wait until every rank's step kernel has finished writing; returns the buffer holding all observations of
        this step and switches the attached env to the other buffer for the next step."""
        full = self.bufs[self.parity]
        self.handles[self.parity].barrier()
        self.parity ^= 1
        if self.env is not None:
            self.env.set_obs_peers(self.peer_ptrs[self.parity], self.first)
        return full

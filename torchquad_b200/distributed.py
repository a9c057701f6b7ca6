"""Multi-GPU layer: one process per GPU, `torch.distributed` (NCCL over NVLink 5 / NVSwitch on the GPU box,
gloo in the CPU tests).  The reference has nothing here ("spawn one process per GPU").

Design (DESIGN.md "Multi-GPU"):
  * every sample is a pure function of (seed, call, row / cube+index), so ranks take disjoint contiguous
    row (MC, VEGAS) or leading-axis slab (Newton-Cotes) ranges of the *same* sample set;
  * Monte Carlo / Newton-Cotes: no data-path collective, one all-reduce of the fp64 partial sums;
  * VEGAS: per iteration one all-reduce over the statistics [weights | counts | JF | JF2]; map and
    stratification state is then updated redundantly and stays bit-identical on all ranks.
"""
import torch

_state = {"enabled": False, "group": None}


def enable(group=None):
    """Shard subsequent integrations over the ranks of `group` (default process group)."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("torch.distributed is not initialised; call init_process_group first")
    _state["enabled"] = True
    _state["group"] = group


def disable():
    _state["enabled"] = False
    _state["group"] = None


def is_enabled():
    return _state["enabled"]


def rank_and_world():
    if not _state["enabled"]:
        return 0, 1
    import torch.distributed as dist

    return dist.get_rank(_state["group"]), dist.get_world_size(_state["group"])


def shard_range(total, rank=None, world=None):
    """Contiguous balanced split of range(total): rank r gets [total*r//R, total*(r+1)//R)."""
    if rank is None or world is None:
        rank, world = rank_and_world()
    return (total * rank) // world, (total * (rank + 1)) // world


def _skew(b):
    """Rotation of a round's blocks among the ranks (include/tqb200.h, tq_fused_vegas_sharded); ints or int64 tensors."""
    return b + (b >> 3) + (b >> 6) + (b >> 9) + (b >> 12) + (b >> 15) + (b >> 18)


def cube_shard(n_cubes, rank=None, world=None):
    """Deal of the VEGAS hypercubes to the ranks: (log2 block, cubes owned by this rank), or None when there are too few
    cubes to share (every rank then runs the whole problem redundantly, no collective).  Blocks of up to 4096 cubes
    spread the regions VEGAS concentrates its samples on over all ranks; round b of `world` consecutive blocks gives rank r
    the block at position (r + skew(b)) % world -- the rotation keeps a rank from owning a fixed digit of the cube index
    (a slab of one dimension) when world and the block size are powers of N_strat."""
    if rank is None or world is None:
        rank, world = rank_and_world()
    if world == 1 or n_cubes < 16 * world:
        return None
    lb = min(12, (n_cubes // (8 * world)).bit_length() - 1)
    block = 1 << lb
    full, rem = divmod(n_cubes, block)       # full blocks, cubes of the trailing partial block
    rounds, last = divmod(full, world)       # full rounds, full blocks of the last (incomplete) round
    pos = (rank + _skew(rounds)) % world     # this rank's position in the last round
    owned = rounds + (1 if pos < last else 0)
    n_local = owned * block + (rem if pos == last else 0)
    return lb, n_local


def global_cube_ids(n_local, lb, rank, world, device=None):
    """Global ids of this rank's cubes in local order (int64 tensor), the mapping of `cube_shard`."""
    l = torch.arange(n_local, dtype=torch.int64, device=device)
    b = l >> lb
    return ((b * world + (rank + _skew(b)) % world) << lb) + (l & ((1 << lb) - 1))


def all_reduce_sum_(*tensors):
    """In-place sum over ranks of every tensor (no-op when not enabled)."""
    if not _state["enabled"]:
        return
    import torch.distributed as dist

    for t in tensors:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=_state["group"])


def pack_all_reduce_sum_(tensors):
    """Sum several same-dtype tensors over ranks with ONE collective (flatten -> all-reduce -> scatter back)."""
    if not _state["enabled"] or not tensors:
        return
    import torch.distributed as dist

    flat = torch.cat([t.reshape(-1) for t in tensors])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=_state["group"])
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].reshape(t.shape))
        off += n

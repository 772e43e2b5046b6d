"""torch.distributed plumbing for the multi-GPU path: one process per GPU (torchrun), the
library's own NCCL communicator is bootstrapped from a unique id that rank 0 creates and
torch.distributed broadcasts.  Sharding helpers are pure host arithmetic."""
import os


def shard_range(n, world, rank):
    """Contiguous row shard [lo, hi) of n items for `rank` of `world` (first n % world ranks get one more)."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def broadcast_bytes(payload, src=0):
    """Broadcast a bytes object from `src` to every rank through torch.distributed."""
    import torch.distributed as td
    box = [payload if td.get_rank() == src else None]
    td.broadcast_object_list(box, src=src)
    return box[0]


def init_comm(ctx):
    """Give `ctx` a NCCL communicator spanning the torch.distributed world."""
    import torch.distributed as td
    world, rank = td.get_world_size(), td.get_rank()
    if world == 1:
        return
    uid = ctx.comm_unique_id() if rank == 0 else None
    uid = broadcast_bytes(uid, src=0)
    ctx.comm_init(world, rank, uid)

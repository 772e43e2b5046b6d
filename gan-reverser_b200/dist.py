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


def file_unique_id(ctx, world, rank, path, timeout_s=120.0):
    """The NCCL unique-id hand-off without torch.distributed (what `lua/ganrev.lua` comm_init_file does for one
    `th apply_r.lua` per GPU): rank 0 writes the id to `path + '.tmp'` and renames it (atomic), the others poll."""
    import time
    if rank == 0:
        uid = ctx.comm_unique_id()
        with open(path + ".tmp", "wb") as f:
            f.write(uid)
        os.replace(path + ".tmp", path)
        return uid
    t0 = time.time()
    while True:
        try:
            with open(path, "rb") as f:
                uid = f.read()
            if uid:
                return uid
        except FileNotFoundError:
            pass
        if time.time() - t0 > timeout_s:
            raise TimeoutError(f"no NCCL unique id at {path}")
        time.sleep(0.02)


def init_comm_file(ctx, world, rank, path, timeout_s=120.0):
    if world == 1:
        return
    ctx.comm_init(world, rank, file_unique_id(ctx, world, rank, path, timeout_s))

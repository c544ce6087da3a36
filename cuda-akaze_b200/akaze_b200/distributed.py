"""Multi-GPU plumbing (one process per GPU, torch.distributed).

detect+describe shards by frame: frames are independent units, so there is NO data-path collective
 — each rank runs its contiguous block of frames.
Matching shards the TRAIN set: every rank matches all queries against its contiguous train range and
produces per-query partial results in the associative form of akz_match(finalize=0); one all_gather of
nq x 16 bytes per rank (NCCL over NVLink on the GPU box, gloo in the CPU tests) is the only exchange,
followed by akz_match_merge(finalize=1) on every rank.
"""
import torch


def shard_bounds(n, world, rank):
    """Contiguous, balanced split of range(n): the first n % world ranks get one extra unit."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_partials(part, group=None):
    """all_gather a (nq, 4) int32 tensor from every rank -> (world, nq, 4)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    outs = [torch.empty_like(part) for _ in range(world)]
    dist.all_gather(outs, part.contiguous(), group=group)
    return torch.stack(outs)


def match_sharded(q, t_local, t_base, mode, local_fn, merge_fn, group=None):
    """q: all queries (replicated); t_local: this rank's train range starting at global index t_base.
    local_fn(q, t_local, t_base, mode) -> (nq, 4) partial; merge_fn(parts(world, nq, 4), mode) -> (nq, 4) final."""
    part = local_fn(q, t_local, t_base, mode)
    parts = gather_partials(part, group)
    return merge_fn(parts, mode)


def comm_init_from_group(ctx, group=None):
    """Create the context's NCCL communicator for the ranks of a torch.distributed group: rank 0 makes the id
    (akz_comm_unique_id), the group broadcasts its 128 bytes, every rank calls akz_comm_init."""
    import torch.distributed as dist
    import akaze_b200 as ab
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = ctx.device if backend == "nccl" else torch.device("cpu")
    idt = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        idt = torch.frombuffer(bytearray(ab.comm_unique_id()), dtype=torch.uint8).to(dev)
    dist.broadcast(idt, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(world, rank, bytes(idt.cpu().numpy().tobytes()))


def match_sharded_gpu(ctx, q, t_local, t_base, mode, group=None):
    """Python-side variant kept for comparison: akz_match(finalize=0) -> torch all_gather -> akz_match_merge(finalize=1), with
    two host synchronisations.  The production path is Context.match_sharded (akz_match_sharded: the gather runs inside the
    library on the context's stream)."""
    def local_fn(q_, t_, base_, mode_):
        r = ctx.match(q_, t_, mode_, t_index_base=base_, finalize=False)
        ctx.sync()
        return r

    def merge_fn(parts, mode_):
        r = ctx.match_merge(parts.contiguous(), mode_, finalize=True)
        ctx.sync()
        return r

    return match_sharded(q, t_local, t_base, mode, local_fn, merge_fn, group)

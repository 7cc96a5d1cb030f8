"""Multi-GPU plumbing: the batch is sharded across ranks (instances are independent — nothing in the
reference couples two controllers) and the only collective is ONE all-gather of the optimal inputs u*
per control step.  One process per GPU, torch.distributed (NCCL on GPUs, gloo in CPU tests)."""


def shard_range(batch, rank, world):
    """Contiguous shard [start, end) of `batch` instances for `rank` (sizes differ by at most one)."""
    base, rem = divmod(batch, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def allgather_outputs(full, start, end, group=None):
    """In-place all-gather: `full` is the [B, nu] buffer on every rank whose rows [start, end) this rank has
    already written (the solver's epilogue writes u* straight into that slice, include/bmpc.h bmpc_bind_output);
    afterwards every rank holds all rows.  Equal shards use all_gather_into_tensor (one NCCL call, no copies)."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    B = full.shape[0]
    if B % world == 0:
        dist.all_gather_into_tensor(full, full[start:end], group=group)
    else:
        rank = dist.get_rank(group)
        parts = [full[slice(*shard_range(B, r, world))] for r in range(world)]
        dist.all_gather(parts, parts[rank].clone(), group=group)
    return full

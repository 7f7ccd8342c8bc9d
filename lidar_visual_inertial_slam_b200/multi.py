"""Multi-GPU plumbing for the replay / bench drivers: one process per GPU, independent sequences
partitioned round-robin by sequence id, NO collective on the data path (a single
registration does not shard).  torch.distributed is used only for the barrier around the timed
region and for the max-over-ranks / sum-over-ranks of the timing scalars."""
import os

import torch
import torch.distributed as dist


def rank_info():
    """(rank, world_size, local_rank) from the torchrun environment (1 process = 1 GPU)."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")),
            int(os.environ.get("LOCAL_RANK", "0")))


def partition_sequences(n_sequences, rank, world):
    """sequence ids owned by `rank`: round-robin by id, as lvreg_replay does per GPU thread"""
    return list(range(rank, n_sequences, world))


def sequence_seed(base_seed, sequence_id):
    """seed = 0x5EED0000 + sequence_id"""
    return base_seed + sequence_id


def init(backend=None, device=None):
    rank, world, local_rank = rank_info()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return rank, world, local_rank


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def reduce_timing(elapsed_ms, units, device="cpu"):
    """max over ranks of the elapsed time, sum over ranks of the units processed.
    Returns (max_ms, total_units); whole-job throughput = total_units / max_ms."""
    if not (dist.is_available() and dist.is_initialized()):
        return float(elapsed_ms), float(units)
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def throughput(total_units, max_ms):
    return total_units / (max_ms * 1e-3) if max_ms > 0 else 0.0

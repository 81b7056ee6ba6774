"""Host-side plumbing for the sharded n-body path (one process per GPU).  torch.distributed is only the messenger:
the data path's collective is NCCL inside libee_b200.so (csrc/ee_nbody.cu).  Works on gloo/CPU tensors as well, which
is how tests/test_sharding_cpu.py covers it without GPUs."""
from typing import Optional, Tuple

import numpy as np


def shard_ranges(n: int, world: int, rank: int, exchange: str) -> Tuple[Tuple[int, int], Tuple[int, int]]:
    """(target range, source range) owned by `rank`; must match NBodyEngine's constructor (csrc/ee_nbody.cu).
    allgather: targets sharded, all sources.  allreduce: all targets, sources sharded."""
    if n % world:
        raise ValueError("n must be divisible by the number of ranks")
    per = n // world
    if world == 1:
        return (0, n), (0, n)
    if exchange == "allgather":
        return (rank * per, (rank + 1) * per), (0, n)
    if exchange == "allreduce":
        return (0, n), (rank * per, (rank + 1) * per)
    raise ValueError("unknown exchange %r" % exchange)


def broadcast_unique_id(dist, uid: Optional[bytes], device="cpu") -> bytes:
    """Rank 0 passes the 128-byte ncclUniqueId from ee_nccl_unique_id(); every rank returns the same bytes."""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        assert uid is not None and len(uid) == 128
        buf.copy_(torch.frombuffer(bytearray(uid), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().numpy().tobytes())


def connect_peers(dist, prop, device="cpu") -> None:
    """All-gather every rank's 512-byte CUDA-IPC blob and hand the table to the engine (NVLink peer path)."""
    import torch
    mine = torch.frombuffer(bytearray(prop.p2p_export()), dtype=torch.uint8).to(device)
    parts = [torch.zeros(512, dtype=torch.uint8, device=device) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, mine)
    prop.p2p_connect(b"".join(bytes(p.cpu().numpy().tobytes()) for p in parts))
    dist.barrier()


def max_over_ranks(dist, x: float, device="cpu") -> float:
    """Multi-GPU timings are reported as the max over ranks."""
    if dist is None:
        return float(x)
    import torch
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def whole_job_rate(units_per_rank_step: int, world: int, steps: int, seconds_max_over_ranks: float, sharded: bool) -> float:
    """body-steps/s of the whole job: a sharded system advances `units` bodies per step in total (strong scaling);
    replicas advance units x world."""
    total = units_per_rank_step * steps * (1 if sharded else world)
    return total / seconds_max_over_ranks

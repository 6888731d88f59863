"""Pattern sharding shared by every multi-GPU form of the likelihood:
rank r of R holds the contiguous pattern slice [floor(r*P/R), floor((r+1)*P/R)) and the only
exchange per evaluation is the sum of one double per rank."""
from __future__ import annotations


def shard_bounds(npatterns: int, rank: int, world: int):
    """Same formula as Data::sliceBounds (strom_b200/host/data.hpp)."""
    return npatterns * rank // world, npatterns * (rank + 1) // world


def sum_over_ranks(scalar_tensor):
    """All-reduce (sum) of the one-element tensor holding this rank's log-likelihood, in place.
    NCCL over NVLink when the tensor is on a GPU, gloo on the CPU (tests)."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(scalar_tensor, op=dist.ReduceOp.SUM)
    return scalar_tensor

"""Multi-GPU host logic: one process per GPU, the MSM is sharded by contiguous
base ranges, each rank runs the full Pippenger pipeline on its slice, and the per-rank partial
results (one 144-byte Jacobian point each) are exchanged with ONE all-gather and folded under
the group law.  EC addition is not an NCCL reduction op, hence gather + fold instead of
all-reduce.  Works over NCCL (GPU tensors) and gloo (CPU tensors, used by the tests)."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous [lo, hi) of n terms owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def all_gather_fold(partial, fold):
    """partial: uint8 tensor (144 B Jacobian record) on the backend's device.
    fold(bytes-like of world*144) -> combined point.  Every rank gets the result."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return fold(partial.cpu().numpy())
    out = torch.empty(partial.numel() * world, dtype=partial.dtype, device=partial.device)
    dist.all_gather_into_tensor(out, partial.contiguous())
    return fold(out.cpu().numpy())


def sharded_msm_g1(bases_shard, scalars_shard):
    """Per-rank entry point on the GPU path: local MSM on this rank's shard, all-gather over
    NCCL, fold on the GPU (dg_fold_g1).  Returns the 144-byte Jacobian result on every rank."""
    from . import lib
    import numpy as np
    part = lib.msm(bases_shard, scalars_shard)
    dev = torch.device('cuda', torch.cuda.current_device())
    t = torch.from_numpy(np.array(part)).to(dev)
    return all_gather_fold(t, lambda parts: bytes(lib.fold(parts)))

"""Multi-GPU layout of the path: objects are independent,
so each rank (one process per GPU) owns a contiguous block of them, the template
banks are replicated in every GPU's HBM, and there is NO data-path collective.
The only communication is the gather of the fixed-size per-object result
records at the end (torch.distributed: NCCL over NVLink on GPUs, gloo on CPU)."""
import numpy as np


def block_range(nobj, rank, world):
    """[start, stop) of the objects rank `rank` of `world` owns: contiguous
    blocks whose sizes differ by at most one."""
    base, extra = divmod(int(nobj), int(world))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_records(local, nobj, device=None):
    """All ranks pass their block's records (n_local, R) float64; every rank gets
    the (nobj, R) array in global object order.  Works without an initialised
    process group (single process)."""
    import torch
    import torch.distributed as dist
    local = np.ascontiguousarray(local, dtype=np.float64)
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        assert len(local) == nobj
        return local
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [block_range(nobj, r, world) for r in range(world)]
    assert len(local) == sizes[rank][1] - sizes[rank][0]
    nmax = max(b - a for a, b in sizes)
    R = local.shape[1]
    dev = device if device is not None else ('cuda' if dist.get_backend() == 'nccl' else 'cpu')
    buf = torch.zeros((nmax, R), dtype=torch.float64, device=dev)
    buf[:len(local)] = torch.from_numpy(local).to(dev)
    out = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(out, buf)
    return np.concatenate([o[:b - a].cpu().numpy() for o, (a, b) in zip(out, sizes)])

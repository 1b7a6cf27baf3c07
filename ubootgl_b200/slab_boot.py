"""Bootstrap helpers for the row-slab runs: one process per GPU, launched by
torchrun.  torch.distributed is plumbing only (rendezvous, the all-gather of
the 64-byte IPC blobs, barriers and the max-over-ranks of timings); halo rows
never go through it -- they move by peer stores inside libubgl's kernels."""
import os

import numpy as np


def init_distributed(backend=None):
    """Initialise torch.distributed from the torchrun environment."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        kw = {}
        if backend == "nccl":
            torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
            kw["device_id"] = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group(backend=backend, rank=int(os.environ.get("RANK", "0")),
                                world_size=int(os.environ.get("WORLD_SIZE", "1")), **kw)
    return dist.get_rank(), dist.get_world_size()


def shutdown(barrier=True):
    """Tear the process group down before the interpreter exits: a gloo process that exits with the group
    alive can die in a helper thread's destructor ("terminate called without an active exception"), and
    NCCL warns about leaked resources."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        if barrier:
            try:
                dist.barrier()
            except Exception:
                pass
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def blob_exchange():
    """callable(bytes) -> list[bytes]: all-gather of one small blob per rank."""
    import torch
    import torch.distributed as dist

    def exchange(blob):
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return [blob]
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
        mine = torch.frombuffer(bytearray(blob), dtype=torch.uint8).to(dev)
        out = [torch.empty_like(mine) for _ in range(dist.get_world_size())]
        dist.all_gather(out, mine)
        return [bytes(t.cpu().numpy().tobytes()) for t in out]

    return exchange


def allreduce_max(x):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allreduce_sum(x):
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier():
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()


def gather_rows(own_lo, rows, H):
    """Assemble a global (H, w) array on every rank from each rank's own rows."""
    import torch
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return rows.copy()
    objs = [None] * dist.get_world_size()
    dist.all_gather_object(objs, (int(own_lo), np.ascontiguousarray(rows)))
    out = np.zeros((H, rows.shape[1]), np.float32)
    for lo, r in objs:
        out[lo:lo + r.shape[0]] = r
    return out

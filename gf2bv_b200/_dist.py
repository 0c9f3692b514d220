"""Host-side plumbing of the one-process-per-GPU launch (torch.distributed is only the
rendezvous: the solver's own exchanges are NCCL calls inside libgf2b200).

Works on whichever backend the process group was initialised with (nccl on the GPU
box, gloo in the CPU tests)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def row_range(m: int, rank: int, world: int) -> tuple[int, int]:
    """Global rows [r0, r1) held by `rank`: the same split gf2b200_system_create uses."""
    return m * rank // world, m * (rank + 1) // world


def _device():
    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0) -> bytes:
    """Every rank gets rank `src`'s `nbytes` bytes (used for the 128-byte ncclUniqueId)."""
    t = torch.zeros(nbytes, dtype=torch.uint8, device=_device())
    if dist.get_rank() == src:
        assert payload is not None and len(payload) == nbytes
        t = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(_device())
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def all_max(value: float) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(value: int) -> int:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return int(value)
    t = torch.tensor([value], dtype=torch.int64, device=_device())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())

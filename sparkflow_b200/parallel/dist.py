"""Process-group plumbing. ``torch.distributed`` is used ONLY for bootstrap / control (rendezvous,
IPC-handle exchange, barriers) and by the gloo CPU transport; the GPU push/pull data path never
touches NCCL (the NCCL build of the same semantics lives in ``baseline/`` as the thing to beat)."""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Any, List, Optional

import torch


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    control_group: Any = None          # gloo group for object broadcasts / barriers

    @property
    def is_master(self) -> bool:
        return self.rank == 0


_CTX: Optional[DistContext] = None


def in_spmd_launch() -> bool:
    return int(os.environ.get("WORLD_SIZE", "1")) > 1 and "RANK" in os.environ


def get_context(init: bool = True) -> DistContext:
    """The SPMD context (one process per worker, e.g. under torchrun) or the trivial single-process one."""
    global _CTX
    import torch.distributed as dist

    if _CTX is not None and (_CTX.world == 1 or dist.is_initialized()):
        return _CTX
    if not in_spmd_launch() and not dist.is_initialized():
        _CTX = DistContext()
        return _CTX
    if not dist.is_initialized():
        if not init:
            return DistContext()
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        os.environ.setdefault("GLOO_SOCKET_IFNAME", "lo")
        use_cuda = torch.cuda.is_available()
        local_rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        if use_cuda:
            torch.cuda.set_device(local_rank % torch.cuda.device_count())
        dist.init_process_group(backend="cpu:gloo,cuda:nccl" if use_cuda else "gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", rank))
    ctrl = dist.new_group(backend="gloo")
    _CTX = DistContext(rank=rank, world=world, local_rank=local_rank, control_group=ctrl)
    return _CTX


def broadcast_object(ctx: DistContext, obj: Any, src: int = 0) -> Any:
    if ctx.world == 1:
        return obj
    import torch.distributed as dist

    box: List[Any] = [obj if ctx.rank == src else None]
    dist.broadcast_object_list(box, src=src, group=ctx.control_group)
    return box[0]


def barrier(ctx: DistContext) -> None:
    if ctx.world > 1:
        import torch.distributed as dist

        dist.barrier(group=ctx.control_group)


def all_gather_object(ctx: DistContext, obj: Any) -> List[Any]:
    if ctx.world == 1:
        return [obj]
    import torch.distributed as dist

    out: List[Any] = [None] * ctx.world
    dist.all_gather_object(out, obj, group=ctx.control_group)
    return out

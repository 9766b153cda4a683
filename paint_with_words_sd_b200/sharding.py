"""Image sharding across the GPUs of one box.

The unit of work is one image (its cond + uncond pair, so the CFG combine stays local).  Image i goes to
rank i mod G.  Each rank holds a full replica of the weights, delivered by ONE broadcast of a single
flat buffer at init (NCCL over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU tests); there is no
collective on the per-step path.  Final latents (32 KB per image at 512^2) are optionally gathered once
at the end.  The reference has no multi-device support at all (batch fixed to 1, gradio_pww.py:30-45
loops over seeds in Python), so this is a new capability rather than a port.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch
import torch.distributed as dist


def env_world():
    """(rank, local_rank, world_size) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def init_distributed(backend: Optional[str] = None) -> bool:
    """Initialise torch.distributed from the torchrun environment when WORLD_SIZE > 1."""
    rank, local_rank, world = env_world()
    if world == 1 or dist.is_initialized():
        return dist.is_initialized()
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    if backend == "nccl":
        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend, rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))
    else:
        dist.init_process_group(backend, rank=rank, world_size=world)
    return True


def shard_images(num_images: int, rank: int, world: int) -> List[int]:
    """Indices of the images owned by `rank`: i mod world == rank."""
    return [i for i in range(num_images) if i % world == rank]


def broadcast_module_weights(module: torch.nn.Module, src: int = 0) -> int:
    """Replicate parameters and buffers of `module` from `src` with a single broadcast of one flat buffer.
    Returns the number of bytes broadcast (0 when not distributed)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return 0
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return 0
    dtype, device = tensors[0].dtype, tensors[0].device
    same = all(t.dtype == dtype for t in tensors)
    if same:
        flat = torch.cat([t.reshape(-1) for t in tensors])
    else:                                   # mixed dtypes: ship raw bytes
        flat = torch.cat([t.contiguous().reshape(-1).view(torch.uint8) for t in tensors])
    dist.broadcast(flat, src=src)
    off = 0
    for t in tensors:
        n = t.numel() if same else t.numel() * t.element_size()
        chunk = flat[off:off + n]
        t.copy_((chunk if same else chunk.view(t.dtype)).reshape(t.shape))
        off += n
    return flat.numel() * flat.element_size()


def gather_latents(local: Dict[int, torch.Tensor], num_images: int) -> Optional[List[torch.Tensor]]:
    """Collect per-image final latents on every rank, ordered by image index (one all_gather_object-free
    all_gather of equally shaped tensors; images a rank does not own are zero-filled slots)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [local[i] for i in range(num_images)]
    world = dist.get_world_size()
    sample = next(iter(local.values()))
    per_rank = (num_images + world - 1) // world
    mine = torch.zeros((per_rank,) + tuple(sample.shape), dtype=sample.dtype, device=sample.device)
    for slot, i in enumerate(sorted(local)):
        mine[slot] = local[i]
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    res = []
    for i in range(num_images):
        res.append(out[i % world][i // world])
    return res

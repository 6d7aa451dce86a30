"""Multi-GPU sharding of the encode path: one process per GPU, no data-path collective.

ASTC blocks are independent (ASTC_Encode.hlsl:561-581 reads one tile and writes one
uint4), so the path shards two ways:

  * a large texture is cut into contiguous block-row BANDS (astc_b200_band): band g's
    input is one contiguous slab of texel rows and its output one contiguous slab of the
    final block buffer, so every rank writes its slice directly (device buffer, pinned
    host buffer or a region of the .astc file) -- nothing is exchanged;
  * a batch of textures / mip chains is dealt out whole, longest-processing-time first.

torch.distributed is used only by callers for rendezvous / barriers / timing; the one
helper here that communicates, gather_blocks(), is an optional convenience for when a
single rank wants the whole compressed texture.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Callable, List, Optional, Sequence

import numpy as np

from . import BLOCK_BYTES, band, block_dim, encode_option, output_size

__all__ = ["Band", "band_plan", "assign_textures", "encode_band", "write_astc_sharded", "gather_blocks"]

ASTC_HEADER_BYTES = 16


@dataclass(frozen=True)
class Band:
    part: int
    y0: int            # first texel row
    rows: int          # texel rows (only the last band can be short of a block multiple)
    byte_offset: int   # where this band's blocks start in the full block buffer
    nbytes: int


def band_plan(width: int, height: int, option: encode_option, parts: int) -> List[Band]:
    """Block rows split evenly into `parts` contiguous bands (some may be empty)."""
    return [Band(g, *band(width, height, option, parts, g)) for g in range(parts)]


def assign_textures(texel_counts: Sequence[int], parts: int) -> List[List[int]]:
    """Whole textures to ranks, longest-processing-time first; returns index lists per rank
    (each list ascending so a rank's outputs keep batch order)."""
    load = [0] * parts
    owner: List[List[int]] = [[] for _ in range(parts)]
    for i in sorted(range(len(texel_counts)), key=lambda i: (-texel_counts[i], i)):
        g = min(range(parts), key=lambda r: (load[r], r))
        owner[g].append(i)
        load[g] += texel_counts[i]
    return [sorted(o) for o in owner]


def _cuda_encode(rows: np.ndarray, option: encode_option) -> np.ndarray:
    from . import encode_astc_host
    return encode_astc_host(rows, option)


def encode_band(rgba: np.ndarray, option: encode_option, rank: int, world: int,
                encode_fn: Optional[Callable[[np.ndarray, encode_option], np.ndarray]] = None):
    """Encode this rank's band of a host image (H, W, 4).  Returns (Band, blocks (n, 16)).
    `encode_fn` defaults to the CUDA path (astc_b200_encode_host on the current device)."""
    h, w = rgba.shape[:2]
    b = band_plan(w, h, option, world)[rank]
    if b.rows == 0:
        return b, np.zeros((0, BLOCK_BYTES), np.uint8)
    blocks = (encode_fn or _cuda_encode)(rgba[b.y0:b.y0 + b.rows], option)
    assert blocks.size == b.nbytes, (blocks.size, b.nbytes)
    return b, blocks.reshape(-1, BLOCK_BYTES)


def write_astc_sharded(path: str, width: int, height: int, option: encode_option, b: Band, blocks: np.ndarray,
                       rank: int) -> None:
    """Every rank writes its slice of ONE .astc file at its byte offset; rank 0 also writes the 16-byte
    header (astc_save.h:52-76).  Safe in any order and without a barrier between the writers: the file
    is created without truncation, every rank sets the same final size, and the ranges are disjoint
    (astc_b200_save_astc_slice).  Readers barrier after the last writer, as for any shared file."""
    from . import save_astc_slice
    d = block_dim(option)
    save_astc_slice(path, d, d, width, height, b.byte_offset, np.ascontiguousarray(blocks, dtype=np.uint8).reshape(-1),
                    write_header=(rank == 0))


def gather_blocks(local: np.ndarray, width: int, height: int, option: encode_option, dst: int = 0):
    """Optional: collect every rank's slice on `dst` (gloo or nccl).  Not on the data path."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(), dist.get_rank()
    plan = band_plan(width, height, option, world)
    longest = max(p.nbytes for p in plan)
    backend = dist.get_backend()
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros(longest, dtype=torch.uint8, device=dev)
    flat = torch.from_numpy(np.ascontiguousarray(local, dtype=np.uint8).reshape(-1))
    buf[: flat.numel()] = flat.to(dev)
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    if rank != dst:
        return None
    out = np.empty(output_size(width, height, option), np.uint8)
    for p, t in zip(plan, parts):
        out[p.byte_offset:p.byte_offset + p.nbytes] = t[: p.nbytes].cpu().numpy()
    return out.reshape(-1, BLOCK_BYTES)


def _parse_cpulist(text: str) -> set:
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int) -> dict:
    """One process per GPU: restrict this process to the CPUs of the GPU's NUMA node, so that the
    pinned staging buffers it allocates afterwards are placed (first touch) in the memory next to
    the GPU's PCIe root port.  With eight ranks each streaming > 1 GiB per call through
    astc_b200_encode_host, buffers on the wrong socket cross the inter-socket link and the
    end-to-end rate collapses.  Returns what was done; never raises (a container may forbid it)."""
    import os
    info = {"bound": False}
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{getattr(p, 'pci_domain_id', 0):04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        info["pci"] = bdf
        base = f"/sys/bus/pci/devices/{bdf}"
        with open(f"{base}/local_cpulist") as f:
            local = _parse_cpulist(f.read())
        try:
            with open(f"{base}/numa_node") as f:
                info["numa_node"] = int(f.read())
        except OSError:
            pass
        allowed = os.sched_getaffinity(0) & local
        if allowed and allowed != os.sched_getaffinity(0):
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
        info["cpus"] = len(os.sched_getaffinity(0))
    except Exception as e:                                   # noqa: BLE001
        info["error"] = repr(e)
    return info

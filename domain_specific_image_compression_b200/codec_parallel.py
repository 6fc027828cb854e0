"""Patch-parallel compress / decompress across ranks: no data-path collective.

Patches are independent objects: rank r of W takes patches r, r+W, r+2W, ... of the batch, runs the ordinary
`CompressionModel.compress` / `.decompress` on its shard, and the per-patch results (byte strings, supports) are gathered on
the HOST (`all_gather_object`) and interleaved back into patch order.  The merged dict has exactly the layout of the
single-process `compress()` (custom_compress, eval_selfcontained_entropy.py:68-74), so it can be handed to either
`decompress()` or `decompress_sharded()`.

Note on reproducibility: the y tables are rebuilt at decode time from sigma/nu predicted by the hyper-synthesis convolutions
(eval_selfcontained_entropy.py:99-114), so encoder and decoder must evaluate those convolutions identically.  cuDNN may pick
different algorithms for different batch sizes; decoding with the same shard layout (same W) — or the same batch as the
encoder — keeps the two sides on the same arithmetic.  This is a property of the reference's float-derived tables, not of
the sharding.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

_PER_PATCH_KEYS = ("strings", "min_y", "max_y", "min_z", "max_z")


def patch_indices(n_patches: int, rank: int, world: int) -> List[int]:
    """Indices of the patches rank `rank` of `world` owns: r, r+W, r+2W, ..."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_patches, world))


def split_compressed(compressed: Dict, idx: Sequence[int]) -> Dict:
    """The sub-dict of a compress() result holding only the patches `idx` (batch dim of shape_y/shape_z updated)."""
    out = {k: [compressed[k][i] for i in idx] for k in _PER_PATCH_KEYS}
    out["shape_y"] = [len(idx)] + list(compressed["shape_y"][1:])
    out["shape_z"] = [len(idx)] + list(compressed["shape_z"][1:])
    return out


def merge_compressed(parts: Sequence[Tuple[Sequence[int], Optional[Dict]]], n_patches: int) -> Dict:
    """Interleave per-rank compress() results back into patch order.  parts = [(indices, dict or None if no patches)]."""
    slots: Dict[str, list] = {k: [None] * n_patches for k in _PER_PATCH_KEYS}
    shape_y = shape_z = None
    for idx, part in parts:
        if not idx:
            continue
        if part is None or len(part["strings"]) != len(idx):
            raise ValueError("a rank returned a result that does not match its patch list")
        for k in _PER_PATCH_KEYS:
            for j, i in enumerate(idx):
                if slots[k][i] is not None:
                    raise ValueError(f"patch {i} was produced twice")
                slots[k][i] = part[k][j]
        sy, sz = list(part["shape_y"][1:]), list(part["shape_z"][1:])
        if shape_y is not None and (sy, sz) != (shape_y, shape_z):
            raise ValueError("ranks disagree on the latent shape")
        shape_y, shape_z = sy, sz
    missing = [i for i, s in enumerate(slots["strings"]) if s is None]
    if missing:
        raise ValueError(f"patches {missing} were not produced by any rank")
    out = dict(slots)
    out["shape_y"] = [n_patches] + shape_y
    out["shape_z"] = [n_patches] + shape_z
    return out


def _rank_world(group, rank, world) -> Tuple[int, int]:
    if rank is not None and world is not None:
        return rank, world
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def _gather_objects(obj, group, world):
    if world == 1:
        return [obj]
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)       # host-side gather of small Python objects; no device collective
    return out


def compress_sharded(model, x: torch.Tensor, tail: int = 10, coder: str = "gpu", group=None, gather: bool = True,
                     rank: Optional[int] = None, world: Optional[int] = None):
    """Every rank passes the same full batch x [B,3,H,W]; rank r codes patches r, r+W, ...
    gather=True: returns the merged dict (identical layout to model.compress(x)) on every rank.
    gather=False: returns (indices, local dict or None) without any communication."""
    r, w = _rank_world(group, rank, world)
    B = x.size(0)
    idx = patch_indices(B, r, w)
    local = model.compress(x[idx], tail=tail, coder=coder) if idx else None
    if not gather:
        return idx, local
    parts = _gather_objects((idx, local), group, w)
    return merge_compressed(parts, B)


def gather_order(n_patches: int, world: int) -> List[int]:
    """Where global patch i sits in the concatenation of the ranks' equal-sized blocks (each padded to ceil(B / W) patches):
    rank r's j-th patch is global patch r + j * W, i.e. patch i is entry i // W of block i % W."""
    per = (n_patches + world - 1) // world
    return [(i % world) * per + i // world for i in range(n_patches)]


def _is_nccl(group) -> bool:
    import torch.distributed as dist
    try:
        return dist.is_available() and dist.is_initialized() and dist.get_backend(group) == "nccl"
    except Exception:
        return False


def _gather_images_nccl(local: Optional[torch.Tensor], idx: Sequence[int], B: int, world: int, group, device) -> Optional[torch.Tensor]:
    """Decoded images are the one thing of this path that is big (402 MB for 128 patches of 512^2): when every rank wants the
    whole batch that IS an exchange step, so it goes over NVLink as one all_gather of equal-sized device blocks instead of being
    pickled through the host (r02m: 1146 ms for 128 patches with all_gather_object).  Returns the batch in patch order on the
    device.  Only called over an NCCL group (other backends keep the object gather)."""
    import torch.distributed as dist
    per = (B + world - 1) // world                                  # patches of rank 0 (the longest list)
    meta = [None] * world
    dist.all_gather_object(meta, None if local is None else tuple(local.shape[1:]), group=group)     # a few bytes
    shape = next((m for m in meta if m is not None), None)
    if shape is None:
        raise ValueError("no rank decoded any patch")
    block = torch.zeros((per,) + tuple(shape), dtype=torch.float32, device=device)
    if local is not None:
        block[: local.shape[0]] = local
    allb = torch.empty((world * per,) + tuple(shape), dtype=torch.float32, device=device)
    dist.all_gather_into_tensor(allb, block, group=group)
    order = torch.tensor(gather_order(B, world), dtype=torch.long, device=device)
    return allb.index_select(0, order)


def decompress_sharded(model, compressed: Dict, coder: str = "gpu", group=None, gather: bool = True,
                       rank: Optional[int] = None, world: Optional[int] = None, device_result: bool = False):
    """Rank r decodes patches r, r+W, ... of a (merged) compress() result.
    gather=True: returns x_hat [B,3,H,W] in patch order on every rank — a HOST tensor (device_result=False, the default) or the
    device tensor itself (device_result=True).  Over an NCCL group the images travel as one device all_gather (NVLink) and are
    copied to the host once; other backends (gloo: the CPU tests) gather pickled objects.
    gather=False: returns (indices, local x_hat on the model's device or None) without any communication."""
    r, w = _rank_world(group, rank, world)
    B = len(compressed["strings"])
    idx = patch_indices(B, r, w)
    local = model.decompress(split_compressed(compressed, idx), coder=coder) if idx else None
    if not gather:
        return idx, local
    if w > 1 and rank is None and _is_nccl(group):
        dev = local.device if local is not None else torch.device("cuda", torch.cuda.current_device())   # an idle rank still takes part
        full = _gather_images_nccl(local, idx, B, w, group, dev)
        if full is not None:
            return full if device_result else full.cpu()
    parts = _gather_objects((idx, None if local is None else local.cpu()), group, w)
    ref = next(p for _, p in parts if p is not None)
    out = torch.empty((B,) + tuple(ref.shape[1:]), dtype=ref.dtype)
    seen = [False] * B
    for ids, part in parts:
        for j, i in enumerate(ids):
            out[i] = part[j]
            seen[i] = True
    if not all(seen):
        raise ValueError("some patches were not decoded by any rank")
    return out

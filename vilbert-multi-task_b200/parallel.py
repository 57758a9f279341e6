"""Multi-GPU layer: one process per GPU, batch sharding with no data-path collective, and the single
all-gather that caption-image retrieval needs.

The reference is single-process / single-GPU (`model.cuda(0)`, worker.py:536; `distributed=False`,
worker.py:481) -- there is nothing to mirror, only the sharding the north_star asks for:

* every pair's forward is independent, except that NLVR2 consumes adjacent samples as one pair
  (`pooled.view(-1, 2048)` [UPSTREAM]; pairing visible at worker.py:266-276) -> shard on even boundaries;
* retrieval (task 7, worker.py:278-284, 359) needs the whole candidate row before its softmax/sort ->
  rank r scores a contiguous block of captions against ALL images, then ONE all-gather of the fp32
  score blocks gives every rank the full [n_captions, n_images] matrix.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int, pair_aligned: bool = False) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of `n` samples for `rank`; sizes differ by at most one unit
    (a unit is 2 samples when `pair_aligned`, so an NLVR2 pair is never split across ranks)."""
    unit = 2 if pair_aligned else 1
    if pair_aligned and n % 2 != 0:
        raise ValueError("pair-aligned sharding needs an even number of samples")
    units = n // unit
    base, extra = divmod(units, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo * unit, hi * unit


def shard_batch(tensors: Sequence[torch.Tensor], rank: int, world: int, pair_aligned: bool = False):
    lo, hi = shard_range(tensors[0].shape[0], rank, world, pair_aligned)
    return [t[lo:hi] for t in tensors]


def all_gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """Concatenate per-rank row blocks (block sizes from `shard_range`) into [n_total, ...] on every rank with ONE
    collective: blocks are padded to the largest block, gathered, and trimmed."""
    if not (dist.is_available() and dist.is_initialized()):
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert local.shape[0] == sizes[rank], (local.shape, sizes, rank)
    m = max(sizes)
    pad = local.new_zeros((m,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out = local.new_empty((world * m,) + tuple(local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous(), group=group)
    return torch.cat([out[r * m: r * m + sizes[r]] for r in range(world)], dim=0)


def retrieval_scores(score_pairs: Callable[[int, torch.Tensor], torch.Tensor], n_captions: int, n_images: int,
                     image_chunk: int = 64, group=None) -> torch.Tensor:
    """Full [n_captions, n_images] score matrix on every rank.

    `score_pairs(c, image_idx)` returns the vil_logit of caption `c` paired with each image in `image_idx`
    (the worker does exactly this for task 7: the text is repeated per image, worker.py:278-284, and
    `vil_logit` is read, worker.py:359).  Captions are sharded contiguously over ranks, images are walked in
    chunks of `image_chunk` pairs per forward; one all-gather at the end.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_range(n_captions, rank, world)
    rows: List[torch.Tensor] = []
    for c in range(lo, hi):
        parts = []
        for s in range(0, n_images, image_chunk):
            idx = torch.arange(s, min(n_images, s + image_chunk))
            parts.append(score_pairs(c, idx).reshape(-1).float())
        rows.append(torch.cat(parts))
    device = rows[0].device if rows else torch.device("cpu")
    local = torch.stack(rows) if rows else torch.empty(0, n_images, device=device)
    return all_gather_rows(local, n_captions, group)


def make_pair_scorer(model, captions, images):
    """Glue for `retrieval_scores`: `captions` = (question[n_cap,T], segment_ids, input_mask), `images` =
    (features[n_img,V,F], spatials[n_img,V,5], image_mask[n_img,V]) resident on the model's device."""
    from . import _lib as L
    q, seg, im = captions
    f, s, vm = images

    def score(c: int, idx: torch.Tensor) -> torch.Tensor:
        n = idx.numel()
        idx = idx.to(f.device)
        task = torch.full((n, 1), 7, dtype=torch.long, device=f.device)
        out = model(q[c:c + 1].repeat(n, 1), f[idx], s[idx], seg[c:c + 1].repeat(n, 1), im[c:c + 1].repeat(n, 1),
                    vm[idx], None, task, select=L.OUT_VIL_LOGIT)
        return out[2]
    return score

"""Folder batches: page-wise data parallelism over the GPUs of one box.

The reference's directory mode is a serial `for` loop over the images of a folder
(ocr_agent/transcribe.py:185-210) that reloads the model after every page (nodes.py:127).  Pages are
independent units -- all reads of one page, their agreement / merge and their CER stay on one rank --
so the path shards with NO data-path collective: page i goes to rank i mod world, every rank keeps a
full weight replica (16.6 GB of 180 GB), and one final gather of the (variable-length) per-page
results closes the job.  `torch.distributed` is used for that gather only (NCCL backend on the GPU
box; `gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Callable, Sequence


def shard_pages(n_pages: int, rank: int, world: int) -> list:
    """Indices of the pages rank `rank` reads: i mod world == rank (round-robin keeps ranks balanced
    to within one page for any folder size)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return list(range(rank, n_pages, world))


def gather_results(local: dict, n_pages: int, group=None) -> list | None:
    """Final gather: every rank contributes {page index: result}; rank 0 returns the list of all
    n_pages results in page order (other ranks return None).  Results are arbitrary picklable
    objects (texts, token-id lists, metric dicts), i.e. variable length."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        merged = dict(local)
    else:
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        bucket = [None] * world if rank == 0 else None
        dist.gather_object(local, bucket, dst=0, group=group)
        if rank != 0:
            return None
        merged = {}
        for part in bucket:
            merged.update(part)
    missing = [i for i in range(n_pages) if i not in merged]
    if missing:
        raise RuntimeError(f"folder gather: pages {missing[:8]}... were read by no rank")
    return [merged[i] for i in range(n_pages)]


def read_folder(pages: Sequence, read_page: Callable, *, rank: int = 0, world: int = 1, pages_per_batch: int = 1,
                group=None):
    """Shard `pages` page-wise, call `read_page(list_of_pages) -> list_of_results` on this rank's share
    in batches of `pages_per_batch`, then gather.  `read_page` is the initial_ocr phase of a batch of
    pages (preprocess all strategies, one batched read, agreement + merge); see bench.py."""
    mine = shard_pages(len(pages), rank, world)
    local = {}
    for i in range(0, len(mine), pages_per_batch):
        idx = mine[i:i + pages_per_batch]
        out = read_page([pages[j] for j in idx])
        if len(out) != len(idx):
            raise RuntimeError("read_page must return one result per page")
        local.update(dict(zip(idx, out)))
    return gather_results(local, len(pages), group)

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

import json
import sys
from pathlib import Path
from typing import Callable, Sequence

IMAGE_EXTENSIONS = {".png", ".jpg", ".jpeg", ".bmp", ".tiff", ".tif", ".webp"}     # transcribe.py:18


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


# ───────────────────────── f1: the folder-batch driver (transcribe.py:185-210) ─────────────────────────
def list_images(input_dir) -> list:
    """transcribe.py:193-195: the images of a folder, sorted."""
    return sorted(f for f in Path(input_dir).iterdir() if f.suffix.lower() in IMAGE_EXTENSIONS)


def match_ground_truth(stem: str, ground_truth_dir):
    """transcribe.py:201-207 / eval_final.py:107-117: `<stem>.md`, then `<stem>.txt`."""
    if ground_truth_dir is None:
        return None
    for ext in (".md", ".txt"):
        cand = Path(ground_truth_dir) / f"{stem}{ext}"
        if cand.exists():
            return cand
    return None


def initial_ocr_page(image_path: str, strategies=None, agreement_threshold=None, tools=None) -> dict:
    """The read phase of one page exactly as `node_initial_ocr` sequences it (nodes.py:76-134, trace and LLM parts
    left out): reads S0 and S1, `compare_versions`, the tiebreaker read S2 when the agreement is below the threshold,
    `merge_versions`, `unload_ocr_model`.  Returns the node's state keys (candidate dicts as nodes.py:53-58 builds
    them) plus the comparison.  This is the per-page function of the folder driver when the reference's own
    `transcribe_single` (which also runs the LLM critic / editor over Ollama) is not what is wanted."""
    if tools is None:
        from . import tools
    cfg = tools.config
    strategy_list = list(strategies if strategies is not None else cfg.PREPROCESSING_STRATEGIES)
    threshold = cfg.AGREEMENT_THRESHOLD if agreement_threshold is None else agreement_threshold
    candidates, used = [], []

    def one_pass(strategy):
        label = "+".join(strategy) if isinstance(strategy, list) else strategy
        if label in used:
            return
        used.append(label)
        text = tools.run_ocr(tools.preprocess_image(image_path, strategy))
        candidates.append({"text": text, "source": f"ocr_{label}", "ocr_params": {"strategy": label}, "score": None})

    one_pass(strategy_list[0] if strategy_list else "original")
    if len(strategy_list) > 1:
        one_pass(strategy_list[1])
    cmp = None
    if len(candidates) >= 2:
        cmp = tools.compare_versions(candidates[0]["text"], candidates[1]["text"])
        if cmp["agreement_rate"] < threshold and len(strategy_list) > 2:
            one_pass(strategy_list[2])
    best = tools.merge_versions([c["text"] for c in candidates])
    tools.unload_ocr_model()
    return {"candidates": candidates, "current_best": best, "strategies_used": used, "comparison": cmp}


def _reference_transcribe_single():
    try:
        from ocr_agent.transcribe import transcribe_single          # the reference's own per-page entry point
    except Exception as e:                                          # pragma: no cover - no reference tree
        raise RuntimeError("transcribe_folder needs `page_fn` or an importable `ocr_agent.transcribe` "
                           f"(the reference package): {e}")
    return transcribe_single


def transcribe_folder(input_dir, output_dir=None, ground_truth_dir=None, *, pages_per_batch: int = 21,
                      page_fn: Callable | None = None, rank: int = 0, world: int = 1, group=None, tools=None,
                      strategies=None, **page_kwargs):
    """Directory mode of `transcribe.main` (transcribe.py:185-210) with the reads batched across pages.

    The reference loops `transcribe_single(image, output_dir, gt)` over the sorted images, each page reading its
    candidates one by one.  Here this rank's pages (page i -> rank i mod world) are taken `pages_per_batch` at a time:
    `tools.prime` preprocesses every configured strategy of those pages and reads ALL their candidates in one batched
    vision / prefill / paged-KV decode pass; then the per-page function runs page by page and finds its
    `preprocess_image` / `run_ocr` calls answered from the cache.  `page_fn(image_path, output_dir, gt_path, **kw)`
    defaults to the reference's UNMODIFIED `transcribe_single` (install this package as `ocr_agent.tools` first:
    `handwritten_ocr_b200.install()`).  Returns, on rank 0, the per-page results in folder order (others: None)."""
    if tools is None:
        from . import tools
    input_dir = Path(input_dir)
    images = list_images(input_dir)
    if not images:
        raise FileNotFoundError(f"No image files found in {input_dir}")
    output_dir = Path(output_dir) if output_dir is not None else input_dir / "results"     # transcribe.py:167-168
    fn = page_fn or _reference_transcribe_single()
    if pages_per_batch > int(tools._options["cache_pages"]):
        tools.configure(cache_pages=pages_per_batch)

    # host pipeline: while the GPU reads batch k, a thread decodes the image files of batch k + 1 (PIL releases the GIL)
    from concurrent.futures import ThreadPoolExecutor
    mine = [images[i] for i in shard_pages(len(images), rank, world)]
    batches = [mine[i:i + pages_per_batch] for i in range(0, len(mine), pages_per_batch)]
    pool = ThreadPoolExecutor(max_workers=4)
    pending = {}

    def submit(k):
        if 0 <= k < len(batches) and k not in pending:
            pending[k] = {str(p): pool.submit(tools._open_array, str(p)) for p in batches[k]}

    submit(0)
    state = {"k": 0}

    def read_batch(batch_images):
        k = state["k"]
        state["k"] += 1
        submit(k)
        decoded = {p: f.result() for p, f in pending.pop(k).items()}
        submit(k + 1)
        tools.prime([str(p) for p in batch_images], strategies, decoded)
        del decoded
        out = []
        for img in batch_images:
            out.append(fn(img, output_dir, match_ground_truth(img.stem, ground_truth_dir), **page_kwargs))
            tools.forget(str(img), delete_files=True)
        return out

    try:
        return read_folder(images, read_batch, rank=rank, world=world, pages_per_batch=pages_per_batch, group=group)
    finally:
        pool.shutdown(wait=False, cancel_futures=True)


# ───────────────────────── f2: the batch evaluator (eval_final.py:94-134) ─────────────────────────
def eval_folder(results_dir, ground_truth_dir=None, output=None, *, lower: bool = False, tools=None,
                textops=None, verbose: bool = False) -> list:
    """Directory mode of `eval_final.main` (eval_final.py:94-134): every `*_transcription.txt` (else `*.txt`) of the
    folder against the ground-truth file of the same stem.  The reference evaluates the files one after another
    (3 pure-Python Levenshtein DPs each); here the 3 x N distances of ALL files go to the GPU in ONE
    `ocrb_levenshtein_batch` launch.  Returns the list `eval_final.main` collects (and writes to `output` as the same
    JSON): per file the `evaluate()` dict plus `"file"`."""
    if tools is None:
        from . import tools
    if textops is None:
        from . import textops
    results_dir = Path(results_dir).resolve()
    txt_files = sorted(results_dir.glob("*_transcription.txt")) or sorted(results_dir.glob("*.txt"))
    if not txt_files:
        raise FileNotFoundError(f"No .txt files found in {results_dir}")
    texts, gts = [], []
    for txt in txt_files:
        stem = txt.stem[: -len("_transcription")] if txt.stem.endswith("_transcription") else txt.stem
        gt_path = match_ground_truth(stem, ground_truth_dir)
        texts.append(txt.read_text(encoding="utf-8"))
        gts.append(tools.parse_ground_truth(gt_path) if gt_path else None)
    with_gt = [i for i, g in enumerate(gts) if g is not None]
    metrics = textops.tier1_metrics_batch([(gts[i], texts[i]) for i in with_gt], lower) if with_gt else []
    by_index = dict(zip(with_gt, metrics))
    results = []
    for i, txt in enumerate(txt_files):
        r = {}
        if i in by_index:
            r["tier1_raw_vs_gt"] = by_index[i]
        r["file"] = str(txt)
        results.append(r)
    if verbose and metrics:
        print(f"Batch Summary ({len(metrics)} files with GT)")
        print(f"  Avg CER: {sum(m['cer'] for m in metrics) / len(metrics):.2%}")
        print(f"  Avg WER: {sum(m['wer_token'] for m in metrics) / len(metrics):.2%}")
    if output is not None:
        with open(output, "w", encoding="utf-8") as f:
            json.dump(results, f, indent=2, ensure_ascii=False)
    return results

"""Drop-in replacement for the hot functions of the reference's `ocr_agent.tools`
(/root/reference/ocr_agent/tools.py): same names, arguments, return values, prints and error
behaviour, so `ocr_agent/nodes.py` and `ocr_agent/graph.py` run unchanged on top (install with
`handwritten_ocr_b200.install()` before `ocr_agent.nodes` is imported -- INTEGRATION.md).

    preprocess_image(image_path, strategy) -> str          tools.py:633
    run_ocr(image_path, params=None) -> str                tools.py:728
    unload_ocr_model() -> None                             tools.py:714
    compare_versions / merge_versions / evaluate / tier1_metrics / cer / wer / levenshtein /
    normalize_text                                         tools.py:51-139,305-350,411-493
    transcribe(image, strategy) -> str                     = run_ocr(preprocess_image(...))  (new)

Everything numerical runs in libocrb200 on the GPU; there is no CPU fallback.  Because nodes.py
asks for reads one at a time (preprocess S0, read, preprocess S1, read, compare, maybe S2 ...), the
first `preprocess_image` of a page speculatively preprocesses all configured GPU strategies and the
first `run_ocr` of that page reads ALL of them in one batch (one vision pass, one prefill, one
paged-KV decode); later calls for the same page return the cached text.
"""
from __future__ import annotations

import gc
import os
import sys
import tempfile
from collections import OrderedDict
from pathlib import Path

import numpy as np
import torch
from PIL import Image, ImageOps

from . import _lib, preprocess
from .vlm_config import EOS
from .textops import (_find_differing_segments, _levenshtein_words, cer, compare_versions, evaluate,  # noqa: F401
                      levenshtein, merge_versions, normalize_text, tier1_metrics, wer)

try:  # the reference's config module when it is importable (same names are read: tools.py:23)
    from ocr_agent import config  # type: ignore
except Exception:  # pragma: no cover - standalone use
    from . import default_config as config

_LOSSLESS = {".png", ".bmp", ".tif", ".tiff", ".ppm", ".pgm"}

# module-level model cache (tools.py:679-680)
_ocr_engine = None
_options = {
    "speculative": True,        # batch all strategies of a page into one read
    "keep_resident": True,      # unload_ocr_model() keeps the 16.6 GB of weights in the 180 GB of HBM
    "checkpoint": os.environ.get("OCRB_CHECKPOINT"),   # dir with HF safetensors; None -> random init
    "vlm_config": None,         # VLMConfig override (tests use the tiny config)
    "max_batch": 64,            # sequences per batched read (paged KV: 57 KB per token per sequence at 7B)
    "seed": 0,
    "force_greedy": False,      # decode greedily even if the checkpoint's generation_config.json asks for sampling
    "cache_pages": 64,          # originals whose preprocessed variants / texts stay cached (LRU)
    "io_threads": 8,            # host threads decoding pages / writing temp files in prime() (folder mode)
    "delete_evicted_files": True,   # remove the temp files of an evicted page (the reference leaks them: tools.py:670)
}
# Cache of preprocessed pages and their transcriptions, keyed by the ORIGINAL image and guarded by the file's
# (mtime_ns, size) signature: a changed file invalidates its entry.  Bounded LRU over originals (`cache_pages`): the
# reference's callers never say when they are done with a page (nodes.py calls unload_ocr_model() between the
# initial reads and the reocr sweep of the SAME page, so that call must not drop the page), hence eviction by age.
class _Page:
    __slots__ = ("sig", "variants", "texts")

    def __init__(self, sig):
        self.sig = sig
        self.variants = {}          # label -> (processed path, device tensor [1,H,W(,3)])
        self.texts = {}             # label -> transcription


_pages: "OrderedDict[str, _Page]" = OrderedDict()
_processed: dict = {}               # processed path -> (original path, label)


def _signature(path: str):
    st = os.stat(path)
    return (st.st_mtime_ns, st.st_size)


def _drop_page(original: str, delete_files: bool = False) -> None:
    page = _pages.pop(original, None)
    if page is None:
        return
    for path, _ in page.variants.values():
        _processed.pop(path, None)
        if delete_files:
            try:
                os.unlink(path)
            except OSError:
                pass


def _page_for(original: str, create: bool = True):
    """The cache entry of an original image (most recently used last); a stale entry (file changed on disk)
    is dropped first."""
    sig = _signature(original)
    page = _pages.get(original)
    if page is not None and page.sig != sig:
        _drop_page(original, _options["delete_evicted_files"])
        page = None
    if page is None:
        if not create:
            return None
        page = _pages[original] = _Page(sig)
        while len(_pages) > max(1, int(_options["cache_pages"])):
            _drop_page(next(iter(_pages)), _options["delete_evicted_files"])
    else:
        _pages.move_to_end(original)
    return page


def configure(**kw) -> None:
    """Set engine options (see _options) before the first run_ocr."""
    for k, v in kw.items():
        if k not in _options:
            raise KeyError(k)
        _options[k] = v


def _label(steps) -> str:
    return "+".join(s for s in steps if s != "original")


def _steps(strategy) -> list:
    return [strategy] if isinstance(strategy, str) else list(strategy)


def _open_array(image_path: str) -> np.ndarray:
    img = Image.open(image_path)
    arr = np.array(img)
    if arr.dtype != np.uint8 or not (arr.ndim == 2 or (arr.ndim == 3 and arr.shape[2] == 3)):
        raise ValueError(f"{image_path}: mode {img.mode} not supported; pages must be RGB or L "
                         "(the reference does not convert modes either: tools.py:656)")
    return arr


def _temp_name(image_path: str, label: str) -> str:
    suffix = Path(image_path).suffix or ".png"
    tmp = tempfile.NamedTemporaryFile(suffix=suffix, delete=False, prefix=f"ocr_{label}_")      # tools.py:670
    tmp.close()
    return tmp.name


def _write(arr: np.ndarray, path: str) -> None:
    # lossless either way; a stored (level 0) PNG instead of Pillow's default level 6 only makes the temp file larger
    # and the save ~30x faster (the file exists for callers that open it; run_ocr reads the cached page)
    kw = {"compress_level": 0} if Path(path).suffix.lower() == ".png" else {}
    Image.fromarray(arr).save(path, **kw)


def _save(arr: np.ndarray, image_path: str, label: str) -> str:
    path = _temp_name(image_path, label)
    _write(arr, path)
    return path


def _wanted_strategies(steps) -> list:
    wanted = [steps]
    if _options["speculative"]:
        for s in getattr(config, "PREPROCESSING_STRATEGIES", []):
            s = _steps(s)
            if _label(s) and all(_label(s) != _label(w) for w in wanted):
                wanted.append(s)
    return wanted


def _preprocess_page(image_path: str, wanted: list, required_label: str | None, decoded=None, pool=None) -> "_Page":
    """Apply every strategy of `wanted` that the cache does not hold yet.  A failure in a SPECULATIVE strategy
    (one the caller did not ask for) is reported and skipped: the reference would never have run it here.
    `decoded`: the already decoded page array (folder mode decodes the next batch on a host thread); `pool`: a thread
    pool the temp files are written on (they are complete when the pool is shut down, i.e. before prime() returns)."""
    page = _page_for(image_path)
    todo = [s for s in wanted if _label(s) not in page.variants]
    if not todo:
        return page
    x = preprocess.to_device(decoded if decoded is not None else _open_array(image_path))
    for s in todo:
        lab = _label(s)
        try:
            y = preprocess.apply_strategy(x, s)
            if pool is None:
                path = _save(y[0].cpu().numpy(), image_path, lab)
            else:
                path = _temp_name(image_path, lab)
                pool.submit(_write, y[0].cpu().numpy(), path)
        except Exception as e:
            if lab == required_label:
                raise
            print(f"  [preprocess] speculative {lab} skipped: {e}", file=sys.stderr)
            continue
        page.variants[lab] = (path, y)
        _processed[path] = (image_path, lab)
    return page


def preprocess_image(image_path: str, strategy) -> str:
    """tools.py:633-673: apply the strategy on the GPU, save a temp file with the input's suffix,
    return its path ("original" / [] return the input path untouched)."""
    steps = _steps(strategy)
    if steps == ["original"] or not steps:
        return image_path
    label = _label(steps)
    print(f"  [preprocess] Applying {label}...")
    page = _preprocess_page(image_path, _wanted_strategies(steps), label)
    path, y = page.variants[label]
    if not os.path.exists(path):
        # somebody removed the temp file: write it again from the cached page
        Image.fromarray(y[0].cpu().numpy()).save(path)
    return path


def _load_ocr_model():
    """tools.py:683-711.  The checkpoint named by config.OLMOCR_MODEL cannot be downloaded offline:
    weights come from a local safetensors directory (OCRB_CHECKPOINT) or are random-initialised at
    the configured dimensions."""
    global _ocr_engine
    if _ocr_engine is not None:
        return _ocr_engine
    if not torch.cuda.is_available():
        raise _lib.OcrbError("run_ocr needs a CUDA device: handwritten-ocr_b200 has no CPU fallback")
    from .engine import OcrEngine
    from .vlm import VLMWeights
    from .vlm_config import VLMConfig
    cfg, tokenizer, gen = checkpoint_metadata(_options["checkpoint"], _options["vlm_config"], _options["force_greedy"])
    dev = torch.device("cuda", torch.cuda.current_device())
    print(f"  [ocr] Loading {getattr(config, 'OLMOCR_MODEL', cfg.name)} on cuda...")
    if _options["checkpoint"]:
        from safetensors.torch import load_file
        sd = {}
        for f in sorted(Path(_options["checkpoint"]).glob("*.safetensors")):
            sd.update(load_file(str(f), device=str(dev)))
        w = VLMWeights.from_state_dict(cfg, sd, free_source=True)
    else:
        w = VLMWeights.random(cfg, dev, seed=_options["seed"])
    _ocr_engine = OcrEngine(w, max_batch=_options["max_batch"], tokenizer=tokenizer,
                            max_new_tokens=int(getattr(config, "OCR_MAX_NEW_TOKENS", 2048)),
                            min_pixels=int(getattr(config, "OCR_MIN_PIXELS", 256 * 256)),
                            max_pixels=int(getattr(config, "OCR_MAX_PIXELS", 1024 * 1024)))
    _ocr_engine.extra_eos = tuple(e for e in gen["eos_token_ids"] if e != EOS)
    print("  [ocr] Model loaded.")
    return _ocr_engine


def checkpoint_metadata(path, vlm_config=None, force_greedy: bool = False):
    """Everything `AutoProcessor.from_pretrained` / `from_pretrained` (tools.py:700-709) read next to the weights, from a
    LOCAL directory (the hub is not reachable): dimensions from `config.json`, the tokenizer from the HF tokenizer files,
    the EOS ids from `generation_config.json` (sampling / beams / repetition penalty are refused unless force_greedy).
    Without a checkpoint: the configured (or 7B-class) dimensions, the synthetic tokenizer, <|im_end|> as EOS.
    Host-only (no CUDA needed)."""
    from .vlm_config import HFTokenizer, VLMConfig, greedy_generation_params
    if not path:
        return vlm_config or VLMConfig.olmocr_7b(), None, {"eos_token_ids": [EOS]}
    path = str(path)
    if not os.path.isdir(path):
        raise FileNotFoundError(f"checkpoint directory {path} does not exist")
    cfg = vlm_config or (VLMConfig.from_pretrained_dir(path) if os.path.exists(os.path.join(path, "config.json"))
                         else VLMConfig.olmocr_7b())
    tokenizer = None
    if any(os.path.exists(os.path.join(path, f)) for f in ("tokenizer.json", "vocab.json", "tokenizer_config.json")):
        tokenizer = HFTokenizer(path)
    try:
        gen = greedy_generation_params(path)
    except NotImplementedError:
        if not force_greedy:
            raise
        gen = {"eos_token_ids": [EOS]}
    if EOS not in gen["eos_token_ids"]:
        raise ValueError(f"generation_config.json does not list <|im_end|> ({EOS}) as an EOS id: {gen['eos_token_ids']}")
    return cfg, tokenizer, gen


def unload_ocr_model():
    """tools.py:714-725.  On a 180 GB B200 the weights stay resident unless keep_resident=False."""
    global _ocr_engine
    if not _options["keep_resident"]:
        _ocr_engine = None
        gc.collect()
        if torch.cuda.is_available():
            torch.cuda.empty_cache()
    print("  [ocr] Model unloaded, memory freed.")


def _load_page_for_model(image_path: str) -> torch.Tensor:
    """HF load_image (image_utils.py:462-501): open, exif_transpose, convert("RGB")."""
    img = ImageOps.exif_transpose(Image.open(image_path)).convert("RGB")
    return preprocess.to_device(np.array(img))


def _read_pending(engine, originals: list, prompt: str, max_new_tokens: int, first=None) -> None:
    """One batched read (per page shape) of every cached variant of `originals` that has no text yet -- at most
    engine.max_batch sequences per read, `first` = (original, label) goes into the first batch."""
    todo = []
    for orig in originals:
        page = _pages.get(orig)
        if page is None:
            continue
        todo += [(orig, lab) for lab in page.variants if lab not in page.texts]
    if first is not None and first in todo:
        todo.remove(first)
        todo.insert(0, first)
    groups: dict = {}
    for orig, lab in todo:
        groups.setdefault(tuple(_pages[orig].variants[lab][1].shape), []).append((orig, lab))
    for same in groups.values():
        for i0 in range(0, len(same), engine.max_batch):
            part = same[i0:i0 + engine.max_batch]
            batch = torch.cat([_pages[o].variants[lab][1] for o, lab in part], 0)
            toks = engine.read_batch(batch, prompt=prompt, max_new_tokens=max_new_tokens)
            for (o, lab), tk in zip(part, toks):
                _pages[o].texts[lab] = engine.detokenize(tk)


def prime(image_paths, strategies=None, decoded: dict | None = None) -> None:
    """Folder mode (transcribe.py:193-209 runs the pages one after another): preprocess the configured strategies of
    SEVERAL pages and read all their candidates in one batch, so that the per-page `preprocess_image` / `run_ocr`
    calls the unmodified `transcribe_single` makes afterwards are cache hits.  `cache_pages` must cover the pages.
    `decoded`: {path: uint8 array} of pages somebody already decoded (see folder.transcribe_folder)."""
    engine = _load_ocr_model()
    strategies = [_steps(s) for s in (strategies if strategies is not None else
                                      getattr(config, "PREPROCESSING_STRATEGIES", []))]
    strategies = [s for s in strategies if _label(s)]
    paths = [str(p) for p in image_paths]
    if len(paths) > int(_options["cache_pages"]):
        raise ValueError(f"prime(): {len(paths)} pages exceed cache_pages={_options['cache_pages']}")
    import contextlib
    import io
    from concurrent.futures import ThreadPoolExecutor
    # host pipeline: decode the image files and write the temp files on threads (PIL releases the GIL in both), the GPU
    # preprocessing of page i runs while pages i+1.. are still being decoded and page i-1's temp files are being written
    decoded = dict(decoded or {})
    with ThreadPoolExecutor(max_workers=int(_options["io_threads"])) as pool:
        arrays = {p: pool.submit(_open_array, p) for p in paths
                  if p not in decoded and any(_label(s) not in (_pages.get(p).variants if p in _pages else {}) for s in strategies)}
        for p in paths:
            arr = decoded.get(p)
            if arr is None and p in arrays:
                arr = arrays.pop(p).result()
            _preprocess_page(p, strategies, None, arr, pool)
    with contextlib.redirect_stdout(io.StringIO()):
        _read_pending(engine, paths, config.OCR_PROMPT, int(config.OCR_MAX_NEW_TOKENS))


def run_ocr(image_path: str, params: dict | None = None) -> str:
    """tools.py:728-771: greedy transcription of the (already preprocessed) image."""
    print(f"  [ocr] Running OCR on {Path(image_path).name}...")
    params = params or {}
    engine = _load_ocr_model()
    prompt = params.get("prompt", config.OCR_PROMPT)
    max_new_tokens = int(params.get("max_new_tokens", config.OCR_MAX_NEW_TOKENS))
    default_call = "prompt" not in params and "max_new_tokens" not in params
    entry = _processed.get(image_path)
    page = _pages.get(entry[0]) if entry is not None else None
    lossless = Path(image_path).suffix.lower() in _LOSSLESS
    if page is not None and lossless and default_call:
        original, label = entry
        if label not in page.texts:
            # one batched read for every preprocessed variant of this page that has no text yet
            _read_pending(engine, [original], prompt, max_new_tokens, first=(original, label))
        result = page.texts[label]
    else:
        x = page.variants[entry[1]][1] if (page is not None and lossless) else _load_page_for_model(image_path)
        toks = engine.read_batch(x, prompt=prompt, max_new_tokens=max_new_tokens)
        result = engine.detokenize(toks[0])
    print(f"  [ocr] Done ({len(result)} chars)")
    return result


def transcribe(image: str, strategy) -> str:
    """The read the pipeline's initial_ocr / reocr nodes perform (nodes.py:41,52) as one call."""
    return run_ocr(preprocess_image(image, strategy))


def forget(image_path: str | None = None, delete_files: bool = False) -> None:
    """Drop cached preprocessed pages / texts (all, or those of one original).  The temp files stay unless asked
    (a path handed out earlier can still be given to run_ocr: it is then read from the file)."""
    for k in ([image_path] if image_path else list(_pages)):
        _drop_page(k, delete_files)


# non-hot names the reference's callers import from tools (agents.py:12, transcribe.py:33,
# eval_final.py:22): re-exported from the reference when it is importable.
try:
    from ocr_agent.tools import call_llm, call_llm_json, parse_ground_truth, parse_json_response  # type: ignore  # noqa: F401
except Exception:  # pragma: no cover
    def call_llm_json(*a, **k):
        raise RuntimeError("call_llm_json is the reference's Ollama client (tools.py:246-299), out of scope "
                           "here; install `ollama` so ocr_agent.tools imports, or supply your own")

    call_llm = call_llm_json

    def parse_ground_truth(file_path):
        p = Path(file_path)
        if not p.exists():
            return None
        raw = p.read_text(encoding="utf-8")
        i = raw.find("## Ground Truth")
        text = raw.strip() if i == -1 else raw[i + len("## Ground Truth"):].strip()
        return text or None

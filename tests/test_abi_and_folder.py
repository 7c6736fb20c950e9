"""CPU checks of the boundary: the C-ABI library exports every symbol include/ocrb200.h declares (no
compute calls), the ctypes table matches the header, and the page-wise folder sharding + final gather
works across two `gloo` ranks."""
import ctypes
import os
import re
import socket

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "ocrb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ocrb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(pkg):
    from handwritten_ocr_b200 import _lib
    assert os.path.exists(_lib.LIB_PATH), "build the library first: python -c 'import __graft_entry__ as g; g.build()'"
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 25
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, f"declared in ocrb200.h but not exported: {missing}"
    assert sorted(_lib.EXPORTS) == syms, "ctypes table (_lib.py) and header disagree"


def test_host_only_entry_points(pkg):
    """Entry points that do no device work can be called without a GPU."""
    from handwritten_ocr_b200 import _lib, preprocess
    L = _lib.load()
    assert L.ocrb_version() >= 1
    assert L.ocrb_skinny_workspace_bytes() > 0
    assert preprocess.smart_resize(768, 1024) == (756, 1036)       # HF smart_resize
    assert preprocess.smart_resize(1024, 768) == (1036, 756)
    with pytest.raises(_lib.OcrbError):
        _lib.call("ocrb_skinny_gemm_bf16", None, 0, None, 0, None, 0, 1, 8, 8, None, None, 0, 0, None, 0.0, None, None)
    assert b"null pointer" in L.ocrb_last_error()


def test_product_code_never_imports_the_oracle():
    pkg_dir = os.path.join(ROOT, "handwritten-ocr_b200")
    for fn in os.listdir(pkg_dir):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg_dir, fn)).read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f"{fn} imports the oracle"


def test_shard_pages_partition():
    from handwritten_ocr_b200.folder import shard_pages
    for n in (0, 1, 7, 256):
        for world in (1, 2, 4, 8):
            parts = [shard_pages(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_pages, out_q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import handwritten_ocr_b200  # noqa: F401
    from handwritten_ocr_b200.folder import read_folder
    pages = [f"page_{i:03d}" for i in range(n_pages)]
    seen = []

    def read_page(batch):
        seen.extend(batch)
        return [{"page": p, "rank": rank, "text": p.upper() * (1 + int(p[-1]))} for p in batch]   # variable length

    res = read_folder(pages, read_page, rank=rank, world=world, pages_per_batch=3)
    out_q.put((rank, seen, res))
    dist.barrier()
    dist.destroy_process_group()


def test_folder_batch_two_ranks_gloo():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_pages, world = 11, 2
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_pages, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got.sort()
    (r0, seen0, res0), (r1, seen1, res1) = got
    assert seen0 == [f"page_{i:03d}" for i in range(0, n_pages, 2)] and seen1 == [f"page_{i:03d}" for i in range(1, n_pages, 2)]
    assert res1 is None and [r["page"] for r in res0] == [f"page_{i:03d}" for i in range(n_pages)]
    assert [r["rank"] for r in res0] == [i % 2 for i in range(n_pages)]

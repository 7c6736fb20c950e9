"""Timing of the text kernels at BASELINE sizes (2k / 8k chars; 350 / 1400 words)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import handwritten_ocr_b200
from handwritten_ocr_b200 import textops as tx, synth
for words in (350, 512, 1400):
    a = synth.text(1, words); b = synth.corrupt(a, 1, 0.06); c = synth.corrupt(a, 2, 0.08)
    ca, cb = tx._codes(tx.normalize_text(a)), tx._codes(tx.normalize_text(b))
    tx.levenshtein_ids_batch([(ca, cb)]); torch.cuda.synchronize()
    t0 = time.perf_counter(); reps = 5
    for _ in range(reps): d = tx.levenshtein_ids_batch([(ca, cb)])
    dt = (time.perf_counter() - t0) / reps
    cells = len(ca) * len(cb)
    t0 = time.perf_counter(); cv = tx.compare_versions(a, b); t_cmp = time.perf_counter() - t0
    t0 = time.perf_counter(); mv = tx.merge_versions([a, b, c]); t_mrg = time.perf_counter() - t0
    t0 = time.perf_counter(); ev = tx.tier1_metrics(a, b); t_ev = time.perf_counter() - t0
    print(f"{words} words, {len(ca)}x{len(cb)} chars: levenshtein {dt*1e3:.2f} ms incl. H2D/D2H ({cells/dt/1e9:.2f} Gcell/s, "
          f"{(4*(len(ca)+len(cb))+4)/dt/1e9:.4f} GB/s algorithmic); compare_versions {t_cmp*1e3:.1f} ms, merge_versions(3) {t_mrg*1e3:.1f} ms, "
          f"tier1_metrics {t_ev*1e3:.1f} ms")

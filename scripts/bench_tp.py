"""BASELINE configs[4]: the large (72B-class) VLM tensor-parallel across the GPUs of one box, 2048-token
transcriptions.  Launch: python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1
--master-port P scripts/bench_tp.py [--new-tokens 2048] [--batch 3].  Rank 0 prints one JSON line."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

ap = argparse.ArgumentParser()
ap.add_argument("--new-tokens", type=int, default=2048)
ap.add_argument("--batch", type=int, default=3)
ap.add_argument("--reps", type=int, default=1)
ap.add_argument("--tiny", action="store_true")
args = ap.parse_args()
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
import handwritten_ocr_b200
from handwritten_ocr_b200 import _lib, engine, preprocess, synth, tp
from handwritten_ocr_b200.vlm_config import VLMConfig
cfg = VLMConfig.tiny() if args.tiny else VLMConfig.qwen72b()
t0 = time.time()
w, lcfg = tp.random_weights_tp(cfg, dev, rank, world, seed=0)
comm = tp.TPComm()
comm.enable_peer_all_reduce(dev, cfg.text.hidden)
eng = engine.OcrEngine(w, max_batch=args.batch, max_new_tokens=args.new_tokens, max_prompt=1600, tp=comm)
pages = preprocess.to_device([synth.page(i)[:, :, 1].copy() for i in range(args.batch)])   # gray candidates of B pages
torch.cuda.synchronize(); dist.barrier()
t_init = time.time() - t0
eng.read_batch(pages, max_new_tokens=min(args.new_tokens, 16))      # warm-up (captures the decode graph per max_new)
for _ in range(args.reps):
    dist.barrier(); torch.cuda.synchronize()
    t1 = time.perf_counter()
    toks = eng.read_batch(pages, max_new_tokens=args.new_tokens)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t1
tm = eng.timings
steps = tm["steps"] - 1
step_ms = tm["decode_ms"] / max(steps, 1)
wbytes = w.decode_weight_bytes()
kvb = lcfg.text.layers * 2 * lcfg.text.kv_heads * lcfg.text.head_dim * 2
avg_ctx = tm["prompt_len"] + steps / 2
alg = wbytes + args.batch * avg_ctx * kvb
peak = 6550.7
t = torch.tensor([step_ms, wall], device=dev, dtype=torch.float64)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "configs[4]: %s tensor-parallel tp%d, B=%d sequences, %d new tokens" % (cfg.name, world, args.batch, args.new_tokens),
                      "decode_step_ms": round(float(t[0]), 4), "decode_tok_per_s": round(args.batch * 1e3 / float(t[0]), 1),
                      "read_wall_s": round(float(t[1]), 3), "vision_ms": round(tm["vision_ms"], 1), "prefill_ms": round(tm["prefill_ms"], 1),
                      "weight_bytes_per_rank_per_step": wbytes, "hbm_gbs_per_rank": round(alg / (float(t[0]) * 1e-3) / 1e9, 1),
                      "hbm_frac_of_measured_peak": round(alg / (float(t[0]) * 1e-3) / 1e9 / peak, 4),
                      "all_reduces_per_step": 2 * lcfg.text.layers, "all_reduce": "nccl" if comm.peer is None else ({1: "fused into the row-parallel GEMM epilogue (flag + pull over peer memory)", 2: "fused into the row-parallel GEMM epilogue (LL push over peer memory)"}[comm.peer.fused] if comm.peer.fused and _lib.load().ocrb_skinny_rowparallel_tp_was_fused() else "one-shot peer-memory kernel after the GEMM"),
                      "lm_head": "logits all-gather" if os.environ.get("OCRB_TP_ARGMAX_GATHER") == "1" else "(max, lowest index) pair exchange", "init_s": round(t_init, 1), "tokens_generated": [len(x) for x in toks]}))
sys.stdout.flush()
eng.close()
dist.barrier()
torch.cuda.synchronize()
os._exit(0)      # skip communicator teardown (hung once with captured NCCL graphs alive); the job is done

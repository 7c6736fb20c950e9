import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d["roofline"]
print(" ".join(sys.argv[1:]), "value", d["value"], "e2e", d["e2e"]["value"], "ms/step", d["ms_per_step"], d["phase_ms_per_step"],
      "decode_step_ms", r["decode_step_ms"], "frac", r["frac"], "kernel_only", r["kernel_only"]["achieved"], r["kernel_only"]["frac"],
      "tok/s", d["decode_phase_tok_per_s"], "launches", d["gpu_launches"], d["clocks"])

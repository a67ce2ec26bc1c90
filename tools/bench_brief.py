"""One-line digest of a bench.py JSON line.  usage: python tools/bench_brief.py bench.json"""
import json, sys
d = json.load(open(sys.argv[1]))
print({k: d.get(k) for k in ("value", "ms_per_step", "n_gpus", "gpu_launches")}, "e2e", round(d["e2e"]["value"]),
      "frac", round(d["roofline"]["frac"], 4), "launch", d["config"].get("launch"), "shared_k", d.get("shared_k") and round(d["shared_k"]["value"]))

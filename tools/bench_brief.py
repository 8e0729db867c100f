"""Print the interesting numbers of a bench.py JSON line (stdin)."""
import json, sys
d = json.loads(sys.stdin.read())
print("value", round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "ms/step", round(d["ms_per_step"], 4), "launches", d["gpu_launches"])
print({k: round(v["ms_per_launch"], 4) for k, v in d["kernels"].items()})
b = d["bulk"]
print("bulk: index", round(b["index_ms"], 3), "knn", round(b["knn_ms"], 3), "cov", round(b["covariance_ms"], 4), "K3 frac", round(b["roofline_K3"]["frac"], 3),
      "batch corr", round(b["batch_correspond_ms"], 3), "batch lin", round(b["batch_linearize_ms"], 4), "K4b frac", round(b["roofline_K4b"]["frac"], 3))
print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in b.items() if k.startswith("prefilter")})
print(b.get("multi_sequence"))

print("e2e_pageable", d.get("e2e_pageable"))
print("e2e_cpp_dropin", d.get("e2e_cpp_dropin"))
print("bulk.all_ranks", b.get("all_ranks"))
print("multi_sequence_8", d.get("multi_sequence_8"))
print("odom_loop_cfg4", d.get("odom_loop_cfg4"))
print("cpu", d.get("cpu_baseline"))
print("per_rank", d.get("per_rank_step_ms_median_min_max_argmax"), "clocks", d.get("clocks"))

"""Quick device timings of the three single-train workloads (rounding cfg3, generic rounding, TT-SVD)."""
import json, sys
sys.path.insert(0, ".")
import bench_extras as b
for f in (b.run_round, b.run_round_generic, b.run_gramsvd, b.run_ttsvd):
    r = f()
    print(f.__name__, json.dumps({k: v for k, v in r.items() if k in ("ms", "ms_per_round", "value", "unit", "launches", "gpu_launches", "ms_per_step", "stats", "last_round")}))
    print("   ", {k: (v if not isinstance(v, (list, dict)) else "...") for k, v in r.items()})

"""Table of tools/nvlink_probe output: best GB/s per (variant, run length), uni- and bidirectional."""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith("{")]
best = {}
for r in rows:
    if "variant" not in r:
        print(r)
        continue
    if r.get("with_local_hbm_load"):
        print("WITH LOCAL HBM LOAD:", r["variant"], r["run_bytes"], "link", r["GBps_per_direction"], "GB/s/dir; local copy",
              r["local_copy_GBps_over_its_run"], "GB/s over its run")
        continue
    k = (r["variant"], r["run_bytes"], r["pitch_mul"], r["bidir"])
    if k not in best or r["GBps_per_direction"] > best[k][0]:
        best[k] = (r["GBps_per_direction"], r["param"])
print(f"{'variant':10s} {'run':>6s} {'pitch':>5s} {'uni GB/s':>9s} {'bidir GB/s':>10s}")
seen = set()
for (v, run, pm, bd) in sorted(best):
    if (v, run, pm) in seen:
        continue
    seen.add((v, run, pm))
    u = best.get((v, run, pm, 0), (0, 0))
    b = best.get((v, run, pm, 1), (0, 0))
    print(f"{v:10s} {run:6d} {pm:5d} {u[0]:9.1f} {b[0]:10.1f}   (param {u[1]}/{b[1]})")

#!/usr/bin/env python
"""Table of the bench / configs JSON lines found under a directory (default gpurun_out/): one row per file
with the cycle time, effective GB/s, roofline and NVLink fractions, e2e and the opt-in switches that were
set -- for the A/B runs of tools/r02_*.sh.   python tools/summarize_runs.py [dir] [prefix]"""
import glob
import json
import os
import sys


def rows(path):
    out = []
    for line in open(path, errors="replace"):
        line = line.strip()
        if not line.startswith("{"):
            continue
        try:
            out.append(json.loads(line))
        except json.JSONDecodeError:
            pass
    return out


def main():
    d = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out"
    prefix = sys.argv[2] if len(sys.argv) > 2 else ""
    files = sorted(glob.glob(os.path.join(d, prefix + "*.json")) + glob.glob(os.path.join(d, prefix + "*.jsonl")))
    print(f"{'file':44s} {'N':>2s} {'ms/step':>9s} {'GB/s':>9s} {'hbm frac':>8s} {'nvl frac':>8s} {'e2e GB/s':>9s}  switches / config")
    for f in files:
        for r in rows(f):
            if "ms_per_step" in r:
                cfg = r.get("config", {})
                print(f"{os.path.basename(f):44s} {r.get('n_gpus', 1):2d} {r['ms_per_step']:9.4f} {r.get('value', 0):9.1f} "
                      f"{(r.get('roofline') or {}).get('frac', 0):8.3f} {(r.get('nvlink') or {}).get('frac', 0):8.3f} "
                      f"{(r.get('e2e') or {}).get('value', 0) or 0:9.1f}  {cfg.get('switches', {})} {cfg.get('backend', '')}")
            elif "config" in r and "fwd_bwd_ms" in r:
                print(f"{os.path.basename(f):44s} {r.get('n_gpus', 0):2d} {r['fwd_bwd_ms']:9.4f} {'':9s} {'':8s} {'':8s} {'':9s}  "
                      f"{r['config']} {r.get('backend', '')} overlap={r.get('overlap_chunks', '')} "
                      f"ok={all(v for k, v in r.items() if k.endswith('_ok'))}")
            elif "what" in r and "gbs" in r:
                print(f"{os.path.basename(f):44s}  1 {r['ms']:9.4f} {r['gbs']:9.1f} {'':8s} {'':8s} {'':9s}  {r['what']} es={r['es']} "
                      f"hint={r.get('cache_hint', '')} tile={r.get('tile', '')}")


if __name__ == "__main__":
    main()

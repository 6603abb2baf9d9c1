"""Table of tools/configs_profile.py JSON lines."""
import json
import sys

for f in sys.argv[1:]:
    for line in open(f):
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        print(f"{d['config']} N={d['n_gpus']} grid={d['grid']} {d['backend']} es={d['element_bytes']} fwd+bwd {d['fwd_bwd_ms']:.3f} ms; worst: {d['furthest_below_roofline']}")
        for k, v in d["stages"].items():
            if v["bound"] == "hbm":
                print(f"    {k:24s} {v['ms']:.4f} ms  HBM {v['GBps']:.0f} GB/s = {v['frac_of_peak']:.2f} of peak  {v['parity']}")
            else:
                print(f"    {k:24s} {v['ms']:.4f} ms  NVLink {v['GBps_per_direction']:.0f} GB/s/dir = {v['frac_of_900']:.2f} of 900 "
                      f"({v['frac_of_dma_737']:.2f} of DMA)  {v['parity']}")

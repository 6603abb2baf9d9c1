"""Print the key numbers of bench.py JSON lines (files may carry other lines before the JSON)."""
import json
import sys

for f in sys.argv[1:]:
    d = None
    for line in open(f):
        if line.startswith("{"):
            try:
                d = json.loads(line)
            except Exception:
                pass
    if d is None:
        print(f, "no JSON line")
        continue
    c = d.get("config", {})
    print(f, "ms/step", round(d.get("ms_per_step", 0), 4), "GB/s", round(d.get("value", 0), 1), "backend", c.get("backend"))
    print("   per transposition:", {k: round(v, 4) for k, v in c.get("transposition_ms", {}).items()},
          "backends:", {k: round(v, 4) for k, v in c.get("backends_ms_per_step", {}).items()})
    r = d.get("roofline", {})
    print("   roofline hbm frac", round(r.get("frac", 0), 3), "exchange", {k: (round(v, 3) if isinstance(v, float) else v)
          for k, v in (r.get("exchange") or {}).items() if k in ("achieved", "frac", "frac_of_measured_dma_737", "avg_launch_ms")})
    if "parity" in d:
        print("   parity:", {k: sorted(set(v.values())) for k, v in d["parity"]["per_backend"].items()}, "peer_error", d["parity"].get("peer_error"))
    e = d.get("e2e") or {}
    print("   e2e", round(e.get("value", 0), 1), "ms", round(e.get("ms_per_step", 0), 2), "switches", c.get("switches"),
          "graph_replays", d.get("graph_replays"), "clocks", d.get("clocks"))

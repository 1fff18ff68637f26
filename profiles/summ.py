#!/usr/bin/env python
"""Print the key fields of bench.py JSON lines found in log files."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path, errors="replace"):
        if not line.startswith("{"):
            continue
        try:
            d = json.loads(line)
        except Exception:
            continue
        c = d.get("config", {})
        print(path.split("/")[-1], "| N", d.get("n_gpus"), c.get("method"), "| value", d.get("value"),
              d.get("unit"), "| ms/step", d.get("ms_per_step"), "| legs", c.get("leg_ms"),
              "| exact", c.get("round_trip_bit_exact"), "| e2e", (d.get("e2e") or {}).get("value"),
              "| roofline", (d.get("roofline") or {}).get("kernel"), (d.get("roofline") or {}).get("frac"))
        for k, v in (d.get("sections") or {}).items():
            print("    section", k, v)
        if d.get("cpu_baseline"):
            print("    cpu", d["cpu_baseline"].get("value"), d["cpu_baseline"].get("cores"))

import json, sys
for line in sys.stdin:
    line = line.strip()
    if line.startswith("{"):
        d = json.loads(line)
        r = d.get("roofline") or {}
        print("%s value=%.3fG ms/step=%.4f kernel_ms=%.4f frac=%.3f cold=%.4f acc=%.3f" % (
            sys.argv[1] if len(sys.argv) > 1 else "", d["value"] / 1e9, d["ms_per_step"], r.get("kernel_ms", 0), r.get("frac", 0),
            d.get("cold_l2_ms_per_step", 0), d.get("accept_rate", 0)))

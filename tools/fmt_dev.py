"""Pretty-print gpu_dev.py JSON lines (stdin) or a dev_*.json file (argv[1])."""
import json
import sys


def show(r):
    p = r["profile"]
    print(r["n"], "nb", r["nb"], "S", r.get("slices"), "cl", r.get("cluster"), "la", r.get("lookahead"),
          "sec=%.3f" % r["sec"], "TF(n3/3)=%.1f" % r["tflops_n3_3"],
          "syrk_ms=%.1f panel_ms=%.1f build_ms=%.1f" % (p["syrk_ms"], p["panel_ms"], p["build_ms"]),
          "syrkTF=%.1f" % r["syrk_tflops"], "rel=%.2e" % r.get("rel_err", float("nan")))


if len(sys.argv) > 1:
    for r in json.load(open(sys.argv[1]))["dense"]:
        show(r)
else:
    for ln in sys.stdin:
        ln = ln.strip()
        if ln.startswith("{"):
            try:
                show(json.loads(ln))
            except Exception:
                print(ln[:200])
        elif ln:
            print(ln[:300])

"""Micro-benchmark of single int8-update launches (b200gp_i8_update_bench) over kernel variants, with the in-kernel
cycle counters (who waits for whom).  Diagnostics only.
  python tools/i8_microbench.py [rows cols K] -- prints one JSON line per variant."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tinygp_b200 import _cabi

SLOTS = ["prod_wait", "prod_total", "mma_wait_full", "mma_wait_tempty", "mma_total", "epi_wait_tfull", "epi_total",
         "ctas", "cta_total"]


def run(ctx, rows, cols, K, S, opts, reps=3, ldq=0, ldc=0):
    ctx.reset_options()
    for k, v in opts.items():
        ctx.set_option(k, v)
    ms = ctypes.c_double()
    dbg = (ctypes.c_ulonglong * 16)()
    try:
        ctx.check(ctx.lib.b200gp_i8_update_bench(ctx.handle, rows, cols, K, S, reps, ldq, ldc, ctypes.byref(ms), dbg))
    except Exception as e:  # noqa: BLE001
        return {"opts": opts, "error": str(e)}
    ops = 2.0 * (S * (S + 1) / 2) * rows * cols * K
    d = dict(zip(SLOTS, list(dbg)[: len(SLOTS)]))
    n = max(1, d["ctas"])
    per = {k: round(v / n / 1e3, 1) for k, v in d.items() if k != "ctas"}     # kilo-cycles per CTA (pair)
    return {"opts": opts, "shape": [rows, cols, K, S, ldq, ldc], "ms": round(ms.value, 3), "pops": round(ops / ms.value / 1e12, 3),
            "ctas": d["ctas"], "kcyc_per_cta": per}


def main():
    a = [int(x) for x in sys.argv[1:6]] if len(sys.argv) >= 4 else [16384, 1024, 16384]
    rows, cols, K = a[:3]
    ldq = a[3] if len(a) > 3 else 0
    ldc = a[4] if len(a) > 4 else 0
    ctx = _cabi.get_context()
    peak = ctx.measure_i8_peak()
    print(json.dumps({"i8_peak_tops": peak}), flush=True)
    variants = [
        {"ozaki_cluster": 21},                                  # 2x1 cluster, unpaired (round-1 default)
        {"ozaki_cluster": 11},
        {"ozaki_pairing": 1, "ozaki_cluster": 21},
        {"ozaki_pairing": 1, "ozaki_cluster": 11},
        {"ozaki_pairing": 2, "ozaki_cluster": 11},              # kc-outer order, single groups
        {"ozaki_cluster": 2},
        {"ozaki_cluster": 2, "ozaki_pairing": 1},
        {"ozaki_cluster": 1},
    ]
    extra = os.environ.get("I8_VARIANTS")
    if extra:
        variants = json.loads(extra)
    for S in ([7] if not os.environ.get("I8_S") else [int(x) for x in os.environ["I8_S"].split(",")]):
        for v in variants:
            print(json.dumps(run(ctx, rows, cols, K, S, v, ldq=ldq, ldc=ldc)), flush=True)


if __name__ == "__main__":
    main()

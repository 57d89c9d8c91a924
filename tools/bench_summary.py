#!/usr/bin/env python
"""Print the essentials of bench.py JSON lines (one file per argument)."""
import json
import signal
signal.signal(signal.SIGPIPE, signal.SIG_DFL)
import sys

for path in sys.argv[1:]:
    try:
        j = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        sys.exit(1)
    print("%s: n_gpus %d value %.0f atoms/s  ms/step %.4f  e2e %.0f (%.4f ms)  launches %s transport %s" % (
        path, j["n_gpus"], j["value"], j["ms_per_step"], j["e2e"]["value"], 1e3 * j["config"]["atoms"] / j["e2e"]["value"] if "config" in j else 0,
        j.get("gpu_launches"), j.get("reduction_transport")))
    if "roofline" in j:
        r = j["roofline"]
        print("  stage_ms", r["stage_ms"], "frac %.4f of cuBLAS %.2f TF, %.4f of DMMA probe" % (r["frac"], r["fp64_dgemm_tflops_measured"], r.get("frac_of_dmma_probe", 0)))
    if "reduction" in j:
        print("  reduction", {k: v for k, v in j["reduction"].items() if k != "note"})
    if "parity" in j:
        print("  parity", j["parity"])
    if "cpu_baseline" in j:
        print("  cpu", j["cpu_baseline"]["value"], j["cpu_baseline"]["cores"])
    for k, v in (j.get("named_configs") or {}).items():
        if not v:
            continue
        r = v.get("roofline", {})
        print("  named %s: %s" % (k, {a: b for a, b in v.items() if a not in ("roofline", "workload", "timing")}))
        if r:
            print("    stage_ms", r["stage_ms"], "gemm frac %.4f" % r["frac"])
    if "clocks" in j:
        print("  clocks", j["clocks"])

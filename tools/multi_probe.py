"""One batch over 1..N devices of this process through ahs_phase_batch_multi: time per call (page-locked inputs), per-device
kernel span, equality with the single-device result, and the time of the largest chain alone (the lower bound of any
split by chains).  usage: python tools/multi_probe.py [workload] [scale]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    lib = api.load_library()
    n = lib.ahs_device_count()
    b = synth.generate(synth.config(workload, scale))
    api.pin_batch(b)
    single = api.phase_batch(b, device=0)
    out = {"workload": workload, "scale": scale, "devices_visible": n, "runs": []}
    g = 1
    while g <= n:
        devs = list(range(g))
        api.phase_batch(b, devices=devs, copy=False).release()
        ts, span = [], 0.0
        for _ in range(4):
            t0 = time.perf_counter(); r = api.phase_batch(b, devices=devs, copy=False); ts.append(time.perf_counter() - t0)
            same = not r.diff(single)
            span = r.timings["ms_total_device"]
            r.release()
        out["runs"].append({"devices": g, "ms_per_call": 1e3 * min(ts), "slowest_device_kernel_span_ms": span, "equal_to_single_device": same})
        g *= 2
    import numpy as np
    big = int(np.argmax(np.diff(b.read_off)))
    one = b.select([big])
    api.phase_batch(one, device=0, copy=False).release()
    t0 = time.perf_counter(); api.phase_batch(one, device=0, copy=False).release()
    out["largest_chain_alone_ms"] = 1e3 * (time.perf_counter() - t0)
    out["largest_chain_reads"] = int(np.diff(b.read_off).max())
    base = out["runs"][0]["ms_per_call"]
    for r in out["runs"]:
        r["speedup"] = base / r["ms_per_call"]
    print(json.dumps(out))


if __name__ == "__main__":
    main()

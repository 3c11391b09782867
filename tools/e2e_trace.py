"""Where the end-to-end time of ahs_phase_batch goes: host trace of the library (AHS_TRACE=1) and
the ctypes call vs the numpy conversion of the result."""
import ctypes as C
import os
import sys
import time

os.environ["AHS_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ahsoka_b200 import api, synth  # noqa: E402

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
batch = synth.generate(synth.config("cfg2", scale))
api.pin_batch(batch)
lib = api.load_library()
for it in range(4):
    s, o = batch.c_struct(), api.BatchOut()
    t0 = time.perf_counter()
    rc = lib.ahs_phase_batch(C.byref(s), C.byref(o), 0)
    t1 = time.perf_counter()
    r = api.result_from_struct(o)
    t2 = time.perf_counter()
    lib.ahs_free_out(C.byref(o))
    t3 = time.perf_counter()
    print(f"call {it}: rc={rc} C call {1e3*(t1-t0):.1f} ms, numpy conversion {1e3*(t2-t1):.1f} ms, free {1e3*(t3-t2):.2f} ms; device {r.timings['ms_total_device']:.1f} h2d {r.timings['ms_h2d']:.1f} d2h {r.timings['ms_d2h']:.1f}", flush=True)

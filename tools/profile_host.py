"""Where does the host time of the reference-facing API go?  cProfile over the e2e loop (C3 and C2)."""
import cProfile
import io
import os
import pstats
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench as B                                             # noqa: E402
from pycusdr_b200.config import loadModularJson               # noqa: E402
from pycusdr_b200.demodulator import UHF                      # noqa: E402

for wl in sys.argv[1:] or ["c3", "c2"]:
    cfg_file, modulation, desc = B.WORKLOADS[wl]
    conf = loadModularJson(os.path.join(B.ROOT, "config", cfg_file))
    cg = conf["GPU"]["UHF"]
    N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
    step = N - ovl
    stream = B.build_stream(conf, modulation, 16, seed=2)
    dem = UHF.Demodulator(conf, B.protocol_for(conf), B.RADIO)
    raw = dem.get_signalBufferHostPointer()
    raw[:] = 0
    blocks = [stream[c * step:(c + 1) * step] for c in range(16)]

    def loop(n):
        for i in range(n):
            raw[ovl:] = blocks[i % 16]
            dem.uploadAndFindCarrier(raw)
            dem.demodulate()
            raw[:ovl] = raw[-ovl:]
    loop(20)
    t0 = time.perf_counter()
    loop(200)
    dt = (time.perf_counter() - t0) / 200
    pr = cProfile.Profile()
    pr.enable()
    loop(200)
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
    print(f"==== {desc}: {dt * 1e3:.3f} ms per chunk through the class API")
    print(s.getvalue()[:4000])

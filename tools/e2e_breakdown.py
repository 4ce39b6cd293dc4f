"""Where does a chunk's wall time go through the reference-facing class API?  Per workload: the caller's fill of the pinned
buffer, uploadAndFindCarrier, demodulate, and -- with the same chunks through the lower-level calls -- the device part
(H2D + graph, until the stream is idle), the SNR means and the stitcher.

    python tools/e2e_breakdown.py [c2 c1 c3]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pycusdr_b200.benchmark import workloads as W              # noqa: E402
from pycusdr_b200.demodulator import UHF                       # noqa: E402
from pycusdr_b200.protocol import loadProtocol                 # noqa: E402


def med(v):
    return float(np.median(v)) * 1e6


for wl in sys.argv[1:] or ["c2", "c1", "c3"]:
    conf, mod, desc = W.load_workload(wl)
    N, ovl, step, fs = W.geometry(conf)
    P = loadProtocol(conf["Main"]["protocols"]["UHF"])(conf=conf)
    ring = 24
    stream = W.build_stream(conf, mod, ring, seed=2)
    blocks = [stream[c * step:(c + 1) * step] for c in range(ring)]
    dem = UHF.Demodulator(conf, P, W.RADIO)
    raw = dem.get_signalBufferHostPointer()
    raw[:] = 0
    n = 120
    t_fill, t_up, t_dem, t_carry = [], [], [], []
    for i in range(n + 10):
        t0 = time.perf_counter()
        raw[ovl:] = blocks[i % ring]
        t1 = time.perf_counter()
        dem.uploadAndFindCarrier(raw)
        t2 = time.perf_counter()
        dem.demodulate()
        t3 = time.perf_counter()
        raw[:ovl] = raw[-ovl:]
        t4 = time.perf_counter()
        if i >= 10:
            t_fill.append(t1 - t0)
            t_up.append(t2 - t1)
            t_dem.append(t3 - t2)
            t_carry.append(t4 - t3)
    total = med(t_fill) + med(t_up) + med(t_dem) + med(t_carry)
    print(f"==== {desc}: class API {total:.1f} us per chunk = fill {med(t_fill):.1f} + uploadAndFindCarrier {med(t_up):.1f} "
          f"+ demodulate {med(t_dem):.1f} + overlap carry {med(t_carry):.1f}  ({step / total:.1f} Msamples/s)")
    # the same chunks through the pieces pcs_chunk_to_bits is made of
    eng, st = dem._engine, dem._stitch
    st.reset()
    t_dev, t_snr, t_st = [], [], []
    clipped = np.empty(0, np.int64)
    for i in range(n + 10):
        raw[ovl:] = blocks[i % ring]
        t0 = time.perf_counter()
        eng.upload()
        res, E, sym, centre, mag = eng.process()
        t1 = time.perf_counter()
        eng.snr_means(dem.doppCyperSymNorm)
        t2 = time.perf_counter()
        st(sym, centre, mag, clipped, res.sp_sym)
        t3 = time.perf_counter()
        raw[:ovl] = raw[-ovl:]
        if i >= 10:
            t_dev.append(t1 - t0)
            t_snr.append(t2 - t1)
            t_st.append(t3 - t2)
    print(f"     pieces: upload + process (H2D, graph, D2H, sync; ctypes included) {med(t_dev):.1f} us, snr_means {med(t_snr):.1f}, "
          f"stitch {med(t_st):.1f}")
    eng.set_profiling(True)
    for i in range(12):
        raw[ovl:] = blocks[i % ring]
        eng.upload()
        eng.process()
    prof = eng.profile()
    eng.set_profiling(False)
    print("     stage events (eager launches): " + ", ".join(f"{k} {v[0] / max(v[1], 1) * 1e3:.1f} us" for k, v in prof.items()))

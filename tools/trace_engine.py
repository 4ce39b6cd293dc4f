"""Timeline of the streaming engine on one GPU (PCS_SHARD_TRACE=1): when each chunk's search starts / ends and its tail ends.
    PCS_SHARD_TRACE=1 python tools/trace_engine.py [workload] [chunks]"""
import os
import sys

os.environ["PCS_SHARD_TRACE"] = "1"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from pycusdr_b200 import sharded
from pycusdr_b200.benchmark import workloads as W
from pycusdr_b200.demodulator import UHF
from pycusdr_b200.protocol import loadProtocol

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 48
conf, mod, desc = W.load_workload(wl)
N, ovl, step, fs = W.geometry(conf)
P = loadProtocol(conf["Main"]["protocols"]["UHF"])(conf=conf)
ring = 16
dev = torch.from_numpy(W.chunks_from_stream(W.build_stream(conf, mod, ring, seed=2), N, ovl, ring)).cuda()
torch.cuda.synchronize()
dem = UHF.Demodulator(conf, P, W.RADIO)
sh = sharded.ShardedStream(dem._engine, 0, 1, lambda o: [o], lag=int(os.environ.get("LAG", "2")))
import time
for i in range(16):                # warm-up: first launches, lazy allocations
    sh.submit(dev[i % ring].data_ptr(), sharded.SRC_DEVICE, lambda c, out: None)
sh.drain(lambda c, out: None)
dem._engine.shard_sync()
t0 = time.perf_counter()
for i in range(n):
    sh.submit(dev[i % ring].data_ptr(), sharded.SRC_DEVICE, lambda c, out: None)
t1 = time.perf_counter()
sh.drain(lambda c, out: None)
dem._engine.shard_sync()
t2 = time.perf_counter()
print(f"host: {n} submits (with their fetches) returned after {1e6 * (t1 - t0) / n:.1f} us per chunk; everything finished after "
      f"{1e6 * (t2 - t0) / n:.1f} us per chunk; launches per chunk {dem._engine.launch_count / (n + 16):.1f}")
first, t = dem._engine.shard_trace()
print(desc, dem._engine.shard_info())
prev_end = None
for i, (a, b, c) in enumerate(t):
    gap = "" if prev_end is None else f" gap since previous search end {1e3 * (a - prev_end):8.1f} us"
    print(f"chunk {first + i:3d}: start {a:9.4f} ms  search end {b:9.4f} (+{1e3 * (b - a):7.1f} us)  tail end {c:9.4f} (+{1e3 * (c - b):7.1f} us){gap}")
    prev_end = b
d = np.diff(t[8:, 1])
print(f"steady state: {1e3 * d.mean():.1f} us per chunk (search end to search end), min {1e3 * d.min():.1f}, max {1e3 * d.max():.1f}")

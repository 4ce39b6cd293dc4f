"""A handful of chunks through the streaming engine on one GPU -- the command ncu profiles (launch list / full capture).
    python tools/ncu_chunks.py [workload] [chunks]"""
import os
import sys

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pycusdr_b200 import sharded
from pycusdr_b200.benchmark import workloads as W
from pycusdr_b200.demodulator import UHF
from pycusdr_b200.protocol import loadProtocol

wl = sys.argv[1] if len(sys.argv) > 1 else "c2"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 10
conf, mod, desc = W.load_workload(wl)
N, ovl, step, fs = W.geometry(conf)
P = loadProtocol(conf["Main"]["protocols"]["UHF"])(conf=conf)
ring = 8
dev = torch.from_numpy(W.chunks_from_stream(W.build_stream(conf, mod, ring, seed=2), N, ovl, ring)).cuda()
torch.cuda.synchronize()
dem = UHF.Demodulator(conf, P, W.RADIO)
sh = sharded.ShardedStream(dem._engine, 0, 1, lambda o: [o], lag=2)
acc = 0
for i in range(n):
    sh.submit(dev[i % ring].data_ptr(), sharded.SRC_DEVICE, lambda c, out: int(out[0].shift))
sh.drain(lambda c, out: int(out[0].shift))
dem._engine.shard_sync()
print(desc, "chunks", n, "launches", dem._engine.launch_count, "shift sum", sum(sh.results.values()))

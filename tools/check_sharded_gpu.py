"""torchrun script: (1) bin-sharded search over NVLink peer memory vs the unsharded path on the same chunks, bit for bit;
(2) host samples in, stitched bits out through ShardedBitStream vs the class API of one process on the same stream
(added after the round's GPU budget ended: part 2 has not run on GPUs yet)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycusdr_b200 import sharded                                                  # noqa: E402
from pycusdr_b200.benchmark import signals as S                                   # noqa: E402
from pycusdr_b200.demodulator import UHF                                          # noqa: E402
from tests.helpers import RADIO, conf_variant, protocol_for                      # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
conf = conf_variant("benchmark/bench_GMSK.json", blockSize=15)
conf["GPU"]["UHF"]["CUDA"]["device"] = local
P = protocol_for(conf)
sig, _ = S.bench_stream("GMSK", 12, seed=4)
N, ovl = 2 ** 15, 2 ** 10
step = N - ovl
nchunks = len(sig) // step - 1
chunks = torch.from_numpy(np.stack([sig[c * step:c * step + N] for c in range(nchunks)]).astype(np.complex64)).cuda()


def collect(out):
    res, E, sym, centre, mag = out
    return {"E": E.copy(), "shift": int(res.shift), "best": float(res.best_idx), "sym": sym.copy(), "centre": centre.copy(),
            "mag": mag.copy(), "timing": tuple(res.timing[:]), "peak": (res.peak_val, res.peak_bin, res.peak_mask, res.peak_offset),
            "timeout": int(res.xchg_timeout)}


def all_gather(obj):
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def gather_object(obj):
    out = [None] * world if rank == 0 else None
    dist.gather_object(obj, out, dst=0)
    return out


dems = [UHF.Demodulator(conf, P, RADIO) for _ in range(2)]
dem = dems[0]
sh = sharded.ShardedPipelines([d._engine for d in dems], rank, world, all_gather)
for rep in range(2):                       # two passes: exercises both parities of the exchange region repeatedly
    for c in range(nchunks):
        sh.enqueue(rep * nchunks + c, chunks[c].data_ptr(), collect=collect)
sh.drain(collect)
merged = sharded.gather_results(sh.results, world, gather_object, rank)
ok = True
if rank == 0:
    ref = UHF.Demodulator(conf, P, RADIO, use_graph=False)
    bad = 0
    for i, got in enumerate(merged):
        ref._engine.enqueue_device(chunks[i % nchunks].data_ptr())
        want = collect(ref._engine.fetch())
        for k in ("E", "sym", "centre", "mag"):
            if not np.array_equal(got[k], want[k]):
                bad += 1
                print(f"chunk {i}: {k} differs")
        for k in ("shift", "best", "timing", "peak"):
            if got[k] != want[k]:
                bad += 1
                print(f"chunk {i}: {k} {got[k]} != {want[k]}")
        if got["timeout"]:
            bad += 1
            print(f"chunk {i}: exchange timeout")
    print(f"sharded check: world {world}, {len(merged)} chunks, {bad} mismatches")
    ok = bad == 0

# ---- host samples in, stitched bits out (ShardedBitStream) vs the class API of one process on the same stream ----
from pycusdr_b200 import _native                                                  # noqa: E402
gloo = dist.new_group(backend="gloo")
pending = []


def send(token, dst, c):
    buf = torch.zeros(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
    buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
    pending.append(dist.isend(buf, dst=dst, tag=c, group=gloo))


def recv(src, c):
    buf = torch.empty(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
    dist.recv(buf, src=src, tag=c, group=gloo)
    return buf.numpy().tobytes()


first = sh.chunks_enqueued
bs = sharded.ShardedBitStream(sh, dem._stitch, rank, world, send, recv, first_chunk=first)
n_host = len(sig) // step
prev_tail = np.zeros(ovl, np.complex64)
for c in range(n_host):
    torch.cuda.synchronize()               # (a real host waits only for the pipeline's previous H2D copy)
    raw = bs.next_buffer()
    raw[:ovl] = prev_tail
    raw[ovl:] = sig[c * step:(c + 1) * step]
    prev_tail = raw[-ovl:].copy()
    bs.submit()
mine = bs.finish()
for w in pending:
    w.wait()
parts = [None] * world if rank == 0 else None
dist.gather_object({c: tuple(a.copy() for a in v) for c, v in mine.items()}, parts, dst=0)
if rank == 0:
    got_bits = {}
    for p in parts:
        got_bits.update(p)
    one = UHF.Demodulator(conf, P, RADIO)
    raw = one.get_signalBufferHostPointer()
    raw[:] = 0
    bad = 0
    for c in range(n_host):
        raw[ovl:] = sig[c * step:(c + 1) * step]
        one.uploadAndFindCarrier(raw)
        want = one.demodulate()[:3]
        raw[:ovl] = raw[-ovl:]
        for name, g, w in zip(("bits", "centres", "trust"), got_bits[first + c], want):
            if not np.array_equal(g, w):
                bad += 1
                print(f"host stream chunk {c}: {name} differ ({len(g)} vs {len(w)})")
    print(f"sharded host stream: world {world}, {n_host} chunks, {bad} mismatches against the single-process class API")
    ok = ok and bad == 0
dist.barrier()
for d in dems:
    d._engine.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

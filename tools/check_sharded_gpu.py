"""torchrun script: bin-sharded search over NVLink peer memory vs the unsharded path on the same chunks, bit for bit."""
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
dist.barrier()
for d in dems:
    d._engine.close()
dist.destroy_process_group()
sys.exit(0 if ok else 1)

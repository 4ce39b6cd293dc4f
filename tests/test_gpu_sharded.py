"""The bin-sharded streaming engine (pcs_shard_*, pycusdr_b200/sharded.py) against the unsharded class API on the same
samples: N > 1 results must be BIT-IDENTICAL to one GPU (E tables, shift, timing, symbols, centres, magnitudes and the
stitched bit stream), because the owner scans the very table a single GPU produces.

The multi-rank cases run `world` processes that all use cuda:0 (CUDA IPC works between processes on one device, the GPU
time-slices between them), so the whole protocol -- IPC regions, chunk broadcast from the ingest rank, peer row stores,
row / data / ack flags, owner rotation, owner-to-owner carry of the bit post-processing -- is exercised by the driver's
single-GPU `pytest -m gpu` run; `tools/check_sharded_gpu.py` repeats it with one GPU per rank."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import signals as S                                      # noqa: E402
from tests.helpers import RADIO, conf_variant, protocol_for          # noqa: E402

pytestmark = pytest.mark.gpu

CASES = {
    # name: (config, blockSize, doppCarrierSteps, modulation, snr, seed)
    "gmsk_fs256": ("benchmark/bench_GMSK.json", 14, 24, "GMSK", 12, 41),      # shifted-filter 256-point search, two lanes
    "bpsk_rot256": ("benchmark/bench_BPSK.json", 13, 10, "BPSK", 10, 42),     # 32 masks: rotate-form kernel, one lane
    "cc11xx_generic": ("CC11xx.json", 16, 12, None, 0, 43),                   # 384-tap filters: generic kernel, one lane
}


def _case(name):
    cfg, bs, D, mod, snr, seed = CASES[name]
    conf = conf_variant(cfg, blockSize=bs, doppCarrierSteps=D)
    N, ovl = 2 ** bs, 2 ** conf["GPU"]["UHF"]["overlap"]
    if mod is None:
        from pycusdr_b200.benchmark import workloads as W
        sig = W.build_stream(conf, None, 20, seed=seed)        # 20 chunks: >= 9 owned chunks per rank on two ranks, so the
                                                               # tail runs from its CUDA graph (captured on an instance's second use)
    else:
        sig, _ = S.bench_stream(mod, snr, n_packets=3, seed=seed, pre_blocks=1)
        sig = sig[:(len(sig) // (N - ovl)) * (N - ovl)]
        sig = sig[:30 * (N - ovl)]      # 30 chunks: the graph-replayed tail is exercised on 1, 2 and 3 ranks
    return conf, N, ovl, np.ascontiguousarray(sig, dtype=np.complex64)


def _unsharded(conf, sig):
    """Per-chunk device outputs and bits of the strictly alternating class API (demodulator_process.py:284-338)."""
    from pycusdr_b200.demodulator import UHF
    dem = UHF.Demodulator(conf, protocol_for(conf), RADIO)
    N, ovl = dem.Nfft, dem.sigOverlap
    step = N - ovl
    raw = dem.get_signalBufferHostPointer()
    raw[:] = 0
    out = []
    for c in range(len(sig) // step):
        raw[ovl:] = sig[c * step:(c + 1) * step]
        dem.uploadAndFindCarrier(raw)
        bits, centres, trust, sp = dem.demodulate()
        last = dem.last
        out.append(dict(E=last["E"].copy(), shift=int(last["shift"]), timing=float(last["timing"][0]), sym=last["sym"].copy(),
                        centres=last["centres"].copy(), mag=last["mag"].copy(), bits=bits, trust=trust, peak=tuple(last["peak"])))
        raw[:ovl] = raw[-ovl:]
    return out


def _compare(got, want, what):
    assert len(got) == len(want)
    for c, (g, w) in enumerate(zip(got, want)):
        np.testing.assert_array_equal(g["E"], w["E"], err_msg=f"{what} chunk {c}: E")
        assert (g["shift"], g["timing"]) == (w["shift"], w["timing"]), f"{what} chunk {c}"
        assert tuple(g["peak"]) == tuple(w["peak"]), f"{what} chunk {c}: peak"
        for k in ("sym", "centres", "mag", "bits", "trust"):
            np.testing.assert_array_equal(g[k], w[k], err_msg=f"{what} chunk {c}: {k}")


def _run_rank(rank, world, name, kind, all_gather, send, recv, lag=2, device=0, before_chunk=None):
    """The per-rank loop of the sharded stream; returns {chunk: result dict} of the chunks this rank owned."""
    import torch
    from pycusdr_b200 import sharded
    from pycusdr_b200.demodulator import UHF
    conf, N, ovl, sig = _case(name)
    conf["GPU"]["UHF"]["CUDA"]["device"] = device
    step = N - ovl
    dem = UHF.Demodulator(conf, protocol_for(conf), RADIO)
    sh = sharded.ShardedStream(dem._engine, rank, world, all_gather, lag=lag)
    info = dem._engine.shard_info()
    assert (info["bin_lo"], info["bin_hi"]) == sh.slices[rank]
    bs = sharded.ShardedBitStream(sh, dem._stitch, rank, world, send, recv)
    raws = {}
    n_chunks = len(sig) // step
    orig = bs._collect

    def collect(c, out):          # keep the raw device outputs of the owned chunks as well
        res, E, sym, centre, mag, means = out
        raws[c] = dict(E=E.copy(), shift=int(res.shift), timing=float(res.timing[0]), sym=sym.copy(), centres=centre.copy(),
                       mag=mag.copy(), peak=(res.peak_val, res.peak_bin, res.peak_mask, res.peak_offset))
        return orig(c, out)
    bs._collect = collect
    prev_tail = np.zeros(ovl, np.complex64)
    dev_keep = []
    for c in range(n_chunks):
        if before_chunk is not None:
            before_chunk(c, dem._engine)
        if rank == 0:
            chunk = np.concatenate((prev_tail, sig[c * step:(c + 1) * step]))
            prev_tail = chunk[-ovl:].copy()
            if kind == sharded.SRC_HOST:
                slot = bs.host_slot()
                slot[:] = chunk
                bs.submit(None, sharded.SRC_HOST)
            else:
                t = torch.from_numpy(chunk).cuda()
                torch.cuda.synchronize()          # the engine's streams do not synchronise with torch's
                dev_keep.append(t)
                bs.submit(t.data_ptr(), sharded.SRC_DEVICE)
        else:
            assert bs.host_slot() is None
            bs.submit(None, kind)
    bs.finish()
    dem._engine.shard_sync()
    mine = {}
    for c, (bits, centres8, trust) in bs.bits.items():
        if c in raws:
            mine[c] = dict(raws[c], bits=bits, trust=trust)
    return mine, set(bs.bits), n_chunks


@pytest.mark.parametrize("name,kind", [("gmsk_fs256", "host"), ("gmsk_fs256", "device"), ("bpsk_rot256", "host"),
                                       ("cc11xx_generic", "device")])
def test_single_rank_engine_equals_the_class_api(name, kind):
    from pycusdr_b200 import sharded
    conf, N, ovl, sig = _case(name)
    want = _unsharded(conf, sig)
    k = sharded.SRC_HOST if kind == "host" else sharded.SRC_DEVICE
    mine, owned, n = _run_rank(0, 1, name, k, lambda o: [o], None, None, lag=3)
    assert owned == set(range(n)) == set(mine) and n == len(want) >= 6
    _compare([mine[c] for c in range(n)], want, name)


def test_graph_replayed_tail_survives_eager_chunks_in_between():
    """The owner's tail is replayed from a CUDA graph from an instance's second use on; with per-stage profiling switched on
    the engine falls back to eager launches, and the flag value the graph reads from device memory must be re-seeded when the
    replays resume (chunks 11-14 eager here, graphs before and after)."""
    from pycusdr_b200 import sharded
    conf, N, ovl, sig = _case("gmsk_fs256")
    want = _unsharded(conf, sig)

    def toggle(c, eng):
        if c == 11:
            eng.set_profiling(True)
        if c == 15:
            eng.set_profiling(False)
    mine, owned, n = _run_rank(0, 1, "gmsk_fs256", sharded.SRC_DEVICE, lambda o: [o], None, None, lag=5, before_chunk=toggle)
    assert n == len(want) >= 25
    _compare([mine[c] for c in range(n)], want, "graph / eager / graph")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q, name, kind):
    import faulthandler
    faulthandler.dump_traceback_later(240, exit=True)
    import torch
    import torch.distributed as dist
    from pycusdr_b200 import _native
    # PCS_TEST_ONE_GPU_PER_RANK=1 on a multi-GPU box: rank r on cuda:r (NVLink between the ranks); default: all on cuda:0
    device = rank if (os.environ.get("PCS_TEST_ONE_GPU_PER_RANK") and torch.cuda.device_count() >= world) else 0
    torch.cuda.set_device(device)
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        pending = []
        cap = _native.Stitcher.STATE_BYTES

        def send(token, dst, c):
            buf = torch.zeros(cap, dtype=torch.uint8)
            buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
            pending.append(dist.isend(buf, dst=dst, tag=c))

        def recv(src, c):
            buf = torch.empty(cap, dtype=torch.uint8)
            dist.recv(buf, src=src, tag=c)
            return buf.numpy().tobytes()
        mine, owned, n = _run_rank(rank, world, name, kind, all_gather, send, recv, lag=1, device=device)
        for w in pending:
            w.wait()
        assert owned == {c for c in range(n) if c % world == rank}
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        if rank == 0:
            merged = {}
            for p in parts:
                merged.update(p)
            q.put(merged)
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("world,name,kind", [(2, "gmsk_fs256", "host"), (3, "gmsk_fs256", "device"), (2, "cc11xx_generic", "host"),
                                             (2, "bpsk_rot256", "device")])
def test_ranks_sharing_one_gpu_reproduce_the_single_gpu_stream(world, name, kind):
    import torch.multiprocessing as mp
    from pycusdr_b200 import sharded
    conf, N, ovl, sig = _case(name)
    want = _unsharded(conf, sig)
    k = sharded.SRC_HOST if kind == "host" else sharded.SRC_DEVICE
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, name, k)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        merged = q.get(timeout=500)
    finally:
        for p in procs:
            p.join(timeout=60)
    assert all(p.exitcode == 0 for p in procs)
    assert sorted(merged) == list(range(len(want)))
    _compare([merged[c] for c in range(len(want))], want, f"{name} world {world}")

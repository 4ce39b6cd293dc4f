"""N > 1 host logic on the CPU: two ``gloo`` ranks drive ``pycusdr_b200.sharded.ShardedStream`` with a NumPy stand-in
for the native engine (chunk "broadcast" from the ingest rank and rows to the owner rank = point-to-point gloo messages,
local rows from the oracle).  The merged per-chunk results must equal what one process computes over all bins."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import oracle as O                                       # noqa: E402
from oracle import signals as S                                      # noqa: E402
from pycusdr_b200 import sharded                                     # noqa: E402
from tests.helpers import RADIO, conf_variant, protocol_for          # noqa: E402

N_CHUNKS = 5


def _conf():
    return conf_variant("benchmark/bench_GMSK.json", blockSize=12, doppCarrierSteps=9)


def _chunks():
    sig, _ = S.get_padded_packet("GMSK", 16, 153600, pad=500, raw_bits=S.createBitSequence(1400, seed=3))
    rng = np.random.RandomState(5)
    sig = (sig + 0.1 * (rng.randn(len(sig)) + 1j * rng.randn(len(sig)))).astype(np.complex64)
    N = 4096
    return [sig[c * 3072:c * 3072 + N] for c in range(N_CHUNKS)]


def _tail(orc, E, pv, po, x):
    """estimate + demod of one chunk from the full tables (what the owner does)."""
    Es = E
    if orc.SUM_ALL_MASKS_PYTHON:
        Es = np.zeros_like(E)
        acc = E[:, 0].copy()
        for m in range(1, E.shape[1]):
            acc = (acc + E[:, m]).astype(np.float32)
        Es[:, 0] = acc
    res = O.find_doppler_est(Es, orc.num_dopplers, orc.doppIdxArrayOffset, orc.SUM_ALL_MASKS_PYTHON)
    lo, hi, hz, shift = O.interpolate_doppler(res[0], orc.doppCyperSymNorm, orc.doppHzLUT)
    orc.X = O.forward_fft(x)
    orc.dopplerIdxlast = shift
    orc.state = O.OverlapState()
    bits = orc.demodulate()[0]
    g = np.unravel_index(np.argmax(pv), pv.shape)
    return {"E": Es, "shift": int(shift), "bits": bits, "peak": (int(g[0]), int(g[1]), int(po[g]))}


class FakeEngine:
    """NumPy stand-in for ``_native.Engine``'s pcs_shard_* interface: rank 0 "broadcasts" the chunk with gloo messages, every
    rank computes its rows with the oracle and sends them to the owner, the owner assembles the table when it fetches."""

    def __init__(self, rank, world):
        conf = _conf()
        self.orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO)
        self.D, self.M = len(self.orc.doppCyperSymNorm), self.orc.num_masks
        self.rank, self.world = rank, world
        self.tables, self.chunks, self.log, self.sends = {}, {}, [], []
        self.slot = np.zeros(4096, np.complex64)      # stand-in for the pinned host ring

    def shard_init(self, rank, world, ring):
        assert (rank, world) == (self.rank, self.world)
        self.lo, self.hi = sharded.bin_partition(self.D, world)[rank]
        return bytes([self.rank]) * 64

    def shard_attach(self, handles):
        assert [h[0] for h in handles] == list(range(self.world)) and all(len(h) == 64 for h in handles)

    def shard_host_slot(self, seq):
        assert self.rank == 0
        return self.slot

    def shard_submit(self, seq, kind, src):
        if self.rank == 0:
            x = self.slot.copy() if src is None else np.asarray(src, dtype=np.complex64)
            assert kind == (sharded.SRC_HOST if src is None else sharded.SRC_DEVICE)
            for r in range(1, self.world):
                self.sends.append(dist.isend(torch.from_numpy(x.view(np.float32).copy()), dst=r, tag=seq * 8 + 7))
        else:
            buf = torch.empty(2 * 4096, dtype=torch.float32)
            dist.recv(buf, src=0, tag=seq * 8 + 7)
            x = buf.numpy().view(np.complex64)
        owner = seq % self.world
        X = O.forward_fft(x)
        E, pv, po = O.search_energy(X, self.orc.masks, self.orc.doppCyperSymNorm[self.lo:self.hi], False, want_peaks=True)
        self.log.append(("push", seq, owner))
        if owner == self.rank:
            t = self.tables.setdefault(seq, [np.zeros((self.D, self.M), np.float32), np.zeros((self.D, self.M), np.float32),
                                             np.zeros((self.D, self.M), np.int32)])
            t[0][self.lo:self.hi], t[1][self.lo:self.hi], t[2][self.lo:self.hi] = E, pv, po
            self.chunks[seq] = x
        else:
            for k, a in enumerate((E, pv, po.view(np.float32))):
                self.sends.append(dist.isend(torch.from_numpy(np.ascontiguousarray(a)), dst=owner, tag=seq * 8 + k))

    def shard_fetch(self, seq):
        t = self.tables.pop(seq)
        for r, (lo, hi) in enumerate(sharded.bin_partition(self.D, self.world)):
            if r == self.rank:
                continue
            for k in range(3):
                buf = torch.empty((hi - lo, self.M), dtype=torch.float32)
                dist.recv(buf, src=r, tag=seq * 8 + k)
                t[k][lo:hi] = buf.numpy() if k < 2 else buf.numpy().view(np.int32)
        self.log.append(("tail", seq))
        return _tail(self.orc, t[0], t[1], t[2], self.chunks.pop(seq))


def _worker(rank, world, port, q, host):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out

        def gather_object(obj):
            out = [None] * world if rank == 0 else None
            dist.gather_object(obj, out, dst=0)
            return out

        eng = FakeEngine(rank, world)
        sh = sharded.ShardedStream(eng, rank, world, all_gather, lag=1)
        owners = []
        for seq, x in enumerate(_chunks()):
            if host:                            # host ingestion: rank 0 writes the samples into the pinned slot
                slot = sh.host_slot()
                assert (slot is None) == (rank != 0)
                if slot is not None:
                    slot[:] = x
                owners.append(sh.submit(None, sharded.SRC_HOST))
            else:                               # only the ingest rank is given the samples
                owners.append(sh.submit(x if rank == 0 else None, sharded.SRC_DEVICE))
            assert len(sh._owned) <= 2          # lag = 1: at most the new chunk and one older one in flight
        sh.drain()
        for w in eng.sends:
            w.wait()
        assert owners == [s % world for s in range(N_CHUNKS)]
        assert sorted(sh.results) == [s for s in range(N_CHUNKS) if s % world == rank]
        merged = sharded.gather_results(sh.results, world, gather_object, rank)
        if rank == 0:
            q.put([{k: (v.tolist() if isinstance(v, np.ndarray) else v) for k, v in r.items()} for r in merged])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_bin_partition_and_owner_schedule():
    assert sharded.bin_partition(256, 8) == [(32 * r, 32 * r + 32) for r in range(8)]
    assert sharded.bin_partition(9, 2) == [(0, 5), (5, 9)]
    assert sharded.bin_partition(65, 4) == [(0, 17), (17, 33), (33, 49), (49, 65)]
    with pytest.raises(ValueError):
        sharded.bin_partition(3, 4)
    assert [sharded.owner_of(s, 4) for s in range(6)] == [0, 1, 2, 3, 0, 1]


def test_owned_chunks_are_collected_lag_rounds_late():
    class E:
        D = 8

        def __init__(self):
            self.calls = []

        def shard_init(self, rank, world, ring): return b"\0" * 64
        def shard_attach(self, handles): pass
        def shard_submit(self, seq, kind, src): self.calls.append(("submit", seq))

        def shard_fetch(self, seq):
            self.calls.append(("fetch", seq))
            return seq
    e = E()
    sh = sharded.ShardedStream(e, 0, 1, lambda o: [o], lag=2)
    for _ in range(5):
        sh.submit(123)
    assert e.calls == [("submit", 0), ("submit", 1), ("submit", 2), ("fetch", 0), ("submit", 3), ("fetch", 1), ("submit", 4)]
    sh.drain()
    assert sorted(sh.results) == [0, 1, 2, 3, 4] and sh.next_seq == 5
    with pytest.raises(ValueError):
        sharded.ShardedStream(E(), 0, 1, lambda o: [o], lag=sharded.RESULT_STAGES)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("host", [False, True])
def test_two_gloo_ranks_reproduce_the_single_process_result(host):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, host)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process over all bins
    conf = _conf()
    orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO)
    assert len(merged) == N_CHUNKS
    for x, got in zip(_chunks(), merged):
        X = O.forward_fft(x)
        E, pv, po = O.search_energy(X, orc.masks, orc.doppCyperSymNorm, False, want_peaks=True)
        want = _tail(orc, E, pv, po, x)
        assert np.array_equal(np.asarray(got["E"], dtype=np.float32), want["E"])      # bit for bit
        assert got["shift"] == want["shift"]
        assert tuple(got["peak"]) == want["peak"]
        assert np.array_equal(np.asarray(got["bits"], dtype=np.uint8), want["bits"])

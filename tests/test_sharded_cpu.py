"""N > 1 host logic on the CPU: two ``gloo`` ranks drive ``pycusdr_b200.sharded.ShardedSearch`` with a NumPy stand-in
for the engine (local rows from the oracle, "peer memory" = point-to-point gloo messages to the owner rank).  The merged
per-chunk results must equal what one process computes over all bins."""
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
    def __init__(self, rank, world, pipe=0):
        self.pipe = pipe
        conf = _conf()
        self.orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO)
        self.D, self.M = len(self.orc.doppCyperSymNorm), self.orc.num_masks
        self.rank, self.world = rank, world
        self.tables = {}
        self.out = None
        self.log = []
        self.host_buffer = np.zeros(4096, np.complex64)      # stand-in for the pinned chunk buffer

    def set_bin_range(self, lo, hi):
        self.lo, self.hi = lo, hi

    def peer_export(self):
        return bytes([self.rank]) * 64

    def peer_attach(self, rank, world, handles):
        assert rank == self.rank and world == self.world
        assert [h[0] for h in handles] == list(range(world)) and all(len(h) == 64 for h in handles)

    def upload_device(self, chunk):
        self.x = chunk

    def upload(self):                       # chunk=None: the samples sit in the engine's pinned host buffer
        self.x = self.host_buffer.copy()

    def _rows(self):
        X = O.forward_fft(self.x)
        return O.search_energy(X, self.orc.masks, self.orc.doppCyperSymNorm[self.lo:self.hi], False, want_peaks=True)

    def enqueue_search_push(self, seq, owner):
        E, pv, po = self._rows()
        self.log.append(("push", seq, owner))
        if owner == self.rank:
            t = self.tables.setdefault(seq, [np.zeros((self.D, self.M), np.float32), np.zeros((self.D, self.M), np.float32),
                                             np.zeros((self.D, self.M), np.int32)])
            t[0][self.lo:self.hi], t[1][self.lo:self.hi], t[2][self.lo:self.hi] = E, pv, po
        else:
            for k, a in enumerate((E, pv, po.view(np.float32))):
                dist.send(torch.from_numpy(np.ascontiguousarray(a)), dst=owner, tag=(seq * 4 + k) * 8 + self.pipe)

    def enqueue_owner_tail(self, seq):
        t = self.tables.pop(seq)
        for r, (lo, hi) in enumerate(sharded.bin_partition(self.D, self.world)):
            if r == self.rank:
                continue
            for k in range(3):
                buf = torch.empty((hi - lo, self.M), dtype=torch.float32)
                dist.recv(buf, src=r, tag=(seq * 4 + k) * 8 + self.pipe)
                t[k][lo:hi] = buf.numpy() if k < 2 else buf.numpy().view(np.int32)
        self.log.append(("tail", seq))
        self.out = _tail(self.orc, t[0], t[1], t[2], self.x)

    def fetch(self):
        out, self.out = self.out, None
        assert out is not None
        return out


def _worker(rank, world, port, q, pipes=1):
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

        if pipes == 1:
            sh = sharded.ShardedSearch(FakeEngine(rank, world), rank, world, all_gather)
            want_owner = [s % world for s in range(N_CHUNKS)]
        else:
            sh = sharded.ShardedPipelines([FakeEngine(rank, world, j) for j in range(pipes)], rank, world, all_gather)
            want_owner = [(s // pipes) % world for s in range(N_CHUNKS)]
        if pipes == 1:
            owners = [sh.enqueue(seq, x) for seq, x in enumerate(_chunks())]
        else:                               # host-buffer ingestion: samples written into the pipeline's pinned buffer
            owners = []
            for seq, x in enumerate(_chunks()):
                sh.pipes[seq % pipes].engine.host_buffer[:] = x
                owners.append(sh.enqueue(seq))
        sh.drain()
        assert owners == want_owner
        assert sorted(sh.results) == [s for s in range(N_CHUNKS) if want_owner[s] == rank]
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


def test_out_of_order_chunks_are_rejected():
    class E:
        D = 8

        def set_bin_range(self, lo, hi): pass
        def peer_export(self): return b"\0" * 64
        def peer_attach(self, *a): pass
    sh = sharded.ShardedSearch(E(), 0, 1, lambda o: [o])
    with pytest.raises(ValueError):
        sh.enqueue(1, None)


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pipes", [1, 2])
def test_two_gloo_ranks_reproduce_the_single_process_result(pipes):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, pipes)) for r in range(world)]
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

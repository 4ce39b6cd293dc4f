"""Exact bit post-processing when consecutive chunks are finished on different ranks (bin sharding):
two ``gloo`` ranks run ``sharded.OrderedStitcher`` over the native stitcher and pass the chunk-to-chunk carry
(dem_base:977-979) point to point; the merged stream must equal what ONE stitcher produces going through the chunks in
order (which tests/test_native_stitch.py pins to the reference's own checkSymbolOverlap).  Host logic only: no GPU."""
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

from pycusdr_b200 import _native, sharded                           # noqa: E402

N, OVL, SPS, N_CHUNKS = 4096, 1024, 16, 20


def _stitcher():
    return _native.Stitcher(nfft=N, overlap=OVL, overlap_offset=20, error_threshold=1000, match_threshold=10,
                            bit_lut=np.array([0, 1], np.uint8), symbol_lut=None)


def _chunks():
    """Symbol tables of consecutive chunks cut from one random bit stream; a per-chunk timing jitter moves the symbols
    next to the window edges in and out of the window, which is what the +-1-bit realignment exists for."""
    rng = np.random.RandomState(7)
    step = N - OVL
    stream = rng.randint(0, 2, size=(N_CHUNKS + 2) * N // SPS).astype(np.int32)
    out = []
    for c in range(N_CHUNKS):
        jitter = int(rng.randint(-14, 15))
        k = np.arange(len(stream))
        centre = k * SPS + 8 - c * step + jitter
        keep = (centre >= 0) & (centre < N)
        sym = stream[keep].copy()
        flips = rng.rand(len(sym)) < 0.01                  # a few symbol errors
        sym[flips] ^= 1
        mag = rng.rand(len(sym)).astype(np.float32) * 1e3
        clipped = np.array([700 + 13 * c, 2000], np.int64) if c % 3 == 0 else np.zeros(0, np.int64)
        out.append((sym, centre[keep].astype(np.int32), mag, clipped, np.float64(SPS)))
    return out


def _sequential():
    st = _stitcher()
    return [st(*args) for args in _chunks()]


def _worker(rank, world, port, q, pipes):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        owner = lambda c: (c // pipes) % world               # noqa: E731  (ShardedPipelines' schedule)
        pending = []

        def send(token, dst, c):
            buf = torch.zeros(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
            buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
            pending.append(dist.isend(buf, dst=dst, tag=c))

        def recv(src, c):
            buf = torch.empty(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
            dist.recv(buf, src=src, tag=c)
            return buf.numpy().tobytes()

        os_ = sharded.OrderedStitcher(_stitcher(), rank, owner, send, recv)
        mine = {c: os_(c, *args) for c, args in enumerate(_chunks()) if owner(c) == rank}
        # the carry of the last chunk is addressed to the owner of a chunk that never comes: take it off the wire
        if owner(N_CHUNKS) == rank and owner(N_CHUNKS - 1) != rank:
            assert len(recv(owner(N_CHUNKS - 1), N_CHUNKS - 1)) == _native.Stitcher.STATE_BYTES
        for w in pending:
            w.wait()
        with pytest.raises(ValueError):
            os_(next(c for c in range(N_CHUNKS) if owner(c) != rank), *_chunks()[0])
        parts = [None] * world if rank == 0 else None
        dist.gather_object({c: [a.tolist() for a in v] for c, v in mine.items()}, parts, dst=0)
        if rank == 0:
            merged = {}
            for p in parts:
                merged.update(p)
            q.put([merged[c] for c in range(N_CHUNKS)])
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_the_carry_depends_on_its_own_chunk_only():
    """What makes the scheme possible: whatever carry a stitcher holds when it processes a chunk, the carry it holds
    afterwards is the same."""
    chunks = _chunks()
    a, b = _stitcher(), _stitcher()
    a(*chunks[0])
    a(*chunks[1])
    b(*chunks[5])
    b(*chunks[1])
    assert a.get_state() == b.get_state() and len(a.get_state()) > 8 + 21
    b.reset()
    assert b.get_state() == b"\0" * 8
    b.set_state(a.get_state())
    x, y = a(*chunks[2]), b(*chunks[2])
    assert all(np.array_equal(u, v) for u, v in zip(x, y))


def test_the_stream_needs_the_realignment():
    seq = _sequential()
    fresh = []
    for args in _chunks():
        fresh.append(_stitcher()(*args))
    assert sum(len(s[0]) != len(f[0]) for s, f in zip(seq, fresh)) >= 2     # the carry changes some chunks by one bit


class _Res:
    xchg_timeout = 0
    sp_sym = float(SPS)


class _StubEngine:
    """Stands in for ``_native.Engine`` in the host protocol: the "device" returns the symbol table of the chunk that is
    fetched (chunks before ``first`` belong to an earlier, device-resident phase and have no symbols)."""
    D = 64

    def __init__(self, rank, first, chunks):
        self.rank, self.first, self.chunks = rank, first, chunks
        self.slot = np.zeros(8, np.complex64)

    def shard_init(self, rank, world, ring): return b"\0" * 64
    def shard_attach(self, handles): pass
    def shard_host_slot(self, seq): return self.slot
    def shard_submit(self, seq, kind, src): pass

    def shard_fetch(self, seq):
        c = seq - self.first
        if c < 0:
            return (_Res(), None, None, None, None, None)
        sym, centre, mag, _, _ = self.chunks[c]
        return (_Res(), None, sym, centre, mag, None)


def _stream_worker(rank, world, port, q, lag, prior):
    import faulthandler
    faulthandler.dump_traceback_later(120, exit=True)       # a protocol deadlock must fail the test, not hang it
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        chunks = [(a, b, c, None, e) for a, b, c, _, e in _chunks()]
        sh = sharded.ShardedStream(_StubEngine(rank, prior, chunks), rank, world, all_gather, lag=lag)
        for i in range(prior):                  # device-resident chunks before the host stream starts (as in bench.py)
            sh.submit(1, sharded.SRC_DEVICE, collect=lambda c, out: None)
        sh.drain(lambda c, out: None)
        assert sh.chunks_enqueued == prior
        pending = []
        cap = _native.Stitcher.state_capacity(OVL, 20, SPS // 2, 7)

        def send(token, dst, c):
            assert len(token) <= cap
            buf = torch.zeros(cap, dtype=torch.uint8)
            buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
            pending.append(dist.isend(buf, dst=dst, tag=c))

        def recv(src, c):
            buf = torch.empty(cap, dtype=torch.uint8)
            dist.recv(buf, src=src, tag=c)
            return buf.numpy().tobytes()
        bs = sharded.ShardedBitStream(sh, _stitcher(), rank, world, send, recv)
        assert bs.first == prior
        for k in range(N_CHUNKS):
            assert (bs.host_slot() is None) == (rank != 0)
            assert bs.submit() == prior + k
            if k == N_CHUNKS // 2:
                bs.drain()                      # a drain in the middle of the stream must neither block nor lose the carry
        mine = bs.finish()
        for w in pending:
            w.wait()
        assert sorted(mine) == [c for c in range(prior, prior + N_CHUNKS) if bs.owner_of_chunk(c) == rank]
        parts = [None] * world if rank == 0 else None
        dist.gather_object({c: [a.tolist() for a in v] for c, v in mine.items()}, parts, dst=0)
        if rank == 0:
            merged = {}
            for p in parts:
                merged.update(p)
            q.put([merged[c] for c in range(prior, prior + N_CHUNKS)])
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
@pytest.mark.parametrize("world,lag,prior", [(2, 3, 5), (4, 2, 7), (3, 0, 0), (4, 1, 16)])
def test_host_stream_over_the_sharded_engine_is_exact_and_terminates(world, lag, prior):
    """``ShardedBitStream``: the whole N > 1 host protocol (what ``bench.py --gpus N`` runs for its e2e figure) with stub
    engines: owned chunks collected 0..3 owner rounds late, a device-resident phase of any length before the host stream,
    a drain in the middle.  The carries travel from owner to owner; nobody may wait in a cycle."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_stream_worker, args=(r, world, port, q, lag, prior)) for r in range(world)]
    for p in procs:
        p.start()
    merged = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    chunks = _chunks()
    st = _stitcher()
    want = [st(a, b, c, (), e) for a, b, c, _, e in chunks]
    assert len(merged) == N_CHUNKS
    for got, exp in zip(merged, want):
        for g, e in zip(got, exp):
            assert np.array_equal(np.asarray(g, dtype=np.uint8), e)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pipes", [1, 2])
def test_two_gloo_ranks_produce_the_single_process_bit_stream(pipes):
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
    want = _sequential()
    assert len(merged) == len(want) == N_CHUNKS
    for got, exp in zip(merged, want):
        for g, e in zip(got, exp):
            assert np.array_equal(np.asarray(g, dtype=np.uint8), e)

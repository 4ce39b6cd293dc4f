"""Doppler-bin sharding of the search across GPUs, one process per GPU.

Rank 0 is the ingest rank: it alone receives the radio's samples (one SigFIFO per radio in the reference,
sigFIFO.py:147-181) and the native engine copies every chunk from its HBM into the peers' chunk rings over NVLink.
Every rank searches a contiguous slice of the Doppler bins; the three per-(bin, mask) tables of the slice (energy, peak
value, peak offset) are stored by the search kernel's finishing CTAs directly into the exchange region of the chunk's
*owner* rank over NVLink peer memory; the owner alone then runs the tail of the chunk -- Doppler estimate, demodulation
at the selected bin, timing recovery, symbol decisions -- on the gathered table, which is the very table a single GPU
would have produced, so the results are bit-identical to the unsharded path.  Owners rotate round-robin over the
chunks, so the part of the work that does not shard costs each rank 1/world of a chunk, and no collective is on the
per-chunk path (``pcs_shard_*`` in include/pycusdr_b200.h; flag protocol in csrc/shard.inc).  ``torch.distributed`` (any
backend) is used once, to exchange the 64-byte CUDA IPC handles, by the bit post-processing to pass a <= 200-byte carry from
owner to owner, and at the end to gather the per-chunk results on rank 0.

The classes here are host logic only: they drive an *engine* object with the interface of ``pycusdr_b200._native.Engine``
(``D``, ``shard_init``, ``shard_attach``, ``shard_host_slot``, ``shard_submit``, ``shard_fetch``), which is what lets the
world_size-2 ``gloo`` tests in ``tests/test_sharded_cpu.py`` run them without a GPU.
"""
from collections import deque

SRC_DEVICE, SRC_HOST = 1, 2          # include/pycusdr_b200.h: PCS_SRC_DEVICE / PCS_SRC_HOST
RESULT_STAGES = 8                    # result stages the engine keeps per owner (csrc/shard.inc: PCS_SHARD_STAGES)


def bin_partition(num_bins, world):
    """Contiguous, balanced slices ``[(lo, hi), ...]`` of ``range(num_bins)``, one per rank, none empty."""
    if world < 1 or num_bins < world:
        raise ValueError(f"cannot shard {num_bins} Doppler bins over {world} ranks")
    base, extra = divmod(num_bins, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def owner_of(seq, world):
    """Rank that runs the non-sharded tail (estimate + demod) of chunk ``seq``."""
    return seq % world


class ShardedStream:
    """Per-rank driver of the native engine.  ``all_gather`` is a callable ``obj -> [obj of rank 0, obj of rank 1, ...]``
    (e.g. a thin wrapper over ``torch.distributed.all_gather_object``).  ``lag`` = owned chunks kept in flight before the
    oldest one is collected (at most ``RESULT_STAGES - 1``): collecting late keeps the host off the device's critical
    path -- by the time chunk ``c`` is fetched, ``lag * world`` later chunks have been enqueued."""

    def __init__(self, engine, rank, world, all_gather, ring=0, lag=3):
        if not 0 <= lag < RESULT_STAGES:
            raise ValueError(f"lag must be in [0, {RESULT_STAGES - 1}]")
        self.engine, self.rank, self.world, self.lag = engine, rank, world, lag
        self.slices = bin_partition(engine.D, world)
        handles = all_gather(engine.shard_init(rank, world, ring))
        if len(handles) != world:
            raise RuntimeError("handle exchange returned %d entries for %d ranks" % (len(handles), world))
        engine.shard_attach(handles)
        self._owned = deque()          # chunks whose tail this rank has enqueued but not collected yet
        self.results = {}              # seq -> whatever ``collect`` made of engine.shard_fetch(seq)
        self._next_seq = 0

    @property
    def next_seq(self):
        """Sequence number of the chunk the next ``submit`` enqueues."""
        return self._next_seq

    chunks_enqueued = next_seq

    def host_slot(self):
        """Ingest rank: the pinned buffer (complex64[nfft]) the NEXT chunk's samples are written into before
        ``submit(kind=SRC_HOST)``; ``None`` on the other ranks."""
        return self.engine.shard_host_slot(self._next_seq) if self.rank == 0 else None

    def submit(self, src=None, kind=SRC_DEVICE, collect=None):
        """Enqueue the next chunk on this rank (call on EVERY rank, in the same order).  ``src`` matters on the ingest
        rank only: a device pointer (``SRC_DEVICE``), a host array (``SRC_HOST``), or ``None`` with ``SRC_HOST`` when the
        samples were written into ``host_slot()``.  Never waits for another rank; if this rank owns the chunk, owned chunks
        older than ``lag`` owner rounds are collected first, through ``collect(seq, fetch_tuple)`` if given.  Returns the
        chunk's owner."""
        seq = self._next_seq
        owner = owner_of(seq, self.world)
        if owner == self.rank:
            while len(self._owned) > self.lag:
                self._collect_one(collect)
        self.engine.shard_submit(seq, kind, src)
        self._next_seq = seq + 1
        if owner == self.rank:
            self._owned.append(seq)
        return owner

    def _collect_one(self, collect):
        seq = self._owned.popleft()
        out = self.engine.shard_fetch(seq)
        head = out[0] if isinstance(out, tuple) else out
        if getattr(head, "xchg_timeout", 0):
            raise RuntimeError(f"chunk {seq}: the exchange failed on rank {self.rank} (code {head.xchg_timeout}: 1 = a "
                               "peer's rows or the chunk never arrived, 2 = a flag overran); its results are not valid")
        self.results[seq] = collect(seq, out) if collect is not None else out

    def drain(self, collect=None):
        """Collect the results of every owned chunk still in flight, in chunk order (synchronises with the device)."""
        while self._owned:
            self._collect_one(collect)
        return self.results


class OrderedStitcher:
    """Exact bit post-processing of a stream whose chunks are finished on different ranks.

    ``checkSymbolOverlap`` (dem_base:863-988) realigns a chunk by one bit against the carry of the chunk before it
    (``poswinP`` / ``posSymEnd``, dem_base:977-979).  That carry is a function of the earlier chunk's own symbols only,
    so the owner of chunk ``c`` can produce it without knowing anything about chunk ``c - 1``: it post-processes its
    chunk once to obtain the carry and sends it to the owner of chunk ``c + 1`` straight away, then takes the carry of
    chunk ``c - 1`` from that chunk's owner and post-processes again, now with the right alignment.  The bits that come
    out are exactly what one process stitching every chunk in order produces; no rank ever waits on a chain longer than
    one message.  When a rank owns consecutive chunks the carry is already in place and one pass is enough.

    ``stitcher`` has the interface of ``_native.Stitcher`` (``__call__``, ``get_state``, ``set_state``, ``reset``);
    ``owner_of_chunk(c)`` gives the rank that finishes global chunk ``c``; ``send(token, dst, c)`` and
    ``recv(src, c)`` move the carry of chunk ``c`` (``bytes``) between ranks (e.g. ``torch.distributed`` point-to-point
    on a ``gloo`` group)."""

    def __init__(self, stitcher, rank, owner_of_chunk, send, recv, first_chunk=0):
        self.stitcher, self.rank, self.owner_of_chunk = stitcher, rank, owner_of_chunk
        self.send, self.recv = send, recv
        self.first = first_chunk     # the chunk the stream starts with: it has no carry to wait for
        self._last = None            # chunk whose carry the stitcher currently holds
        self._local = {}             # carries of chunks this rank owns whose successor it owns too
        stitcher.reset()

    def __call__(self, c, sym, centre, mag, clipped, sp_sym):
        if self.owner_of_chunk(c) != self.rank:
            raise ValueError(f"chunk {c} belongs to rank {self.owner_of_chunk(c)}")
        st = self.stitcher
        if c < self.first:
            raise ValueError(f"chunk {c} precedes the first chunk of the stream ({self.first})")
        if c > self.first and self._last != c - 1 and self.owner_of_chunk(c - 1) == self.rank:
            st.set_state(self._local.pop(c - 1))
            self._last = c - 1
        in_place = c == self.first or self._last == c - 1
        if c == self.first:
            st.reset()
        out = st(sym, centre, mag, clipped, sp_sym)          # exact if the right carry was in place; it is the carry pass otherwise
        token = st.get_state()
        nxt = self.owner_of_chunk(c + 1)
        if nxt == self.rank:
            self._local[c] = token
        else:
            self.send(token, nxt, c)
        if not in_place:
            st.set_state(self.recv(self.owner_of_chunk(c - 1), c - 1))
            out = st(sym, centre, mag, clipped, sp_sym)
        self._last = c
        self._local.pop(c - 2, None)
        return out


class ShardedBitStream:
    """Host samples in on rank 0, the stitched bit stream out, on N GPUs: ``ShardedStream`` for the device work and an
    ``OrderedStitcher`` for the bit post-processing of the chunks this rank owns.

        bs = ShardedBitStream(stream, stitcher, rank, world, send, recv)
        for block in blocks:                              # the loop runs on EVERY rank; only rank 0 touches samples
            buf = bs.host_slot()                          # rank 0: pinned slot of the next chunk (None elsewhere)
            if buf is not None:
                buf[:ovl] = previous tail; buf[ovl:] = block      # demodulator_process.py:287,337
            bs.submit()
        bs.finish()
        bs.bits                                           # {chunk number: (bits, centres, trust)} of the owned chunks

    ``send(token, dst, c)`` / ``recv(src, c)`` move the carry of chunk ``c`` (<= ``Stitcher.state_capacity`` bytes)
    between ranks.  The last chunk's carry is addressed to the owner of a chunk that never comes; ``finish`` takes it off
    the wire."""

    def __init__(self, stream, stitcher, rank, world, send, recv, first_chunk=None):
        self.sh, self.rank, self.world = stream, rank, world
        self.first = stream.next_seq if first_chunk is None else first_chunk
        self._recv = recv
        self.ordered = OrderedStitcher(stitcher, rank, self.owner_of_chunk, send, recv, self.first)
        self.bits = {}
        self.info = {}                 # {chunk number: the engine's result block}

    def owner_of_chunk(self, c):
        return owner_of(c, self.world)

    @property
    def next(self):
        return self.sh.next_seq

    def host_slot(self):
        return self.sh.host_slot()

    def _collect(self, c, out):
        res, sym, centre, mag = out[0], out[2], out[3], out[4]
        self.info[c] = res
        self.bits[c] = self.ordered(c, sym, centre, mag, (), float(res.sp_sym))
        return c

    def submit(self, src=None, kind=SRC_HOST):
        """Enqueue the next chunk (H2D on the ingest rank included).  Returns its chunk number."""
        c = self.sh.next_seq
        self.sh.submit(src, kind, collect=self._collect)
        return c

    def drain(self):
        """Finish every owned chunk in flight (in chunk order); the stream can go on afterwards."""
        self.sh.drain(self._collect)

    def finish(self):
        self.drain()
        last = self.sh.next_seq - 1
        if last >= self.first and self.owner_of_chunk(last + 1) == self.rank and self.owner_of_chunk(last) != self.rank:
            self._recv(self.owner_of_chunk(last), last)
        return self.bits


def gather_results(results, world, gather_object, rank):
    """Merge the per-rank ``{seq: result}`` dicts on rank 0, ordered by chunk. ``gather_object(obj)`` returns the
    list of all ranks' objects on rank 0 (``None`` elsewhere)."""
    parts = gather_object(results)
    if rank != 0:
        return None
    merged = {}
    for part in parts:
        for seq, r in part.items():
            if seq in merged:
                raise RuntimeError(f"chunk {seq} reported by two ranks")
            merged[seq] = r
    return [merged[s] for s in sorted(merged)]

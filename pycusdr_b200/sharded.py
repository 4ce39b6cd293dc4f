"""Doppler-bin sharding of the search across GPUs, one process per GPU.

Every rank holds the chunk and the full shift table and searches a contiguous slice of the Doppler bins.  The
three per-(bin, mask) tables of the slice (energy, peak value, peak offset) are stored by the search kernels
directly into the exchange region of the chunk's *owner* rank over NVLink peer memory
(``pcs_enqueue_search_push``); the owner alone then runs the tail of the chunk -- Doppler estimate, demodulation at
the selected bin, timing recovery, symbol decisions -- on the gathered table, which is the very table a single GPU
would have produced, so the results are bit-identical to the unsharded path.  Owners rotate round-robin over the
chunks, so the part of the work that does not shard costs each rank 1/world of a chunk, and no collective is on
the per-chunk path.  ``torch.distributed`` (any backend) is used once, to exchange the 64-byte CUDA IPC handles,
and at the end to gather the per-chunk results on rank 0.

The classes here are host logic only: they drive an *engine* object with the interface of
``pycusdr_b200._native.Engine`` (``set_bin_range``, ``peer_export``, ``peer_attach``, ``upload_device``,
``enqueue_search_push``, ``enqueue_owner_tail``, ``fetch``), which is what lets the world_size-2 ``gloo`` test in
``tests/test_sharded_cpu.py`` run them without a GPU.
"""
from collections import deque


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


class ShardedSearch:
    """Per-rank driver. ``all_gather`` is a callable ``obj -> [obj of rank 0, obj of rank 1, ...]`` (e.g. a thin
    wrapper over ``torch.distributed.all_gather_object``)."""

    def __init__(self, engine, rank, world, all_gather):
        self.engine, self.rank, self.world = engine, rank, world
        self.slices = bin_partition(engine.D, world)
        lo, hi = self.slices[rank]
        engine.set_bin_range(lo, hi)
        handles = all_gather(engine.peer_export())
        if len(handles) != world:
            raise RuntimeError("handle exchange returned %d entries for %d ranks" % (len(handles), world))
        engine.peer_attach(rank, world, handles)
        self._owned = deque()          # chunks whose tail this rank has enqueued but not collected yet
        self.results = {}              # seq -> whatever ``collect`` made of engine.fetch()
        self._next_seq = 0

    def enqueue(self, seq, chunk=None, collect=None):
        """Enqueue chunk ``seq``. ``chunk`` is a device pointer / array understood by the engine, or ``None`` when the
        samples have been written into the engine's pinned host buffer (``engine.host_buffer``): the H2D copy is then
        part of the enqueued work, as in ``uploadToGPU`` (dem_base:548-558).  Chunks must come in order.
        If this rank owns the chunk its tail is enqueued too; the results of the previously owned chunk are collected
        first (the engine has one result staging area), through ``collect(fetch_tuple)`` if given."""
        if seq != self._next_seq:
            raise ValueError(f"chunks must be enqueued in order: expected {self._next_seq}, got {seq}")
        self._next_seq += 1
        owner = owner_of(seq, self.world)
        if owner == self.rank:
            self.drain(collect)
        if chunk is None:
            self.engine.upload()
        else:
            self.engine.upload_device(chunk)
        self.engine.enqueue_search_push(seq, owner)
        if owner == self.rank:
            self.engine.enqueue_owner_tail(seq)
            self._owned.append(seq)
        return owner

    @property
    def next_seq(self):
        """Sequence number the next ``enqueue`` must carry."""
        return self._next_seq

    def drain(self, collect=None):
        """Collect the results of every owned chunk still in flight (synchronises with the device)."""
        while self._owned:
            seq = self._owned.popleft()
            out = self.engine.fetch()
            head = out[0] if isinstance(out, tuple) else out
            if getattr(head, "xchg_timeout", 0):
                raise RuntimeError(f"chunk {seq}: a peer's rows never reached rank {self.rank} (exchange timed out); "
                                   "its results are not valid")
            self.results[seq] = collect(out) if collect is not None else out
        return self.results


class ShardedPipelines:
    """``P`` independent ``ShardedSearch`` pipelines per rank (one engine, pair of CUDA streams and exchange region
    each): chunk ``c`` goes to pipeline ``c % P`` as its local chunk ``c // P``.  With P = 2 the search kernel of one
    chunk fills the SMs that the last, partially filled wave and the reduction kernels of the previous chunk leave
    idle -- the same "two chunks in flight" the single-GPU path uses."""

    def __init__(self, engines, rank, world, all_gather):
        self.pipes = [ShardedSearch(e, rank, world, all_gather) for e in engines]
        self.rank, self.world = rank, world
        self.slices = self.pipes[0].slices

    def enqueue(self, chunk_no, chunk=None, collect=None):
        P = len(self.pipes)
        return self.pipes[chunk_no % P].enqueue(chunk_no // P, chunk, collect)

    def drain(self, collect=None):
        """Collect every owned chunk still in flight, in increasing chunk order across the pipelines (each holds at most
        one).  The order matters to a collector that carries state from chunk to chunk across ranks (OrderedStitcher):
        pipeline-by-pipeline draining can leave two ranks waiting for each other's carries.  ``collect`` may be one
        callable or one per pipeline."""
        P = len(self.pipes)
        order = sorted((p._owned[0] * P + j, j) for j, p in enumerate(self.pipes) if p._owned)
        for _, j in order:
            self.pipes[j].drain(collect[j] if isinstance(collect, (list, tuple)) else collect)
        return self.results

    @property
    def chunks_enqueued(self):
        """Number of chunks enqueued so far = the chunk number the next ``enqueue`` must carry."""
        return sum(p.next_seq for p in self.pipes)

    @property
    def results(self):
        """{global chunk number: result} of the chunks this rank owned."""
        P = len(self.pipes)
        return {q * P + j: r for j, p in enumerate(self.pipes) for q, r in p.results.items()}


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
    """Host samples in, the stitched bit stream out, on N GPUs: ``ShardedPipelines`` for the device work and an
    ``OrderedStitcher`` for the bit post-processing of the chunks this rank owns.

        bs = ShardedBitStream(pipelines, stitcher, rank, world, send, recv, first_chunk=pipelines.chunks_enqueued)
        for block in blocks:                              # on EVERY rank, the same samples
            buf = bs.next_buffer()                        # pinned chunk buffer of the pipeline the next chunk goes to
            buf[:ovl] = previous tail; buf[ovl:] = block  # (after that pipeline's earlier H2D copy has completed)
            bs.submit()
        bs.finish()
        bs.bits                                           # {chunk number: (bits, centres, trust)} of the owned chunks

    ``send(token, dst, c)`` / ``recv(src, c)`` move the ≤ 1 KB carry of chunk ``c`` between ranks.  The last chunk's carry
    is addressed to the owner of a chunk that never comes; ``finish`` takes it off the wire."""

    def __init__(self, pipelines, stitcher, rank, world, send, recv, first_chunk=0):
        self.sh, self.rank, self.world = pipelines, rank, world
        self.K = len(pipelines.pipes)
        self.first = self.next = first_chunk
        self._recv = recv
        self.ordered = OrderedStitcher(stitcher, rank, self.owner_of_chunk, send, recv, first_chunk)
        self._owned = [deque() for _ in range(self.K)]
        self.bits = {}

    def owner_of_chunk(self, c):
        return owner_of(c // self.K, self.world)          # ShardedPipelines: chunk c = local chunk c // K of pipeline c % K

    @property
    def next_pipe(self):
        return self.next % self.K

    def next_buffer(self):
        return self.sh.pipes[self.next_pipe].engine.host_buffer

    def _collector(self, pipe):
        def collect(out):
            c = self._owned[pipe].popleft()
            res, _, sym, centre, mag = out
            self.bits[c] = self.ordered(c, sym, centre, mag, (), float(res.sp_sym))
            return c
        return collect

    def submit(self):
        """The next chunk's samples are in ``next_buffer()``: enqueue it (H2D included).  Returns its chunk number."""
        c, pipe = self.next, self.next_pipe
        if self.owner_of_chunk(c) == self.rank:
            self._owned[pipe].append(c)
        self.sh.enqueue(c, None, collect=self._collector(pipe))
        self.next += 1
        return c

    def drain(self):
        """Finish every owned chunk in flight (in chunk order); the stream can go on afterwards."""
        self.sh.drain([self._collector(j) for j in range(self.K)])

    def finish(self):
        self.drain()
        last = self.next - 1
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

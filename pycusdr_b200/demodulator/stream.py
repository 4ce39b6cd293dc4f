"""Streaming front end of the UHF demodulator: the work of the reference's ``SigFIFO`` ring
buffer and of the chunk loop in ``demodulator_process.py:284-338`` done natively (``pcs_ingest_*``).

    dem = StreamDemodulator(conf, protocol, radioName, inflight=2)
    for block in zmq_blocks:                 # any block size
        for out in dem.push(block):          # zero or more finished chunks, in order
            demodOut.send_pyobj({**header, **out})
    for out in dem.flush(): ...

Every element is the per-chunk part of the dict the reference's process loop sends to the decoder
(``demodulator_process.py:259-309``): ``doppler``, ``doppler_std``, ``SNR``, ``data`` (uint8 bits), ``trust``,
``spSymEst``.  The numbers are identical to feeding the same samples through ``uploadAndFindCarrier`` /
``demodulate`` chunk by chunk with the overlap carry of ``demodulator_process.py:337``; what changes is that the
overlap never returns to the host, the H2D copy of a chunk overlaps the kernels of the previous ones, and
``inflight`` chunks are processed concurrently (one handle, CUDA stream and graph each).
"""
import numpy as np

from .. import _native
from .UHF import Demodulator as UHFDemodulator


class StreamDemodulator(UHFDemodulator):

    def __init__(self, conf, protocol, radioName, inflight=2, **kw):
        super().__init__(conf, protocol, radioName, **kw)
        if inflight < 1:
            raise ValueError("inflight must be >= 1")
        self._engines = [self._engine] + [_native.Engine(**self._engine_kwargs) for _ in range(inflight - 1)]
        self._ingest = _native.Ingest(self._engines, self.Nfft, self.sigOverlap, device=self._engine_kwargs["device"])

    def __del__(self):
        ing = getattr(self, "_ingest", None)
        if ing is not None:
            ing.close()
        for e in getattr(self, "_engines", [])[1:]:
            e.close()
        super().__del__()

    def _finish(self, out):
        res, E, sym, centre, mag, sig, noise = out
        doppler, doppler_std, _, SNR = self._finish_search(res, E, (sig, noise))
        self._pending = (res, sym, centre, mag)
        bits, centres, trust, spSym = self._Demodulator__demodulate()
        return {"doppler": doppler, "doppler_std": doppler_std, "SNR": SNR, "data": bits, "centres": centres,
                "trust": trust, "spSymEst": spSym}

    def push(self, samples):
        """Append complex64 samples; returns the chunks that have finished so far (possibly none)."""
        self._ingest.push(samples)
        done = []
        while True:
            out = self._ingest.pop(block=False)
            if out is None:
                return done
            done.append(self._finish(out))

    def flush(self):
        """Wait for every submitted chunk (samples short of a full chunk stay buffered)."""
        done = []
        while self._ingest.pending() > 0:
            done.append(self._finish(self._ingest.pop(block=True)))
        return done

"""Synthetic workloads of SURVEY.md 8(d) (BASELINE.json configs C1..C5): one place for bench.py, the parity tests and the
tools so that "the benchmarked configuration" and "the tested configuration" are the same bytes.

    conf, modulation, desc = load_workload("c2")
    stream = build_stream(conf, modulation, n_chunks, seed=2)        # complex64 samples, n_chunks * (N - overlap)
    chunks = chunks_from_stream(stream, N, overlap, n_chunks)        # what the chunk loop hands the demodulator

The chunk loop is the reference's (demodulator_process.py:287,337): chunk c = last 2^overlap samples of chunk c-1 followed
by block c of the stream; the first chunk's overlap is zero-filled."""
import hashlib
import os

import numpy as np

from ..config import loadModularJson
from . import signals as S

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
RADIO = "UHF-H"

WORKLOADS = {
    # name: (config file, modulation, description)
    "c1": ("CC11xx.json", None, "C1 CC11xx FSK-2 7416 baud x128, N=2^16, D=64, M=8"),
    "c2": ("c2_base_2p18_256bins.json", "GMSK", "C2 GMSK 9600 baud x16, N=2^18, D=256, M=8"),
    "c3": ("benchmark/bench_GMSK.json", "GMSK", "C3 bench_GMSK, N=2^15, D=64, M=8"),
    "c4": ("c4_sband_2p20_4096bins.json", "GMSK", "C4 wide search, N=2^20, D=4096, M=8"),
}


def load_workload(name):
    cfg_file, modulation, desc = WORKLOADS[name]
    return loadModularJson(os.path.join(ROOT, "config", cfg_file)), modulation, desc


def geometry(conf):
    """(Nfft, overlap, new samples per chunk, sample rate) of a configuration."""
    cg, cr = conf["GPU"]["UHF"], conf["Radios"]["Rx"][RADIO]
    N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
    return N, ovl, N - ovl, cr["baud"] * cr["samplesPerSym"]


def build_stream(conf, modulation, n_chunks, seed):
    """Synthetic sample stream for ``n_chunks`` chunks (SURVEY 8(d) inputs): back-to-back benchmark packets (or, for
    ``modulation is None``, the C1 CC11xx-style FSK-2 packet mixed to the radio's offset + 7 kHz) plus AWGN."""
    cr = conf["Radios"]["Rx"][RADIO]
    N, ovl, step, fs = geometry(conf)
    need = n_chunks * step
    sps, baud = cr["samplesPerSym"], cr["baud"]
    rng = np.random.RandomState(seed)
    if modulation is None:      # C1: FSK-2 packet, CC11xx style, Es/N0 15 dB
        bits = S.createBitSequence(400, seed=123)
        sig = S.modulateFSK(bits, sps)
        one = np.concatenate((np.zeros(4096, np.complex64), sig, np.zeros(4096, np.complex64)))
        f0 = cr["frequencyOffset_Hz"] + 7000.0
        snr_r = 15 - 10 * np.log10(sps)
    else:
        one, _ = S.get_padded_packet(modulation, sps, fs, offset_freq=cr["frequencyOffset_Hz"])
        one = one.astype(np.complex64)
        f0 = None
        snr_r = S.bench_snr_to_awgn_snr(modulation, 12.0, baud, fs)
    reps = need // len(one) + 1
    clean = np.tile(one, reps)[:need]
    if f0 is not None:
        clean = clean * np.exp(2j * np.pi * f0 / fs * np.arange(need)).astype(np.complex64)
    p_sig = np.mean(np.abs(one) ** 2)
    noise_p = p_sig * 10 ** (-snr_r / 10)
    out = np.empty(need, dtype=np.complex64)
    amp = np.float32(np.sqrt(noise_p / 2))
    for a in range(0, need, 1 << 22):       # blockwise: keeps the float64 temporaries small
        n = min(1 << 22, need - a)
        out[a:a + n] = clean[a:a + n] + amp * (rng.randn(n) + 1j * rng.randn(n))
    return out


def chunks_from_stream(stream, N, ovl, n_chunks):
    """[n_chunks, N] array: chunk c = overlap tail of chunk c-1 + new block c (demodulator_process.py:287,337)."""
    step = N - ovl
    out = np.zeros((n_chunks, N), dtype=np.complex64)
    for c in range(n_chunks):
        lo = c * step - ovl
        if lo < 0:
            out[c, ovl:] = stream[:step]
        else:
            out[c] = stream[lo:lo + N]
    return out


class BitsDigest:
    """Running SHA-256 over the per-chunk outputs a demodulator hands to decoder_process (bits) plus the spectrum shift
    and timing bin of every chunk: both arms of bench.py print it over the same chunk indices, so "same results as the
    reference" is checkable from the two JSON lines alone."""

    def __init__(self):
        self._h = hashlib.sha256()
        self.chunks = 0
        self.bits = 0
        self.shift_sum = 0

    def add(self, bits, shift, timing_bin=None):
        b = np.ascontiguousarray(bits, dtype=np.uint8)
        self._h.update(np.int64(len(b)).tobytes())
        self._h.update(b.tobytes())
        self._h.update(np.int64(int(shift)).tobytes())
        if timing_bin is not None:
            self._h.update(np.int64(int(timing_bin)).tobytes())
        self.chunks += 1
        self.bits += len(b)
        self.shift_sum += int(shift)

    def hexdigest(self):
        return self._h.hexdigest()[:16]

    def summary(self):
        return {"sha": self.hexdigest(), "chunks": self.chunks, "bits": self.bits, "shift_sum": self.shift_sum}

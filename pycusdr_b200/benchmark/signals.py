"""Synthetic test-signal generators for benchmarks and tests.

Restates the generators of the reference's BER benchmark, ``examples/benchmark/create_signals.py``
(``createBitSequence`` :13-26, ``modulateBPSK`` :45-60, ``modulateFSK`` :64-79, ``modulateGFSK2``
:84-98, ``modulateGMSK`` :101-113, ``awgn`` :116-142, ``get_padded_packet`` :179-201) and the
SNR -> noise scaling of ``bench_modem.py:203-242``.  ``tests/test_oracle_golden.py`` checks them
against vectors produced by the reference code itself.
"""
import numpy as np

from ..lib.filters import gaussianFilter, rrcosfilter


def createBitSequence(n_bits, seed=None):
    if seed:
        state = np.random.get_state()
        np.random.seed(seed)
    bits = np.random.randint(0, 2, n_bits)
    if seed:
        np.random.set_state(state)
    return bits


def packetData():
    return createBitSequence(10000, seed=123)


def encodeNRZS(bits):
    out = np.zeros(len(bits), dtype=np.uint8)
    out[0] = bits[0]
    for i in range(1, len(bits)):
        out[i] = out[i - 1] if bits[i] == 1 else (~out[i - 1] & 1)
    return out


def modulateBPSK(raw_bits, sps):
    nrzs = encodeNRZS(np.concatenate(([1, 0, 1], raw_bits))).astype(float) * 2 - 1
    taps = rrcosfilter(0.5, 6, sps)
    taps = taps / np.sum(taps)
    return np.convolve(taps, np.repeat(nrzs, sps)).astype(np.complex64)


def modulateFSK(raw_bits, sps):
    step = np.ones(sps) / sps * np.pi
    lut = np.array([-step, step])
    phase = np.cumsum(lut[raw_bits]) - (raw_bits[0] * 2 - 1) * np.pi / 2
    return np.exp(1j * np.mod(phase, 2 * np.pi)).astype(np.complex64)


def modulateGFSK2(raw_bits, sps):
    taps = gaussianFilter(1, 1, sps, 4 * sps)
    phase = np.convolve(taps, np.repeat(raw_bits * 2 - 1, sps))
    return np.exp(1j * np.cumsum(phase) / sps * np.pi).astype(np.complex64)


def modulateGMSK(raw_bits, sps):
    taps = gaussianFilter(1, 0.5, sps, 4 * sps)
    phase = np.convolve(taps, np.repeat(raw_bits * 2 - 1, sps))
    return np.exp(1j * np.cumsum(phase) / sps * np.pi / 2).astype(np.complex64)


MODULATORS = {"BPSK": modulateBPSK, "GMSK": modulateGMSK, "FSK": modulateFSK, "GFSK": modulateGFSK2}


def awgn(sig, snr, measured=True):
    """Adds white Gaussian noise from the *global* NumPy RNG (seed it before calling)."""
    if measured:
        sigp = 10 * np.log10(np.linalg.norm(np.abs(sig), 2) ** 2 / len(sig))
        snr = snr - sigp
    noiseP = 10 ** (-snr / 10)
    if np.iscomplexobj(sig):
        return sig + np.sqrt(noiseP / 2) * (np.random.randn(len(sig)) + 1j * np.random.randn(len(sig)))
    return sig + np.sqrt(noiseP) * np.random.randn(len(sig))


def get_padded_packet(modulation, spSym=16, fs=9600 * 16, offset_freq=None, raw_bits=(), pad=10000):
    """Packet padded with ``pad`` zeros on both sides and mixed to ``offset_freq`` (default fs/4)."""
    if offset_freq is None:
        offset_freq = fs / 4
    if len(raw_bits) == 0:
        raw_bits = packetData()
    if modulation not in MODULATORS:
        raise TypeError("Only supports GMSK, FSK, GFSK and BPSK")
    sig = MODULATORS[modulation](np.asarray(raw_bits), spSym)
    full = np.concatenate((np.zeros(pad), sig, np.zeros(pad)))
    full = full * np.exp(1j * 2 * np.pi * offset_freq / fs * np.arange(len(full)))
    return full, raw_bits


BENCH_BW = {  # occupied bandwidth used for the SNR -> Eb/N0 bookkeeping (bench_modem.py:203-209)
    "GMSK": lambda baud: baud / 0.7,
    "BPSK": lambda baud: baud * 1.5,
    "FSK": lambda baud: 2 * baud + 2 * (baud / 2),
    "GFSK": lambda baud: 2 * baud + 2 * (baud / 2),
}


def bench_snr_to_awgn_snr(modulation, snr_db, baud=9600, fs=9600 * 16):
    """SNR_r handed to awgn() for a bench "SNR" (bench_modem.py:225-242)."""
    return snr_db + 10 * np.log10(BENCH_BW[modulation](baud) / fs)


def ebn0_db(modulation, snr_db, baud=9600):
    """bench_modem.py:249."""
    return snr_db + 10 * np.log10(BENCH_BW[modulation](baud) / baud)


def doppler_rate(sig, rate_hz_per_s, fs):
    """Extension: linear Doppler rate applied as exp(j*pi*r*t^2)."""
    t = np.arange(len(sig)) / fs
    return sig * np.exp(1j * np.pi * rate_hz_per_s * t * t)


def bench_stream(modulation, snr_db, n_packets=1, seed=1000, spSym=16, baud=9600, offset_freq=None,
                 rate_hz_per_s=0.0, pre_blocks=5, chunk=2 ** 14):
    """The sample stream ``bench_modem.SendSignal.sendToModem`` publishes (bench_modem.py:57-104):
    ``pre_blocks`` chunks of real noise, then ``n_packets`` padded packets each with its own AWGN
    realisation (power measured over the padded packet), then ``pre_blocks`` chunks of noise again.
    Deterministic for a given ``seed``; the global NumPy RNG state is restored."""
    fs = baud * spSym
    sig, bits = get_padded_packet(modulation, spSym, fs, offset_freq)
    sig = sig.astype(np.complex64)
    if rate_hz_per_s:
        sig = doppler_rate(sig, rate_hz_per_s, fs).astype(np.complex64)
    snr_r = bench_snr_to_awgn_snr(modulation, snr_db, baud, fs)
    state = np.random.get_state()
    np.random.seed(seed)
    parts = [(np.sqrt(0.1) * np.random.randn(chunk)).astype(np.complex64) for _ in range(pre_blocks)]
    for _ in range(n_packets):
        parts.append(awgn(sig, snr_r).astype(np.complex64))
    parts += [(np.sqrt(0.1) * np.random.randn(chunk)).astype(np.complex64) for _ in range(pre_blocks)]
    np.random.set_state(state)
    return np.concatenate(parts), bits

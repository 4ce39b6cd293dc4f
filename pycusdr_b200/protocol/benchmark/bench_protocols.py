"""Benchmark protocols (demodulator-facing part): GMSK, FSK, GFSK, BPSK.

Reference: ``pyCuSDR/protocol/benchmark/bench_GMSK.py:40-80``, ``bench_FSK.py:33-81``,
``bench_GFSK.py:42-90``, ``bench_BPSK.py:47-75,84-199,245-262``.  All four sum the masks
before the Doppler search (``SUM_ALL_MASKS_PYTHON = True``).
"""
import numpy as np
from scipy import signal
from ...lib.filters import gmskMod, rrcosfilter
from ..protocolBase import ProtocolBase
from ..FSK2_base import fsk_phase_templates


class Bench_base(ProtocolBase):
    name = "bench_base_class"
    SUM_ALL_MASKS_PYTHON = True

    def get_symbolLUT2(self, maskLen):
        return self._centre_bit_LUT(maskLen), []


def _hamming_weight(templates):
    w = signal.get_window("hamming", len(templates[0]))
    return [t * w for t in templates]


class Bench_GMSK(Bench_base):
    name = "bench_GMSK"

    def get_filter(self, Nfft, spSym, maskSize):
        templates = []
        for m in self._get_xcorrMasks(maskSize):
            wave, _, n = gmskMod(m, spSym)
            templates.append(wave[n // 2:-n // 2 + 1])
        masks = self._pad_and_conj_fft(_hamming_weight(templates), Nfft)
        return masks.shape[0], masks


class Bench_FSK(Bench_base):
    name = "bench_FSK"

    def get_filter(self, Nfft, spSym, maskSize):
        # +-pi radians per symbol (bench_FSK.py:44): same as FSK2 with nCycles = 0.5
        templates = fsk_phase_templates(self._get_xcorrMasks(maskSize), spSym, 0.5)
        masks = self._pad_and_conj_fft(templates, Nfft)
        return masks.shape[0], masks


class Bench_GFSK(Bench_base):
    """The reference's bench_GFSK overrides GFSK2.get_filter with the plain FSK bank
    (``bench_GFSK.py:42-73``, Hamming weighting commented out at ``:63``)."""
    name = "bench_GFSK"

    def get_filter(self, Nfft, spSym, maskSize):
        templates = fsk_phase_templates(self._get_xcorrMasks(maskSize), spSym, 0.5)
        masks = self._pad_and_conj_fft(templates, Nfft)
        return masks.shape[0], masks


def _nrzs_lut(maskLen):
    """symbolLUT[s, 0, :] = successors of s that decode to 1, [s, 1, :] = to 0
    (reference ``bench_BPSK.py:84-199``)."""
    if maskLen == 5:
        same, flip = ([0, 1, 2, 3], [4, 5, 6, 7]), ([12, 13, 14, 15], [8, 9, 10, 11])
        rows = [same] * 4 + [flip] * 8 + [same] * 4
    elif maskLen == 4:
        same, flip = ([0, 1], [2, 3]), ([6, 7], [4, 5])
        rows = [same] * 2 + [flip] * 4 + [same] * 2
    else:
        raise Exception(f"bench_BPSK: Invalid mask length ({maskLen})")
    return np.array(rows, dtype=int)


class Bench_BPSK(Bench_base):
    name = "bench_BPSK"

    def get_filter(self, Nfft, spSym, maskSize):
        self.num_masks = int(2 ** (maskSize - 1))
        taps = rrcosfilter(0.5, 6, spSym)
        taps = taps / np.sum(taps)
        n = len(taps)
        templates = []
        for m in self._get_xcorrMasks(maskSize) * 2 - 1:
            shaped = np.convolve(np.repeat(m, spSym), taps)
            templates.append(shaped[n // 2:-n // 2 + 1])
        masks = self._pad_and_conj_fft(templates, Nfft)
        return masks.shape[0], masks

    def get_symbolLUT2(self, maskLen):
        return None, _nrzs_lut(maskLen)

"""GFSK-2 matched-filter bank (reference ``pyCuSDR/protocol/GFSK2_base.py:22-61``)."""
import numpy as np
from scipy import signal
from ..lib.filters import gaussianFilter
from .protocolBase import ProtocolBase

BT = 1.0  # bandwidth-time product of the Gaussian pulse


class GFSK2(ProtocolBase):
    name = "GFSK2 Base"

    def get_filter(self, Nfft, spSym, maskSize):
        taps = gaussianFilter(1, BT, spSym, 4 * spSym) * np.pi / spSym  # half a period / symbol
        n = len(taps)
        templates = []
        for m in self._get_xcorrMasks(maskSize):
            phase = np.convolve(np.repeat(m * 2 - 1, spSym), taps)
            wave = np.exp(1j * np.cumsum(phase))
            templates.append(wave[n // 2:-n // 2 + 1])
        self._weight_filters(templates)
        masks = self._pad_and_conj_fft(templates, Nfft)
        return masks.shape[0], masks

    def _weight_filters(self, filters):
        w = signal.get_window("hamming", len(filters[0]))
        for i in range(len(filters)):
            filters[i] = filters[i] * w

"""FSK-2 matched-filter bank (reference ``pyCuSDR/protocol/FSK2_base.py:17-46``)."""
import numpy as np
from .protocolBase import ProtocolBase


def fsk_phase_templates(patterns, spSym, nCycles):
    """Continuous-phase FSK templates exp(j*phase) for each +-1 bit pattern.

    Each symbol ramps the phase by ``+-2*pi*nCycles``; the first symbol starts at ``-+pi/2``.
    The ramp restarts from the *last sample* of the previous symbol plus one step, exactly as
    the reference builds it (``FSK2_base.py:27-36``, ``bench_FSK.py:46-58``).
    """
    ramp = np.linspace(1 / spSym, 1, spSym) * np.pi * 2 * nCycles
    out = []
    for p in patterns:
        p = p * 2 - 1
        ph = np.empty(len(p) * spSym)
        ph[:spSym] = p[0] * ramp - p[0] * np.pi / 2
        for j in range(1, len(p)):
            ph[j * spSym:(j + 1) * spSym] = ph[j * spSym - 1] + p[j] * ramp
        out.append(np.exp(1j * ph))
    return out


class FSK2(ProtocolBase):
    name = "FSK2 Basee"

    def get_filter(self, Nfft, spSym, maskSize, nCycles=0.5):
        templates = fsk_phase_templates(self._get_xcorrMasks(maskSize), spSym, nCycles)
        masks = self._pad_and_conj_fft(templates, Nfft)
        return masks.shape[0], masks

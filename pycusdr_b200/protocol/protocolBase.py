"""Filter/LUT side of the pyCuSDR protocol plugin interface.

Only the part of ``ProtocolBase`` the demodulator hot path calls is provided
(reference ``pyCuSDR/protocol/protocolBase.py:27-75``; call sites
``demodulator/demodulator_base.py:123-127,196,207``):

* ``get_filter(Nfft, spSym, maskSize) -> (num_masks, complex64[num_masks, Nfft])``
  the conjugated spectra of the zero-padded matched-filter templates,
* ``get_symbolLUT2(maskSize) -> (bitLUT | None, symbolLUT)``,
* ``name`` and the optional ``SUM_ALL_MASKS_PYTHON`` flag.

Framing, packet and modulator methods belong to the decoder / Tx side and are out of
scope; a reference protocol object can be passed to
``pycusdr_b200.demodulator`` unchanged because only the members above are touched.
"""
import numpy as np


class ProtocolBase:
    name = "ProtocolBase"

    def __init__(self, **args):
        self.conf = args.get("conf")

    def get_filter(self, Nfft, spSym=None, maskSize=0):
        raise NotImplementedError("Sub class needs to implement this method")

    def get_symbolLUT2(self, maskLen):
        raise NotImplementedError("Sub class needs to implement this method")

    # -- helpers shared by the concrete protocols -------------------------------------
    @staticmethod
    def _get_xcorrMasks(maskLen):
        """All 2**maskLen bit patterns, MSB first, as a float array (row i = binary repr of i).

        Reference ``protocol/benchmark/bench_base.py:52-58`` / ``protocol/CC11xx.py:81-87``.
        """
        idx = np.arange(2 ** maskLen)[:, None]
        shifts = np.arange(maskLen - 1, -1, -1)[None, :]
        return ((idx >> shifts) & 1).astype(np.float64)

    @staticmethod
    def _centre_bit_LUT(maskLen):
        """bitLUT[m] = middle bit of pattern m (reference ``bench_GMSK.py:66-80``)."""
        return ProtocolBase._get_xcorrMasks(maskLen)[:, int(maskLen / 2)]

    @staticmethod
    def _pad_and_conj_fft(templates, Nfft):
        """conj(FFT_Nfft(template)) as complex64 rows (reference ``bench_GMSK.py:56-59``)."""
        out = np.empty((len(templates), Nfft), dtype=np.complex64)
        for i, tpl in enumerate(templates):
            out[i] = np.conj(np.fft.fft(tpl, Nfft)).astype(np.complex64)
        return out

"""CC11xx (FSK-2 / GFSK-2) protocol: demodulator-facing part only.

Reference ``pyCuSDR/protocol/CC11xx.py:30-112``: ``modIDX = 0`` selects the FSK2 filter
bank; all masks are summed before the Doppler search (``:55``).  Packet framing, CRC16 and
PN9 de-whitening (``:115`` onwards) belong to the decoder and are out of scope.
"""
import numpy as np
from .FSK2_base import FSK2


class CC11xx(FSK2):
    name = "CC11xx FSK-2"
    SUM_ALL_MASKS_PYTHON = True

    def get_symbolLUT2(self, maskLen):
        bitLUT = self._centre_bit_LUT(maskLen)
        half = np.arange(2 ** (maskLen - 1))
        symLUT = np.stack([half * 2 + 1, half * 2], axis=1).astype(int)
        return bitLUT, np.concatenate((symLUT, symLUT), axis=0)

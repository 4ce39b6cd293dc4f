"""UHF backend: forward FFT + Doppler search, then demodulation at the found bin
(reference pyCuSDR/demodulator/UHF.py:5-20; input thresholding is disabled there, :14)."""
from .demodulator_base import Demodulator as Demodulator_base


class Demodulator(Demodulator_base):

    def uploadAndFindCarrier(self, samples):
        if self.fused and self.one_call and self._stitch is not None:
            return self.chunkToBits(samples)      # same results, one native call for the whole chunk
        self.uploadToGPU(samples, direct=True)      # (no input thresholding on this backend: the samples are only read)
        return self.findUHF(samples)

    def demodulate(self):
        return self.demodulateUHF()

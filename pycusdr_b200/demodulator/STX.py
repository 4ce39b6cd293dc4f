"""STX backend: thresholded input, no Doppler search, demodulation at the fixed IF offset bin
(reference pyCuSDR/demodulator/STX.py:6-25, dem_base:758-761)."""
from .demodulator_base import Demodulator as Demodulator_base


class Demodulator(Demodulator_base):

    def uploadAndFindCarrier(self, samples):
        samples = self._as_chunk_buffer(samples)
        if self.native_threshold and self.Nfft <= 2 ** 22:
            self.thresholdAndUpload(samples)        # clipping on the device, same in-place result
        else:
            self.thresholdInput(samples)
            self.uploadToGPU(samples)
        return 0, 0, self.clippedPeakIPure, 0

    def demodulate(self):
        return self.demodulateSTX()

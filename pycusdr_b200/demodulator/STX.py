"""STX backend: thresholded input, no Doppler search, demodulation at the fixed IF offset bin
(reference pyCuSDR/demodulator/STX.py:6-25, dem_base:758-761)."""
from .demodulator_base import Demodulator as Demodulator_base


class Demodulator(Demodulator_base):

    def uploadAndFindCarrier(self, samples):
        samples = self._as_chunk_buffer(samples)
        if self.native_threshold and self.Nfft <= 2 ** 22:
            # opt-in (native_threshold=True): clipping on the device.  Same algorithm and summation order as the NumPy
            # statement, but |x| is CUDA's hypotf where the reference uses np.abs, whose complex64 SIMD kernels are not
            # correctly rounded and differ from host to host: clip levels and clipped samples agree to 1e-6 relative,
            # borderline clipped indices (and with them trust bytes) may differ.  The default keeps the NumPy statement.
            self.thresholdAndUpload(samples)
        else:
            self.thresholdInput(samples)
            self.uploadToGPU(samples)
        return 0, 0, self.clippedPeakIPure, 0

    def demodulate(self):
        return self.demodulateSTX()

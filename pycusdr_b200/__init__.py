"""pycusdr_b200 -- B200-native demodulator hot path of pyCuSDR behind the reference's own API.

``pycusdr_b200.demodulator.UHF.Demodulator`` / ``.STX.Demodulator`` are drop-ins for the reference's
``demodulator.UHF/STX.Demodulator`` (pyCuSDR/demodulator_process.py:21-36,242); the CUDA kernels are
reached through the C ABI in ``include/pycusdr_b200.h`` (``pycusdr_b200._native``).
"""
__version__ = "0.1.0"
LOG_NAME = "pyCuSDR"

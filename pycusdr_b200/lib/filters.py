"""Pulse-shaping filters used to build the matched-filter templates.

Host-side (NumPy, float64) helpers for the protocol plugins' ``get_filter``.
They restate the behaviour of the reference's ``pyCuSDR/lib/filters.py``
(``rrcosfilter`` :6-57, ``gaussianFilter`` :60-84) and ``pyCuSDR/lib/gmskmod.py``
(``gmskMod`` :10-43) so that the filter spectra handed to the CUDA path are
numerically the same as the reference's.  ``tests/test_protocols.py`` pins them
against vectors generated from the reference (``tests/golden/protocol_filters.npz``).
"""
import numpy as np

__all__ = ["rrcosfilter", "gaussianFilter", "gmskMod"]


def rrcosfilter(beta, span, spsym):
    """Root-raised-cosine taps, unit energy, ``span*spsym + 1`` taps.

    Mirrors reference ``lib/filters.py:6-57`` (a rcosdesign(..., 'sqrt') clone):
    closed forms at t = 0 and at the |4*beta*t| = 1 singularities, the general
    expression elsewhere.
    """
    half = span * spsym / 2
    t = np.arange(-half, half + 1) / spsym
    taps = np.zeros(t.size)

    at_zero = t == 0
    at_pole = np.abs(np.abs(4 * beta * t) - 1) < np.sqrt(np.finfo(float).eps)
    regular = ~(at_zero | at_pole)

    taps[at_zero] = -1 / (np.pi * spsym) * (np.pi * (beta - 1) - 4 * beta)
    if at_pole.any():
        taps[at_pole] = 1 / (2 * np.pi * spsym) * (
            np.pi * (beta + 1) * np.sin(np.pi * (beta + 1) / (4 * beta))
            - 4 * beta * np.sin(np.pi * (beta - 1) / (4 * beta))
            + np.pi * (beta - 1) * np.cos(np.pi * (beta - 1) / (4 * beta))
        )
    tr = t[regular]
    taps[regular] = (
        -4 * beta / spsym
        * (np.cos((1 + beta) * np.pi * tr) + np.sin((1 - beta) * np.pi * tr) / (4 * beta * tr))
        / (np.pi * ((4 * beta * tr) ** 2 - 1))
    )
    return taps / np.sqrt(np.sum(taps ** 2))


def gaussianFilter(gain, BT, spSym, nTaps):
    """Gaussian pulse taps normalised to ``1/gain`` DC gain (reference ``lib/filters.py:60-84``)."""
    a = np.sqrt(np.log(2) / 2) / BT
    t = np.linspace(-0.5 * nTaps, 0.5 * nTaps - 1, nTaps) / spSym
    taps = np.sqrt(np.pi) / a * np.exp(-(np.pi ** 2 * t ** 2) / a ** 2)
    taps /= np.sum(taps) * gain
    return taps


def gmskMod(bits, spSym, bw=0.5, nTaps=None, gain=1):
    """GMSK-modulate ``bits`` (0/1 or +-1). Returns (waveform, phase, filter length).

    Reference ``lib/gmskmod.py:10-43``: NRZ map, Gaussian-filter the upsampled bits with
    taps scaled to pi/2 per symbol, integrate the phase.
    """
    bits = np.asarray(bits)
    if not bits.min() < 0:
        bits = bits * 2 - 1
    if nTaps is None:
        nTaps = 4 * spSym
    taps = gaussianFilter(gain, bw, spSym, nTaps) * np.pi / 2 / spSym
    phase = np.cumsum(np.convolve(taps, np.repeat(bits, spSym)))
    return np.exp(1j * phase), phase, len(taps)

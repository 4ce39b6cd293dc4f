"""Drop-in for the reference's ``demodulator`` package (pyCuSDR/demodulator/__init__.py:3-5):
exposes ``log`` and the ``UHF`` / ``STX`` backends, both classes named ``Demodulator``."""
from .demodulator_base import log  # noqa: F401
from . import UHF  # noqa: F401
from . import STX  # noqa: F401
from . import stream  # noqa: F401,E402  (extension: native streaming ingest)

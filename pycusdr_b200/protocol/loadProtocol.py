"""Protocol registry (reference ``pyCuSDR/protocol/loadProtocol.py:3-20``)."""


def loadProtocol(protocolName):
    if protocolName == "CC11xx":
        from .CC11xx import CC11xx as cls
    elif protocolName == "bench_GMSK":
        from .benchmark import Bench_GMSK as cls
    elif protocolName == "bench_BPSK":
        from .benchmark import Bench_BPSK as cls
    elif protocolName == "bench_FSK":
        from .benchmark import Bench_FSK as cls
    elif protocolName == "bench_GFSK":
        from .benchmark import Bench_GFSK as cls
    else:
        raise ImportError("Protocol %s does not exist" % (protocolName))
    return cls

"""Diagnostic: per-chunk comparison of the CUDA demodulator and the oracle on a bench stream."""
import sys
import numpy as np
sys.path.insert(0, ".")
from oracle import oracle as O, signals as S
from tests.helpers import RADIO, load_conf, protocol_for
from pycusdr_b200.demodulator import UHF

mod, cfg, snr, seed = sys.argv[1], sys.argv[2], float(sys.argv[3]), int(sys.argv[4])
conf = load_conf(cfg)
P = protocol_for(conf)
dem, orc = UHF.Demodulator(conf, P, RADIO), O.OracleDemodulator(conf, P, RADIO)
sig, bits = S.bench_stream(mod, snr, seed=seed)
N, ovl = dem.Nfft, dem.sigOverlap
step = N - ovl
rawd, rawo = dem.get_signalBufferHostPointer(), orc.get_signalBufferHostPointer()
rawd[:] = 0; rawo[:] = 0
for c in range(len(sig) // step):
    rawd[ovl:] = sig[c*step:(c+1)*step]; rawo[ovl:] = sig[c*step:(c+1)*step]
    a = dem.uploadAndFindCarrier(rawd); b = orc.uploadAndFindCarrier(rawo)
    ba = dem.demodulate(); bb = orc.demodulate()
    ld, lo = dem.last, orc.last
    same = len(ba[0]) == len(bb[0]) and np.array_equal(ba[0], bb[0])
    ndiff = -1 if len(ba[0]) != len(bb[0]) else int(np.sum(ba[0] != bb[0]))
    symdiff = int(np.sum(ld['sym'] != lo['sym'])) if len(ld['sym']) == len(lo['sym']) else -1
    cdiff = int(np.sum(ld['centres'] != lo['centres'])) if len(ld['centres']) == len(lo['centres']) else -1
    print(c, 'same' if same else 'DIFF', 'nbits', len(ba[0]), len(bb[0]), 'bitdiff', ndiff, 'symdiff', symdiff, 'centrediff', cdiff,
          'best', ld['res'][0], lo['res'][0], 'shift', ld['shift'], lo['shift'], 'timing', ld['timing'][:2], lo['timing'][:2],
          'spSym', ld['spSym'], lo['spSym'], 'off', ld['codeOffset'], lo['codeOffset'], 'snr', a[3], b[3])
    rawd[:ovl] = rawd[-ovl:]; rawo[:ovl] = rawo[-ovl:]

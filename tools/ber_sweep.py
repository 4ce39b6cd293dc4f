"""BER sweep of the benchmark protocols (examples/benchmark procedure: one 10 000-bit packet per AWGN realisation,
bench_modem.py:198-249) through three implementations on the same sample streams:
  cuda     the product path (pycusdr_b200.demodulator.UHF)
  oracle   the NumPy restatement (oracle/oracle.py)
  refgpu   the reference's own cuda_kernels.cu + cuFFT (oracle/ref_gpu)
Writes gpurun_out/ber_sweep.json and a Markdown table.  Run on the GPU box:  python tools/ber_sweep.py [packets]"""
import json
import math
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as O                                    # noqa: E402
from oracle import signals as S                                   # noqa: E402
from oracle.ref_gpu.driver import RefGpuDemodulator               # noqa: E402
from pycusdr_b200.demodulator import UHF                          # noqa: E402
from tests.helpers import RADIO, load_conf, protocol_for          # noqa: E402
from tests.test_ber_sweep import CFG, best_alignment_errors, wilson   # noqa: E402

packets = int(sys.argv[1]) if len(sys.argv) > 1 else 2
SNRS = [0, 2, 4, 6, 8, 10, 12]
out = {"packets_per_point": packets, "bits_per_packet": 10000, "points": []}
for mod in ("GMSK", "FSK", "GFSK", "BPSK"):
    conf = load_conf(CFG[mod])
    P = protocol_for(conf)
    for snr in SNRS:
        errs = {"cuda": 0, "oracle": 0, "refgpu": 0}
        nbits = 0
        for pk in range(packets):
            sig, tx = S.bench_stream(mod, snr, seed=1000 + 13 * pk + snr)
            impls = {"cuda": UHF.Demodulator(conf, P, RADIO), "oracle": O.OracleDemodulator(conf, P, RADIO),
                     "refgpu": RefGpuDemodulator(conf, P, RADIO)}
            for name, dem in impls.items():
                bits = np.concatenate([c["data"] for c in O.run_stream(dem, sig)])
                errs[name] += best_alignment_errors(bits, tx)
            impls["refgpu"].close()
            del impls
            nbits += len(tx)
        lo, hi = wilson(errs["refgpu"], nbits)
        row = {"modulation": mod, "bench_snr_db": snr, "ebn0_db": round(S.ebn0_db(mod, snr), 2), "bits": nbits,
               **{f"ber_{k}": v / nbits for k, v in errs.items()}, "refgpu_wilson95": [lo, hi],
               "cuda_within_ci": bool(lo - 1e-4 <= errs["cuda"] / nbits <= hi + 1e-4)}
        out["points"].append(row)
        print(row, flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
with open(os.path.join(ROOT, "gpurun_out", "ber_sweep.json"), "w") as f:
    json.dump(out, f, indent=1)
with open(os.path.join(ROOT, "gpurun_out", "ber_sweep.md"), "w") as f:
    f.write("| modulation | bench SNR [dB] | Eb/N0 [dB] | bits | BER cuda | BER oracle | BER reference kernels | Wilson 95 % of the reference | cuda within |\n|---|---|---|---|---|---|---|---|---|\n")
    for r in out["points"]:
        f.write(f"| {r['modulation']} | {r['bench_snr_db']} | {r['ebn0_db']} | {r['bits']} | {r['ber_cuda']:.2e} | {r['ber_oracle']:.2e} | "
                f"{r['ber_refgpu']:.2e} | [{r['refgpu_wilson95'][0]:.2e}, {r['refgpu_wilson95'][1]:.2e}] | {r['cuda_within_ci']} |\n")
print("all within CI:", all(r["cuda_within_ci"] for r in out["points"]))

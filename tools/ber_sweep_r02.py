"""BER sweep at the survey's sample size (examples/benchmark/bench_modem.py:198-249): >= 100 packets of
10 000 bits per Eb/N0 point, four benchmark protocols, bench "SNR" 0..12 dB, the CUDA path against the REFERENCE'S OWN
KERNELS (cuda_kernels.cu compiled unmodified + cuFFT, oracle/ref_gpu) on the same B200 and the same sample streams; plus a
linear Doppler rate of +-50 / +-200 Hz/s (extension) on a subset.  Per chunk it also counts spectrum-shift and
timing-bin disagreements and differing symbol decisions.  (The NumPy oracle is too slow for this size; it stays the spot
check of tests/test_ber_sweep.py.)

    python tools/ber_sweep_r02.py [packets_per_point] [packets_per_rate_point]     # on the GPU box
Writes gpurun_out/ber_sweep_r02.json and .md."""
import json
import os
import sys
import time

import numpy as np
import scipy.signal

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import signals as S                                   # noqa: E402
from oracle.ref_gpu.driver import RefGpuDemodulator               # noqa: E402
from pycusdr_b200.demodulator import UHF                          # noqa: E402
from tests.helpers import RADIO, load_conf, protocol_for          # noqa: E402
from tests.test_ber_sweep import CFG, wilson                      # noqa: E402


def alignment_errors(rx, tx):
    """Fewest bit errors of the transmitted packet against any alignment of the received stream (what the decoder's sync
    search achieves, decoder.py:96-113), either polarity; FFT cross-correlation to find the candidates."""
    L = len(tx)
    if len(rx) < L:
        return L
    a = rx.astype(np.float32) * 2 - 1
    b = tx.astype(np.float32) * 2 - 1
    corr = scipy.signal.fftconvolve(a, b[::-1], mode="valid")
    best = L
    for o in np.argsort(-np.abs(corr))[:3]:
        e = int(np.sum(rx[o:o + L] != tx))
        best = min(best, e, L - e)
    return best


class Pair:
    """The CUDA path and the reference's kernels fed the same chunks; keeps per-chunk agreement counters."""

    def __init__(self, mod):
        self.conf = load_conf(CFG[mod])
        P = protocol_for(self.conf)
        self.dem = UHF.Demodulator(self.conf, P, RADIO)
        self.ref = RefGpuDemodulator(self.conf, P, RADIO)
        self.ref.inspect = False
        self.N, self.ovl = self.dem.Nfft, self.dem.sigOverlap
        self.rd, self.rr = self.dem.get_signalBufferHostPointer(), self.ref.get_signalBufferHostPointer()
        self.rd[:] = 0
        self.rr[:] = 0
        self.first = True
        self.stats = dict(chunks=0, shift_diff=0, timing_diff=0, symbols=0, symbol_diff=0, bits_len_diff=0)

    def run(self, sig):
        step = self.N - self.ovl
        bd, br = [], []
        for c in range(len(sig) // step):
            blk = sig[c * step:(c + 1) * step]
            self.rd[self.ovl:] = blk
            self.rr[self.ovl:] = blk
            self.dem.uploadAndFindCarrier(self.rd)
            self.ref.uploadAndFindCarrier(self.rr)
            a, b = self.dem.demodulate()[0], self.ref.demodulate()[0]
            ld, lr = self.dem.last, self.ref.last
            st = self.stats
            if not self.first:       # the very first chunk carries the zero-filled overlap (ties of rounding noise)
                st["chunks"] += 1
                st["shift_diff"] += int(ld["shift"] != lr["shift"])
                st["timing_diff"] += int(ld["timing"][0] != lr["timing"][0])
                if len(ld["sym"]) == len(lr["sym"]):
                    st["symbols"] += len(lr["sym"])
                    st["symbol_diff"] += int(np.sum(ld["sym"] != lr["sym"]))
                else:
                    st["bits_len_diff"] += 1
            self.first = False
            bd.append(a)
            br.append(b)
            self.rd[:self.ovl] = self.rd[-self.ovl:]
            self.rr[:self.ovl] = self.rr[-self.ovl:]
        return np.concatenate(bd), np.concatenate(br)

    def close(self):
        self.ref.close()


def main():
    packets = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    rate_packets = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    snrs = list(range(0, 13, 2))
    out = {"packets_per_point": packets, "bits_per_packet": 10000, "points": [], "rate_points": [], "agreement": {}}
    t0 = time.time()
    for mod in ("GMSK", "FSK", "GFSK", "BPSK"):
        pair = Pair(mod)
        for snr in snrs:
            ec = er = nbits = 0
            for pk in range(packets):
                sig, tx = S.bench_stream(mod, snr, seed=100000 + 977 * pk + snr)
                a, b = pair.run(sig)
                ec += alignment_errors(a, tx)
                er += alignment_errors(b, tx)
                nbits += len(tx)
            lo, hi = wilson(er, nbits)
            row = {"modulation": mod, "bench_snr_db": snr, "ebn0_db": round(S.ebn0_db(mod, snr), 2), "packets": packets,
                   "bits": nbits, "errors_cuda": ec, "errors_reference": er, "ber_cuda": ec / nbits, "ber_reference": er / nbits,
                   "reference_wilson95": [lo, hi], "cuda_within_ci": bool(lo - 1e-6 <= ec / nbits <= hi + 1e-6)}
            out["points"].append(row)
            print(row, f"[{time.time() - t0:.0f} s]", flush=True)
        for snr in (6, 10):
            for rate in (50.0, -50.0, 200.0, -200.0):
                ec = er = nbits = 0
                for pk in range(rate_packets):
                    sig, tx = S.bench_stream(mod, snr, seed=200000 + 977 * pk + snr + int(abs(rate)), rate_hz_per_s=rate)
                    a, b = pair.run(sig)
                    ec += alignment_errors(a, tx)
                    er += alignment_errors(b, tx)
                    nbits += len(tx)
                lo, hi = wilson(er, nbits)
                row = {"modulation": mod, "bench_snr_db": snr, "rate_hz_per_s": rate, "packets": rate_packets, "bits": nbits,
                       "errors_cuda": ec, "errors_reference": er, "ber_cuda": ec / nbits, "ber_reference": er / nbits,
                       "reference_wilson95": [lo, hi], "cuda_within_ci": bool(lo - 1e-6 <= ec / nbits <= hi + 1e-6)}
                out["rate_points"].append(row)
                print(row, f"[{time.time() - t0:.0f} s]", flush=True)
        out["agreement"][mod] = dict(pair.stats)
        print(mod, pair.stats, flush=True)
        pair.close()
    out["wall_s"] = time.time() - t0
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "ber_sweep_r02.json"), "w") as f:
        json.dump(out, f, indent=1)
    with open(os.path.join(ROOT, "gpurun_out", "ber_sweep_r02.md"), "w") as f:
        f.write(f"# BER sweep, round 2: {packets} packets x 10 000 bits per point, CUDA path vs the reference's own kernels on the same B200\n\n")
        f.write("`python tools/ber_sweep_r02.py` (examples/benchmark procedure, bench_modem.py:198-249; one AWGN realisation per packet, "
                "seeds 100000 + 977 * packet + SNR).  `errors` = bit errors against the transmitted packet at the best alignment.\n\n")
        f.write("| modulation | bench SNR [dB] | Eb/N0 [dB] | bits | errors CUDA | errors reference | BER CUDA | BER reference | Wilson 95 % of the reference | CUDA within |\n|---|---|---|---|---|---|---|---|---|---|\n")
        for r in out["points"]:
            f.write(f"| {r['modulation']} | {r['bench_snr_db']} | {r['ebn0_db']} | {r['bits']} | {r['errors_cuda']} | {r['errors_reference']} | "
                    f"{r['ber_cuda']:.3e} | {r['ber_reference']:.3e} | [{r['reference_wilson95'][0]:.3e}, {r['reference_wilson95'][1]:.3e}] | {r['cuda_within_ci']} |\n")
        f.write(f"\n## Linear Doppler rate (extension, SURVEY F6): {rate_packets} packets per point\n\n")
        f.write("| modulation | bench SNR [dB] | rate [Hz/s] | bits | errors CUDA | errors reference | Wilson 95 % of the reference | CUDA within |\n|---|---|---|---|---|---|---|---|\n")
        for r in out["rate_points"]:
            f.write(f"| {r['modulation']} | {r['bench_snr_db']} | {r['rate_hz_per_s']:+.0f} | {r['bits']} | {r['errors_cuda']} | {r['errors_reference']} | "
                    f"[{r['reference_wilson95'][0]:.3e}, {r['reference_wilson95'][1]:.3e}] | {r['cuda_within_ci']} |\n")
        f.write("\n## Per-chunk agreement with the reference's kernels over the whole sweep\n\n")
        f.write("| modulation | chunks | spectrum shift differs | timing bin differs | symbol decisions compared | differ | chunks with a different symbol count |\n|---|---|---|---|---|---|---|\n")
        for mod, st in out["agreement"].items():
            f.write(f"| {mod} | {st['chunks']} | {st['shift_diff']} | {st['timing_diff']} | {st['symbols']} | {st['symbol_diff']} | {st['bits_len_diff']} |\n")
        ok = all(r["cuda_within_ci"] for r in out["points"] + out["rate_points"])
        f.write(f"\nAll points inside the reference's Wilson 95 % interval: **{ok}**.  Wall time {out['wall_s']:.0f} s.\n")
    print("all within CI:", all(r["cuda_within_ci"] for r in out["points"] + out["rate_points"]))


if __name__ == "__main__":
    main()

"""COMPARISON VARIANT (BASELINE.json north_star; SURVEY.md 2.2(ii) bar 2, 7 step 8): rows a6-a8 as a batched inverse cuFFT
with load / store callbacks (pycusdr_b200/csrc/variant/; classic cufftXtSetCallback callbacks -- the LTO flavour fails at
plan creation on this image, see cufft_variant.cu) on the same inputs as the product's fused search kernel.
Reports the variant's device time per search, the product's, and the agreement of the energies.  Never the product path.

    python tools/cufft_variant.py [--workload c2|c3|c1] [--reps 5] [--bins-per-batch 64]      # on the GPU box
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LIB = os.path.join(ROOT, "pycusdr_b200", "libpcs_cufft_variant.so")


def ensure_built():
    """The variant links the static cuFFT (~ 290 MB) and is therefore built where it runs, not shipped."""
    if not os.path.exists(LIB):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "pycusdr_b200", "csrc"), "variant"], check=True,
                       stdout=subprocess.DEVNULL)
    return LIB


def run(conf, chunk, reps=5, bins_per_batch=64):
    """Returns (variant dict, E of the variant [D, M] per mask, E of the product [D, M] reference layout)."""
    from pycusdr_b200.benchmark.workloads import RADIO
    from pycusdr_b200.demodulator import UHF
    from pycusdr_b200.protocol import loadProtocol
    protocol = loadProtocol(conf["Main"]["protocols"]["UHF"])(conf=conf)
    dem = UHF.Demodulator(conf, protocol, RADIO, fused=False)
    N, D, M = dem.Nfft, dem.doppIdxArrayLen, dem.num_masks
    masks = np.ascontiguousarray(protocol.get_filter(N, dem.spsym, dem.confGPU["xcorrMaskSize"])[1], dtype=np.complex64)
    shifts = np.ascontiguousarray(dem.doppCyperSymNorm, dtype=np.int32)
    x = np.ascontiguousarray(chunk, dtype=np.complex64)
    lib = C.CDLL(ensure_built())
    lib.pcsv_search.restype = C.c_int
    lib.pcsv_search.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                C.POINTER(C.c_float)]
    lib.pcsv_last_error.restype = C.c_char_p
    E = np.zeros((D, M), dtype=np.float32)
    ms = C.c_float(0)
    rc = lib.pcsv_search(x.ctypes.data, N, masks.ctypes.data, M, shifts.ctypes.data, D, int(bins_per_batch), int(reps),
                         E.ctypes.data, C.byref(ms))
    if rc != 0:
        raise RuntimeError(f"cuFFT variant failed ({rc}): {lib.pcsv_last_error().decode()}")
    # the product on the same chunk
    raw = dem.get_signalBufferHostPointer()
    raw[:] = x
    dem.uploadToGPU(raw)
    dem._engine.set_profiling(True)
    for _ in range(reps + 1):
        res, Eo = dem._engine.search()
    prof = dem._engine.profile()
    dem._engine.set_profiling(False)
    ours_ms = (prof["search"][0] + prof["block_spectra"][0] + prof["reduce"][0]) / max(prof["search"][1], 1)
    Eo = Eo.copy()
    # reference layout: SUM mode folds the masks into column 0 in mask order (kern:453-464)
    if dem.SUM_ALL_MASKS_PYTHON:
        acc = E[:, 0].copy()
        for m in range(1, M):
            acc = (acc + E[:, m]).astype(np.float32)
        Ev = np.zeros_like(E)
        Ev[:, 0] = acc
    else:
        Ev = E
    err = float(np.max(np.abs(Ev.astype(np.float64) - Eo)) / np.max(np.abs(Eo)))
    step = N - dem.sigOverlap
    out = {"what": "LABELLED COMPARISON VARIANT, not the product path: a6-a8 as batched inverse cuFFT (C2C, in place, "
                   f"{bins_per_batch} bins x {M} masks per plan) with a load callback (shift x filter product) and a store callback "
                   "(|y|^2 / 2^18, float atomics); kern:339-373, 421-480",
           "nfft": N, "doppler_bins": D, "masks": M, "ms_per_search": float(ms.value),
           "msamples_per_s_search_only": step / (ms.value * 1e-3) / 1e6,
           "product_ms_per_search": ours_ms, "product_msamples_per_s_search_only": step / (ours_ms * 1e-3) / 1e6,
           "speedup_of_the_product": float(ms.value) / ours_ms, "energy_rel_err_vs_product": err}
    return out, E, Eo


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c2")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--bins-per-batch", type=int, default=64)
    args = ap.parse_args()
    from pycusdr_b200.benchmark import workloads as W
    conf, mod, desc = W.load_workload(args.workload)
    N, ovl, step, fs = W.geometry(conf)
    chunk = W.chunks_from_stream(W.build_stream(conf, mod, 2, seed=2), N, ovl, 2)[1]
    out, E, Eo = run(conf, chunk, args.reps, args.bins_per_batch)
    out["workload"] = desc
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"cufft_variant_{args.workload}.json"), "w") as f:
        json.dump(out, f, indent=1)
    assert out["energy_rel_err_vs_product"] < 1e-4, out["energy_rel_err_vs_product"]


if __name__ == "__main__":
    main()

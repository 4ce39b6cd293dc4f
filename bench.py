#!/usr/bin/env python
"""Headline benchmark: Doppler-searched Msamples/s of the demodulator hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4|c5]

A *step* is one batch of ``--chunks-per-step`` (16) consecutive chunks of the synthetic stream through the whole per-chunk
path: chunk spectrum, Doppler search over all bins x masks, Doppler estimate, demod surface at the found bin, timing
recovery, symbol decisions, result copy to the host.  ``value`` counts the NEW samples per chunk (Nfft - 2^overlap, the
reference's own rate convention, pyCuSDR/demodulator_process.py:333) with the chunks resident in the ingest GPU's HBM
when the timed region starts; ``value_with_h2d`` is the same loop fed from pinned host memory (H2D inside);
``e2e`` is the metric from host samples to the stitched bit stream through the repo's public API: at N = 1 the
reference-facing class (``UHF.Demodulator.uploadAndFindCarrier`` + ``demodulate``, strictly alternating), at N > 1
``sharded.ShardedBitStream`` with ONE ingest rank (rank 0 alone holds host samples).  In both the samples wait in
page-locked host memory (a sample ring registered once; chunks are overlapping windows of it) and the H2D copy of every
chunk is inside the timed region; ``e2e.caller_fill_loop`` is the same figure with the reference's caller unchanged (it
first copies every block into the one pinned chunk buffer), ``e2e_stream`` the streaming API on one GPU.

Workload (BASELINE.json configs[1]): GMSK 9600 baud x 16 samples/symbol, 2^18-sample chunks,
256 Doppler bins, 8 matched filters, back-to-back benchmark packets with AWGN at "SNR" 12 dB, seed 2.

Every N runs the same native engine (pcs_shard_*, csrc/shard.inc): Doppler bins sharded over the ranks, the chunk
broadcast from rank 0's HBM over NVLink by the copy engines, the rows of the [D, M] tables stored by the search kernels
straight into the chunk owner's memory, owner-only tail; no collective on the data path -> strong scaling of one stream.
The line carries ``parity_vs_single_gpu``: rank 0 re-runs the timed chunks on one GPU through the unsharded path and the
run FAILS unless every chunk's shift, timing bin and symbol tables are identical.

--impl reference: the reference has no CPU implementation of this path and PyCUDA cannot be installed offline, so this
arm runs the reference's own device code (cuda_kernels.cu compiled unmodified for sm_100a, oracle/ref_gpu) + cuFFT with
the reference's launch sequence on the same B200; without a GPU it times the NumPy/SciPy port instead.  Both arms print
``verify``: a SHA-256 over the bits, spectrum shift and timing bin of chunks 1..6 of the stream through the class API.
"""
import argparse
import json
import math
import os
import sys
import threading
import time
import zlib

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # the engine's streams must not alias onto one hardware queue

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from pycusdr_b200.benchmark import workloads as W      # noqa: E402

RADIO = W.RADIO
VERIFY_CHUNKS = 6          # chunks 1..6 of the stream go into the ``verify`` digest of both arms


def protocol_for(conf):
    from pycusdr_b200.protocol import loadProtocol
    return loadProtocol(conf["Main"]["protocols"]["UHF"])(conf=conf)


def alg_counts(N, D, M, S):
    """F_alg [FLOP] and B_alg / B_unfused [bytes] per chunk exactly as SURVEY.md 8(d) defines them."""
    P = D * M * N
    F = (D * M + M + 1.5) * 5 * N * math.log2(N) + 10 * P + 8 * M * N
    B = 8 * N + 8 * M * N + 4 * D * M + 12 * S
    return F, B, 32 * P


def search_kernel_counts(N, D, M):
    """Algorithmic work of the dominant kernel alone (the fused shift x filter -> inverse FFT -> |.|^2 ->
    sum/arg-max): the D*M inverse transforms, the products and the |.|^2 accumulation of 8(d)'s F_alg, and its
    compulsory bytes (chunk + filter spectra in, 12 bytes per (bin, mask) out)."""
    P = D * M * N
    return D * M * 5 * N * math.log2(N) + 10 * P, 8 * N + 8 * M * N + 12 * D * M


def bench_config(desc, N, ovl, D, M, cps):
    """The ``config`` object: identical keys and values in both arms."""
    return {"workload": desc, "nfft": N, "overlap": ovl, "doppler_bins": D, "masks": M, "chunks_per_step": cps,
            "samples_per_step": (N - ovl) * cps}


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU with NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # no NVML: report it instead of inventing clocks
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for n, bit in names.items():
                    if mask & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml_unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def oracle_baseline(conf, chunk, budget_s=20.0, workers=None):
    """Times the NumPy/SciPy oracle (search + demod) on a bounded sample of one chunk.
    This is the only place bench.py executes oracle/ code."""
    from oracle import oracle as O
    workers = workers or os.cpu_count()
    orc = O.OracleDemodulator(conf, protocol_for(conf), RADIO, fft_workers=workers)
    D = len(orc.doppCyperSymNorm)
    X = O.forward_fft(chunk)
    t0 = time.perf_counter()
    O.search_energy(X, orc.masks, orc.doppCyperSymNorm[:2], orc.SUM_ALL_MASKS_PYTHON, workers=workers)
    per_bin = (time.perf_counter() - t0) / 2
    nb = int(max(2, min(D, budget_s / max(per_bin, 1e-6))))
    shifts = orc.doppCyperSymNorm[:nb]
    t0 = time.perf_counter()
    X = O.forward_fft(chunk)
    O.search_energy(X, orc.masks, shifts, orc.SUM_ALL_MASKS_PYTHON, workers=workers)
    t_search = (time.perf_counter() - t0) * D / nb
    orc.X = X
    orc.dopplerIdxlast = int(orc.doppCyperSymNorm[D // 2])
    t0 = time.perf_counter()
    orc.demodulate()
    t_demod = time.perf_counter() - t0
    new_samples = orc.Nfft - orc.sigOverlap
    return {"value": new_samples / (t_search + t_demod) / 1e6, "unit": "Msamples/s", "cores": int(workers),
            "kind": "port",
            "sample": f"one chunk: {nb} of {D} Doppler bins searched (scaled linearly to {D}) + full demod; "
                      f"scipy.fft workers={workers}; {t_search + t_demod:.2f} s/chunk extrapolated"}


def cpu_baselines(conf, chunk, budget_s):
    """NumPy/SciPy port on all host cores, plus the single-worker figure SURVEY 8(d) asks for."""
    cpu = oracle_baseline(conf, chunk, budget_s=budget_s)
    one = oracle_baseline(conf, chunk, budget_s=budget_s / 2, workers=1)
    cpu["single_worker"] = {"value": one["value"], "unit": one["unit"], "cores": 1, "sample": one["sample"]}
    cpu["host_cpus"] = {"os_cpu_count": os.cpu_count(), "affinity": len(os.sched_getaffinity(0))}
    return cpu


def verify_digest(dem, stream, N, ovl):
    """Chunks 0..VERIFY_CHUNKS of the stream through the class contract from a fresh state; digest over chunks 1.. (chunk
    0 carries the zero-filled overlap, whose first decisions are ties of rounding noise in ANY implementation)."""
    step = N - ovl
    raw = dem.get_signalBufferHostPointer()
    raw[:] = 0
    dg = W.BitsDigest()
    for c in range(VERIFY_CHUNKS + 1):
        raw[ovl:] = stream[c * step:(c + 1) * step]
        dem.uploadAndFindCarrier(raw)
        bits = dem.demodulate()[0]
        if c > 0:
            dg.add(bits, dem.last["shift"], dem.last["timing"][0])
        raw[:ovl] = raw[-ovl:]
    out = dg.summary()
    out["what"] = f"sha256 over bits + spectrum shift + timing bin of chunks 1..{VERIFY_CHUNKS} of the stream (seed 2), class API"
    return out


def run_reference(args, conf, desc):
    """--impl reference.  The reference has NO CPU implementation of this path: its per-chunk work is 8 CUDA kernels +
    cuFFT.  On a GPU box this arm therefore runs the reference's own device code -- cuda_kernels.cu compiled
    unmodified into oracle/_ref/ and launched with the reference's call sequence, launch shapes, zero-copy pinned
    buffers and blocking D2H copies (oracle/ref_gpu/driver.py) -- through the same class contract, chunk by chunk.
    Without a GPU or the cubin it times the NumPy/SciPy port on the host cores instead.  Either way the line carries a
    cpu_baseline (the NumPy port on a bounded sample)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N, ovl, step, fs = W.geometry(conf)
    modulation = W.WORKLOADS[args.workload][1]
    cps = args.chunks_per_step
    cpu = None
    gpu_ok = False
    protocol = protocol_for(conf)
    cg = conf["GPU"]["UHF"]
    D = conf["Radios"]["Rx"][RADIO]["doppCarrierSteps"]
    M = protocol.get_filter(4096, conf["Radios"]["Rx"][RADIO]["samplesPerSym"], cg["xcorrMaskSize"])[0]
    try:
        import torch
        from oracle.ref_gpu import driver as R
        gpu_ok = torch.cuda.is_available() and R.available(M, cg["bitWindowWidth"],
                                                           bool(getattr(protocol, "SUM_ALL_MASKS_PYTHON", False)), 0)
    except Exception:       # no torch / no cubin: CPU arm
        gpu_ok = False
    ring = 32
    stream = W.build_stream(conf, modulation, max(ring, VERIFY_CHUNKS + 1), seed=2)
    verify = None
    if gpu_ok:
        dem = R.RefGpuDemodulator(conf, protocol, RADIO)
        verify = verify_digest(dem, stream, N, ovl)
        dem.close()
        dem = R.RefGpuDemodulator(conf, protocol, RADIO)
        dem.inspect = False                        # stock call sequence: no extra D2H of the energy table
        raw = dem.get_signalBufferHostPointer()
        raw[:] = 0
        blocks = [stream[c * step:(c + 1) * step] for c in range(ring)]
        nbits = 0
        k = 0
        for _ in range(args.warmup * cps):
            raw[ovl:] = blocks[k % ring]
            dem.uploadAndFindCarrier(raw)
            dem.demodulate()
            raw[:ovl] = raw[-ovl:]
            k += 1
        dem.sync()
        t0 = time.perf_counter()
        for _ in range(args.steps * cps):
            raw[ovl:] = blocks[k % ring]
            dem.uploadAndFindCarrier(raw)
            bits = dem.demodulate()[0]
            nbits += len(bits)
            raw[:ovl] = raw[-ovl:]
            k += 1
        dem.sync()
        dt = time.perf_counter() - t0
        launches = dem.launches
        dem.close()
        v = step * cps * args.steps / dt / 1e6
        steps, warm = args.steps, args.warmup
        kind = ("reference device code (cuda_kernels.cu unmodified, sm_100a cubin) + cuFFT on this B200, reference launch "
                "sequence incl. zero-copy pinned input and blocking D2H copies")
        if not args.no_cpu_baseline:
            cpu = cpu_baselines(conf, W.chunks_from_stream(stream, N, ovl, 2)[1], budget_s=8.0)
        extra = {"reference_arm": "gpu", "gpu_launches": launches, "bits_per_step": nbits / max(steps, 1)}
    else:
        chunk = W.chunks_from_stream(stream, N, ovl, 2)[1]
        steps, warm = max(1, min(args.steps, 2)), 1
        budget = 40.0 / (steps + warm)
        vals = []
        for i in range(warm + steps):
            r = oracle_baseline(conf, chunk, budget_s=budget)
            if i >= warm:
                vals.append(r)
        v = float(np.mean([r["value"] for r in vals]))
        cpu = dict(vals[-1])
        cpu["value"] = v
        kind = "NumPy/SciPy port of the reference algorithm on the host cores (no GPU or no reference cubin here)"
        extra = {"reference_arm": "cpu_port"}
    line = {"impl": "reference", "metric": "doppler_searched_msamples_per_s", "value": v, "unit": "Msamples/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": step * cps / v / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(desc, N, ovl, D, M, cps), "what": kind, "cpu_baseline": cpu, "verify": verify,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line))


def run_c5(args):
    """BASELINE config 5: 64 concurrent satellite channels (32 bench_GMSK + 32 bench_FSK, C3-sized chunks), one handle
    and one CUDA stream per channel on this GPU; under torchrun the channels are sharded over the ranks with no exchange at
    all (total work fixed, 64 / world channels per GPU).  A step = one chunk of every channel."""
    import torch
    from pycusdr_b200.config import loadModularJson
    from pycusdr_b200.demodulator import UHF
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n_ch, ring = 64, 8
    mine = [c for c in range(n_ch) if c % world == rank]
    # chunks in flight per GPU: a C3-sized chunk is ~ 0.2 ms of dependent small kernels, so a rank with few channels keeps
    # more steps of each channel in flight (64 chunks per GPU whatever the world size)
    depth = max(2, min(8, n_ch // max(len(mine), 1)))
    dems, ptrs, keep = [], [], []
    step_samples = 0
    for c in mine:
        mod = "GMSK" if c < 32 else "FSK"
        conf = loadModularJson(os.path.join(ROOT, "config", "benchmark", f"bench_{mod}.json"))
        conf["GPU"]["UHF"]["CUDA"]["device"] = local
        N, ovl, step_samples, _ = W.geometry(conf)
        stream = W.build_stream(conf, mod, ring, seed=5000 + c)
        dev = torch.from_numpy(W.chunks_from_stream(stream, N, ovl, ring)).cuda()
        keep.append(dev)
        ptrs.append([dev[i].data_ptr() for i in range(ring)])
        dems.append([UHF.Demodulator(conf, protocol_for(conf), RADIO) for _ in range(depth)])
    engs = [[d._engine for d in pair] for pair in dems]          # ``depth`` handles per channel: step i on handle i % depth
    streams = [torch.cuda.ExternalStream(e.stream) for pair in engs for e in pair]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def enqueue(i):
        for pair, p in zip(engs, ptrs):
            pair[i % depth].enqueue_device(p[i % ring])

    def collect(i):
        acc = 0
        for pair in engs:
            acc += int(pair[i % depth].fetch()[0].shift)
        return acc

    def run(first, count):
        """one chunk of every channel per step; the results of step i - (depth - 1) are collected after step i is enqueued"""
        acc = 0
        done = first
        for i in range(first, first + count):
            enqueue(i)
            if i - done >= depth - 1:
                acc += collect(done)
                done += 1
        while done < first + count:
            acc += collect(done)
            done += 1
        return acc
    warm = max(args.warmup, 4, depth)
    run(0, warm)
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = sum(e.launch_count for pair in engs for e in pair)
    ev0.record(streams[0])
    for st in streams[1:]:
        st.wait_event(ev0)
    checksum = run(warm, args.steps)
    for st in streams[1:]:
        streams[0].wait_stream(st)
    ev1.record(streams[0])
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    launches = sum(e.launch_count for pair in engs for e in pair) - l0
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    if rank == 0:
        value = n_ch * step_samples / (ms_step * 1e-3) / 1e6
        print(json.dumps({
            "metric": "doppler_searched_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C5 64 channels (32 bench_GMSK + 32 bench_FSK), N=2^15, D=64, M=8, one chunk per channel "
                                   "per step", "channels": n_ch, "channels_per_gpu": len(mine), "samples_per_step": n_ch * step_samples,
                       "x_real_time_per_channel": value * 1e6 / n_ch / 153600.0,
                       "l2": f"{ring} distinct chunks per channel", "parallelism": "channels sharded over GPUs, no exchange"},
            "clocks": clocks, "gpu_launches": int(launches), "launch_mode": f"cuda_graph per channel, {depth} handles per channel (steps pipelined)",
            "checksum": checksum}))
    if dist is not None:
        dist.destroy_process_group()


def chunk_check(res, sym, centre, mag):
    """Per-chunk fingerprint of the device results: spectrum shift, timing bin, number of symbols, CRC-32 of the symbol,
    centre and magnitude tables.  Cheap enough to compute for every chunk inside the timed region."""
    crc = zlib.crc32(np.ascontiguousarray(sym).tobytes())
    crc = zlib.crc32(np.ascontiguousarray(centre).tobytes(), crc)
    crc = zlib.crc32(np.ascontiguousarray(mag).tobytes(), crc)
    return (int(res.shift), int(res.timing[0]), int(res.n_sym), crc)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=25)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(W.WORKLOADS) + ["c5"])
    ap.add_argument("--chunks-per-step", type=int, default=16, help="consecutive chunks of the stream that make one step")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 12)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--log2-block", type=int, default=0)
    ap.add_argument("--groups-per-cta", type=int, default=0, help="tuning knob of the 256-point search kernel")
    ap.add_argument("--search-form", type=int, default=0,
                    help="256-point search: 0 shifted filters (default), 1 / 2 rotate-the-chunk comparison variants")
    ap.add_argument("--items-per-cta", type=int, default=0, help="tuning knob of the shifted-filter search kernel")
    ap.add_argument("--warps20", action="store_true", help="tuning knob: 96-register build of the shifted-filter kernel")
    ap.add_argument("--ring", type=int, default=0, help="chunk ring depth of the sharded engine (0 = choose)")
    ap.add_argument("--lag", type=int, default=-1, help="owned chunks kept in flight before the oldest is collected (0..7); "
                    "default 2 on several GPUs (= 2 * world chunks), 6 on one")
    ap.add_argument("--watchdog-s", type=int, default=600, help="abort the process after this many seconds")
    ap.add_argument("--doppler-bins", type=int, default=0,
                    help="experiment: override doppCarrierSteps (e.g. one rank's slice of the bins on a single GPU); the "
                         "line then no longer measures the named workload and says so")
    args = ap.parse_args()
    # a rank that stops making progress must end the run instead of holding the other ranks (and the box) forever
    import faulthandler
    faulthandler.dump_traceback_later(args.watchdog_s, exit=True)

    if args.workload == "c5":
        return run_c5(args)
    conf, modulation, desc = W.load_workload(args.workload)
    if args.doppler_bins:
        conf["Radios"]["Rx"][RADIO]["doppCarrierSteps"] = args.doppler_bins
        desc += f" [EXPERIMENT: {args.doppler_bins} Doppler bins]"
    if args.impl == "reference":
        return run_reference(args, conf, desc)

    import torch
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = gloo = None
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        gloo = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=180))
    conf["GPU"]["UHF"]["CUDA"]["device"] = local

    from pycusdr_b200 import _native, sharded
    from pycusdr_b200.demodulator import UHF
    cr = conf["Radios"]["Rx"][RADIO]
    N, ovl, step_samples, fs = W.geometry(conf)
    protocol = protocol_for(conf)
    cps = args.chunks_per_step
    knobs = dict(groups_per_cta=args.groups_per_cta, search_form=args.search_form, items_per_cta=args.items_per_cta,
                 warps20=args.warps20, log2_block=args.log2_block)

    def barrier():
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    def all_gather(obj):
        if dist is None:
            return [obj]
        out = [None] * world
        dist.all_gather_object(out, obj, group=gloo)
        return out

    # ---- workload: ring of distinct chunks larger than L2, resident in the INGEST GPU's HBM (rank 0) ----
    ring = max(8, min(128, (320 << 20) // (8 * N)))
    stream = host_chunks = dev_chunks = None
    ptrs = [None] * ring
    if rank == 0:
        stream = W.build_stream(conf, modulation, ring, seed=2)
        host_chunks = W.chunks_from_stream(stream, N, ovl, ring)
        dev_chunks = torch.from_numpy(host_chunks).cuda()
        torch.cuda.synchronize()
        ptrs = [dev_chunks[i].data_ptr() for i in range(ring)]
    ring_bytes = ring * 8 * N

    dem = UHF.Demodulator(conf, protocol, RADIO, **knobs)
    eng = dem._engine
    D, M = eng.D, eng.M
    plan = eng.plan()
    S_nom = N // cr["samplesPerSym"]
    if args.lag < 0:
        args.lag = 2 if world > 1 else 6
    sh = sharded.ShardedStream(eng, rank, world, all_gather, ring=args.ring, lag=args.lag)
    info = eng.shard_info()
    lo, hi = sh.slices[rank]
    tstream = torch.cuda.Stream()
    checks = {}

    def collect(c, out):
        checks[c] = chunk_check(out[0], out[2], out[3], out[4])
        return None

    def run_chunks(count, kind=sharded.SRC_DEVICE):
        for _ in range(count):
            c = sh.next_seq
            src = None
            if rank == 0:
                if kind == sharded.SRC_DEVICE:
                    src = ptrs[c % ring]
                else:                         # pinned slots were filled before the timed region
                    sh.host_slot()
            sh.submit(src, kind, collect)

    def timed(count, kind=sharded.SRC_DEVICE):
        """``count`` chunks between two events on an otherwise idle stream; everything this rank enqueued has finished when
        the second one is recorded.  Returns milliseconds (max over ranks)."""
        sh.drain(collect)
        eng.shard_sync()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(tstream)
        run_chunks(count, kind)
        sh.drain(collect)
        eng.shard_sync()
        ev1.record(tstream)
        ev1.synchronize()
        ms = ev0.elapsed_time(ev1)
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # ---- warm-up, then the timed region ----
    run_chunks(max(args.warmup, 3) * cps)
    sh.drain(collect)
    first_timed = sh.next_seq
    launches0 = eng.launch_count
    sampler = ClockSampler(local)
    sampler.start()
    ms_total = timed(args.steps * cps)
    clocks = sampler.stop()
    launches = eng.launch_count - launches0
    if dist is not None:
        t = torch.tensor([launches], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        launches = int(t.item())
    n_timed = args.steps * cps
    ms_step = ms_total / args.steps
    value = step_samples * cps / (ms_step * 1e-3) / 1e6

    # ---- the same loop fed from pinned host memory on the ingest rank (H2D inside the timed region) ----
    if rank == 0:
        for k in range(info["ring"]):
            eng.shard_host_slot(sh.next_seq + k)[:] = host_chunks[(sh.next_seq + k) % ring]
    h2d_chunks = min(n_timed, 8 * cps)
    run_chunks(info["ring"], sharded.SRC_HOST)
    ms_h2d = timed(h2d_chunks, sharded.SRC_HOST)
    value_h2d = step_samples * h2d_chunks / (ms_h2d * 1e-3) / 1e6

    # ---- parity: rank 0 re-runs the timed chunks on ONE GPU through the unsharded path ----
    all_checks = all_gather({c: v for c, v in checks.items() if first_timed <= c < first_timed + n_timed})
    parity = None
    dem1 = None
    if rank == 0:
        merged = {}
        for part in all_checks:
            merged.update(part)
        n_par = min(n_timed, 1024)
        dem1 = UHF.Demodulator(conf, protocol, RADIO, **knobs)
        e1 = dem1._engine
        bad = []
        for c in range(first_timed, first_timed + n_par):
            e1.enqueue_device(ptrs[c % ring])
            res, _, sym, centre, mag = e1.fetch()
            if merged.get(c) != chunk_check(res, sym, centre, mag):
                bad.append(c)
        parity = {"ok": len(merged) == n_timed and not bad, "chunks_checked": n_par, "chunks_timed": n_timed,
                  "results_collected": len(merged), "mismatches": bad[:8],
                  "what": "shift, timing bin, symbol count and CRC-32 of the symbol / centre / magnitude tables of every timed "
                          "chunk vs the unsharded path (pcs_enqueue_device + pcs_fetch) on rank 0"}

    barrier()           # nobody submits further chunks while rank 0 is still re-running the timed ones

    # ---- per-stage device times: chunks one at a time with CUDA events around every stage ----
    eng.set_profiling(True)
    for _ in range(2 * max(world, 4)):
        run_chunks(1)
        sh.drain(collect)
        eng.shard_sync()
    prof = eng.profile()
    eng.set_profiling(False)
    barrier()

    # ---- end to end: host samples -> stitched bits ----
    e2e = None
    e2e_stream = None
    verify = None
    e2e_steps = args.e2e_steps or min(args.steps, 12)
    if world == 1 and not args.no_e2e:
        verify = verify_digest(dem1, stream, N, ovl)
        n_e2e = e2e_steps * cps
        d2h = 88 + 4 * D * M + 12 * eng.max_sym + 2 * 8 * (2 * 5 + 147)

        # (a) the samples sit in a page-locked ring (what a receiver thread fills; chunk c is the window
        #     [c * step, c * step + N) of it, overlap included) and every call copies its chunk to the GPU from there
        ext = np.concatenate((np.zeros(ovl, np.complex64), stream))
        reg = dem1.registerHostMemory(ext)
        windows = [ext[c * step_samples:c * step_samples + N] for c in range(ring)]

        def loop_ring(first, count):
            nb = 0
            for i in range(first, first + count):
                dem1.uploadAndFindCarrier(windows[i % ring])
                nb += len(dem1.demodulate()[0])
            return nb

        # (b) the reference's caller, unchanged (demodulator_process.py:287-337): fill the one pinned chunk buffer, call,
        #     carry the overlap -- 8 N bytes of host memcpy per chunk on top of (a)
        raw = dem1.get_signalBufferHostPointer()
        blocks = [stream[c * step_samples:(c + 1) * step_samples] for c in range(ring)]

        def loop_fill(first, count):
            nb = 0
            for i in range(first, first + count):
                raw[ovl:] = blocks[i % ring]
                dem1.uploadAndFindCarrier(raw)
                nb += len(dem1.demodulate()[0])
                raw[:ovl] = raw[-ovl:]
            return nb

        def timed_loop(loop):
            dem1._stitch.reset()
            raw[:] = 0
            loop(0, min(5, ring))
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            nb = loop(5, n_e2e)
            torch.cuda.synchronize()
            return time.perf_counter() - t0, nb
        dt_fill, nbits_fill = timed_loop(loop_fill)
        dt, nbits = timed_loop(loop_ring)

        # (c) the streaming API every N > 1 run uses for its e2e figure (sharded.ShardedBitStream), here on one GPU: the same
        #     windows, chunks in flight instead of strictly alternating calls
        sh.drain(collect)
        bs = sharded.ShardedBitStream(sh, dem._stitch, 0, 1, None, None)

        def run_stream(first, count):
            for i in range(first, first + count):
                bs.submit(windows[i % ring], sharded.SRC_HOST)
            bs.drain()
        run_stream(0, 8)
        eng.shard_sync()
        nb0 = sum(len(v[0]) for v in bs.bits.values())
        t0 = time.perf_counter()
        run_stream(8, n_e2e)
        eng.shard_sync()
        dt_stream = time.perf_counter() - t0
        e2e_stream = {"value": step_samples * n_e2e / dt_stream / 1e6, "unit": "Msamples/s",
                      "ms_per_step": dt_stream / e2e_steps * 1e3,
                      "bits_per_step": (sum(len(v[0]) for v in bs.bits.values()) - nb0) / e2e_steps,
                      "api": "sharded.ShardedBitStream on one GPU (the API of the N > 1 e2e figure): host samples -> stitched "
                             "bits with chunks in flight, same page-locked windows"}
        reg.close()
        e2e = {"value": step_samples * n_e2e / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * N * cps,
               "d2h_bytes_per_step": int(d2h) * cps, "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "bits_per_step": nbits / e2e_steps,
               "api": "demodulator.UHF.Demodulator.uploadAndFindCarrier(chunk) + demodulate(), strictly alternating; the chunks "
                      "are overlapping windows of a page-locked sample ring (Demodulator.registerHostMemory), copied to the "
                      "GPU from there inside the call",
               "caller_fill_loop": {"value": step_samples * n_e2e / dt_fill / 1e6, "unit": "Msamples/s",
                                    "ms_per_step": dt_fill / e2e_steps * 1e3, "bits_per_step": nbits_fill / e2e_steps,
                                    "what": "the reference's caller unchanged (demodulator_process.py:287-337): every block is "
                                            "first copied into the one pinned chunk buffer by the caller (8 N bytes of host "
                                            "memcpy per chunk), then the same two calls"}}
    elif not args.no_e2e:
        # ONE ingest rank: rank 0 alone touches host samples (sigFIFO.py:147-181); H2D once, NVLink broadcast by the engine;
        # tail + D2H + bit post-processing on each chunk's owner, the chunk-to-chunk carry of checkSymbolOverlap passed from
        # owner to owner over gloo (sharded.OrderedStitcher) so that the bit stream is the one a single process produces
        def e2e_sharded():
            n_e2e = e2e_steps * cps
            cap = _native.Stitcher.state_capacity(ovl, dem.overlapOffset, dem.spsymMin, dem.windowWidth)
            sends = []

            def send(token, dst, c):
                buf = torch.zeros(cap, dtype=torch.uint8)
                buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
                sends.append(dist.isend(buf, dst=dst, tag=c % 30000, group=gloo))

            def recv(src, c):
                buf = torch.empty(cap, dtype=torch.uint8)
                dist.recv(buf, src=src, tag=c % 30000, group=gloo)
                return buf.numpy().tobytes()
            sh.drain(collect)
            bs = sharded.ShardedBitStream(sh, dem._stitch, rank, world, send, recv)
            windows = reg = None
            if rank == 0:        # the radio's samples: a page-locked ring; chunk c = its window [c * step, c * step + N)
                ext = np.concatenate((np.zeros(ovl, np.complex64), stream))
                reg = _native.host_register(ext)
                windows = [ext[c * step_samples:c * step_samples + N] for c in range(ring)]
            state = {"k": 0}

            def run(count):
                for _ in range(count):
                    src = None
                    if rank == 0:
                        src = windows[state["k"] % ring]
                        state["k"] += 1
                    bs.submit(src, sharded.SRC_HOST)
                bs.drain()                                         # in chunk order: the carries travel from chunk to chunk
            run(2 * world)
            eng.shard_sync()
            dist.barrier(group=gloo)                  # (gloo: it times out instead of hanging if a rank has dropped out)
            nb0 = sum(len(v[0]) for v in bs.bits.values())
            t0 = time.perf_counter()
            run(n_e2e)
            eng.shard_sync()
            dt = time.perf_counter() - t0
            bs.finish()
            for w in sends:
                w.wait()
            if reg is not None:
                reg.close()
            return dt, sum(len(v[0]) for v in bs.bits.values()) - nb0, n_e2e

        # a failure on any rank (or a carry that never arrives: the gloo group times out) must not take the device-resident
        # line with it: every rank reports afterwards, and the figure is dropped everywhere if one of them failed
        try:
            dt, nbits, n_e2e = e2e_sharded()
            failed = 0
        except Exception as exc:                       # noqa: BLE001
            print(f"bench: N > 1 end-to-end phase failed on rank {rank}: {exc!r}", file=sys.stderr)
            dt, nbits, n_e2e, failed = 0.0, 0, 1, 1
        agg = torch.tensor([float(failed), dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        nb = torch.tensor([nbits], device="cuda", dtype=torch.int64)
        dist.all_reduce(nb)
        if agg[0].item() == 0:
            dt = float(agg[1].item())
            e2e = {"value": step_samples * n_e2e / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * N * cps,
                   "d2h_bytes_per_step": int(88 + 4 * D * M + 12 * eng.max_sym + 2 * 8 * 157) * cps, "steps": e2e_steps,
                   "ms_per_step": dt / e2e_steps * 1e3, "bits_per_step": int(nb.item()) / e2e_steps,
                   "api": "sharded.ShardedBitStream: host samples on rank 0 ONLY (windows of a page-locked sample ring, one H2D per chunk), NVLink broadcast "
                          "of the chunk by the copy engines, bin-sharded search on all ranks, tail + D2H on the chunk's owner, bit "
                          "post-processing on the owner with the chunk-to-chunk carry passed owner to owner over gloo",
                   "note": "max over ranks of the wall time between barriers"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    if parity is not None and not parity["ok"]:
        print(json.dumps({"error": "sharded results differ from the single-GPU path", "parity_vs_single_gpu": parity}))
        raise SystemExit(3)

    # ---- roofline of the dominant kernel (search) ----
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        with open(peaks_file) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    fp32_peak = _native.measure_fp32_peak(local)
    Dl = hi - lo
    k_flop, k_bytes = search_kernel_counts(N, Dl, M)
    F_alg, B_alg, B_unfused = alg_counts(N, D, M, S_nom)
    s_ms, s_cnt = prof["search"]
    search_ms = s_ms / max(s_cnt, 1)
    kname = ("search_fs256_kernel" if (args.search_form == 0 and M <= 16) else "search_os256_kernel") \
        if plan["log2_block"] == 8 else "search_os_kernel"
    bank = eng.bank_factor()
    if bank[0]:             # long filters that are combinations of a few basis segments (bank_factor.cu): R transforms per item
        kname = "search_fb_kernel"
    roof = {
        "bound": "fp32", "kernel": kname,
        "achieved": k_flop / (search_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": k_flop / (search_ms * 1e-3) / 1e12 / fp32_peak,
        "peak_source": "measured FMA loop on this GPU (pcs_measure_fp32_peak); nominal 74.4",
        "kernel_ms": search_ms, "kernel_share_of_step": search_ms * cps / ms_step,
        "kernel_ms_note": "one launch of the search kernel alone (CUDA events on its stream, chunks one at a time after the "
                          "timed region); share = kernel_ms x chunks_per_step / ms_per_step" + (
                              f"; this rank's slice of {Dl} of {D} bins" if world > 1 else ""),
        "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": k_bytes,
        "hbm_fraction": k_bytes / (search_ms * 1e-3) / 1e9 / hbm_peak,
        "surface_equiv_hbm_fraction": 32.0 * Dl * M * N / (search_ms * 1e-3) / 1e9 / hbm_peak,
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "chunk_bound_ms": max(F_alg / (fp32_peak * 1e12), B_alg / (hbm_peak * 1e9)) * 1e3,
        "chunk_frac": max(F_alg / (fp32_peak * 1e12), B_alg / (hbm_peak * 1e9)) * 1e3 * cps / ms_step / world,
        "traffic": None,
    }
    try:        # DRAM bytes and FMA-pipe activity of one launch of this kernel from the committed ncu --set full capture
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(roof["kernel"])
        if t and t["workload"] == args.workload and world == 1 and not args.doppler_bins:
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = f"ncu --set full, capture {t['capture']} (profiles/)"
            if "pipe_fma_cycles_active_pct" in t and "duration_us" in t:
                # what the hardware executes: FMA-pipe busy cycles the capture measured, rescaled to the live kernel time
                roof["executed_fma_pipe_frac"] = t["pipe_fma_cycles_active_pct"] / 100.0 * t["duration_us"] / (search_ms * 1e3)
                roof["executed_note"] = ("frac uses the survey's 5 N log2 N flop convention for the reference's Nfft-point transforms; "
                                         "executed_fma_pipe_frac = sm__pipe_fma_cycles_active of the capture x its duration / live kernel_ms")
    except (OSError, ValueError, KeyError):
        pass
    stages = {k: (v[0] / max(v[1], 1)) for k, v in prof.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = cpu_baselines(conf, host_chunks[1], budget_s=10.0)

    # ---- labelled comparison variant (north_star): a6-a8 as batched cuFFT with load / store callbacks, same chunk ----
    variants = {}
    if world == 1 and not args.no_variants and not args.doppler_bins:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import cufft_variant
            del dem1, dev_chunks                      # the variant needs up to 1 GiB for its in-place batch buffer
            torch.cuda.empty_cache()
            v, _, _ = cufft_variant.run(conf, host_chunks[1], reps=3, bins_per_batch=64)
            variants["cufft_callback"] = v
        except Exception as exc:                      # noqa: BLE001  (a comparison variant must not take the line with it)
            variants["cufft_callback"] = {"unavailable": repr(exc)[:300]}

    line = {
        "metric": "doppler_searched_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": bench_config(desc, N, ovl, D, M, cps),
        "details": {"x_real_time": value * 1e6 / fs, "ms_per_chunk": ms_step / cps,
                    "path": {1: "overlap_save", 2: "full", 3: "parseval"}.get(plan["path"]),
                    "block": 2 ** plan["log2_block"], "valid_per_block": plan["valid_per_block"],
                    "bank_factor": ({"segment_taps": bank[1], "segments": bank[2], "basis_filters": bank[3]} if bank[0] else None),
                    "l2": f"ring of {ring} distinct chunks = {ring_bytes >> 20} MiB (> 126 MiB L2) in rank 0's HBM, one per chunk",
                    "engine": {"ring": info["ring"], "lanes": info["lanes"], "result_stages": info["stages"], "lag": args.lag},
                    "parallelism": "single GPU, streaming engine (two search lanes + tail stream)" if world == 1 else (
                        f"doppler bins sharded over {world} GPUs; chunk broadcast from rank 0's HBM over NVLink (copy engines); rows "
                        f"stored by the search kernels into the chunk owner's memory; owner-only tail, owners round-robin; no "
                        f"collective on the data path")},
        "clocks": clocks, "gpu_launches": int(launches), "launches_per_chunk": launches / n_timed,
        "launch_mode": "eager: 3-4 launches per chunk and rank + the owner's tail",
        "value_with_h2d": {"value": value_h2d, "unit": "Msamples/s", "chunks": h2d_chunks,
                           "note": "same loop, every chunk copied from pinned host memory on the ingest rank inside the timed region"},
        "parity_vs_single_gpu": parity["ok"], "parity": parity,
        "stage_ms": stages, "roofline": roof, "e2e": e2e, "e2e_stream": e2e_stream, "cpu_baseline": cpu, "verify": verify,
        "variants": variants,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: Doppler-searched Msamples/s of the demodulator hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]

A *step* is one chunk through the whole per-chunk path (chunk spectrum, Doppler search over all
bins x masks, Doppler estimate, demod surface at the found bin, timing recovery, symbol decisions,
result copy to the host).  ``value`` counts the NEW samples per chunk (Nfft - 2^overlap, the
reference's own rate convention, pyCuSDR/demodulator_process.py:333) with the chunks already in HBM;
``e2e`` is the same metric through the reference-facing Python class with host buffers (pinned H2D,
D2H and the host-side stitching inside the timed region).

Workload (BASELINE.json configs[1]): GMSK 9600 baud x 16 samples/symbol, 2^18-sample
chunks, 256 Doppler bins, 8 matched filters, back-to-back benchmark packets with AWGN at "SNR" 12 dB, seed 2.

N > 1 (torchrun, one rank per GPU): Doppler bins are sharded over the ranks, every rank holds the chunk,
the [D, M] energy/peak tables are all-gathered with NCCL and every rank finishes the (cheap) estimate +
demod redundantly -> strong scaling of the same workload.

--impl reference: the reference has no CPU implementation of this path and PyCUDA cannot be installed offline, so
this arm runs the reference's own device code (cuda_kernels.cu compiled unmodified for sm_100a, oracle/ref_gpu) +
cuFFT with the reference's launch sequence on the same B200; without a GPU it times the NumPy/SciPy port instead.
"""
import argparse
import json
import math
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (config file, modulation, description)
    "c1": ("CC11xx.json", None, "C1 CC11xx FSK-2 7416 baud x128, N=2^16, D=64, M=8"),
    "c2": ("c2_base_2p18_256bins.json", "GMSK", "C2 GMSK 9600 baud x16, N=2^18, D=256, M=8"),
    "c3": ("benchmark/bench_GMSK.json", "GMSK", "C3 bench_GMSK, N=2^15, D=64, M=8"),
    "c4": ("c4_sband_2p20_4096bins.json", "GMSK", "C4 wide search, N=2^20, D=4096, M=8"),
}
RADIO = "UHF-H"


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


def build_stream(conf, modulation, n_chunks, seed):
    """Synthetic sample stream for ``n_chunks`` chunks (SURVEY 8(d) inputs)."""
    from pycusdr_b200.benchmark import signals as S
    cg = conf["GPU"]["UHF"]
    cr = conf["Radios"]["Rx"][RADIO]
    N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
    need = n_chunks * (N - ovl)
    sps, baud = cr["samplesPerSym"], cr["baud"]
    fs = sps * baud
    rng = np.random.RandomState(seed)
    if modulation is None:      # C1: FSK-2 packet, CC11xx style, Es/N0 15 dB
        bits = S.createBitSequence(400, seed=123)
        sig = S.modulateFSK(bits, sps)
        one = np.concatenate((np.zeros(4096, np.complex64), sig, np.zeros(4096, np.complex64)))
        f0 = cr["frequencyOffset_Hz"] + 7000.0
        snr_r = 15 - 10 * np.log10(sps)
    else:
        one, _ = S.get_padded_packet(modulation, sps, fs, offset_freq=cr["frequencyOffset_Hz"])
        one = one.astype(np.complex64)
        f0 = None
        snr_r = S.bench_snr_to_awgn_snr(modulation, 12.0, baud, fs)
    reps = need // len(one) + 1
    clean = np.tile(one, reps)[:need]
    if f0 is not None:
        clean = clean * np.exp(2j * np.pi * f0 / fs * np.arange(need)).astype(np.complex64)
    p_sig = np.mean(np.abs(one) ** 2)
    noise_p = p_sig * 10 ** (-snr_r / 10)
    out = np.empty(need, dtype=np.complex64)
    amp = np.float32(np.sqrt(noise_p / 2))
    for a in range(0, need, 1 << 22):       # blockwise: keeps the float64 temporaries small
        n = min(1 << 22, need - a)
        out[a:a + n] = clean[a:a + n] + amp * (rng.randn(n) + 1j * rng.randn(n))
    return out


def chunks_from_stream(stream, N, ovl, n_chunks):
    """[n_chunks, N] array: chunk c = overlap tail of chunk c-1 + new block c (demodulator_process.py:287,337)."""
    step = N - ovl
    out = np.zeros((n_chunks, N), dtype=np.complex64)
    for c in range(n_chunks):
        lo = c * step - ovl
        if lo < 0:
            out[c, ovl:] = stream[:step]
        else:
            out[c] = stream[lo:lo + N]
    return out


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
    O.search_energy(X, orc.masks, orc.doppCyperSymNorm[:4], orc.SUM_ALL_MASKS_PYTHON, workers=workers)
    per_bin = (time.perf_counter() - t0) / 4
    nb = int(max(4, min(D, budget_s / max(per_bin, 1e-6))))
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
    cg = conf["GPU"]["UHF"]
    N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
    step = N - ovl
    modulation = WORKLOADS[args.workload][1]
    cpu = None
    gpu_ok = False
    try:
        import torch
        from oracle.ref_gpu import driver as R
        protocol = protocol_for(conf)
        gpu_ok = torch.cuda.is_available() and R.available(
            protocol.get_filter(4096, conf["Radios"]["Rx"][RADIO]["samplesPerSym"], cg["xcorrMaskSize"])[0],
            cg["bitWindowWidth"], bool(getattr(protocol, "SUM_ALL_MASKS_PYTHON", False)), 0)
    except Exception as e:       # no torch / no cubin: CPU arm
        gpu_ok = False
        why = repr(e)
    n_chunks = args.warmup + args.steps + 1
    ring = min(n_chunks, 32)
    stream = build_stream(conf, modulation, ring, seed=2)
    if gpu_ok:
        dem = R.RefGpuDemodulator(conf, protocol, RADIO)
        raw = dem.get_signalBufferHostPointer()
        raw[:] = 0
        blocks = [stream[c * step:(c + 1) * step] for c in range(ring)]
        nbits = 0
        for i in range(args.warmup):
            raw[ovl:] = blocks[i % ring]
            dem.uploadAndFindCarrier(raw)
            dem.demodulate()
            raw[:ovl] = raw[-ovl:]
        dem.sync()
        t0 = time.perf_counter()
        for i in range(args.steps):
            raw[ovl:] = blocks[(args.warmup + i) % ring]
            dem.uploadAndFindCarrier(raw)
            bits = dem.demodulate()[0]
            nbits += len(bits)
            raw[:ovl] = raw[-ovl:]
        dem.sync()
        dt = time.perf_counter() - t0
        launches = dem.launches
        dem.close()
        v = step * args.steps / dt / 1e6
        steps, warm = args.steps, args.warmup
        kind = ("reference device code (cuda_kernels.cu unmodified, sm_100a cubin) + cuFFT on this B200, reference launch "
                "sequence incl. zero-copy pinned input and blocking D2H copies")
        cpu = oracle_baseline(conf, chunks_from_stream(stream, N, ovl, 2)[1], budget_s=10.0)
        extra = {"reference_arm": "gpu", "gpu_launches": launches, "bits_per_step": nbits / max(steps, 1)}
    else:
        chunk = chunks_from_stream(stream, N, ovl, 2)[1]
        steps, warm = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
        budget = 60.0 / (steps + warm)
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
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": step / v / 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "nfft": N, "overlap": ovl, "samples_per_step": step, "what": kind},
            "cpu_baseline": cpu,
            "e2e": {"value": v, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    line.update(extra)
    print(json.dumps(line))


def run_c5(args):
    """BASELINE config 5: 64 concurrent satellite channels (32 bench_GMSK + 32 bench_FSK, C3-sized chunks), one handle
    and one CUDA stream per channel on this GPU; under torchrun the channels are sharded over the ranks with no exchange at
    all (weak-scaling-free: total work fixed, 64 / world channels per GPU).  A step = one chunk of every channel."""
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
    dems, ptrs, keep = [], [], []
    step_samples = 0
    for c in mine:
        mod = "GMSK" if c < 32 else "FSK"
        conf = loadModularJson(os.path.join(ROOT, "config", "benchmark", f"bench_{mod}.json"))
        conf["GPU"]["UHF"]["CUDA"]["device"] = local
        cg = conf["GPU"]["UHF"]
        N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
        step_samples = N - ovl
        stream = build_stream(conf, mod, ring, seed=5000 + c)
        dev = torch.from_numpy(chunks_from_stream(stream, N, ovl, ring)).cuda()
        keep.append(dev)
        ptrs.append([dev[i].data_ptr() for i in range(ring)])
        dems.append([UHF.Demodulator(conf, protocol_for(conf), RADIO) for _ in range(2)])
    engs = [[d._engine for d in pair] for pair in dems]          # two handles per channel: step i on handle i % 2
    streams = [torch.cuda.ExternalStream(e.stream) for pair in engs for e in pair]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def enqueue(i):
        for pair, p in zip(engs, ptrs):
            pair[i % 2].enqueue_device(p[i % ring])

    def collect(i):
        acc = 0
        for pair in engs:
            acc += int(pair[i % 2].fetch()[0].shift)
        return acc

    def run(first, count):
        """one chunk of every channel per step; the results of step i - 1 are collected after step i is enqueued"""
        acc = 0
        for i in range(first, first + count):
            enqueue(i)
            if i > first:
                acc += collect(i - 1)
        return acc + collect(first + count - 1)
    warm = max(args.warmup, 4)
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
            "clocks": clocks, "gpu_launches": int(launches), "launch_mode": "cuda_graph per channel, two handles per channel (steps pipelined)",
            "checksum": checksum}))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS) + ["c5"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 200)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--log2-block", type=int, default=0)
    ap.add_argument("--groups-per-cta", type=int, default=0, help="tuning knob of the 256-point search kernel")
    ap.add_argument("--xb-smem", action="store_true", help="tuning knob: block spectrum in shared memory")
    ap.add_argument("--search-form", type=int, default=0,
                    help="256-point search: 0 shifted filters (default), 1 / 2 rotate-the-chunk comparison variants")
    ap.add_argument("--items-per-cta", type=int, default=0, help="tuning knob of the shifted-filter search kernel")
    ap.add_argument("--warps20", action="store_true", help="tuning knob: 96-register build of the shifted-filter kernel")
    ap.add_argument("--allow-unvalidated", action="store_true", help="N > 1: accept --inflight > 2")
    ap.add_argument("--watchdog-s", type=int, default=900, help="N > 1: abort the process after this many seconds")
    ap.add_argument("--doppler-bins", type=int, default=0,
                    help="experiment: override doppCarrierSteps (e.g. one rank's slice of the bins on a single GPU); the "
                         "line then no longer measures the named workload and says so")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: how the per-rank Doppler-bin tables reach the estimate (NVLink peer stores | NCCL all-gather)")
    ap.add_argument("--inflight", type=int, default=0,
                    help="chunks in flight for the device-resident figure (one handle + stream each; SURVEY 8(d) allows >= 2)")
    args = ap.parse_args()
    if args.inflight <= 0:      # measured: 3 handles in flight on one GPU, 2 sharded pipelines per rank on several
        args.inflight = 3 if int(os.environ.get("WORLD_SIZE", "1")) == 1 else 2
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        if args.inflight > 2 and not args.allow_unvalidated:
            # round 1: `--gpus 8 --inflight 3` made no progress for six minutes and the GPU budget ended with that run.
            # Reproduced afterwards on the CPU (tests/test_ordered_stitch_cpu.py): the host path drained its pipelines
            # one by one, so two ranks could wait for each other's chunk-to-chunk carries whenever the pipelines held
            # different numbers of chunks (always with three).  Fixed (ShardedPipelines.drain goes in chunk order), but
            # three pipelines have not run on GPUs since, hence the clamp.
            print("bench: more than two sharded pipelines per rank has not been re-validated on GPUs; using --inflight 2 "
                  "(--allow-unvalidated overrides)", file=sys.stderr)
            args.inflight = 2
        # a rank that stops making progress must end the run instead of holding the other ranks (and the box) forever
        import faulthandler
        faulthandler.dump_traceback_later(args.watchdog_s, exit=True)

    if args.workload == "c5":
        return run_c5(args)
    from pycusdr_b200.config import loadModularJson
    cfg_file, modulation, desc = WORKLOADS[args.workload]
    conf = loadModularJson(os.path.join(ROOT, "config", cfg_file))
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
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    conf["GPU"]["UHF"]["CUDA"]["device"] = local

    from pycusdr_b200 import _native
    from pycusdr_b200.demodulator import UHF
    cg, cr = conf["GPU"]["UHF"], conf["Radios"]["Rx"][RADIO]
    N, ovl = 2 ** cg["blockSize"], 2 ** cg["overlap"]
    step_samples = N - ovl
    fs = cr["baud"] * cr["samplesPerSym"]
    protocol = protocol_for(conf)

    # ---- workload: ring of distinct chunks larger than L2 ----
    ring = max(8, min(128, (320 << 20) // (8 * N)))
    stream = build_stream(conf, modulation, ring, seed=2)
    host_chunks = chunks_from_stream(stream, N, ovl, ring)
    dev_chunks = torch.from_numpy(host_chunks).cuda()
    ring_bytes = dev_chunks.numel() * 8

    knobs = dict(groups_per_cta=args.groups_per_cta, xb_smem=args.xb_smem, search_form=args.search_form,
                 items_per_cta=args.items_per_cta, warps20=args.warps20)
    dem = UHF.Demodulator(conf, protocol, RADIO, log2_block=args.log2_block, **knobs)
    eng = dem._engine
    D, M = eng.D, eng.M
    plan = eng.plan()
    S_nom = N // cr["samplesPerSym"]

    sh = None
    if world > 1 and args.exchange == "p2p":
        # bin sharding, exchange through NVLink peer memory, owner-only tail (pycusdr_b200/sharded.py)
        from pycusdr_b200 import sharded

        def all_gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
        K = max(1, args.inflight)
        extra = [UHF.Demodulator(conf, protocol, RADIO, log2_block=args.log2_block, **knobs) for _ in range(K - 1)]
        engs = [eng] + [d._engine for d in extra]
        p2p_error = None
        try:
            sh = sharded.ShardedPipelines(engs, rank, world, all_gather)
        except Exception as e:        # e.g. CUDA IPC not permitted between these processes
            sh, p2p_error = None, repr(e)
        ok = torch.tensor([0 if sh is None else 1], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 0:       # every rank falls back together to the NCCL exchange on a fresh handle
            if rank == 0:
                print(f"bench: peer-memory exchange unavailable ({p2p_error}); using the NCCL all-gather path", file=sys.stderr)
            sh = None
            args.exchange = "nccl"
            dem = UHF.Demodulator(conf, protocol, RADIO, log2_block=args.log2_block, **knobs)
            eng = dem._engine
        else:
            lo, hi = sh.slices[rank]
            streams = [torch.cuda.ExternalStream(e.stream) for e in engs]
            timing_stream = streams[0]
    if world > 1 and sh is None:   # bin sharding with NCCL all-gathers and a replicated tail (comparison variant / fallback)
        per = (D + world - 1) // world
        lo, hi = min(rank * per, D), min((rank + 1) * per, D)
        eng.set_bin_range(lo, max(hi, lo + 1))
        ts = torch.cuda.current_stream()
        eng.set_stream(ts.cuda_stream)
        pe, pv, po = eng.shard_buffers()

        def as_tensor(ptr, dtype, typestr):
            class _W:
                __cuda_array_interface__ = {"shape": (D * M,), "typestr": typestr, "data": (ptr, False), "version": 2}
            return torch.as_tensor(_W(), device=f"cuda:{local}")
        tabs = [as_tensor(pe, torch.float32, "<f4"), as_tensor(pv, torch.float32, "<f4"), as_tensor(po, torch.int32, "<i4")]
        even = (D % world == 0)
        stage = [torch.empty(per * M, dtype=t.dtype, device=t.device) for t in tabs]
        gath = [torch.empty(per * M * world, dtype=t.dtype, device=t.device) for t in tabs]

        def one_step(ptr):
            eng.upload_device(ptr)
            eng.enqueue_search_local()
            for t, s, g in zip(tabs, stage, gath):
                s[:(hi - lo) * M].copy_(t[lo * M:hi * M])
                dist.all_gather_into_tensor(g, s)
                t.copy_(g[:D * M]) if even else [t[r * per * M:min((r + 1) * per, D) * M].copy_(
                    g[r * per * M:r * per * M + (min((r + 1) * per, D) - r * per) * M]) for r in range(world)]
            eng.enqueue_estimate_and_demod(True)
            return eng.fetch()
        timing_stream = ts
    elif world == 1:
        def one_step(ptr):
            eng.enqueue_device(ptr)
            return eng.fetch()
        timing_stream = torch.cuda.ExternalStream(eng.stream)
        # chunks in flight: handle k % K takes chunk k, so the latency-bound tail of one chunk (estimate, demod, timing,
        # symbol decisions, result copies) overlaps the search kernel of the next one
        K = max(1, args.inflight)
        extra = [UHF.Demodulator(conf, protocol, RADIO, log2_block=args.log2_block, **knobs)
                 for _ in range(K - 1)]
        engs = [eng] + [d._engine for d in extra]
        streams = [torch.cuda.ExternalStream(e.stream) for e in engs]

    ptrs = [dev_chunks[i].data_ptr() for i in range(ring)]
    checksum = 0
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def consume(out):
        return int(out[0].shift) + int(out[2][:16].sum())

    if sh is not None:
        unit = world * K
        warm = -(-max(args.warmup, unit) // unit) * unit          # whole owner rounds of every pipeline
        for i in range(warm):
            sh.enqueue(i, ptrs[i % ring], collect=consume)
        sh.drain(consume)
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        launches0 = sum(e.launch_count for e in engs)
        sampler = ClockSampler(local)
        sampler.start()
        ev0.record(streams[0])
        for st in streams[1:]:
            st.wait_event(ev0)
        for i in range(args.steps):
            sh.enqueue(warm + i, ptrs[(warm + i) % ring], collect=consume)
        sh.drain(consume)
        for st in streams[1:]:
            streams[0].wait_stream(st)
        ev1.record(streams[0])
        torch.cuda.synchronize()
        launches = sum(e.launch_count for e in engs) - launches0
        checksum = sum(v for k, v in sh.results.items() if k >= warm)
        ck = torch.tensor([checksum], device="cuda", dtype=torch.int64)
        dist.all_reduce(ck)
        checksum = int(ck.item())
    elif world > 1:
        for i in range(args.warmup):
            one_step(ptrs[i % ring])
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        launches0 = eng.launch_count
        sampler = ClockSampler(local)
        sampler.start()
        ev0.record(timing_stream)
        for i in range(args.steps):
            checksum += consume(one_step(ptrs[(args.warmup + i) % ring]))
        ev1.record(timing_stream)
        torch.cuda.synchronize()
        launches = eng.launch_count - launches0
    else:
        def pipeline(first, count, timed):
            """chunk k on handle k % K; the result of chunk k - K + 1 is collected right after chunk k is enqueued."""
            acc = 0
            if timed:
                ev0.record(streams[0])
                for st in streams[1:]:
                    st.wait_event(ev0)
            for i in range(count):
                e = engs[i % K]
                if i >= K:
                    acc += consume(e.fetch())
                e.enqueue_device(ptrs[(first + i) % ring])
            for i in range(count, count + min(K, count)):
                acc += consume(engs[i % K].fetch())
            if timed:
                for st in streams[1:]:
                    streams[0].wait_stream(st)
                ev1.record(streams[0])
            return acc
        pipeline(0, max(args.warmup, 2 * K), False)      # also lets every handle capture its graph
        torch.cuda.synchronize()
        launches0 = sum(e.launch_count for e in engs)
        sampler = ClockSampler(local)
        sampler.start()
        checksum = pipeline(args.warmup, args.steps, True)
        torch.cuda.synchronize()
        launches = sum(e.launch_count for e in engs) - launches0
    clocks = sampler.stop()
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = step_samples / (ms_step * 1e-3) / 1e6
    if sh is None:
        res = one_step(ptrs[0])[0]

    # ---- per-stage device times: same chunks again with CUDA events around every stage on the handle's stream
    #      (launched kernel by kernel; the timed region above replays them as one CUDA graph) ----
    eng.set_profiling(True)
    if sh is not None:
        base = warm + args.steps
        for i in range(unit * max(2, min(args.steps, 64) // unit)):
            sh.enqueue(base + i, ptrs[(base + i) % ring], collect=consume)
        sh.drain(consume)
    else:
        for i in range(min(args.steps, 100)):
            one_step(ptrs[(args.warmup + i) % ring])
    torch.cuda.synchronize()
    prof = eng.profile()
    eng.set_profiling(False)

    # ---- labelled variant: Parseval energies (no inverse transforms, no peak) on the same chunks, one in flight ----
    variants = {}
    if world == 1 and not args.no_variants:
        demv = UHF.Demodulator(conf, protocol, RADIO, path=_native.PATH_PARSEVAL)
        ev, sv = demv._engine, torch.cuda.ExternalStream(demv._engine.stream)
        for i in range(10):
            ev.enqueue_device(ptrs[i % ring])
            ev.fetch()
        torch.cuda.synchronize()
        nv = min(args.steps, 200)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(sv)
        ck = 0
        for i in range(nv):
            ev.enqueue_device(ptrs[(10 + i) % ring])
            ck += int(ev.fetch()[0].shift)
        b.record(sv)
        torch.cuda.synchronize()
        msv = a.elapsed_time(b) / nv
        variants["parseval"] = {
            "value": step_samples / (msv * 1e-3) / 1e6, "unit": "Msamples/s", "ms_per_step": msv, "chunks_in_flight": 1,
            "note": "LABELLED ALTERNATIVE, not the headline: E[d,m] by Parseval (SURVEY F2), identical estimate / shift / bits, "
                    "no (peak, bin, offset) output", "shift_checksum": ck}
        del demv

    # ---- extension: the same host-buffer workload through the native streaming ingest (not the e2e figure) ----
    if world == 1 and not args.no_variants:
        from pycusdr_b200.demodulator.stream import StreamDemodulator
        sd = StreamDemodulator(conf, protocol, RADIO, inflight=2)
        zmq_block = 1 << 16                       # samples per push, like a ZMQ message
        flat = stream[:ring * step_samples]
        nb = 0
        for a in range(0, 6 * step_samples, zmq_block):      # warm-up (graph capture on both handles)
            nb += sum(len(o["data"]) for o in sd.push(flat[a:a + zmq_block]))
        sd.flush()
        n_s = min(args.steps, 100) * step_samples
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbits = 0
        pos = 6 * step_samples
        for a in range(0, n_s, zmq_block):
            i = (pos + a) % (len(flat) - zmq_block)
            nbits += sum(len(o["data"]) for o in sd.push(flat[i:i + zmq_block]))
        nbits += sum(len(o["data"]) for o in sd.flush())
        dt = time.perf_counter() - t0
        variants["stream_ingest"] = {
            "value": n_s / dt / 1e6, "unit": "Msamples/s", "chunks_in_flight": 2, "push_block_samples": zmq_block,
            "bits": int(nbits), "api": "demodulator.stream.StreamDemodulator.push (host samples in, stitched bits out)",
            "note": "extension beyond the reference's strictly alternating API: H2D, kernels and host stitching overlap"}
        del sd

    # ---- e2e through the reference-facing class, host buffers ----
    e2e_steps = args.e2e_steps or min(args.steps, 200)
    e2e = None
    if world == 1:
        raw = dem.get_signalBufferHostPointer()
        raw[:] = 0
        blocks = [stream[c * step_samples:(c + 1) * step_samples] for c in range(ring)]
        for c in range(min(5, ring)):
            raw[ovl:] = blocks[c]
            dem.uploadAndFindCarrier(raw)
            dem.demodulate()
            raw[:ovl] = raw[-ovl:]
        torch.cuda.synchronize()
        nbits = 0
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            raw[ovl:] = blocks[(5 + i) % ring]
            dem.uploadAndFindCarrier(raw)
            bits, centres, trust, spSym = dem.demodulate()
            nbits += len(bits)
            raw[:ovl] = raw[-ovl:]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        d2h = 88 + 4 * D * M + 12 * eng.max_sym + 2 * 8 * (max(int(res.sig_len), 0))
        e2e = {"value": step_samples * e2e_steps / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * N,
               "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
               "bits_per_step": nbits / e2e_steps,
               "api": "demodulator.UHF.Demodulator.uploadAndFindCarrier + demodulate (pinned chunk buffer)"}

    # ---- e2e at N > 1: host samples on every rank -> pinned buffer -> H2D -> sharded search -> owner tail -> owner D2H
    # + bit post-processing on the owner, with the chunk-to-chunk carry of checkSymbolOverlap passed from owner to owner
    # (sharded.OrderedStitcher) so that the bit stream is the one a single process produces.  Every rank ingests the whole
    # chunk ("one H2D per GPU"). ----
    if sh is not None:
        def e2e_sharded():
            import datetime
            e2e_n = min(args.steps, 200) // (world * K) * (world * K) or world * K
            blocks = [stream[c * step_samples:(c + 1) * step_samples] for c in range(ring)]
            dems = [dem] + extra
            bufs = [d.get_signalBufferHostPointer() for d in dems]
            for b in bufs:
                b[:] = 0
            gloo = dist.new_group(backend="gloo", timeout=datetime.timedelta(seconds=120))
            sends = []

            def send(token, dst, c):
                buf = torch.zeros(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
                buf[:len(token)] = torch.frombuffer(bytearray(token), dtype=torch.uint8)
                sends.append(dist.isend(buf, dst=dst, tag=c, group=gloo))

            def recv(src, c):
                buf = torch.empty(_native.Stitcher.STATE_BYTES, dtype=torch.uint8)
                dist.recv(buf, src=src, tag=c, group=gloo)
                return buf.numpy().tobytes()
            bs = sharded.ShardedBitStream(sh, dem._stitch, rank, world, send, recv, first_chunk=sh.chunks_enqueued)

            def run(count):
                for _ in range(count):
                    pipe, i = bs.next_pipe, bs.next
                    streams[pipe].synchronize()           # the pipeline's previous H2D has left the pinned buffer
                    raw = bufs[pipe]
                    raw[:ovl] = bufs[(i - 1) % K][-ovl:]   # overlap carry (demodulator_process.py:337)
                    raw[ovl:] = blocks[i % ring]
                    bs.submit()
                bs.drain()                                # in chunk order: the carries travel from chunk to chunk
            run((-bs.first) % (world * K) + world * K)    # warm-up: up to the next whole owner round, plus one round
            torch.cuda.synchronize()
            dist.barrier(group=gloo)                  # (gloo: it times out instead of hanging if a rank has dropped out)
            nb0 = sum(len(v[0]) for v in bs.bits.values())
            t0 = time.perf_counter()
            run(e2e_n)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            bs.finish()
            for w in sends:
                w.wait()
            return dt, sum(len(v[0]) for v in bs.bits.values()) - nb0, e2e_n

        # a failure on any rank (or a carry that never arrives: the gloo group times out after 120 s) must not take the
        # device-resident line with it: no NCCL collective runs inside the phase, every rank reports afterwards, and the
        # figure is dropped everywhere if one of them failed
        try:
            dt, nbits, e2e_n = e2e_sharded()
            failed = 0
        except Exception as exc:                       # noqa: BLE001
            print(f"bench: N > 1 end-to-end phase failed on rank {rank}: {exc!r}", file=sys.stderr)
            dt, nbits, e2e_n, failed = 0.0, 0, 1, 1
        agg = torch.tensor([float(failed), dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(agg, op=dist.ReduceOp.MAX)
        nb = torch.tensor([nbits], device="cuda", dtype=torch.int64)
        dist.all_reduce(nb)
        if agg[0].item() == 0:
            dt = float(agg[1].item())
            e2e = {"value": step_samples * e2e_n / dt / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 8 * N * world,
                   "d2h_bytes_per_step": int(88 + 4 * D * M + 12 * eng.max_sym), "steps": e2e_n, "ms_per_step": dt / e2e_n * 1e3,
                   "bits_per_step": int(nb.item()) / e2e_n,
                   "api": "sharded.ShardedBitStream on every rank: samples in each pipeline's pinned buffer, H2D + bin-sharded search "
                          "on all ranks (ShardedPipelines), tail + D2H on the chunk's owner, bit post-processing on the owner with "
                          "the chunk-to-chunk carry passed owner to owner over gloo (OrderedStitcher)",
                   "note": "max over ranks of the wall time between barriers"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (search) ----
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm_peak, hbm_src = 6650.0, "fallback"
    if os.path.exists(peaks_file):
        with open(peaks_file) as f:
            hbm_peak, hbm_src = float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    fp32_peak = _native.measure_fp32_peak(local)
    Dl = (eng.D if world == 1 else (hi - lo))
    k_flop, k_bytes = search_kernel_counts(N, Dl, M)
    F_alg, B_alg, B_unfused = alg_counts(N, D, M, S_nom)
    s_ms, s_cnt = prof["search"]
    search_ms = s_ms / max(s_cnt, 1)
    roof = {
        "bound": "fp32", "kernel": ("search_fs256_kernel" if (args.search_form == 0 and not args.xb_smem and M <= 16)
                                     else "search_os256_kernel") if plan["log2_block"] == 8 else "search_os_kernel",
        "achieved": k_flop / (search_ms * 1e-3) / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
        "frac": k_flop / (search_ms * 1e-3) / 1e12 / fp32_peak,
        "peak_source": "measured FMA loop on this GPU (pcs_measure_fp32_peak); nominal 74.4",
        "kernel_ms": search_ms, "kernel_share_of_step": search_ms / ms_step,
        "algorithmic_flop_per_launch": k_flop, "algorithmic_bytes_per_launch": k_bytes,
        "hbm_fraction": k_bytes / (search_ms * 1e-3) / 1e9 / hbm_peak,
        "surface_equiv_hbm_fraction": 32.0 * Dl * M * N / (search_ms * 1e-3) / 1e9 / hbm_peak,
        "hbm_peak_gbs": hbm_peak, "hbm_peak_source": hbm_src,
        "chunk_bound_ms": max(F_alg / (fp32_peak * 1e12), B_alg / (hbm_peak * 1e9)) * 1e3,
        "chunk_frac": max(F_alg / (fp32_peak * 1e12), B_alg / (hbm_peak * 1e9)) * 1e3 / ms_step if world == 1 else None,
        "traffic": None,
    }
    if roof["kernel"] == "search_fs256_kernel" and world == 1:
        # what the hardware executes (DESIGN.md "Instruction mix"): 489 FMA-pipe lane-slots per filtered 256-point transform
        # and lane (SASS count), 16 lanes per transform, M transforms per (bin, block) item
        slots = float(plan["num_blocks"]) * Dl * M * 16 * 489
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        roof["executed_fma_pipe_frac"] = slots / (148 * 128 * clk * 1e6 * search_ms * 1e-3)
        roof["executed_note"] = ("frac uses the survey's 5 N log2 N flop convention for the reference's Nfft-point transforms; "
                                 "executed_fma_pipe_frac = FMA-pipe lane-slots the kernel's SASS issues / slots available in kernel_ms")
    try:        # DRAM bytes of one launch of this kernel from the committed ncu --set full capture (same workload, 1 GPU)
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            t = json.load(f).get(roof["kernel"])
        if t and t["workload"] == args.workload and world == 1:
            roof["traffic"] = t["dram_bytes_per_launch"]
            roof["traffic_source"] = f"ncu --set full, capture {t['capture']} (profiles/)"
    except (OSError, ValueError, KeyError):
        pass
    stages = {k: (v[0] / max(v[1], 1)) for k, v in prof.items()}

    cpu = None
    if not args.no_cpu_baseline and world == 1:
        cpu = oracle_baseline(conf, host_chunks[1], budget_s=15.0)

    line = {
        "metric": "doppler_searched_msamples_per_s", "value": value, "unit": "Msamples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "nfft": N, "overlap": ovl, "doppler_bins": D, "masks": M,
                   "samples_per_step": step_samples, "x_real_time": value * 1e6 / fs,
                   "path": {1: "overlap_save", 2: "full", 3: "parseval"}.get(plan["path"]),
                   "block": 2 ** plan["log2_block"], "valid_per_block": plan["valid_per_block"],
                   "l2": f"ring of {ring} distinct chunks = {ring_bytes >> 20} MiB (> 126 MiB L2), one per step",
                   "parallelism": "single GPU" if world == 1 else (
                       f"doppler bins sharded over {world} GPUs, rows pushed to the chunk's owner over NVLink peer memory, "
                       f"owner-only tail, owners round-robin" if sh is not None else
                       f"doppler bins sharded over {world} GPUs + NCCL all-gather, replicated tail")},
        "clocks": clocks, "gpu_launches": int(launches), "launches_per_step": launches / args.steps,
        "launch_mode": "cuda_graph" if world == 1 else ("eager, NVLink peer stores" if sh is not None else "eager + NCCL"),
        "chunks_in_flight": (K if (world == 1 or sh is not None) else 1),
        "stage_ms": stages, "roofline": roof, "e2e": e2e, "cpu_baseline": cpu, "checksum": checksum, "variants": variants,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

"""TEST INFRASTRUCTURE -- drives the reference's OWN device code on a B200 without PyCUDA.

The reference reaches the GPU through PyCUDA prepared kernels and a ctypes cuFFT binding, neither of which
installs offline, and its SM-cores table stops at sm_75 so its launch shapes collapse to zero
threads on newer parts (F9).  This module loads the cubins that ``oracle/ref_gpu/Makefile`` compiles from the
unmodified ``pyCuSDR/demodulator/cuda_kernels.cu`` and issues the launch sequence of
``pyCuSDR/demodulator/demodulator_base.py`` through the CUDA driver API and libcufft with ctypes:

    uploadToGPU          dem_base:548-558     cufftExecC2C forward on the mapped pinned buffers
    __findUHF            dem_base:567-605     setArrayToZeros, multInputVectorWithShiftedMasksDopp, batched inverse
                                              cuFFT, blockAbsSumAtomic, findDopplerEst, 8-byte D2H
    __demodulate         dem_base:776-803     multInputVectorWithShiftedMask, inverse cuFFT x M,
                                              sumXCorrBuffMasks, R2C, findCodeRateAndPhase, findCentres, D2H
with the launch shapes of dem_base:353-390,485-491,996 evaluated for 128 cores per SM (the one patch: the entry
``(10, 0): 128`` the cores table lacks).  The host-side arithmetic between the launches is the oracle's
(``oracle/oracle.py``), which is pinned to the reference's host code by ``tests/test_oracle_golden.py``.

Uses: (1) ``tools/make_refgpu_golden.py`` runs it on a B200 and commits the outputs as
``tests/golden/refgpu_*.npz`` -- the fixtures that pin the oracle's device-side steps; (2) GPU parity tests
compare the product kernels with it directly; (3) ``bench.py --impl reference`` times it as "the reference's
path on the same B200".  It is never imported by the product package.

Deliberate deviation: ``GPU_magnitude`` is allocated with 4 bytes per symbol (the reference allocates 1 byte per
symbol although the kernel stores floats, dem_base:472 vs kern:143, and so writes out of bounds); the bytes read
back are the same.
"""
import ctypes as C
import os

import numpy as np

from .. import oracle as O

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(os.path.dirname(HERE), "_ref")

CUFFT_C2C, CUFFT_R2C = 0x29, 0x2A
CUFFT_FORWARD, CUFFT_INVERSE = -1, 1
CU_MEMHOSTALLOC_DEVICEMAP = 0x02


class RefGpuError(RuntimeError):
    pass


def _load(names):
    last = None
    for n in names:
        try:
            return C.CDLL(n)
        except OSError as e:
            last = e
    raise RefGpuError(f"cannot load any of {names}: {last}")


class _Cuda:
    """The handful of driver-API and cuFFT entry points the reference's call sequence needs."""

    def __init__(self):
        self.cu = _load(["libcuda.so.1", "libcuda.so"])
        site = [os.path.join(p, "nvidia", "cufft", "lib", "libcufft.so.11") for p in __import__("sys").path]
        self.fft = _load(["libcufft.so.11", "libcufft.so", "/usr/local/cuda/lib64/libcufft.so.11"] + site)
        self.cu.cuGetErrorString.argtypes = [C.c_int, C.POINTER(C.c_char_p)]

    def ck(self, rc, what):
        if rc != 0:
            s = C.c_char_p()
            self.cu.cuGetErrorString(rc, C.byref(s))
            raise RefGpuError(f"{what} failed: CUresult {rc} ({(s.value or b'?').decode()})")

    def ckf(self, rc, what):
        if rc != 0:
            raise RefGpuError(f"{what} failed: cufftResult {rc}")


def cubin_path(num_masks, window_width, sum_all, code_search_mask_offset):
    return os.path.join(REF_DIR, f"refkern_m{num_masks}_w{window_width}_s{int(bool(sum_all))}_c{code_search_mask_offset}.cubin")


def available(num_masks=8, window_width=7, sum_all=True, code_search_mask_offset=0):
    return os.path.exists(cubin_path(num_masks, window_width, sum_all, code_search_mask_offset))


class RefGpuDemodulator:
    """``demodulator.UHF.Demodulator`` / ``demodulator.STX.Demodulator`` of the reference, device part verbatim."""

    CORES_PER_SM = 128       # the missing (10, 0) entry of lib/cudaConvertSMVer2Cores.py

    def __init__(self, conf, protocol, radioName, backend="UHF", device=None):
        self.backend = backend
        self.protocol = protocol
        self.confRadio = cr = conf["Radios"]["Rx"][radioName]
        self.confGPU = cg = conf["GPU"][cr["CUDA_settings"]]
        # ---- host-side set-up: dem_base:84-174 (same statements as the oracle twin) ----
        self.Nfft = N = 2 ** cg["blockSize"]
        self.sigOverlap = 2 ** cg["overlap"]
        self.sigOverlapWin = int(self.sigOverlap / 2)
        self.peakThresholdScale = cg["peakThresholdScale"]
        self.overlapOffset = cg.get("symbol_check_overlap_offset", O.SYMBOL_CHECK_OVERLAP_OFFSET)
        self.symbol_check_error_threshold = cg.get("symbol_check_error_threshold", O.SYMBOL_CHECK_ERROR_THRESHOLD)
        self.symbol_check_match_threshold = self.overlapOffset - cg.get(
            "symbol_check_match_num_errors_allowed", O.SYMBOL_CHECK_MATCH_NUM_ERRORS_ALLOWED)
        self.spsym = cr["samplesPerSym"]
        self.spsymMin = int(self.spsym / 2)
        self.baudRate = cr["baud"]
        self.sampleRate = self.baudRate * self.spsym
        self.windowWidth = cg["bitWindowWidth"]
        self.CODE_SEARCH_MASK_OFFSET = 0
        self.SUM_ALL_MASKS_PYTHON = bool(getattr(protocol, "SUM_ALL_MASKS_PYTHON", False))
        g = O.doppler_grid(conf, radioName, N)
        self.num_dopplers = g["num_dopplers"]
        self.doppIdxArrayOffset = g["element_offset"]
        self.doppCyperSymNorm = g["shifts"]
        self.doppIdxArrayLen = D = len(g["shifts"])
        self.doppHzLUT = g["doppHzLUT"]
        self.doppOffsetIdx = g["doppOffsetIdx"]
        self.centreFreqOffset = g["centreFreqOffset"]
        self.num_masks, masks = protocol.get_filter(N, self.spsym, cg["xcorrMaskSize"])
        M = self.num_masks
        self.masks = np.ascontiguousarray(masks, dtype=np.complex64)
        self.bitLUT, self.symbolLUT = protocol.get_symbolLUT2(cg["xcorrMaskSize"])
        self.iHigh, self.iLow = O.code_rate_band(N, self.spsym)
        self.numThreads = cg["CUDA"]["numThreads"]
        self.batchSize = cg["CUDA"].get("batchSize", 0)
        self.num_streams = cg["CUDA"].get("streams", 3)
        self.clippedPeakIPure = []
        self.state = O.OverlapState()
        self.dopplerIdxlast = 0
        self.last = {}
        self.launches = 0
        self.inspect = True      # keep E[D,M] of every chunk in .last (one extra D2H the reference does not do; the timed
                                 # reference arm of bench.py switches it off)

        # ---- device set-up: dem_base:177-221 ----
        path = cubin_path(M, self.windowWidth, self.SUM_ALL_MASKS_PYTHON, self.CODE_SEARCH_MASK_OFFSET)
        if not os.path.exists(path):
            raise RefGpuError(f"{path} missing: run `make -C oracle/ref_gpu` where /root/reference is mounted")
        self.api = api = _Cuda()
        cu = api.cu
        api.ck(cu.cuInit(0), "cuInit")
        dev = C.c_int()
        api.ck(cu.cuDeviceGet(C.byref(dev), cg["CUDA"]["device"] if device is None else device), "cuDeviceGet")
        self.dev = dev
        self.ctx = C.c_void_p()
        api.ck(cu.cuDevicePrimaryCtxRetain(C.byref(self.ctx), dev), "cuDevicePrimaryCtxRetain")
        self._current()
        sm = C.c_int()
        api.ck(cu.cuDeviceGetAttribute(C.byref(sm), 16, dev), "cuDeviceGetAttribute(MULTIPROCESSOR_COUNT)")
        self.CUDA_NUM_SMX = sm.value
        self.CUDA_NUM_THREADS = self.CUDA_NUM_SMX * self.CORES_PER_SM              # dem_base:266-271
        self.CUDA_NUM_WARPS = int(self.CUDA_NUM_THREADS / 32)
        with open(path, "rb") as f:
            image = f.read()
        self.module = C.c_void_p()
        api.ck(cu.cuModuleLoadData(C.byref(self.module), image), "cuModuleLoadData")
        self.fn = {}
        for name in ("setArrayToZeros", "multInputVectorWithShiftedMasksDopp", "blockAbsSumAtomic", "findDopplerEst",
                     "multInputVectorWithShiftedMask", "sumXCorrBuffMasks", "findCodeRateAndPhase", "findCentres"):
            f = C.c_void_p()
            api.ck(cu.cuModuleGetFunction(C.byref(f), self.module, name.encode()), f"cuModuleGetFunction({name})")
            self.fn[name] = f
        self._allocs, self._plans, self._streams, self._hosts = [], [], [], []
        c8, f4, i4 = 8, 4, 4
        self.bufXcorr = self._alloc(c8 * N * D * M)                               # dem_base:436
        self.bufDoppSum = self._alloc(f4 * D * M)
        self.bufDoppResult = self._alloc(f4 * 2)
        self.bufDoppIdx = self._alloc(i4 * D)
        self.bufFindDoppTmp = self._alloc(f4 * M)
        self.bufBitsMask = self._alloc(c8 * M * N)
        self.sigTime_host, self.sigTime = self._host_mapped(N)                    # dem_base:456-460
        self.sigFreq_host, self.sigFreq = self._host_mapped(N)
        nsym_max = int(N / self.spsymMin)
        self.symbols = self._alloc(i4 * nsym_max)
        self.centres = self._alloc(i4 * nsym_max)
        self.magnitude = self._alloc(f4 * nsym_max)                               # see module docstring
        self.bufCodeAndPhase = self._alloc(f4 * N)
        self.bufCodeAndPhaseOut = self._alloc(c8 * N)
        self.bufCodeAndPhaseResult = self._alloc(f4 * 3)
        self._htod(self.bufBitsMask, self.masks)
        self._htod(self.bufDoppIdx, np.ascontiguousarray(self.doppCyperSymNorm, dtype=np.int32))
        # FFT plans: dem_base:279-338, 501
        self.planDemod = self._plan(N, CUFFT_C2C, M)
        self.planFwd = self._plan(N, CUFFT_C2C, 1)
        fftBatch = D * M
        if self.batchSize == 0:
            self.FFTNoBatches = 1
            self.planDopplers = self._plan(N, CUFFT_C2C, fftBatch)
        else:
            self.FFTNoBatches = int(fftBatch / self.batchSize)
            if self.FFTNoBatches != fftBatch / self.batchSize:
                raise Exception("FFT batch size has to be an integer divider of xcorrNumMasks * doppCarrierSteps")
            self.FFTBatchSize = int(fftBatch / self.FFTNoBatches)
            self.fftLoops = int(np.ceil(fftBatch / self.num_streams / self.FFTBatchSize))
            self.planDopplers = []
            for _ in range(self.num_streams):
                s = C.c_void_p()
                api.ck(cu.cuStreamCreate(C.byref(s), 0), "cuStreamCreate")   # blocking stream, like pycuda.driver.Stream()
                self._streams.append(s)
                p = self._plan(N, CUFFT_C2C, self.FFTBatchSize)
                api.ckf(api.fft.cufftSetStream(p, s), "cufftSetStream")
                self.planDopplers.append(p)
            self.buffAddr = [self.bufXcorr.value + i * c8 * N * self.FFTBatchSize for i in range(self.FFTNoBatches)]
        self.planR2C = self._plan(N, CUFFT_R2C, 1)
        # launch shapes: dem_base:361-390, 485-491
        self.gVecMasks, self.bVecMasks = (int(N / self.numThreads), 1), (int(self.numThreads), 1, 1)
        self.gAbsSum, self.bAbsSum = (2, int(D)), (128, 1, 1)
        self.gDopp, self.bDopp = (1, 1), (M, 1, 1)
        self.gZero, self.bZero = (int(D), 1), (int(M), 1, 1)
        self.gVecMasks2, self.bVecMasks2 = (int(N / 256), 1), (256, 1, 1)
        self.bCodeMax = (int(min(32 * self.CUDA_NUM_WARPS, 1024)), 1, 1)
        self.gCodeMax = (1, 1)
        self.bXCorrSum = (int(self.CUDA_NUM_THREADS / self.CUDA_NUM_SMX), 1, 1)
        self.gXCorrSum = (int(self.CUDA_NUM_SMX * 4), 1)
        self.bCentres = (256, 1, 1)
        self.sync()

    # ---- thin driver-API helpers --------------------------------------------------------------------
    def _current(self):
        self.api.ck(self.api.cu.cuCtxSetCurrent(self.ctx), "cuCtxSetCurrent")

    def _alloc(self, nbytes):
        p = C.c_uint64()
        self.api.ck(self.api.cu.cuMemAlloc_v2(C.byref(p), C.c_size_t(max(int(nbytes), 4))), f"cuMemAlloc({nbytes})")
        self._allocs.append(p)
        return p

    def _host_mapped(self, n):
        hp = C.c_void_p()
        self.api.ck(self.api.cu.cuMemHostAlloc(C.byref(hp), C.c_size_t(8 * n), CU_MEMHOSTALLOC_DEVICEMAP), "cuMemHostAlloc")
        self._hosts.append(hp)
        dp = C.c_uint64()
        self.api.ck(self.api.cu.cuMemHostGetDevicePointer_v2(C.byref(dp), hp, 0), "cuMemHostGetDevicePointer")
        arr = np.ctypeslib.as_array(C.cast(hp, C.POINTER(C.c_float)), shape=(2 * n,)).view(np.complex64)
        arr[:] = 0
        return arr, dp

    def _plan(self, n, kind, batch):
        p = C.c_int()
        self.api.ckf(self.api.fft.cufftPlan1d(C.byref(p), int(n), int(kind), int(batch)), f"cufftPlan1d({n},{batch})")
        self._plans.append(p)
        return p

    def _htod(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self.api.ck(self.api.cu.cuMemcpyHtoD_v2(dptr, arr.ctypes.data_as(C.c_void_p), C.c_size_t(arr.nbytes)), "cuMemcpyHtoD")

    def _dtoh(self, arr, dptr):
        self.api.ck(self.api.cu.cuMemcpyDtoH_v2(arr.ctypes.data_as(C.c_void_p), dptr, C.c_size_t(arr.nbytes)), "cuMemcpyDtoH")

    def _launch(self, name, grid, block, *args):
        """prepared_call(grid, block, *args) on the NULL stream."""
        holders = []
        for a in args:
            if isinstance(a, C.c_uint64):
                holders.append(C.c_uint64(a.value))
            elif isinstance(a, (np.int32, int)):
                holders.append(C.c_int32(int(a)))
            elif isinstance(a, (np.float32, float)):
                holders.append(C.c_float(float(a)))
            else:
                raise TypeError(type(a))
        params = (C.c_void_p * len(holders))(*[C.cast(C.byref(h), C.c_void_p) for h in holders])
        gx, gy = grid
        bx, by, bz = block
        self.api.ck(self.api.cu.cuLaunchKernel(self.fn[name], gx, gy, 1, bx, by, bz, 0, None, params, None),
                    f"cuLaunchKernel({name}, grid {grid}, block {block})")
        self.launches += 1

    def _exec_c2c(self, plan, src, dst, direction):
        self.api.ckf(self.api.fft.cufftExecC2C(plan, C.c_uint64(int(src)), C.c_uint64(int(dst)), direction), "cufftExecC2C")

    def sync(self):
        self.api.ck(self.api.cu.cuCtxSynchronize(), "cuCtxSynchronize")

    def close(self):
        if getattr(self, "api", None) is None:
            return
        api, cu = self.api, self.api.cu
        try:
            self._current()
            cu.cuCtxSynchronize()
            for p in self._plans:
                api.fft.cufftDestroy(p)
            for s in self._streams:
                cu.cuStreamDestroy_v2(s)
            for p in self._allocs:
                cu.cuMemFree_v2(p)
            self.sigTime_host = self.sigFreq_host = None
            for h in self._hosts:
                cu.cuMemFreeHost(h)
            cu.cuModuleUnload(self.module)
            cu.cuDevicePrimaryCtxRelease_v2(self.dev)
        finally:
            self.api = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- the reference's public methods ---------------------------------------------------------------
    def get_signalBufferHostPointer(self):
        return self.sigTime_host

    def uploadToGPU(self, samples=None):
        self._current()
        if samples is not None and samples is not self.sigTime_host:
            self.sigTime_host[:] = samples
        self._exec_c2c(self.planFwd, self.sigTime.value, self.sigFreq.value, CUFFT_FORWARD)     # dem_base:557

    def search_device(self):
        """Device part of __findUHF (dem_base:571-605). Returns float32[2]."""
        N, D = np.int32(self.Nfft), np.int32(self.doppIdxArrayLen)
        self._launch("setArrayToZeros", self.gZero, self.bZero, self.bufDoppSum)
        self._launch("multInputVectorWithShiftedMasksDopp", self.gVecMasks, self.bVecMasks, self.bufXcorr, self.sigFreq,
                     self.bufBitsMask, self.bufDoppIdx, N, D)
        if self.FFTNoBatches == 1:
            self._exec_c2c(self.planDopplers, self.bufXcorr.value, self.bufXcorr.value, CUFFT_INVERSE)
        else:
            for j in range(self.fftLoops):
                for i in range(self.num_streams):
                    if j * self.num_streams + i < self.FFTNoBatches:
                        a = self.buffAddr[j * self.num_streams + i]
                        self._exec_c2c(self.planDopplers[i], a, a, CUFFT_INVERSE)
        self._launch("blockAbsSumAtomic", self.gAbsSum, self.bAbsSum, self.bufDoppSum, self.bufXcorr, N)
        self._launch("findDopplerEst", self.gDopp, self.bDopp, self.bufDoppResult, self.bufFindDoppTmp, self.bufDoppSum,
                     np.int32(self.num_dopplers), np.int32(self.doppIdxArrayOffset))
        best = np.empty(2, dtype=np.float32)
        self._dtoh(best, self.bufDoppResult)
        return best

    def energies(self):
        E = np.empty((self.doppIdxArrayLen, self.num_masks), dtype=np.float32)
        self._dtoh(E, self.bufDoppSum)
        return E

    def uploadAndFindCarrier(self, samples):
        self._current()
        if self.backend == "STX":                                                # STX.py:8-25
            self.clippedPeakIPure, _ = O.threshold_input(samples, self.peakThresholdScale, self.Nfft)
            self.uploadToGPU(samples)
            return 0, 0, self.clippedPeakIPure, 0
        self.uploadToGPU(samples)
        best = self.search_device()
        self.last.update(res=best.copy())
        if self.inspect:
            self.last.update(E=self.energies())
        try:
            lo, hi, hz, shift = O.interpolate_doppler(best[0], self.doppCyperSymNorm, self.doppHzLUT)
            self.dopplerIdxlast = shift
            self.sync()                                                          # computeSNR: dem_base:641
            SNR = O.compute_snr(self.sigFreq_host, self.doppCyperSymNorm, lo, hi, 5, self.Nfft)
            freqOffset = hz - self.centreFreqOffset
            sdev_Hz = best[1] / self.Nfft * self.sampleRate
        except ValueError:
            self.dopplerIdxlast = 0
            freqOffset, sdev_Hz, SNR = 0.0, 0.0, 0.0
        return freqOffset, sdev_Hz, self.clippedPeakIPure, SNR

    def demod_device(self, shift):
        """Device part of __demodulate (dem_base:776-803, 711-752, 991-1009)."""
        N = self.Nfft
        self._launch("multInputVectorWithShiftedMask", self.gVecMasks2, self.bVecMasks2, self.bufXcorr, self.sigFreq,
                     self.bufBitsMask, np.int32(shift))
        self._exec_c2c(self.planDemod, self.bufXcorr.value, self.bufXcorr.value, CUFFT_INVERSE)
        self._launch("sumXCorrBuffMasks", self.gXCorrSum, self.bXCorrSum, self.bufCodeAndPhase, self.bufXcorr, np.int32(N))
        self.api.ckf(self.api.fft.cufftExecR2C(self.planR2C, self.bufCodeAndPhase, self.bufCodeAndPhaseOut), "cufftExecR2C")
        self._launch("findCodeRateAndPhase", self.gCodeMax, self.bCodeMax, self.bufCodeAndPhaseResult,
                     self.bufCodeAndPhaseOut, np.int32(self.iHigh), np.int32(self.iLow - self.iHigh))
        r = np.empty(3, dtype=np.float32)
        self._dtoh(r, self.bufCodeAndPhaseResult)
        spSym, codeOffset = O.code_rate_host(r, N)
        sp = spSym if spSym >= self.spsymMin else self.spsymMin                  # dem_base:994-995
        grid = (int(np.ceil(N / sp / self.bCentres[0])), 1)
        self._launch("findCentres", grid, self.bCentres, self.symbols, self.centres, self.magnitude, self.bufXcorr,
                     np.float32(sp), np.float32(codeOffset), np.int32(N), np.int32(0))
        S = int(N / sp)
        sym = np.empty(S, dtype=np.int32)
        self._dtoh(sym, self.symbols)
        centres = np.empty(S, dtype=np.int32)
        self._dtoh(centres, self.centres)
        mag = np.empty(S, dtype=np.float32)
        self._dtoh(mag, self.magnitude)
        return r, spSym, codeOffset, sym, centres, mag

    def demod_surface(self):
        """y[m, :] left in bufXcorr by the last demod_device call."""
        y = np.empty((self.num_masks, self.Nfft), dtype=np.complex64)
        self._dtoh(y, self.bufXcorr)
        return y

    def demodulate(self):
        self._current()
        if self.backend == "STX":
            self.dopplerIdxlast = self.doppOffsetIdx                             # dem_base:758-761
        r, spSym, codeOffset, sym, centres, mag = self.demod_device(int(self.dopplerIdxlast))
        trust = O.trust_from_magnitudes(mag, len(sym))                           # dem_base:1005-1007
        self.last.update(shift=int(self.dopplerIdxlast), timing=r, spSym=spSym, codeOffset=codeOffset,
                         sym=sym, centres=centres, mag=mag)
        dataBits, symErr = O.extract_bits(sym, self.bitLUT, self.symbolLUT)
        cW, bW, tW = O.check_symbol_overlap(self.state, len(symErr), centres, dataBits, trust, self.Nfft,
                                            self.sigOverlapWin, self.overlapOffset,
                                            self.symbol_check_error_threshold, self.symbol_check_match_threshold)
        tW = O.tag_clipped_peaks(tW, cW, self.clippedPeakIPure, spSym, self.Nfft)
        return bW.astype(np.uint8), cW.astype(np.uint8), tW.astype(np.uint8), spSym

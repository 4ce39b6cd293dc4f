// Native sample ingest: the job of the reference's SigFIFO ring buffer + the chunk loop of
// demodulator_process.py:284-338 (raw[ovl:] = getBlock(); process; raw[:ovl] = raw[-ovl:]) as a pipeline:
//   * samples arrive in arbitrary block sizes (what the ZMQ SUB socket delivers, sigFIFO.py:147-181) and are appended
//     to a pinned staging ring;
//   * every complete block of (nfft - overlap) new samples becomes a chunk in a device buffer whose first `overlap`
//     samples are a device-to-device copy of the previous chunk's tail (the overlap carry never goes back to the host);
//     the host-to-device copy runs on a copy stream and overlaps the kernels of the chunks already in flight;
//   * chunks go round-robin to the handles given at creation (one CUDA stream + graph each), results come back in
//     chunk order through pcs_ingest_pop.
// The per-chunk numbers are identical to pushing the same samples through pcs_upload / pcs_process chunk by chunk.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <deque>
#include <vector>

#include "../../include/pycusdr_b200.h"

int pcs_fail_msg(int code, const char* msg);

#define S_TRY(expr)                                                                     \
    do {                                                                                \
        cudaError_t e__ = (expr);                                                       \
        if (e__ != cudaSuccess) return pcs_fail_msg(PCS_ERR_CUDA, cudaGetErrorString(e__)); \
    } while (0)

namespace {
struct Record {
    int64_t chunk;
    pcs_result res;
    std::vector<float> E, mag, sig, noise;
    std::vector<int32_t> sym, centre;
};
}  // namespace

struct pcs_ingest {
    std::vector<pcs_handle*> handles;
    std::vector<int64_t> pending;          // chunk number in flight on each handle, -1 = none
    int device = 0, nfft = 0, overlap = 0, step = 0, D = 0, M = 0, max_sym = 0;
    int K = 0, R = 0;                      // device chunk buffers, pinned staging slots
    float2* d_chunks = nullptr;            // [K][nfft]
    float2* h_ring = nullptr;              // [R][step]
    cudaStream_t copy_stream = nullptr;
    std::vector<cudaEvent_t> ev_ready, ev_done, ev_h2d;
    std::vector<bool> done_valid, h2d_valid;
    int64_t next_chunk = 0, next_pop = 0;
    int fill = 0, slot = 0;                // samples in the staging slot being filled
    std::deque<Record> ready;
};

static int collect(pcs_ingest* s, int hi) {
    // fetch the finished chunk of handle hi into the ordered queue
    pcs_handle* h = s->handles[hi];
    Record r;
    r.chunk = s->pending[hi];
    r.E.resize((size_t)s->D * s->M);
    r.sym.resize(s->max_sym);
    r.centre.resize(s->max_sym);
    r.mag.resize(s->max_sym);
    if (int rc = pcs_fetch(h, &r.res, r.E.data(), r.sym.data(), r.centre.data(), r.mag.data())) return rc;
    const int n = r.res.n_sym > 0 ? (r.res.n_sym < s->max_sym ? r.res.n_sym : s->max_sym) : 0;
    r.sym.resize(n);
    r.centre.resize(n);
    r.mag.resize(n);
    const int wl = r.res.sig_len > 0 ? r.res.sig_len : 0;
    r.sig.resize((size_t)2 * wl);
    r.noise.resize((size_t)2 * wl);
    if (wl > 0)
        if (int rc = pcs_snr_windows(h, r.sig.data(), r.noise.data())) return rc;
    s->pending[hi] = -1;
    // keep the queue ordered by chunk (handles finish in order per handle; across handles the older chunk is fetched first
    // by construction, but insert defensively)
    auto it = s->ready.end();
    while (it != s->ready.begin() && (it - 1)->chunk > r.chunk) --it;
    s->ready.insert(it, std::move(r));
    return 0;
}

static int submit_slot(pcs_ingest* s) {
    const int64_t c = s->next_chunk;
    const int b = (int)(c % s->K), hi = (int)(c % (int64_t)s->handles.size());
    pcs_handle* h = s->handles[hi];
    // the handle has one result staging area: its previous chunk must be collected first (back-pressure)
    if (s->pending[hi] >= 0)
        if (int rc = collect(s, hi)) return rc;
    cudaStream_t hs = reinterpret_cast<cudaStream_t>((uintptr_t)pcs_stream(h));
    float2* dst = s->d_chunks + (size_t)b * s->nfft;
    // buffer b was last read by chunk c - K
    if (s->done_valid[b]) S_TRY(cudaStreamWaitEvent(s->copy_stream, s->ev_done[b], 0));
    if (c == 0) {
        S_TRY(cudaMemsetAsync(dst, 0, sizeof(float2) * s->overlap, s->copy_stream));
    } else {
        const float2* prev = s->d_chunks + (size_t)((c - 1) % s->K) * s->nfft;
        S_TRY(cudaMemcpyAsync(dst, prev + (s->nfft - s->overlap), sizeof(float2) * s->overlap, cudaMemcpyDeviceToDevice,
                              s->copy_stream));
    }
    S_TRY(cudaMemcpyAsync(dst + s->overlap, s->h_ring + (size_t)s->slot * s->step, sizeof(float2) * s->step,
                          cudaMemcpyHostToDevice, s->copy_stream));
    S_TRY(cudaEventRecord(s->ev_h2d[s->slot], s->copy_stream));
    s->h2d_valid[s->slot] = true;
    S_TRY(cudaEventRecord(s->ev_ready[b], s->copy_stream));
    S_TRY(cudaStreamWaitEvent(hs, s->ev_ready[b], 0));
    if (int rc = pcs_enqueue_device(h, dst)) return rc;
    S_TRY(cudaEventRecord(s->ev_done[b], hs));
    s->done_valid[b] = true;
    s->pending[hi] = c;
    s->next_chunk++;
    // next staging slot; wait until its previous H2D has left the host buffer
    s->slot = (s->slot + 1) % s->R;
    s->fill = 0;
    if (s->h2d_valid[s->slot]) S_TRY(cudaEventSynchronize(s->ev_h2d[s->slot]));
    return 0;
}

extern "C" {

int pcs_ingest_create(pcs_handle* const* handles, int32_t n_handles, int32_t nfft, int32_t overlap, int32_t num_bins,
                      int32_t num_masks, int32_t device, pcs_ingest** out) {
    if (!handles || !out || n_handles < 1) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    if (overlap < 0 || overlap >= nfft) return pcs_fail_msg(PCS_ERR_INVALID, "overlap must be in [0, nfft)");
    pcs_ingest* s = new pcs_ingest();
    s->handles.assign(handles, handles + n_handles);
    s->pending.assign(n_handles, -1);
    s->device = device; s->nfft = nfft; s->overlap = overlap; s->step = nfft - overlap; s->D = num_bins; s->M = num_masks;
    s->max_sym = pcs_max_symbols(handles[0]);
    s->K = n_handles + 2;
    s->R = n_handles + 2;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_chunks, sizeof(float2) * (size_t)s->K * nfft);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&s->h_ring, sizeof(float2) * (size_t)s->R * s->step, cudaHostAllocDefault);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->copy_stream, cudaStreamNonBlocking);
    s->ev_ready.resize(s->K); s->ev_done.resize(s->K); s->ev_h2d.resize(s->R);
    s->done_valid.assign(s->K, false); s->h2d_valid.assign(s->R, false);
    for (int i = 0; i < s->K && e == cudaSuccess; ++i) {
        e = cudaEventCreateWithFlags(&s->ev_ready[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&s->ev_done[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < s->R && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&s->ev_h2d[i], cudaEventDisableTiming);
    if (e != cudaSuccess) {
        pcs_ingest_destroy(s);
        return pcs_fail_msg(PCS_ERR_CUDA, cudaGetErrorString(e));
    }
    *out = s;
    return PCS_OK;
}

int pcs_ingest_destroy(pcs_ingest* s) {
    if (!s) return PCS_OK;
    cudaSetDevice(s->device);
    if (s->copy_stream) { cudaStreamSynchronize(s->copy_stream); cudaStreamDestroy(s->copy_stream); }
    for (cudaEvent_t e : s->ev_ready) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_done) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : s->ev_h2d) if (e) cudaEventDestroy(e);
    if (s->d_chunks) cudaFree(s->d_chunks);
    if (s->h_ring) cudaFreeHost(s->h_ring);
    delete s;
    return PCS_OK;
}

int pcs_ingest_push(pcs_ingest* s, const void* samples, int64_t n, int32_t* chunks_submitted) {
    if (!s || (!samples && n > 0)) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    S_TRY(cudaSetDevice(s->device));
    const float2* src = reinterpret_cast<const float2*>(samples);
    int submitted = 0;
    while (n > 0) {
        const int64_t room = s->step - s->fill;
        const int64_t take = n < room ? n : room;
        memcpy(s->h_ring + (size_t)s->slot * s->step + s->fill, src, sizeof(float2) * (size_t)take);
        s->fill += (int)take;
        src += take;
        n -= take;
        if (s->fill == s->step) {
            if (int rc = submit_slot(s)) return rc;
            ++submitted;
        }
    }
    if (chunks_submitted) *chunks_submitted = submitted;
    return PCS_OK;
}

int pcs_ingest_pending(const pcs_ingest* s, int64_t* submitted, int64_t* popped) {
    if (!s) return pcs_fail_msg(PCS_ERR_INVALID, "null stream");
    if (submitted) *submitted = s->next_chunk;
    if (popped) *popped = s->next_pop;
    return PCS_OK;
}

int pcs_ingest_pop(pcs_ingest* s, int32_t block, pcs_result* res, float* E_out, int32_t* sym, int32_t* centre, float* mag,
                   float* sig_win, float* noise_win, int32_t* ready) {
    if (!s || !ready) return pcs_fail_msg(PCS_ERR_INVALID, "null argument");
    S_TRY(cudaSetDevice(s->device));
    *ready = 0;
    if (s->next_pop >= s->next_chunk) return PCS_OK;                       // nothing submitted beyond what was popped
    if (s->ready.empty() || s->ready.front().chunk != s->next_pop) {
        const int hi = (int)(s->next_pop % (int64_t)s->handles.size());
        if (s->pending[hi] != s->next_pop) return pcs_fail_msg(PCS_ERR_STATE, "stream bookkeeping lost a chunk");
        if (!block) {
            cudaStream_t hs = reinterpret_cast<cudaStream_t>((uintptr_t)pcs_stream(s->handles[hi]));
            const cudaError_t q = cudaStreamQuery(hs);
            if (q == cudaErrorNotReady) return PCS_OK;
            if (q != cudaSuccess) return pcs_fail_msg(PCS_ERR_CUDA, cudaGetErrorString(q));
        }
        if (int rc = collect(s, hi)) return rc;
    }
    Record& r = s->ready.front();
    if (res) *res = r.res;
    if (E_out) memcpy(E_out, r.E.data(), sizeof(float) * r.E.size());
    if (sym) memcpy(sym, r.sym.data(), sizeof(int32_t) * r.sym.size());
    if (centre) memcpy(centre, r.centre.data(), sizeof(int32_t) * r.centre.size());
    if (mag) memcpy(mag, r.mag.data(), sizeof(float) * r.mag.size());
    if (sig_win) memcpy(sig_win, r.sig.data(), sizeof(float) * r.sig.size());
    if (noise_win) memcpy(noise_win, r.noise.data(), sizeof(float) * r.noise.size());
    s->ready.pop_front();
    s->next_pop++;
    *ready = 1;
    return PCS_OK;
}

}  // extern "C"

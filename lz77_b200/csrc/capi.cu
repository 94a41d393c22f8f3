// capi.cu -- the C ABI of liblz77b200.so (see include/lz77_b200.h).
//
// Thin shim between C host code and the sm_100a kernels: parameter checks that
// mirror the reference CLI (main.c:35-38,102-114), device scratch management,
// host<->device copies for the host-buffer entry points, CUDA-event timing of
// every stage.  No CPU implementation of the codec lives here: without a CUDA
// device every compute entry point fails with LZ77_E_NODEVICE.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include <atomic>
#include <mutex>

#include "context.cuh"

using namespace lz77;

namespace lz77 {

namespace {
Context g_ctx[kMaxDevices + 1];            // [kMaxDevices]: the placeholder before any init
std::atomic<Context *> g_default{nullptr};  // last context initialised by any thread
thread_local Context *t_ctx = nullptr;      // the calling thread's binding
std::mutex g_init_mutex;
}  // namespace

Context &ctx()
{
    if (t_ctx) return *t_ctx;
    Context *d = g_default.load(std::memory_order_acquire);
    return d ? *d : g_ctx[kMaxDevices];
}

int fail_cuda(cudaError_t rc, const char *what)
{
    Context &c = ctx();
    snprintf(c.err, sizeof c.err, "%s: %s", what, cudaGetErrorString(rc));
    cudaGetLastError();  // clear the sticky-free error
    if (rc == cudaErrorMemoryAllocation) return LZ77_E_NOMEM;
    if (rc == cudaErrorNoDevice || rc == cudaErrorInsufficientDriver) return LZ77_E_NODEVICE;
    return LZ77_E_CUDA;
}

int grow(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return LZ77_OK;
    if (*buf) {
        cudaFree(*buf);
        *buf = nullptr;
        *cap = 0;
    }
    need = (need + (size_t)(1 << 20) - 1) & ~(size_t)((1 << 20) - 1);
    cudaError_t rc = cudaMalloc(buf, need);
    if (rc != cudaSuccess) return fail_cuda(rc, "cudaMalloc");
    *cap = need;
    return LZ77_OK;
}

// Resolve and validate (sb, la) the way encode() + main.c do: -1 selects the
// default (lz77.c:65-66); la in 2..255 and sb in 0..65535 (main.c:35-38).
// sb == 0 makes the reference divide by zero (tree.c:66) and is
// rejected here.
int make_params(int sb, int la, Params *P)
{
    if (sb == -1) sb = LZ77_DEFAULT_SB;
    if (la == -1) la = LZ77_DEFAULT_LA;
    if (sb < 1 || sb > LZ77_MAX_SB || la < 1 || la > LZ77_MAX_LA) return LZ77_E_ARG;
    P->sb = sb;
    P->la = la;
    P->ob = bitof(sb);
    P->lb = bitof(la);
    P->tbits = P->ob + P->lb + 8;
    int cap = (1 << P->ob) - 1;  // Appendix B2: off == 2^ob does not fit the field
    P->window = sb < cap ? sb : cap;
    P->block = lz77_gpu_block_size(sb);
    P->block_shift = bitof((int)P->block);
    P->tile_shift = P->block_shift < 17 ? P->block_shift : 17;
    P->history = ctx().history ? 1 : 0;
    P->fused_pack = ctx().fused_pack ? 1 : 0;
    return LZ77_OK;
}

float ms_between(cudaEvent_t a, cudaEvent_t b)
{
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, a, b) != cudaSuccess) {
        cudaGetLastError();
        return 0.f;
    }
    return ms;
}

// header parse, lz77.c:157-158; the header always travels through the host
int read_header(const unsigned char hdr[4], long n_in, Params *P, long long *n_tokens)
{
    if (n_in < 4) return LZ77_E_STREAM;
    const int sb = hdr[0] | (hdr[1] << 8);
    const int la = hdr[2] | (hdr[3] << 8);
    if (sb < 1 || la < 1 || la > LZ77_MAX_LA) return LZ77_E_STREAM;  // bitof(0) is undefined
    if (make_params(sb, la, P) != LZ77_OK) return LZ77_E_STREAM;
    // lz77.c:271-280: a short read ends the stream, so trailing bits < T are padding
    *n_tokens = ((long long)(n_in - 4) * 8) / P->tbits;
    return LZ77_OK;
}

}  // namespace lz77

#define g (lz77::ctx())

namespace {

// default chunk of the pipelined host entry points, measured on the bench workload
// (encode + decode end to end): 4 / 8 / 16 / 32 MiB -> 17.5 / 15.9 / 16.2 / 16.9 ms; the
// large-window encoder wants more blocks per launch
long long host_chunk_bytes(const Params &P)
{
    if (g.host_chunk != 0) return g.host_chunk;
    return P.window > 8191 ? (16ll << 20) : (8ll << 20);
}

// Every exit of a pipelined host path -- also an error return in the middle of it -- first
// waits for the asynchronous copies that read or write the caller's buffers.
struct StreamDrain {
    Context &c;
    explicit StreamDrain(Context &ctx_) : c(ctx_) {}
    ~StreamDrain()
    {
        cudaStreamSynchronize(c.copy_in);
        cudaStreamSynchronize(c.copy_out);
        cudaStreamSynchronize(c.aux);
        cudaStreamSynchronize(c.aux2);
        cudaStreamSynchronize(c.hi);
        cudaStreamSynchronize(c.stream);
    }
};

// the first n events of the pool (created on demand, destroyed at shutdown)
cudaEvent_t *pool_events(size_t n)
{
    Context &c = g;
    while (c.pool.size() < n) {
        cudaEvent_t e;
        if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        c.pool.push_back(e);
    }
    return c.pool.data();
}

}  // namespace

extern "C" {

int lz77_bitof(int n) { return bitof(n); }

int lz77_token_bits(int sb, int la) { return bitof(sb) + bitof(la) + 8; }

long lz77_gpu_encode_bound(long n_in, int sb, int la)
{
    if (sb == -1) sb = LZ77_DEFAULT_SB;
    if (la == -1) la = LZ77_DEFAULT_LA;
    const long t = lz77_token_bits(sb, la);
    return 4 + (n_in * t + 7) / 8;
}

long lz77_gpu_block_size(int sb)
{
    if (sb == -1) sb = LZ77_DEFAULT_SB;
    return sb <= 8191 ? 65536L : 524288L;
}

long lz77_gpu_segment_size(int sb, int la)
{
    Params P;
    if (make_params(sb, la, &P) != LZ77_OK) return kSegBytes;
    return parse_segment_bytes(P.window, P.la, P.fused_pack != 0);
}

int lz77_gpu_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

namespace {

void destroy_context(Context &c)
{
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaDeviceSynchronize();
    comm_release(c);
    for (auto &e : c.ev) cudaEventDestroy(e);
    for (auto &e : c.pool) cudaEventDestroy(e);
    if (c.scratch) cudaFree(c.scratch);
    if (c.jump) cudaFree(c.jump);
    bigwin_release(c.device);
    if (c.stage_in) cudaFree(c.stage_in);
    if (c.stage_out) cudaFree(c.stage_out);
    if (c.xfer) cudaFree(c.xfer);
    if (c.user_in) cudaFree(c.user_in);
    if (c.user_out) cudaFree(c.user_out);
    if (c.pinned) cudaFreeHost(c.pinned);
    if (c.pinned_totals) cudaFreeHost(c.pinned_totals);
    if (c.copy_in) cudaStreamDestroy(c.copy_in);
    if (c.copy_out) cudaStreamDestroy(c.copy_out);
    if (c.aux) cudaStreamDestroy(c.aux);
    if (c.aux2) cudaStreamDestroy(c.aux2);
    if (c.hi) cudaStreamDestroy(c.hi);
    cudaStreamDestroy(c.own_stream);
    c = Context();
}

int create_context(Context &c, int device)
{
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    c.stream = c.own_stream;
    for (auto &e : c.ev) CK(cudaEventCreate(&e));
    CK(cudaMallocHost((void **)&c.pinned, 256));
    CK(cudaStreamCreateWithFlags(&c.copy_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c.copy_out, cudaStreamNonBlocking));
    {
        // The long kernels of the chunked host pipelines (search, tile decode) run on
        // low-priority streams and the short ones that follow each chunk (token-count
        // scan, bit-packer, token scan) on a high-priority stream: otherwise the queued
        // search kernels of later chunks keep every SM slot busy and the packers -- and
        // with them the D2H copies -- only run after the last search.
        int lo_pri = 0, hi_pri = 0;
        CK(cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri));
        CK(cudaStreamCreateWithPriority(&c.aux, cudaStreamNonBlocking, lo_pri));
        CK(cudaStreamCreateWithPriority(&c.aux2, cudaStreamNonBlocking, lo_pri));
        CK(cudaStreamCreateWithPriority(&c.hi, cudaStreamNonBlocking, hi_pri));
    }
    CK(cudaHostAlloc((void **)&c.pinned_totals, kMaxHostChunks * sizeof(unsigned long long),
                     cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void **)&c.pinned_totals_dev, c.pinned_totals, 0));
    c.device = device;
    memset(&c.last, 0, sizeof c.last);
    memset(&c.comm_last, 0, sizeof c.comm_last);
    c.ready = true;
    return LZ77_OK;
}

}  // namespace

// Destroys every context of the process.  No other thread may be inside the library.
void lz77_gpu_shutdown(void)
{
    lz77_mgpu_shutdown();  // the worker threads of the single-process pool use these contexts
    std::lock_guard<std::mutex> lock(g_init_mutex);
    for (int d = 0; d < kMaxDevices; d++) destroy_context(g_ctx[d]);
    g_default.store(nullptr, std::memory_order_release);
    t_ctx = nullptr;
}

int lz77_gpu_init(int device)
{
    std::lock_guard<std::mutex> lock(g_init_mutex);
    int n = lz77_gpu_device_count();
    if (n <= 0) {
        snprintf(g.err, sizeof g.err, "no CUDA device visible");
        return LZ77_E_NODEVICE;
    }
    if (device < 0 || device >= n || device >= kMaxDevices) return LZ77_E_ARG;
    Context &c = g_ctx[device];
    t_ctx = &c;  // errors of the calls below land in this context
    if (!c.ready) {
        int rc = create_context(c, device);
        if (rc != LZ77_OK) {
            t_ctx = nullptr;
            return rc;
        }
    }
    g_default.store(&c, std::memory_order_release);
    return LZ77_OK;
}

const char *lz77_gpu_strerror(int rc)
{
    switch (rc) {
    case LZ77_OK: return "ok";
    case LZ77_E_ARG: return "bad argument";
    case LZ77_E_SPACE: return "output buffer too small";
    case LZ77_E_STREAM: return "malformed stream";
    case LZ77_E_NOMEM: return "out of memory";
    case LZ77_E_NODEVICE: return "no CUDA device / library not initialised";
    case LZ77_E_CUDA: return "CUDA error";
    }
    return "unknown error";
}

const char *lz77_gpu_last_error(void) { return g.err; }

void *lz77_gpu_host_alloc(long n)
{
    void *p = nullptr;
    if (n <= 0) n = 1;
    if (cudaMallocHost(&p, (size_t)n) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}

void lz77_gpu_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

void lz77_gpu_set_timing(int enabled) { g.timing = enabled != 0; }

void lz77_gpu_set_host_chunk(long bytes)
{
    // <= 0: no chunking (one H2D, the kernels, one D2H)
    g.host_chunk = bytes > 0 ? bytes : (1ll << 62);
}

void lz77_gpu_set_history(int enabled) { g.history = enabled != 0; }

void lz77_gpu_set_fused_pack(int enabled) { g.fused_pack = enabled != 0; }

int lz77_gpu_set_jump_piece(long bytes)
{
    if (bytes != 0 && (bytes < (1L << 20) || bytes > (256L << 20))) return LZ77_E_ARG;
    g.jump_piece = bytes;
    return LZ77_OK;
}

int lz77_gpu_set_stream(void *cuda_stream)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    cudaStreamSynchronize(g.stream);
    g.stream = cuda_stream ? (cudaStream_t)cuda_stream : g.own_stream;
    return LZ77_OK;
}

int lz77_gpu_last_timing(struct lz77_timing *t)
{
    if (!t) return LZ77_E_ARG;
    *t = g.last;
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// encode
// ---------------------------------------------------------------------------

int lz77_gpu_encode_device(const void *d_in, long n_in, int sb, int la, void *d_out, long out_cap,
                           long *n_out, long *n_tokens)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    Params P;
    if (make_params(sb, la, &P) != LZ77_OK || n_in < 0 || !d_out || (n_in > 0 && !d_in) || !n_out)
        return LZ77_E_ARG;
    // the kernels move 128-bit words and TMA bulk copies: 16-byte aligned buffers only
    if ((((uintptr_t)d_in) | ((uintptr_t)d_out)) & 15) return LZ77_E_ARG;
    const long bound = lz77_gpu_encode_bound(n_in, P.sb, P.la);
    if (out_cap < ((bound + 15) & ~15L)) return LZ77_E_SPACE;
    CK(cudaSetDevice(g.device));
    int rc = grow(&g.scratch, &g.scratch_cap, encode_scratch_bytes(n_in, P));
    if (rc) return rc;

    unsigned long long *d_total = nullptr;
    StageEvents se = {{g.ev[0], g.ev[1], g.ev[2], g.ev[3]}};
    CK(launch_encode((const uint8_t *)d_in, n_in, 0, P, g.scratch, (uint32_t *)d_out, &d_total,
                     g.stream, g.timing ? &se : nullptr));
    CK(cudaMemcpyAsync(g.pinned, d_total, 8, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    const unsigned long long k = g.pinned[0];
    *n_out = 4 + (long)((k * (unsigned long long)P.tbits + 7) / 8);
    if (n_tokens) *n_tokens = (long)k;

    memset(&g.last, 0, sizeof g.last);
    g.last.launches = encode_launch_count(n_in, P);
    g.last.n_tokens = (long)k;
    if (g.timing) {
        g.last.enc_search_ms = ms_between(g.ev[0], g.ev[1]);
        g.last.enc_scan_ms = ms_between(g.ev[1], g.ev[2]);
        g.last.enc_pack_ms = ms_between(g.ev[2], g.ev[3]);
    }
    return LZ77_OK;
}

int lz77_gpu_encode(const unsigned char *in, long n_in, int sb, int la, unsigned char *out,
                    long out_cap, long *n_out)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    Params P;
    if (make_params(sb, la, &P) != LZ77_OK || n_in < 0 || !out || (n_in > 0 && !in) || !n_out)
        return LZ77_E_ARG;
    CK(cudaSetDevice(g.device));
    const long bound = lz77_gpu_encode_bound(n_in, P.sb, P.la);
    const size_t in_cap = ((size_t)n_in + 15) & ~(size_t)15;
    const size_t o_cap = ((size_t)bound + 15) & ~(size_t)15;
    int rc = grow(&g.stage_in, &g.stage_in_cap, in_cap + 16);
    if (rc) return rc;
    rc = grow(&g.stage_out, &g.stage_out_cap, o_cap + 16);
    if (rc) return rc;

    // Large inputs: chunks of whole blocks flow through three streams so the H2D
    // copy of chunk c+1 and the D2H copy of chunk c-1 overlap the kernels of
    // chunk c.  The running token count stays on the device between chunks.
    const long long granule = encode_chunk_granule();
    long long chunk = host_chunk_bytes(P) / granule * granule;
    if (chunk < granule) chunk = granule;
    // chunk boundaries: equal chunks (8 or 16 MiB unless lz77_gpu_set_host_chunk() says otherwise)
    std::vector<long long> bounds(1, 0);
    for (long long pos = chunk; pos < (long long)n_in; pos += chunk) bounds.push_back(pos);
    bounds.push_back(n_in);
    const long long n_chunks = (long long)bounds.size() - 1;
    if (n_chunks >= 2 && n_chunks <= kMaxHostChunks) {
        rc = grow(&g.scratch, &g.scratch_cap, encode_scratch_bytes(n_in, P));
        if (rc) return rc;
        const EncodePlan pl = encode_plan(g.scratch, n_in, P);
        cudaEvent_t *evp = pool_events((size_t)(3 * n_chunks));
        if (!evp) return LZ77_E_CUDA;
        cudaEvent_t *ev_in = evp, *ev_done = evp + n_chunks, *ev_parse = evp + 2 * n_chunks;
        // (the large-window search keeps its bucket tables in one scratch area, so its
        // chunks must not overlap)
        const bool split_search = P.window <= 8191;
        // the copy streams start after whatever the compute stream still has queued
        StreamDrain drain(g);
        // (fused pack: the chunks' kernels run on two streams and look back across launches;
        // their shared state is zeroed here, in front of all of them)
        if (pl.fused) CK(launch_parse_bucket_fused_reset(n_in, pl.fused, g.stream));
        CK(cudaEventRecord(g.ev[4], g.stream));
        CK(cudaStreamWaitEvent(g.copy_in, g.ev[4], 0));
        CK(cudaStreamWaitEvent(g.copy_out, g.ev[4], 0));
        CK(cudaStreamWaitEvent(g.aux, g.ev[4], 0));
        CK(cudaStreamWaitEvent(g.aux2, g.ev[4], 0));
        CK(cudaStreamWaitEvent(g.hi, g.ev[4], 0));
        for (long long c = 0; c < n_chunks; c++) {
            const long long lo = bounds[c];
            const long long len = bounds[c + 1] - lo;
            CK(cudaMemcpyAsync((char *)g.stage_in + lo, in + lo, (size_t)len,
                               cudaMemcpyHostToDevice, g.copy_in));
            CK(cudaEventRecord(ev_in[c], g.copy_in));
            if (split_search) {
                // searches of consecutive chunks alternate between two streams so the
                // tail of one overlaps the head of the next; scan + pack follow in order
                cudaStream_t ps = (c & 1) ? g.aux2 : g.aux;
                CK(cudaStreamWaitEvent(ps, ev_in[c], 0));
                CK(launch_encode_chunk((const uint8_t *)g.stage_in, 0, lo, len, c == 0, P, pl,
                                       (uint32_t *)g.stage_out, ps, nullptr, 1,
                                       &g.pinned_totals_dev[c], (int)(c & 63)));
                CK(cudaEventRecord(ev_parse[c], ps));
                CK(cudaStreamWaitEvent(g.hi, ev_parse[c], 0));
                CK(launch_encode_chunk((const uint8_t *)g.stage_in, 0, lo, len, c == 0, P, pl,
                                       (uint32_t *)g.stage_out, g.hi, nullptr, 2,
                                       &g.pinned_totals_dev[c], (int)(c & 63)));
            } else {
                CK(cudaStreamWaitEvent(g.hi, ev_in[c], 0));
                CK(launch_encode_chunk((const uint8_t *)g.stage_in, 0, lo, len, c == 0, P, pl,
                                       (uint32_t *)g.stage_out, g.hi, nullptr, 0,
                                       &g.pinned_totals_dev[c], (int)(c & 63)));
            }
            CK(cudaEventRecord(ev_done[c], g.hi));
        }
        CK(cudaEventRecord(g.ev[5], g.hi));  // behind the last chunk's bit-packer
        long done_bytes = 0;
        int result = LZ77_OK;
        unsigned long long k = 0;
        for (long long c = 0; c < n_chunks; c++) {
            CK(cudaEventSynchronize(ev_done[c]));
            k = g.pinned_totals[c];
            const unsigned long long bits = 32ull + k * (unsigned long long)P.tbits;
            // words below the one holding the next token's first bit are final
            long final_bytes = c + 1 == n_chunks ? (long)((bits + 7) / 8) : (long)(bits / 32 * 4);
            if (final_bytes > out_cap) {
                result = LZ77_E_SPACE;
                break;
            }
            if (final_bytes > done_bytes) {
                CK(cudaStreamWaitEvent(g.copy_out, ev_done[c], 0));
                CK(cudaMemcpyAsync(out + done_bytes, (char *)g.stage_out + done_bytes,
                                   (size_t)(final_bytes - done_bytes), cudaMemcpyDeviceToHost,
                                   g.copy_out));
                done_bytes = final_bytes;
            }
        }
        CK(cudaStreamSynchronize(g.copy_out));
        CK(cudaStreamSynchronize(g.hi));
        CK(cudaStreamSynchronize(g.stream));
        CK(cudaStreamSynchronize(g.aux));
        CK(cudaStreamSynchronize(g.aux2));
        if (result != LZ77_OK) return result;
        *n_out = done_bytes;
        memset(&g.last, 0, sizeof g.last);
        g.last.launches = (int)n_chunks * encode_launch_count(chunk, P);
        g.last.n_tokens = (long)k;
        // pipelined call: the stages overlap, so only the span of the compute stream
        // (first H2D issued .. last bit-packer done) is reported, as the search time
        if (g.timing) g.last.enc_search_ms = ms_between(g.ev[4], g.ev[5]);
        return LZ77_OK;
    }

    CK(cudaEventRecord(g.ev[4], g.stream));
    if (n_in > 0) CK(cudaMemcpyAsync(g.stage_in, in, (size_t)n_in, cudaMemcpyHostToDevice, g.stream));
    CK(cudaEventRecord(g.ev[5], g.stream));
    long n = 0, k = 0;
    rc = lz77_gpu_encode_device(g.stage_in, n_in, P.sb, P.la, g.stage_out, (long)o_cap, &n, &k);
    if (rc) return rc;
    if (n > out_cap) return LZ77_E_SPACE;
    CK(cudaEventRecord(g.ev[6], g.stream));
    CK(cudaMemcpyAsync(out, g.stage_out, (size_t)n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev[7], g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *n_out = n;
    if (g.timing) {
        g.last.h2d_ms = ms_between(g.ev[4], g.ev[5]);
        g.last.d2h_ms = ms_between(g.ev[6], g.ev[7]);
    }
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// decode
// ---------------------------------------------------------------------------

namespace {

#ifndef LZ77_JUMP_WINDOW_MAX
#define LZ77_JUMP_WINDOW_MAX 512
#endif
constexpr int kJumpWindowMax = LZ77_JUMP_WINDOW_MAX;  // windows up to this decode by pointer jumping

// runs pass 1; on success *n_out is the decoded size
int decode_scan_device(const void *d_in, long n_in, Params *P, long long *n_tokens, long *n_out,
                       bool *cross_block)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    if (n_in < 0 || !d_in || !n_out) return LZ77_E_ARG;
    if (((uintptr_t)d_in) & 15) return LZ77_E_ARG;  // 16-byte aligned buffers only
    if (n_in < 4) return LZ77_E_STREAM;
    CK(cudaSetDevice(g.device));
    unsigned char hdr[4];
    CK(cudaMemcpyAsync(g.pinned, d_in, 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    memcpy(hdr, g.pinned, 4);
    int rc = read_header(hdr, n_in, P, n_tokens);
    if (rc) return rc;
    memset(&g.last, 0, sizeof g.last);
    g.last.n_tokens = (long)*n_tokens;
    if (*n_tokens == 0) {
        *n_out = 0;
        return LZ77_OK;
    }
    rc = grow(&g.scratch, &g.scratch_cap, decode_scratch_bytes(*n_tokens, *P));
    if (rc) return rc;
    DecodeInfo *d_info = nullptr;
    if (g.timing) CK(cudaEventRecord(g.ev[0], g.stream));
    CK(launch_decode_scan((const uint32_t *)d_in, n_in, *n_tokens, *P, g.scratch, &d_info,
                          g.stream));
    if (g.timing) CK(cudaEventRecord(g.ev[1], g.stream));
    CK(cudaMemcpyAsync(g.pinned, d_info, sizeof(DecodeInfo), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    const DecodeInfo *info = (const DecodeInfo *)g.pinned;
    *n_out = (long)info->n_out;
    // Pointer jumping also for windows so small that nearly every source is still being
    // written by a neighbouring warp of the tile decoder: it does not care
    // how near a source is.
    if (cross_block) *cross_block = info->cross_block != 0 || P->window <= kJumpWindowMax;
    g.last.launches = decode_launch_count(false);
    if (g.timing) g.last.dec_scan_ms = ms_between(g.ev[0], g.ev[1]);
    return LZ77_OK;
}

}  // namespace

namespace {

constexpr long long kMaxDecodeChunks = 48;  // one ticket slot per tile launch (decode.cu)

// Host decode of a large stream as a pipeline over chunks of the compressed
// input: the H2D copy of chunk c+1, the token scan of chunk c, the tile decode
// of the output tiles chunk c completed and the D2H copy of the tiles before
// them all run concurrently (copy_in / compute / aux / copy_out streams).  The
// stage_in buffer has already been sized by the caller.
int decode_pipelined(const unsigned char *in, long n_in, unsigned char *out, long out_cap,
                     long *n_out, long long chunk_bytes, long long n_chunks)
{
    Params P;
    long long K = 0;
    int rc = read_header(in, n_in, &P, &K);
    if (rc) return rc;
    memset(&g.last, 0, sizeof g.last);
    g.last.n_tokens = (long)K;
    if (K == 0) {
        *n_out = 0;
        return LZ77_OK;
    }
    const long long tile_bytes = 1LL << P.tile_shift;
    long long max_out = K << P.lb;  // len + 1 <= 2^lb
    if (max_out > out_cap) max_out = out_cap;
    const size_t o_cap = (((size_t)max_out + tile_bytes) + 15) & ~(size_t)15;
    rc = grow(&g.stage_out, &g.stage_out_cap, o_cap + 16);
    if (rc) return rc;
    rc = grow(&g.scratch, &g.scratch_cap, decode_scratch_bytes(K, P));
    if (rc) return rc;

    cudaEvent_t *evp = pool_events((size_t)(3 * n_chunks));
    if (!evp) return LZ77_E_CUDA;
    cudaEvent_t *ev_in = evp, *ev_scan = evp + n_chunks, *ev_tiles = evp + 2 * n_chunks;
    const size_t in_cap = ((size_t)n_in + 15) & ~(size_t)15;
    StreamDrain drain(g);
    CK(cudaMemsetAsync((char *)g.stage_in + (in_cap - 16), 0, 32, g.stream));
    CK(cudaEventRecord(g.ev[4], g.stream));
    CK(cudaStreamWaitEvent(g.copy_in, g.ev[4], 0));
    CK(cudaStreamWaitEvent(g.copy_out, g.ev[4], 0));
    CK(cudaStreamWaitEvent(g.aux, g.ev[4], 0));
    CK(cudaStreamWaitEvent(g.hi, g.ev[4], 0));

    // the scan kernels write the output position behind their last token straight into
    // pinned memory; a chunk that completes no scan chunk leaves the sentinel
    for (long long c = 0; c < n_chunks; c++) g.pinned_totals[c] = ~0ull;
    // queue every H2D copy and every scan; the scans need nothing from the host
    const long long granule = decode_scan_granule();
    DecodeInfo *d_info = nullptr;
    long long tok_lo = 0;
    for (long long c = 0; c < n_chunks; c++) {
        const long long lo = c * chunk_bytes;
        const long long hi = (c + 1 == n_chunks) ? n_in : lo + chunk_bytes;
        CK(cudaMemcpyAsync((char *)g.stage_in + lo, in + lo, (size_t)(hi - lo),
                           cudaMemcpyHostToDevice, g.copy_in));
        CK(cudaEventRecord(ev_in[c], g.copy_in));
        CK(cudaStreamWaitEvent(g.hi, ev_in[c], 0));
        long long tok_hi = K;
        if (c + 1 < n_chunks) {
            tok_hi = ((hi - 4) * 8) / P.tbits / granule * granule;  // whole tokens, whole scan chunks
            if (tok_hi < tok_lo) tok_hi = tok_lo;
        }
        CK(launch_decode_scan_range((const uint32_t *)g.stage_in, hi, K, tok_lo, tok_hi, P,
                                    g.scratch, &d_info, g.hi, &g.pinned_totals_dev[c]));
        tok_lo = tok_hi;
        CK(cudaEventRecord(ev_scan[c], g.hi));
    }

    // as the scans finish: decode the tiles they completed, copy them back
    long long tiles_done = 0, n_total = 0, pos_seen = 0;
    bool cross = false;  // a match left its block: pointer jumping instead of tiles
    int result = LZ77_OK;
    for (long long c = 0; c < n_chunks; c++) {
        CK(cudaEventSynchronize(ev_scan[c]));
        if (g.pinned_totals[c] != ~0ull) {
            pos_seen = (long long)(g.pinned_totals[c] & ~(1ull << 63));
            cross = (g.pinned_totals[c] >> 63) != 0 ||  // sticky: the flag is never cleared
                    P.window <= kJumpWindowMax;
        }
        const long long pos = pos_seen;
        const bool last = c + 1 == n_chunks;
        long long tile_end;
        if (last) {
            n_total = pos;
            if (pos > out_cap) {
                result = LZ77_E_SPACE;
                break;
            }
            tile_end = (pos + tile_bytes - 1) >> P.tile_shift;
        } else {
            tile_end = pos > 0 ? (pos - 1) >> P.tile_shift : 0;  // tiles wholly scanned
            const long long cap_tiles = (long long)(max_out >> P.tile_shift);
            if (tile_end > cap_tiles) tile_end = cap_tiles;        // never past the caller's buffer
            if (tile_end < tiles_done) tile_end = tiles_done;
        }
        if (tile_end > tiles_done) {
            CK(cudaStreamWaitEvent(g.aux, ev_scan[c], 0));
            if (cross) {
                const long long piece = decode_jump_piece(max_out, P, g.jump_piece);
                if ((rc = grow(&g.jump, &g.jump_cap, decode_jump_scratch_bytes(piece)))) return rc;
                CK(launch_decode_jump_range((const uint32_t *)g.stage_in, n_in, K,
                                            tiles_done << P.tile_shift,
                                            last ? pos : tile_end << P.tile_shift, last, P,
                                            g.scratch, g.jump, piece, (uint8_t *)g.stage_out,
                                            g.aux));
            } else {
                CK(launch_decode_tiles_range((const uint32_t *)g.stage_in, n_in, K, tiles_done,
                                             tile_end, last, pos, (int)c, 0, P, g.scratch,
                                             (uint8_t *)g.stage_out, g.aux));
            }
            CK(cudaEventRecord(ev_tiles[c], g.aux));
            CK(cudaStreamWaitEvent(g.copy_out, ev_tiles[c], 0));
            const long long b_lo = tiles_done << P.tile_shift;
            long long b_hi = tile_end << P.tile_shift;
            if (last && b_hi > pos) b_hi = pos;
            CK(cudaMemcpyAsync(out + b_lo, (char *)g.stage_out + b_lo, (size_t)(b_hi - b_lo),
                               cudaMemcpyDeviceToHost, g.copy_out));
            tiles_done = tile_end;
        }
    }
    CK(cudaStreamSynchronize(g.aux));
    CK(cudaStreamSynchronize(g.copy_out));
    CK(cudaStreamSynchronize(g.hi));
    CK(cudaStreamSynchronize(g.stream));
    if (result == LZ77_OK) {
        CK(cudaMemcpy(g.pinned, d_info, sizeof(DecodeInfo), cudaMemcpyDeviceToHost));
        if (((const DecodeInfo *)g.pinned)->error) result = LZ77_E_STREAM;
    }
    *n_out = n_total;
    g.last.launches = (int)(2 * n_chunks);
    return result;
}

}  // namespace

int lz77_gpu_decode_size_device(const void *d_in, long n_in, long *n_out)
{
    Params P;
    long long k = 0;
    return decode_scan_device(d_in, n_in, &P, &k, n_out, nullptr);
}

int lz77_gpu_slice_tokens_device(const void *d_in, long n_in, long tok_lo, long tok_hi,
                                 void *d_out, long out_cap, long *n_out)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    if (n_in < 0 || !d_in || !d_out || !n_out) return LZ77_E_ARG;
    if ((((uintptr_t)d_in) | ((uintptr_t)d_out)) & 15) return LZ77_E_ARG;
    if (n_in < 4) return LZ77_E_STREAM;
    CK(cudaSetDevice(g.device));
    CK(cudaMemcpyAsync(g.pinned, d_in, 4, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    unsigned char hdr[4];
    memcpy(hdr, g.pinned, 4);
    Params P;
    long long K = 0;
    int rc = read_header(hdr, n_in, &P, &K);
    if (rc) return rc;
    if (tok_lo < 0 || tok_hi < tok_lo || tok_hi > K) return LZ77_E_ARG;
    const long long bytes = 4 + ((long long)(tok_hi - tok_lo) * P.tbits + 7) / 8;
    const long long words = (bytes + 3) / 4;
    if (words * 4 > out_cap) return LZ77_E_SPACE;
    CK(launch_slice_tokens((const uint32_t *)d_in, n_in, tok_lo, tok_hi, P, (uint32_t *)d_out,
                           words, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *n_out = (long)bytes;
    return LZ77_OK;
}

int lz77_gpu_token_at_device(const void *d_in, long n_in, long pos, long *tok, long *tok_pos)
{
    Params P;
    long long k = 0;
    long n = 0;
    if (!tok || !tok_pos) return LZ77_E_ARG;
    int rc = decode_scan_device(d_in, n_in, &P, &k, &n, nullptr);
    if (rc) return rc;
    if (pos < 0 || pos > n) return LZ77_E_ARG;
    if (k == 0 || pos == n) {
        *tok = (long)k;
        *tok_pos = n;
        return LZ77_OK;
    }
    // the result lands behind the DecodeInfo copy in the pinned staging area
    long long *d_result = (long long *)((char *)g.scratch + 64);
    CK(launch_token_at((const uint32_t *)d_in, n_in, k, pos, P, g.scratch, d_result, g.stream));
    CK(cudaMemcpyAsync(g.pinned, d_result, 16, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    *tok = (long)((const long long *)g.pinned)[0];
    *tok_pos = (long)((const long long *)g.pinned)[1];
    return LZ77_OK;
}

int lz77_gpu_decode_device(const void *d_in, long n_in, void *d_out, long out_cap, long *n_out)
{
    Params P;
    long long k = 0;
    long n = 0;
    bool cross = false;
    int rc = decode_scan_device(d_in, n_in, &P, &k, &n, &cross);
    if (rc) return rc;
    *n_out = n;
    if (n == 0) return LZ77_OK;
    if (!d_out || (((uintptr_t)d_out) & 15)) return LZ77_E_ARG;
    if (out_cap < n) return LZ77_E_SPACE;
    const long long piece = decode_jump_piece(n, P, g.jump_piece);
    if (cross && (rc = grow(&g.jump, &g.jump_cap, decode_jump_scratch_bytes(piece)))) return rc;
    if (g.timing) CK(cudaEventRecord(g.ev[2], g.stream));
    CK(launch_decode_copy((const uint32_t *)d_in, n_in, k, n, cross, P, g.scratch, g.jump, piece,
                          (uint8_t *)d_out, g.stream));
    if (g.timing) CK(cudaEventRecord(g.ev[3], g.stream));
    // the error flag is final only after pass 2
    DecodeInfo *d_info = (DecodeInfo *)g.scratch;
    CK(cudaMemcpyAsync(g.pinned, d_info, sizeof(DecodeInfo), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.last.launches = decode_launch_count(true);
    if (g.timing) g.last.dec_copy_ms = ms_between(g.ev[2], g.ev[3]);
    if (((const DecodeInfo *)g.pinned)->error) return LZ77_E_STREAM;
    return LZ77_OK;
}

int lz77_gpu_decode_size(const unsigned char *in, long n_in, long *n_out)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    if (n_in < 0 || !in || !n_out) return LZ77_E_ARG;
    if (n_in < 4) return LZ77_E_STREAM;
    CK(cudaSetDevice(g.device));
    const size_t in_cap = ((size_t)n_in + 15) & ~(size_t)15;
    int rc = grow(&g.stage_in, &g.stage_in_cap, in_cap + 16);
    if (rc) return rc;
    CK(cudaMemsetAsync((char *)g.stage_in + (in_cap - 16), 0, 32, g.stream));
    CK(cudaMemcpyAsync(g.stage_in, in, (size_t)n_in, cudaMemcpyHostToDevice, g.stream));
    return lz77_gpu_decode_size_device(g.stage_in, n_in, n_out);
}

int lz77_gpu_decode(const unsigned char *in, long n_in, unsigned char *out, long out_cap,
                    long *n_out)
{
    if (!g.ready) return LZ77_E_NODEVICE;
    if (n_in < 0 || !in || !n_out) return LZ77_E_ARG;
    if (n_in < 4) return LZ77_E_STREAM;
    CK(cudaSetDevice(g.device));
    const size_t in_cap = ((size_t)n_in + 15) & ~(size_t)15;
    int rc = grow(&g.stage_in, &g.stage_in_cap, in_cap + 16);
    if (rc) return rc;
    {
        // large streams: chunked pipeline (see decode_pipelined); very large ones use
        // bigger chunks so that the number of tile launches stays bounded
        Params hp;
        long long hk = 0;
        long long chunk_bytes = read_header(in, n_in, &hp, &hk) == LZ77_OK ? host_chunk_bytes(hp)
                                                                           : (8ll << 20);
        if ((n_in + chunk_bytes - 1) / chunk_bytes > kMaxDecodeChunks)
            chunk_bytes = ((n_in + kMaxDecodeChunks - 1) / kMaxDecodeChunks + 0xfffff) & ~0xfffffLL;
        const long long n_chunks = (n_in + chunk_bytes - 1) / chunk_bytes;
        if (n_chunks >= 2 && n_chunks <= kMaxDecodeChunks && out)
            return decode_pipelined(in, n_in, out, out_cap, n_out, chunk_bytes, n_chunks);
    }
    CK(cudaEventRecord(g.ev[4], g.stream));
    CK(cudaMemsetAsync((char *)g.stage_in + (in_cap - 16), 0, 32, g.stream));
    CK(cudaMemcpyAsync(g.stage_in, in, (size_t)n_in, cudaMemcpyHostToDevice, g.stream));
    CK(cudaEventRecord(g.ev[5], g.stream));

    Params P;
    long long k = 0;
    long n = 0;
    bool cross = false;
    rc = decode_scan_device(g.stage_in, n_in, &P, &k, &n, &cross);
    if (rc) return rc;
    const float scan_ms = g.last.dec_scan_ms;
    *n_out = n;
    if (n == 0) return LZ77_OK;
    if (!out) return LZ77_E_ARG;
    if (out_cap < n) return LZ77_E_SPACE;
    rc = grow(&g.stage_out, &g.stage_out_cap, (((size_t)n + 15) & ~(size_t)15) + 16);
    if (rc) return rc;
    const long long piece = decode_jump_piece(n, P, g.jump_piece);
    if (cross && (rc = grow(&g.jump, &g.jump_cap, decode_jump_scratch_bytes(piece)))) return rc;
    if (g.timing) CK(cudaEventRecord(g.ev[2], g.stream));
    CK(launch_decode_copy((const uint32_t *)g.stage_in, n_in, k, n, cross, P, g.scratch, g.jump,
                          piece, (uint8_t *)g.stage_out, g.stream));
    if (g.timing) CK(cudaEventRecord(g.ev[3], g.stream));
    CK(cudaMemcpyAsync(g.pinned, g.scratch, sizeof(DecodeInfo), cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev[6], g.stream));
    CK(cudaMemcpyAsync(out, g.stage_out, (size_t)n, cudaMemcpyDeviceToHost, g.stream));
    CK(cudaEventRecord(g.ev[7], g.stream));
    CK(cudaStreamSynchronize(g.stream));
    g.last.launches = decode_launch_count(true);
    g.last.dec_scan_ms = scan_ms;
    if (g.timing) {
        g.last.dec_copy_ms = ms_between(g.ev[2], g.ev[3]);
        g.last.h2d_ms = ms_between(g.ev[4], g.ev[5]);
        g.last.d2h_ms = ms_between(g.ev[6], g.ev[7]);
    }
    if (((const DecodeInfo *)g.pinned)->error) return LZ77_E_STREAM;
    return LZ77_OK;
}

}  // extern "C"

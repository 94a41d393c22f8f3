// comm.cu -- one input, one stream, several GPUs (include/lz77_b200.h, "several GPUs").
//
// The reference is one sequential loop (lz77.c:89-135, 164-195); what makes the path
// shard is this encoder's independent blocks and the format's fixed-width tokens
// (lz77.c:246-252): a rank encodes a run of whole blocks on its own and the payloads
// concatenate; a rank decodes a run of whole blocks from any cut of the token array
// that starts on a block boundary.  The only exchange is moving buffers -- NCCL
// point-to-point over NVLink, grouped -- plus 8-byte all-gathers of counts.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): inside a process that already
// loaded a copy (torch) that copy is used; the library itself keeps loading on
// machines without NCCL, where the sharded entry points return LZ77_E_COMM.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>

#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#include "context.cuh"

using namespace lz77;

namespace lz77 {

// ---------------------------------------------------------------------------
// NCCL, bound at run time (the ABI of these entry points is stable across 2.x)
// ---------------------------------------------------------------------------

namespace {

struct NcclId {
    char internal[LZ77_COMM_ID_BYTES];
};
typedef void *NcclComm;
constexpr int kNcclUint8 = 1, kNcclUint64 = 5;  // ncclDataType_t

struct Nccl {
    void *lib = nullptr;
    int (*GetUniqueId)(NcclId *) = nullptr;
    int (*CommInitRank)(NcclComm *, int, NcclId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*Send)(const void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
    int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl g_nccl;
std::once_flag g_nccl_once;

const Nccl &nccl()
{
    std::call_once(g_nccl_once, [] {
        Nccl &n = g_nccl;
        n.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!n.lib) n.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!n.lib) return;
        bool all = true;
        auto sym = [&](const char *name) {
            void *p = dlsym(n.lib, name);
            if (!p) all = false;
            return p;
        };
        n.GetUniqueId = (decltype(n.GetUniqueId))sym("ncclGetUniqueId");
        n.CommInitRank = (decltype(n.CommInitRank))sym("ncclCommInitRank");
        n.CommDestroy = (decltype(n.CommDestroy))sym("ncclCommDestroy");
        n.Send = (decltype(n.Send))sym("ncclSend");
        n.Recv = (decltype(n.Recv))sym("ncclRecv");
        n.AllGather = (decltype(n.AllGather))sym("ncclAllGather");
        n.Broadcast = (decltype(n.Broadcast))sym("ncclBroadcast");
        n.GroupStart = (decltype(n.GroupStart))sym("ncclGroupStart");
        n.GroupEnd = (decltype(n.GroupEnd))sym("ncclGroupEnd");
        n.GetErrorString = (decltype(n.GetErrorString))sym("ncclGetErrorString");
        n.ok = all;
    });
    return g_nccl;
}

int fail_nccl(int rc, const char *what)
{
    Context &c = ctx();
    const Nccl &n = nccl();
    snprintf(c.err, sizeof c.err, "%s: %s", what,
             n.GetErrorString ? n.GetErrorString(rc) : "NCCL error");
    return LZ77_E_COMM;
}

#define NK(call)                                       \
    do {                                               \
        int rc_ = (call);                              \
        if (rc_ != 0) return fail_nccl(rc_, #call);    \
    } while (0)

// device / pinned scratch of the small exchanges (unsigned long long slots)
constexpr int kSendSlot = 0;     // this rank's contribution (<= 8 values)
constexpr int kGatherSlot = 64;  // world x (<= 8) values
constexpr size_t kCommBytes = 4096;

inline size_t round16(size_t b) { return (b + 15) & ~(size_t)15; }

// every rank contributes n values; out[r * n + i] is value i of rank r.  dev_first, when
// set, is a device pointer whose 8 bytes replace vals[0] (a count a kernel just wrote).
int allgather_u64(Context &c, const unsigned long long *vals, int n,
                  const unsigned long long *dev_first, unsigned long long *out)
{
    const Nccl &nc = nccl();
    cudaStream_t st = c.stream;
    memcpy(c.comm_host + kSendSlot, vals, (size_t)n * 8);
    CK(cudaMemcpyAsync(c.comm_dev + kSendSlot, c.comm_host + kSendSlot, (size_t)n * 8,
                       cudaMemcpyHostToDevice, st));
    if (dev_first)
        CK(cudaMemcpyAsync(c.comm_dev + kSendSlot, dev_first, 8, cudaMemcpyDeviceToDevice, st));
    NK(nc.AllGather(c.comm_dev + kSendSlot, c.comm_dev + kGatherSlot, (size_t)n, kNcclUint64,
                    (NcclComm)c.nccl_comm, st));
    CK(cudaMemcpyAsync(c.comm_host + kGatherSlot, c.comm_dev + kGatherSlot,
                       (size_t)n * c.world * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(out, c.comm_host + kGatherSlot, (size_t)n * c.world * 8);
    c.comm_last.collectives++;
    return LZ77_OK;
}

int bcast_u64(Context &c, unsigned long long *vals, int n, int root)
{
    const Nccl &nc = nccl();
    cudaStream_t st = c.stream;
    if (c.rank == root) {
        memcpy(c.comm_host + kSendSlot, vals, (size_t)n * 8);
        CK(cudaMemcpyAsync(c.comm_dev + kSendSlot, c.comm_host + kSendSlot, (size_t)n * 8,
                           cudaMemcpyHostToDevice, st));
    }
    NK(nc.Broadcast(c.comm_dev + kSendSlot, c.comm_dev + kSendSlot, (size_t)n, kNcclUint64, root,
                    (NcclComm)c.nccl_comm, st));
    CK(cudaMemcpyAsync(c.comm_host + kSendSlot, c.comm_dev + kSendSlot, (size_t)n * 8,
                       cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    memcpy(vals, c.comm_host + kSendSlot, (size_t)n * 8);
    c.comm_last.collectives++;
    return LZ77_OK;
}

// Every rank reports a status; all ranks return the first non-zero one (rank order), so
// a failure on one rank (an allocation, a malformed slice) ends the call on every rank
// instead of leaving the others inside a collective.
int agree(Context &c, int my_rc)
{
    unsigned long long v = (unsigned long long)(long long)my_rc, all[kMaxDevices * 4];
    int rc = allgather_u64(c, &v, 1, nullptr, all);
    if (rc != LZ77_OK) return rc;
    for (int r = 0; r < c.world; r++)
        if ((long long)all[r] != 0) return (int)(long long)all[r];
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// device-side bit shift: dst bits [d0, d0 + nbits) = src bits [s0, s0 + nbits)
// (LSB-first bit numbering of 32-bit little-endian words, bitio.c:203-239).  Words of
// dst that the range covers completely are stored; a first word shared with the bits
// in front of d0 is OR-ed into what is there; a last word that starts inside the
// range is stored with zeros behind the range (the next append ORs into it).  This is
// the <= 7-bit seam merge between the payloads of two ranks when T % 8 != 0.
// ---------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
lz77_copy_bits_kernel(uint32_t *__restrict__ dst, long long d0, const uint32_t *__restrict__ src,
                      long long s0, long long nbits)
{
    const long long w_first = d0 >> 5, w_last = (d0 + nbits - 1) >> 5;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long w = w_first + (long long)blockIdx.x * blockDim.x + threadIdx.x; w <= w_last;
         w += stride) {
        const long long rel = (w << 5) - d0;  // range position of bit 0 of this word
        const int lead = rel < 0 ? (int)(-rel) : 0;                    // bits in front of the range
        const long long left = nbits - rel;                            // range bits from bit 0 on
        const int end = left >= 32 ? 32 : (int)left;                   // first bit behind the range
        const long long sb = s0 + (rel < 0 ? 0 : rel);
        const uint32_t lo = src[sb >> 5], hi = src[(sb >> 5) + 1];
        uint32_t v = __funnelshift_r(lo, hi, (int)(sb & 31)) << lead;
        if (end < 32) v &= (1u << end) - 1u;
        if (lead > 0)
            atomicOr(dst + w, v);
        else
            dst[w] = v;
    }
}

cudaError_t launch_copy_bits(uint32_t *dst, long long d0, const uint32_t *src, long long s0,
                             long long nbits, cudaStream_t st)
{
    if (nbits <= 0) return cudaSuccess;
    const long long words = ((d0 + nbits - 1) >> 5) - (d0 >> 5) + 1;
    long long blocks = (words + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    lz77_copy_bits_kernel<<<(unsigned)blocks, 256, 0, st>>>(dst, d0, src, s0, nbits);
    return cudaGetLastError();
}

void shard_range(long long n_bytes, int world, long long block, int rank, long long *lo,
                 long long *hi)
{
    const long long n_blocks = (n_bytes + block - 1) / block;
    const long long base = n_blocks / world, extra = n_blocks % world;
    const long long lo_blk = base * rank + (rank < extra ? rank : extra);
    const long long cnt = base + (rank < extra ? 1 : 0);
    *lo = lo_blk * block < n_bytes ? lo_blk * block : n_bytes;
    *hi = (lo_blk + cnt) * block < n_bytes ? (lo_blk + cnt) * block : n_bytes;
}

struct CommTimer {  // scatter / compute / gather spans of one call, on the call's stream
    Context &c;
    explicit CommTimer(Context &ctx_) : c(ctx_) {}
    void mark(int i) { cudaEventRecord(c.comm_ev[i], c.stream); }
    void finish()
    {
        c.comm_last.scatter_ms = ms_between(c.comm_ev[0], c.comm_ev[1]);
        c.comm_last.compute_ms = ms_between(c.comm_ev[1], c.comm_ev[2]);
        c.comm_last.gather_ms = ms_between(c.comm_ev[2], c.comm_ev[3]);
        c.comm_last.total_ms = ms_between(c.comm_ev[0], c.comm_ev[3]);
    }
};

}  // namespace

void comm_release(Context &c)
{
    if (c.nccl_comm && nccl().ok) nccl().CommDestroy((NcclComm)c.nccl_comm);
    c.nccl_comm = nullptr;
    c.rank = 0;
    c.world = 1;
    if (c.comm_dev) cudaFree(c.comm_dev);
    if (c.comm_host) cudaFreeHost(c.comm_host);
    c.comm_dev = c.comm_host = nullptr;
    for (auto &e : c.comm_ev)
        if (e) cudaEventDestroy(e);
    for (auto &e : c.comm_ev) e = nullptr;
}

}  // namespace lz77

extern "C" {

int lz77_shard_range(long n_bytes, int world, long block, int rank, long *lo, long *hi)
{
    if (n_bytes < 0 || world < 1 || block < 1 || rank < 0 || rank >= world || !lo || !hi)
        return LZ77_E_ARG;
    long long a = 0, b = 0;
    shard_range(n_bytes, world, block, rank, &a, &b);
    *lo = (long)a;
    *hi = (long)b;
    return LZ77_OK;
}

int lz77_comm_get_unique_id(void *id)
{
    if (!id) return LZ77_E_ARG;
    const Nccl &nc = nccl();
    if (!nc.ok) {
        snprintf(ctx().err, sizeof ctx().err, "libnccl.so.2 not found");
        return LZ77_E_COMM;
    }
    NcclId nid;
    NK(nc.GetUniqueId(&nid));
    memcpy(id, &nid, sizeof nid);
    return LZ77_OK;
}

int lz77_comm_init(const void *id, int rank, int world)
{
    Context &c = ctx();
    if (!c.ready) return LZ77_E_NODEVICE;
    if (!id || world < 1 || world > kMaxDevices || rank < 0 || rank >= world) return LZ77_E_ARG;
    const Nccl &nc = nccl();
    if (!nc.ok) {
        snprintf(c.err, sizeof c.err, "libnccl.so.2 not found");
        return LZ77_E_COMM;
    }
    CK(cudaSetDevice(c.device));
    comm_release(c);
    NcclId nid;
    memcpy(&nid, id, sizeof nid);
    NcclComm comm = nullptr;
    NK(nc.CommInitRank(&comm, world, nid, rank));
    c.nccl_comm = comm;
    c.rank = rank;
    c.world = world;
    CK(cudaMalloc((void **)&c.comm_dev, kCommBytes));
    CK(cudaMallocHost((void **)&c.comm_host, kCommBytes));
    for (auto &e : c.comm_ev) CK(cudaEventCreate(&e));
    memset(&c.comm_last, 0, sizeof c.comm_last);
    return LZ77_OK;
}

void lz77_comm_destroy(void)
{
    Context &c = ctx();
    if (!c.ready) return;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    comm_release(c);
}

int lz77_comm_last_stats(struct lz77_comm_stats *s)
{
    if (!s) return LZ77_E_ARG;
    *s = ctx().comm_last;
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// encode: scatter -> per-rank encode -> count all-gather -> gather into ONE stream
// ---------------------------------------------------------------------------

int lz77_gpu_encode_sharded_device(const void *d_in, long n_in, int sb, int la, void *d_out,
                                   long out_cap, long *n_out, long *n_tokens, int root)
{
    Context &c = ctx();
    if (!c.ready) return LZ77_E_NODEVICE;
    if (!n_out) return LZ77_E_ARG;
    if (c.world == 1 || !c.nccl_comm)
        return lz77_gpu_encode_device(d_in, n_in, sb, la, d_out, out_cap, n_out, n_tokens);
    if (root < 0 || root >= c.world) return LZ77_E_ARG;
    const Nccl &nc = nccl();
    NcclComm comm = (NcclComm)c.nccl_comm;
    cudaStream_t st = c.stream;
    CK(cudaSetDevice(c.device));
    memset(&c.comm_last, 0, sizeof c.comm_last);
    CommTimer tm(c);
    int rc;

    // 1. root checks its arguments and broadcasts the job
    Params P;
    unsigned long long meta[4] = {0, 0, 0, 0};
    if (c.rank == root) {
        int arg = LZ77_OK;
        if (make_params(sb, la, &P) != LZ77_OK || n_in < 0 || !d_out || (n_in > 0 && !d_in))
            arg = LZ77_E_ARG;
        else if ((((uintptr_t)d_in) | ((uintptr_t)d_out)) & 15)
            arg = LZ77_E_ARG;
        else if ((size_t)out_cap < round16((size_t)lz77_gpu_encode_bound(n_in, P.sb, P.la)))
            arg = LZ77_E_SPACE;
        meta[0] = (unsigned long long)n_in;
        meta[1] = arg == LZ77_OK ? ((unsigned long long)P.sb | ((unsigned long long)P.la << 16) |
                                    ((unsigned long long)P.history << 32) |
                                    ((unsigned long long)P.fused_pack << 33))
                                 : 0;
        meta[2] = (unsigned long long)(long long)arg;
    }
    if ((rc = bcast_u64(c, meta, 3, root))) return rc;
    if ((long long)meta[2] != 0) return (int)(long long)meta[2];
    const long long N = (long long)meta[0];
    if (make_params((int)(meta[1] & 0xffff), (int)((meta[1] >> 16) & 0xffff), &P) != LZ77_OK)
        return LZ77_E_ARG;
    P.history = (int)((meta[1] >> 32) & 1);  // root's settings count
    P.fused_pack = (int)((meta[1] >> 33) & 1);
    const int T = P.tbits, world = c.world, rank = c.rank;

    // 2. runs of whole blocks; every rank allocates, then all agree to go on
    long long lo[kMaxDevices], hi[kMaxDevices];
    for (int r = 0; r < world; r++) shard_range(N, world, P.block, r, &lo[r], &hi[r]);
    const long long len = hi[rank] - lo[rank];
    const bool direct = rank == root && lo[rank] == 0;  // root's run is the head of the stream
    // history mode (lz77_gpu_set_history on root; it travels in the job): a run also needs
    // the window in front of it, so the scatter sends that halo along
    long long halo[kMaxDevices];
    for (int r = 0; r < world; r++) {
        const long long want = P.history ? (long long)((P.window + 15) & ~15) : 0;
        halo[r] = lo[r] < want ? lo[r] : want;
    }
    const size_t my_bound = round16((size_t)lz77_gpu_encode_bound(len, P.sb, P.la));
    rc = LZ77_OK;
    if (rank != root) rc = grow(&c.stage_in, &c.stage_in_cap, (size_t)(len + halo[rank]) + 64);
    if (!rc && !direct) rc = grow(&c.stage_out, &c.stage_out_cap, my_bound + 64);
    if (!rc) rc = grow(&c.scratch, &c.scratch_cap, encode_scratch_bytes(len, P));
    if ((rc = agree(c, rc))) return rc;

    // 3. scatter: grouped point-to-point, the runs differ in size
    tm.mark(0);
    NK(nc.GroupStart());
    if (rank == root) {
        for (int r = 0; r < world; r++)
            if (r != root && hi[r] > lo[r]) {
                NK(nc.Send((const char *)d_in + lo[r] - halo[r], (size_t)(hi[r] - lo[r] + halo[r]),
                           kNcclUint8, r, comm, st));
                c.comm_last.sent_bytes += (long)(hi[r] - lo[r] + halo[r]);
            }
    } else if (len > 0) {
        NK(nc.Recv(c.stage_in, (size_t)(len + halo[rank]), kNcclUint8, root, comm, st));
        c.comm_last.recv_bytes += (long)(len + halo[rank]);
    }
    NK(nc.GroupEnd());
    c.comm_last.collectives++;
    tm.mark(1);

    // 4. every rank encodes its run
    const uint8_t *src = rank == root ? (const uint8_t *)d_in + lo[rank]
                                      : (const uint8_t *)c.stage_in + halo[rank];
    const long long pre = rank == root ? lo[rank] : halo[rank];
    uint32_t *dst = direct ? (uint32_t *)d_out : (uint32_t *)c.stage_out;
    unsigned long long *d_total = nullptr;
    CK(launch_encode(src, len, pre, P, c.scratch, dst, &d_total, st, nullptr));
    tm.mark(2);

    // 5. token counts -> bit offset of every payload: 32 + T * sum(K_before)
    unsigned long long zero = 0, K[kMaxDevices];
    if ((rc = allgather_u64(c, &zero, 1, d_total, K))) return rc;
    unsigned long long bit[kMaxDevices + 1];
    bit[0] = kHeaderBits;
    for (int r = 0; r < world; r++) bit[r + 1] = bit[r] + K[r] * (unsigned long long)T;
    const unsigned long long k_total = (bit[world] - kHeaderBits) / (unsigned long long)T;
    auto payload_bytes = [&](int r) { return (size_t)((K[r] * (unsigned long long)T + 7) / 8); };

    // 6. gather the payloads into one stream on root
    if (T % 8 == 0) {
        NK(nc.GroupStart());
        if (rank == root) {
            for (int r = 0; r < world; r++)
                if (r != root && K[r] > 0) {
                    NK(nc.Recv((char *)d_out + bit[r] / 8, payload_bytes(r), kNcclUint8, r, comm, st));
                    c.comm_last.recv_bytes += (long)payload_bytes(r);
                }
        } else if (K[rank] > 0) {
            NK(nc.Send((const char *)c.stage_out + 4, payload_bytes(rank), kNcclUint8, root, comm, st));
            c.comm_last.sent_bytes += (long)payload_bytes(rank);
        }
        NK(nc.GroupEnd());
        c.comm_last.collectives++;
        if (rank == root && !direct) {
            CK(cudaMemcpyAsync(d_out, c.stage_out, 4, cudaMemcpyDeviceToDevice, st));
            if (K[rank] > 0)
                CK(cudaMemcpyAsync((char *)d_out + bit[rank] / 8, (const char *)c.stage_out + 4,
                                   payload_bytes(rank), cudaMemcpyDeviceToDevice, st));
        }
    } else {
        // payloads start inside a byte: root receives them side by side and shifts each
        // into place on the device, in stream order (a seam word is stored by the payload
        // that ends in it and OR-ed into by the one that follows)
        size_t slot[kMaxDevices], need = 0;
        for (int r = 0; r < world; r++) {
            slot[r] = need;
            if (r != root) need += round16(payload_bytes(r) + 16);
        }
        rc = rank == root ? grow(&c.xfer, &c.xfer_cap, need + 64) : LZ77_OK;
        if ((rc = agree(c, rc))) return rc;
        NK(nc.GroupStart());
        if (rank == root) {
            for (int r = 0; r < world; r++)
                if (r != root && K[r] > 0) {
                    NK(nc.Recv((char *)c.xfer + slot[r], payload_bytes(r), kNcclUint8, r, comm, st));
                    c.comm_last.recv_bytes += (long)payload_bytes(r);
                }
        } else if (K[rank] > 0) {
            NK(nc.Send((const char *)c.stage_out + 4, payload_bytes(rank), kNcclUint8, root, comm, st));
            c.comm_last.sent_bytes += (long)payload_bytes(rank);
        }
        NK(nc.GroupEnd());
        c.comm_last.collectives++;
        if (rank == root) {
            if (!direct) CK(cudaMemcpyAsync(d_out, c.stage_out, 4, cudaMemcpyDeviceToDevice, st));
            for (int r = 0; r < world; r++) {
                if (K[r] == 0 || (r == root && direct)) continue;
                const uint32_t *s = r == root ? (const uint32_t *)c.stage_out
                                              : (const uint32_t *)((char *)c.xfer + slot[r]);
                CK(launch_copy_bits((uint32_t *)d_out, (long long)bit[r], s, r == root ? 32 : 0,
                                    (long long)(K[r] * (unsigned long long)T), st));
            }
        }
    }
    tm.mark(3);
    CK(cudaStreamSynchronize(st));
    tm.finish();
    *n_out = (long)((bit[world] + 7) / 8);
    if (n_tokens) *n_tokens = (long)k_total;
    memset(&c.last, 0, sizeof c.last);
    c.last.n_tokens = (long)K[rank];
    c.last.launches = encode_launch_count(len, P);
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// decode: token slices from root -> sums -> split points -> per-rank decode -> gather
// ---------------------------------------------------------------------------

int lz77_gpu_decode_sharded_device(const void *d_in, long n_in, void *d_out, long out_cap,
                                   long *n_out, int root)
{
    Context &c = ctx();
    if (!c.ready) return LZ77_E_NODEVICE;
    if (!n_out) return LZ77_E_ARG;
    if (c.world == 1 || !c.nccl_comm)
        return lz77_gpu_decode_device(d_in, n_in, d_out, out_cap, n_out);
    if (root < 0 || root >= c.world) return LZ77_E_ARG;
    const Nccl &nc = nccl();
    NcclComm comm = (NcclComm)c.nccl_comm;
    cudaStream_t st = c.stream;
    CK(cudaSetDevice(c.device));
    memset(&c.comm_last, 0, sizeof c.comm_last);
    CommTimer tm(c);
    const int world = c.world, rank = c.rank;
    int rc;

    // 1. root reads the header and broadcasts the job
    Params P;
    long long K = 0;
    unsigned long long meta[4] = {0, 0, 0, 0};
    if (rank == root) {
        int arg = LZ77_OK;
        unsigned char hdr[4] = {0, 0, 0, 0};
        if (n_in < 0 || !d_in || (((uintptr_t)d_in) & 15) || (d_out && (((uintptr_t)d_out) & 15)))
            arg = LZ77_E_ARG;
        else if (n_in < 4)
            arg = LZ77_E_STREAM;
        else {
            CK(cudaMemcpyAsync(c.pinned, d_in, 4, cudaMemcpyDeviceToHost, st));
            CK(cudaStreamSynchronize(st));
            memcpy(hdr, c.pinned, 4);
            arg = read_header(hdr, n_in, &P, &K);
        }
        meta[0] = (unsigned long long)n_in;
        meta[1] = hdr[0] | (hdr[1] << 8) | (hdr[2] << 16) | ((unsigned long long)hdr[3] << 24);
        meta[2] = (unsigned long long)(long long)arg;
        meta[3] = (unsigned long long)(d_out ? out_cap : 0);
    }
    if ((rc = bcast_u64(c, meta, 4, root))) return rc;
    if ((long long)meta[2] != 0) return (int)(long long)meta[2];
    const long long n_stream = (long long)meta[0], cap_out = (long long)meta[3];
    unsigned char hdr[4] = {(unsigned char)meta[1], (unsigned char)(meta[1] >> 8),
                            (unsigned char)(meta[1] >> 16), (unsigned char)(meta[1] >> 24)};
    if ((rc = read_header(hdr, (long)n_stream, &P, &K))) return rc;
    const int T = P.tbits;

    // 2. fewer than one block of tokens per rank: root decodes alone
    if (K / world < P.block) {
        unsigned long long res[2] = {0, 0};
        if (rank == root) {
            long n = 0;
            const int drc = lz77_gpu_decode_device(d_in, n_in, d_out, out_cap, &n);
            res[0] = (unsigned long long)(long long)drc;
            res[1] = (unsigned long long)n;
        }
        if ((rc = bcast_u64(c, res, 2, root))) return rc;
        *n_out = (long)res[1];
        return (int)(long long)res[0];
    }

    // 3. even cut of the token array, plus one block of tokens of margin (a token decodes
    //    to at least one byte, so the margin reaches the next block boundary)
    long long k_lo[kMaxDevices], k_hi[kMaxDevices], m_hi[kMaxDevices];
    size_t nb[kMaxDevices], slot[kMaxDevices];
    for (int r = 0; r < world; r++) {
        k_lo[r] = K * r / world;
        k_hi[r] = K * (r + 1) / world;
        m_hi[r] = k_hi[r] + P.block < K ? k_hi[r] + P.block : K;
        nb[r] = (size_t)(((m_hi[r] - k_lo[r]) * T + 7) / 8);
    }
    const long long k_own = k_hi[rank] - k_lo[rank], k_loc = m_hi[rank] - k_lo[rank];
    const size_t n_loc = 4 + nb[rank];
    // xfer: [0, reslice) this rank's run of whole blocks as a standalone stream; behind it,
    // on root when T % 8 != 0, the bit-realigned slices on their way to the other ranks
    const size_t reslice = round16(n_loc + 64);
    size_t need = reslice;
    for (int r = 0; r < world; r++) {
        slot[r] = need;
        if (rank == root && r != root && T % 8 != 0) need += round16(4 + nb[r] + 16);
    }
    rc = grow(&c.stage_in, &c.stage_in_cap, round16(n_loc) + 64);
    if (!rc) rc = grow(&c.xfer, &c.xfer_cap, need + 64);
    if (!rc) rc = grow(&c.scratch, &c.scratch_cap, decode_scratch_bytes(k_loc, P));
    if ((rc = agree(c, rc))) return rc;

    tm.mark(0);
    CK(cudaMemsetAsync((char *)c.stage_in + (round16(n_loc) - 16), 0, 48, st));
    if (rank == root) {
        // own slice (bit-realigned when T % 8 != 0)
        CK(launch_slice_tokens((const uint32_t *)d_in, n_stream, k_lo[rank], m_hi[rank], P,
                               (uint32_t *)c.stage_in, (long long)(n_loc + 3) / 4, st));
        if (T % 8 != 0)
            for (int r = 0; r < world; r++)
                if (r != root)
                    CK(launch_slice_tokens((const uint32_t *)d_in, n_stream, k_lo[r], m_hi[r], P,
                                           (uint32_t *)((char *)c.xfer + slot[r]),
                                           (long long)(4 + nb[r] + 3) / 4, st));
    } else {
        memcpy(c.pinned, hdr, 4);
        CK(cudaMemcpyAsync(c.stage_in, c.pinned, 4, cudaMemcpyHostToDevice, st));
    }
    NK(nc.GroupStart());
    if (rank == root) {
        for (int r = 0; r < world; r++)
            if (r != root) {
                const char *p = T % 8 == 0 ? (const char *)d_in + 4 + k_lo[r] * T / 8
                                           : (const char *)c.xfer + slot[r] + 4;
                NK(nc.Send(p, nb[r], kNcclUint8, r, comm, st));
                c.comm_last.sent_bytes += (long)nb[r];
            }
    } else {
        NK(nc.Recv((char *)c.stage_in + 4, nb[rank], kNcclUint8, root, comm, st));
        c.comm_last.recv_bytes += (long)nb[rank];
    }
    NK(nc.GroupEnd());
    c.comm_last.collectives++;
    tm.mark(1);

    // 4. decoded size of the own tokens -> decoded position of every slice
    DecodeInfo *d_info = nullptr;
    CK(launch_decode_scan_range((const uint32_t *)c.stage_in, (long long)n_loc, k_loc, 0, k_own, P,
                                c.scratch, &d_info, st, nullptr));
    unsigned long long zero = 0, sums[kMaxDevices];
    if ((rc = allgather_u64(c, &zero, 1, &d_info->n_out, sums))) return rc;
    long long pos[kMaxDevices + 1], start[kMaxDevices + 1];
    pos[0] = 0;
    for (int r = 0; r < world; r++) pos[r + 1] = pos[r] + (long long)sums[r];
    const long long n_total = pos[world];
    for (int r = 0; r < world; r++)  // first block boundary at or after the slice's position
        start[r] = (pos[r] + P.block - 1) / P.block * P.block;
    start[world] = n_total;
    *n_out = (long)n_total;
    if (cap_out < n_total) return LZ77_E_SPACE;  // (the same verdict on every rank)

    // 5. the token that starts that block: split points, agreed by all-gather.  A stream
    //    in which no token starts there is not a stream of the block encoder.
    const long long to_boundary = start[rank] - pos[rank];
    CK(launch_decode_scan((const uint32_t *)c.stage_in, (long long)n_loc, k_loc, P, c.scratch,
                          &d_info, st));
    long long *d_result = (long long *)((char *)c.scratch + 64);
    CK(launch_token_at((const uint32_t *)c.stage_in, (long long)n_loc, k_loc, to_boundary, P,
                       c.scratch, d_result, st));
    CK(cudaMemcpyAsync(c.pinned, d_result, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    const long long k_rel = ((const long long *)c.pinned)[0], k_pos = ((const long long *)c.pinned)[1];
    unsigned long long mine[2] = {(unsigned long long)(k_lo[rank] + k_rel),
                                  (unsigned long long)(k_pos == to_boundary ? 1 : 0)},
                       all[kMaxDevices * 2];
    if ((rc = allgather_u64(c, mine, 2, nullptr, all))) return rc;
    long long split[kMaxDevices + 1];
    for (int r = 0; r < world; r++) {
        if (!all[2 * r + 1]) {
            snprintf(c.err, sizeof c.err,
                     "no token starts on a block boundary (rank %d): not a stream of the block encoder", r);
            return LZ77_E_STREAM;
        }
        split[r] = (long long)all[2 * r];
    }
    split[world] = K;

    // 6. every rank decodes its run of whole blocks: root in place, the others into staging
    const long long a = split[rank] - k_lo[rank], b = split[rank + 1] - k_lo[rank];
    const long long size = start[rank + 1] - start[rank];
    rc = LZ77_OK;
    if (b < a || b > k_loc || size < 0) rc = LZ77_E_STREAM;
    void *dst = nullptr;
    if (!rc && rank == root) dst = (char *)d_out + start[rank];
    if (!rc && rank != root && !(rc = grow(&c.stage_out, &c.stage_out_cap, round16((size_t)size) + 64)))
        dst = c.stage_out;
    if (!rc && b > a) {
        const long long n_sl = 4 + ((b - a) * T + 7) / 8;
        CK(cudaMemsetAsync((char *)c.xfer + (round16((size_t)n_sl) - 16), 0, 48, st));
        CK(launch_slice_tokens((const uint32_t *)c.stage_in, (long long)n_loc, a, b, P,
                               (uint32_t *)c.xfer, (n_sl + 3) / 4, st));
        long m = 0;
        rc = lz77_gpu_decode_device(c.xfer, (long)n_sl, dst, (long)((size + 15) & ~15LL), &m);
        if (!rc && m != size) rc = LZ77_E_STREAM;
    } else if (!rc && size != 0) {
        rc = LZ77_E_STREAM;
    }
    tm.mark(2);
    if ((rc = agree(c, rc))) return rc;

    // 7. the plaintext travels to root
    NK(nc.GroupStart());
    if (rank == root) {
        for (int r = 0; r < world; r++)
            if (r != root && start[r + 1] > start[r]) {
                NK(nc.Recv((char *)d_out + start[r], (size_t)(start[r + 1] - start[r]), kNcclUint8, r,
                           comm, st));
                c.comm_last.recv_bytes += (long)(start[r + 1] - start[r]);
            }
    } else if (size > 0) {
        NK(nc.Send(c.stage_out, (size_t)size, kNcclUint8, root, comm, st));
        c.comm_last.sent_bytes += (long)size;
    }
    NK(nc.GroupEnd());
    c.comm_last.collectives++;
    tm.mark(3);
    CK(cudaStreamSynchronize(st));
    tm.finish();
    return LZ77_OK;
}

// ---------------------------------------------------------------------------
// single process, one worker thread per device
// ---------------------------------------------------------------------------

namespace {

struct Pool {
    int n = 0;
    std::vector<std::thread> threads;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    long gen = 0;
    int pending = 0;
    bool stop = false;
    // the job of the current generation
    int kind = 0;  // 1 encode, 2 decode
    const unsigned char *in = nullptr;
    long n_in = 0;
    int sb = 0, la = 0;
    unsigned char *out = nullptr;
    long out_cap = 0;
    long n_out = 0;
    std::vector<int> rc;
    std::vector<int> init_rc;
    char err[256] = {0};
    NcclId id;
};

Pool *g_pool = nullptr;
std::mutex g_pool_mutex;

int run_job(Pool &p, int r)
{
    Context &c = ctx();
    long n = 0;
    if (p.kind == 1) {
        if (r != 0) return lz77_gpu_encode_sharded_device(nullptr, 0, 0, 0, nullptr, 0, &n, nullptr, 0);
        const size_t cap = round16((size_t)lz77_gpu_encode_bound(p.n_in, p.sb, p.la));
        int rc = grow(&c.user_in, &c.user_in_cap, round16((size_t)p.n_in) + 64);
        if (!rc) rc = grow(&c.user_out, &c.user_out_cap, cap + 64);
        if (!rc && p.n_in > 0 &&
            cudaMemcpyAsync(c.user_in, p.in, (size_t)p.n_in, cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
            rc = LZ77_E_CUDA;
        if (!rc) cudaStreamSynchronize(c.stream);
        // (a failed allocation on root still enters the collective, with arguments it rejects)
        rc = lz77_gpu_encode_sharded_device(rc ? nullptr : c.user_in, rc ? -1 : p.n_in, p.sb, p.la,
                                            c.user_out, (long)cap, &n, nullptr, 0);
        if (rc) return rc;
        if (n > p.out_cap) return LZ77_E_SPACE;
        if (cudaMemcpyAsync(p.out, c.user_out, (size_t)n, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
            cudaStreamSynchronize(c.stream) != cudaSuccess)
            return LZ77_E_CUDA;
        p.n_out = n;
        return LZ77_OK;
    }
    if (r != 0) return lz77_gpu_decode_sharded_device(nullptr, 0, nullptr, 0, &n, 0);
    int rc = grow(&c.user_in, &c.user_in_cap, round16((size_t)p.n_in) + 64);
    if (!rc) rc = grow(&c.user_out, &c.user_out_cap, round16((size_t)p.out_cap) + 64);
    if (!rc) {
        cudaMemsetAsync((char *)c.user_in + (round16((size_t)p.n_in) - 16), 0, 48, c.stream);
        if (cudaMemcpyAsync(c.user_in, p.in, (size_t)p.n_in, cudaMemcpyHostToDevice, c.stream) != cudaSuccess)
            rc = LZ77_E_CUDA;
        cudaStreamSynchronize(c.stream);
    }
    rc = lz77_gpu_decode_sharded_device(rc ? nullptr : c.user_in, rc ? -1 : p.n_in, c.user_out,
                                        (long)(p.out_cap & ~15L), &n, 0);
    p.n_out = n;
    if (rc) return rc;
    if (n > 0 && (cudaMemcpyAsync(p.out, c.user_out, (size_t)n, cudaMemcpyDeviceToHost, c.stream) != cudaSuccess ||
                  cudaStreamSynchronize(c.stream) != cudaSuccess))
        return LZ77_E_CUDA;
    return LZ77_OK;
}

void worker(Pool *pp, int r)
{
    Pool &p = *pp;
    int rc = lz77_gpu_init(r);
    if (!rc && p.n > 1) rc = lz77_comm_init(&p.id, r, p.n);
    long seen = 0;
    {
        std::lock_guard<std::mutex> lk(p.m);
        p.init_rc[r] = rc;
        if (rc) snprintf(p.err, sizeof p.err, "%s", ctx().err);
        p.pending--;
    }
    p.cv_done.notify_all();
    for (;;) {
        std::unique_lock<std::mutex> lk(p.m);
        p.cv_job.wait(lk, [&] { return p.stop || p.gen != seen; });
        if (p.stop) break;
        seen = p.gen;
        lk.unlock();
        const int jrc = rc ? rc : run_job(p, r);
        lk.lock();
        p.rc[r] = jrc;
        if (jrc && !p.err[0]) snprintf(p.err, sizeof p.err, "%s", ctx().err);
        p.pending--;
        lk.unlock();
        p.cv_done.notify_all();
    }
    if (!rc) lz77_comm_destroy();
}

int submit(int kind, const unsigned char *in, long n_in, int sb, int la, unsigned char *out,
           long out_cap, long *n_out)
{
    std::lock_guard<std::mutex> guard(g_pool_mutex);
    if (!g_pool) return LZ77_E_NODEVICE;
    Pool &p = *g_pool;
    {
        std::lock_guard<std::mutex> lk(p.m);
        p.kind = kind;
        p.in = in;
        p.n_in = n_in;
        p.sb = sb;
        p.la = la;
        p.out = out;
        p.out_cap = out_cap;
        p.n_out = 0;
        p.err[0] = 0;
        p.pending = p.n;
        p.gen++;
    }
    p.cv_job.notify_all();
    std::unique_lock<std::mutex> lk(p.m);
    p.cv_done.wait(lk, [&] { return p.pending == 0; });
    if (n_out) *n_out = p.n_out;
    for (int r = 0; r < p.n; r++)
        if (p.rc[r]) {
            snprintf(ctx().err, sizeof ctx().err, "%s", p.err);
            return p.rc[r];
        }
    return LZ77_OK;
}

}  // namespace

int lz77_mgpu_init(int n_gpus)
{
    std::lock_guard<std::mutex> guard(g_pool_mutex);
    if (g_pool) return g_pool->n == n_gpus ? LZ77_OK : LZ77_E_ARG;
    const int have = lz77_gpu_device_count();
    if (have <= 0) return LZ77_E_NODEVICE;
    if (n_gpus < 1 || n_gpus > have || n_gpus > kMaxDevices) return LZ77_E_ARG;
    Pool *p = new Pool;
    p->n = n_gpus;
    p->rc.assign(n_gpus, 0);
    p->init_rc.assign(n_gpus, 0);
    if (n_gpus > 1) {
        int rc = lz77_comm_get_unique_id(&p->id);
        if (rc) {
            delete p;
            return rc;
        }
    }
    p->pending = n_gpus;
    for (int r = 0; r < n_gpus; r++) p->threads.emplace_back(worker, p, r);
    {
        std::unique_lock<std::mutex> lk(p->m);
        p->cv_done.wait(lk, [&] { return p->pending == 0; });
    }
    g_pool = p;
    bool init_ok = true;
    for (int r = 0; r < n_gpus; r++) init_ok = init_ok && p->init_rc[r] == 0;
    if (init_ok && n_gpus > 1) {
        // NCCL connects two ranks the first time they talk: do that here (one tiny input that
        // gives every rank a block), not inside the first real call
        const long n = (long)n_gpus * 65536L;
        std::vector<unsigned char> in((size_t)n, 0), out((size_t)lz77_gpu_encode_bound(n, -1, -1) + 64);
        long n_out = 0;
        {
            std::lock_guard<std::mutex> lk(p->m);
            p->kind = 1, p->in = in.data(), p->n_in = n, p->sb = -1, p->la = -1;
            p->out = out.data(), p->out_cap = (long)out.size(), p->n_out = 0;
            p->pending = p->n;
            p->gen++;
        }
        p->cv_job.notify_all();
        std::unique_lock<std::mutex> lk(p->m);
        p->cv_done.wait(lk, [&] { return p->pending == 0; });
        (void)n_out;
    }
    for (int r = 0; r < n_gpus; r++)
        if (p->init_rc[r]) {
            const int rc = p->init_rc[r];
            snprintf(ctx().err, sizeof ctx().err, "%s", p->err);
            // workers that failed stay parked; shut the pool down
            {
                std::lock_guard<std::mutex> lk(p->m);
                p->stop = true;
            }
            p->cv_job.notify_all();
            for (auto &t : p->threads) t.join();
            delete p;
            g_pool = nullptr;
            return rc;
        }
    return LZ77_OK;
}

void lz77_mgpu_shutdown(void)
{
    std::lock_guard<std::mutex> guard(g_pool_mutex);
    if (!g_pool) return;
    {
        std::lock_guard<std::mutex> lk(g_pool->m);
        g_pool->stop = true;
    }
    g_pool->cv_job.notify_all();
    for (auto &t : g_pool->threads) t.join();
    delete g_pool;
    g_pool = nullptr;
}

int lz77_mgpu_encode(const unsigned char *in, long n_in, int sb, int la, unsigned char *out,
                     long out_cap, long *n_out)
{
    if (n_in < 0 || !out || (n_in > 0 && !in) || !n_out) return LZ77_E_ARG;
    return submit(1, in, n_in, sb, la, out, out_cap, n_out);
}

int lz77_mgpu_decode(const unsigned char *in, long n_in, unsigned char *out, long out_cap,
                     long *n_out)
{
    if (n_in < 0 || !in || !n_out) return LZ77_E_ARG;
    return submit(2, in, n_in, 0, 0, out, out_cap, n_out);
}

}  // extern "C"

// context.cuh -- per-device state of the C ABI (capi.cu, comm.cu).
//
// One Context per CUDA device.  lz77_gpu_init(device) binds the CALLING THREAD to the
// context of that device (and makes it the process default for threads that never
// called init), so several host threads can drive several GPUs side by side -- that is
// how the single-process multi-GPU entry points (lz77_mgpu_*) run one rank per device.
// Two threads must not share one device.
#pragma once

#include <stdint.h>

#include <vector>

#include "../../include/lz77_b200.h"
#include "kernels.cuh"

namespace lz77 {

constexpr int kMaxDevices = 16;
constexpr long long kMaxHostChunks = 4096;

struct Context {
    bool ready = false;
    int device = -1;
    cudaStream_t stream = nullptr;      // the stream every call uses
    cudaStream_t own_stream = nullptr;  // created by init; used unless the caller sets one
    cudaStream_t copy_in = nullptr;     // H2D / D2H streams of the chunked host path
    cudaStream_t copy_out = nullptr;
    cudaStream_t aux = nullptr;         // extra compute streams of the chunked host paths
    cudaStream_t aux2 = nullptr;
    cudaStream_t hi = nullptr;          // high-priority stream for the short kernels of a pipeline
    unsigned long long *pinned_totals = nullptr;  // running count per host chunk, written by the
    unsigned long long *pinned_totals_dev = nullptr;  // kernels through this device alias
    std::vector<cudaEvent_t> pool;      // untimed events of the chunked host paths (reused)
    void *scratch = nullptr;
    size_t scratch_cap = 0;
    void *jump = nullptr;  // pointer-jumping state of the cross-block decoder
    size_t jump_cap = 0;
    void *stage_in = nullptr;   // device staging for the host entry points / received shards
    size_t stage_in_cap = 0;
    void *stage_out = nullptr;
    size_t stage_out_cap = 0;
    void *xfer = nullptr;       // sharded paths: token slices in flight, payloads to be merged
    size_t xfer_cap = 0;
    void *user_in = nullptr;    // lz77_mgpu_*: the root's copy of the caller's host buffers
    size_t user_in_cap = 0;
    void *user_out = nullptr;
    size_t user_out_cap = 0;
    unsigned long long *pinned = nullptr;  // small pinned read-back area (256 bytes)
    cudaEvent_t ev[8];
    bool timing = true;
    long long host_chunk = 0;   // lz77_gpu_set_host_chunk(); 0: default
    long long jump_piece = 0;   // lz77_gpu_set_jump_piece(); 0: default
    bool history = false;       // lz77_gpu_set_history()
    bool fused_pack = false;    // lz77_gpu_set_fused_pack()
    lz77_timing last;
    char err[256];

    // communicator of the sharded entry points (comm.cu)
    void *nccl_comm = nullptr;
    int rank = 0, world = 1;
    unsigned long long *comm_dev = nullptr;   // small device area for counts / flags
    unsigned long long *comm_host = nullptr;  // its pinned mirror
    cudaEvent_t comm_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    lz77_comm_stats comm_last;
};

Context &ctx();  // the calling thread's context (a placeholder that is not `ready` before init)

int fail_cuda(cudaError_t rc, const char *what);
int grow(void **buf, size_t *cap, size_t need);
int make_params(int sb, int la, Params *P);
int read_header(const unsigned char hdr[4], long n_in, Params *P, long long *n_tokens);
float ms_between(cudaEvent_t a, cudaEvent_t b);
void comm_release(Context &c);  // comm.cu; called by lz77_gpu_shutdown()

#define CK(call)                                                  \
    do {                                                          \
        cudaError_t rc_ = (call);                                 \
        if (rc_ != cudaSuccess) return lz77::fail_cuda(rc_, #call); \
    } while (0)

}  // namespace lz77

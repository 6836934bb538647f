// Library-level entry points: version, thread-local error string, kernel inventory.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "agx_common.cuh"

namespace agx {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return AGX_ERR_CUDA;
}

// ---- side streams (agx_common.cuh) ------------------------------------------------------------------
namespace {
constexpr int kMaxDev = 16, kSidePerDev = 2;
struct SideRes {
    cudaStream_t s[kSidePerDev];
    cudaEvent_t ev_fork[kSidePerDev], ev_join[kSidePerDev];
    bool ready;
};
SideRes g_side[kMaxDev];
}  // namespace

int SideStream::fork(cudaStream_t main_stream, int which) {
    static const bool off = getenv("AGX_NO_SIDE_STREAMS") != nullptr;
    main = side = main_stream;
    forked = false;
    int dev = 0;
    if (off || which < 0 || which >= kSidePerDev) return AGX_OK;
    AGX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= kMaxDev) return AGX_OK;
    SideRes& R = g_side[dev];
    if (!R.ready) {
        for (int i = 0; i < kSidePerDev; ++i) {
            AGX_CUDA(cudaStreamCreateWithFlags(&R.s[i], cudaStreamNonBlocking));
            AGX_CUDA(cudaEventCreateWithFlags(&R.ev_fork[i], cudaEventDisableTiming));
            AGX_CUDA(cudaEventCreateWithFlags(&R.ev_join[i], cudaEventDisableTiming));
        }
        R.ready = true;
    }
    AGX_CUDA(cudaEventRecord(R.ev_fork[which], main_stream));
    AGX_CUDA(cudaStreamWaitEvent(R.s[which], R.ev_fork[which], 0));
    side = R.s[which];
    forked = true;
    which_ = which;
    return AGX_OK;
}

int SideStream::join() {
    if (!forked) return AGX_OK;
    int dev = 0;
    AGX_CUDA(cudaGetDevice(&dev));
    SideRes& R = g_side[dev];
    AGX_CUDA(cudaEventRecord(R.ev_join[which_], side));
    AGX_CUDA(cudaStreamWaitEvent(main, R.ev_join[which_], 0));
    forked = false;
    return AGX_OK;
}

static std::atomic<uint64_t> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

}  // namespace agx

extern "C" uint64_t agx_launch_count(void) { return agx::g_launches.load(); }

extern "C" int agx_version(void) { return AGX_VERSION; }

extern "C" const char* agx_last_error(void) { return agx::g_err; }

extern "C" int agx_kernel_inventory(char* h_buf, size_t h_buf_bytes) {
    static const char* names =
        "csr_make_keys,radix_hist,radix_scatter,scan_reduce,scan_spine,scan_apply,csr_finalize,"
        "csr_rowptr,coalesce_prepare,coalesce_flag,coalesce_compact,"
        "agg_rows,agg_chunks,agg_chunks_fixup,"
        "gemm_f32,gemm_splitk_reduce,gemm_tf32x3_tc,"
        "sum_arrays,bn_stats,bn_apply,bn_bwd_reduce,bn_bwd_apply,colsum,"
        "log_softmax_nll,log_softmax_nll_bwd,adam_step,head_forward,ce_forward,ce_finish,"
        "smooth_l1,fill_f32,scale_mask,gather_rows,is_identity,transpose,pack_rows,"
        "unpack_rows_add,gat_edge_softmax,sddmm";
    if (!h_buf || h_buf_bytes == 0) {
        agx::set_error("agx_kernel_inventory: null buffer");
        return AGX_ERR_INVALID;
    }
    strncpy(h_buf, names, h_buf_bytes - 1);
    h_buf[h_buf_bytes - 1] = 0;
    return AGX_OK;
}

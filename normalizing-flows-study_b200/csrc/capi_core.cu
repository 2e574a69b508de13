// capi_core.cu -- library-wide C-ABI helpers (status strings, last CUDA error, launch counter).
#include <atomic>
#include <mutex>
#include <string>
#include "nf_common.cuh"

namespace nf {

static std::atomic<int64_t> g_launches{0};
static std::mutex g_err_mu;
static std::string g_last_cuda_error;

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
void count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int nf_set_cuda_error(cudaError_t e) {
    std::lock_guard<std::mutex> lk(g_err_mu);
    g_last_cuda_error = cudaGetErrorString(e);
    return NF_ERR_CUDA;
}

}  // namespace nf

extern "C" int nf_abi_version(void) { return 1; }

extern "C" const char* nf_status_string(int status) {
    switch (status) {
        case NF_OK: return "ok";
        case NF_ERR_BAD_SHAPE: return "bad shape or size argument";
        case NF_ERR_UNSUPPORTED: return "unsupported dtype or configuration";
        case NF_ERR_MISALIGNED: return "pointer not 16-byte aligned";
        case NF_ERR_CUDA: return "CUDA runtime error";
        case NF_ERR_NULL: return "required pointer is NULL";
        case NF_ERR_WORKSPACE: return "workspace or packed buffer too small";
        default: return "unknown status";
    }
}

extern "C" const char* nf_last_cuda_error(void) {
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(nf::g_err_mu);
    copy = nf::g_last_cuda_error;
    return copy.c_str();
}

extern "C" int64_t nf_launch_count(void) { return nf::g_launches.load(std::memory_order_relaxed); }

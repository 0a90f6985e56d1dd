// Error plumbing, device probe and launch accounting for the C ABI.
#include "common.cuh"

namespace asr {

static thread_local std::string g_err;
int64_t g_launches = 0;

void set_error(const std::string &msg) { g_err = msg; }

static int g_dev_checked = 0;
static int g_sm_count = 0;

int ensure_device() {
    if (g_dev_checked == 1) return ASR_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        set_error(std::string("no CUDA device: ") + cudaGetErrorString(e) +
                  " (this library has no CPU fallback)");
        cudaGetLastError();
        return ASR_ERR_CUDA;
    }
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp p;
    e = cudaGetDeviceProperties(&p, dev);
    if (e != cudaSuccess) {
        set_error(std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e));
        return ASR_ERR_CUDA;
    }
    if (p.major != 10) {
        set_error("device is sm_" + std::to_string(p.major) + std::to_string(p.minor) +
                  "; this library is built for sm_100a (B200) only");
        return ASR_ERR_UNSUPPORTED;
    }
    g_sm_count = p.multiProcessorCount;
    g_dev_checked = 1;
    return ASR_OK;
}

int sm_count() { return g_sm_count > 0 ? g_sm_count : 148; }

// Does this process hold an active primary context on `dev`?  (Driver entry point through the runtime: the library
// does not link libcuda.)  Unknown -> true: restoring the caller's device is the conservative answer.
bool device_context_active(int dev) {
    typedef int (*PFN_state)(int, unsigned int *, int *);
    static PFN_state fn = nullptr;
    static bool tried = false;
    if (!tried) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuDevicePrimaryCtxGetState", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_state>(p);
        tried = true;
    }
    if (!fn) return true;
    unsigned int flags = 0;
    int active = 1;
    if (fn(dev, &flags, &active) != 0) return true;      // CUdevice handles are the ordinals
    return active != 0;
}

}  // namespace asr

extern "C" {

const char *asr_last_error(void) { return asr::g_err.c_str(); }
int asr_abi_version(void) { return 2; }
int64_t asr_launch_count(void) { return asr::g_launches; }

}

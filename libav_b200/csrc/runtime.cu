// libav_b200/csrc/runtime.cu -- device selection, sticky error channel, memory helpers of the C-ABI.
// The reference's DSP slots have no error return: CUDA failures are recorded here, never
// swallowed, and there is no CPU fallback anywhere in this library.
#include "common.cuh"
#include "../../include/avdsp_b200.h"
#include <mutex>
#include <string>
#include <utility>
#include <vector>
#include <string.h>

namespace avb {

static std::mutex g_mu;
static std::string g_err;
static bool g_err_is_fault = false;          // g_err holds a CUDA failure (sticks) rather than a refusal (replaceable)
static int g_sms = 0;
static void (*g_log_cb)(int level, const char *msg) = nullptr;

// Two kinds of message share the channel.  A REFUSAL ("format not taken over", bad arguments: the caller falls back to its C path, which
// a drop-in caller does routinely) is kept only until something more important arrives; a FAULT (a CUDA call failed -- the void DSP
// slots have no other way to report it) sticks and replaces a refusal, so that avb200_last_error() never hides a real failure behind
// an earlier benign message.
static void record(const char *where, const char *msg, bool fault)
{
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_err.empty() || (fault && !g_err_is_fault)) { g_err = std::string(where) + ": " + msg; g_err_is_fault = fault; }
    if (g_log_cb) g_log_cb(fault ? 16 /* AV_LOG_ERROR */ : 24 /* AV_LOG_WARNING */, (std::string(where) + ": " + msg).c_str());
}
void set_error_msg(const char *where, const char *msg) { record(where, msg, false); }
void set_error(const char *where, cudaError_t e) { record(where, cudaGetErrorString(e), true); }
// The device avb200_init() chose.  CUDA's current device is PER THREAD and starts at 0: a caller thread the library has not seen yet
// (decoder frame threads, a scaler per worker thread) would otherwise launch onto device 0 with pointers, streams and tensor maps of
// device g_device -- "invalid argument" on every rank but the first.  Every public entry point calls enter() first.
static int g_device = -1;
void enter()
{
    static thread_local int t_dev = -2;
    if (g_device >= 0 && t_dev != g_device) { if (cudaSetDevice(g_device) == cudaSuccess) t_dev = g_device; else cudaGetLastError(); }
}
int check_launch(const char *where)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error(where, e); return -1; }
    return 0;
}
// experiment knobs (bench / profiling only): small string-keyed integer table
static std::mutex g_tune_mu;
static std::vector<std::pair<std::string, int>> g_tune;
int tuning(const char *key)
{
    std::lock_guard<std::mutex> lk(g_tune_mu);
    for (auto &kv : g_tune) if (kv.first == key) return kv.second;
    return 0;
}
int sm_count()
{
    if (g_sms) return g_sms;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
        n = 148;
    g_sms = n;
    return n;
}

}  // namespace avb

using namespace avb;

extern "C" {

void avb200_set_tuning(const char *key, int value)
{
    std::lock_guard<std::mutex> lk(g_tune_mu);
    for (auto &kv : g_tune) if (kv.first == key) { kv.second = value; return; }
    g_tune.emplace_back(key, value);
}

int avb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int avb200_init(int device)
{
    int n = avb200_device_count();
    if (n <= 0) { set_error_msg("avb200_init", "no CUDA device visible (this library has no CPU fallback)"); return -1; }
    if (device < 0 || device >= n) { set_error_msg("avb200_init", "bad device index"); return -1; }
    AVB_CUDA(cudaSetDevice(device), "avb200_init");
    AVB_CUDA(cudaFree(0), "avb200_init");
    cudaDeviceProp p;
    AVB_CUDA(cudaGetDeviceProperties(&p, device), "avb200_init");
    if (p.major != 10) { set_error_msg("avb200_init", "kernels are built for sm_100a only"); return -1; }
    g_sms = p.multiProcessorCount;
    g_device = device;
    return 0;
}

const char *avb200_last_error(void)
{
    static thread_local std::string copy;
    std::lock_guard<std::mutex> lk(g_mu);
    copy = g_err;
    return copy.c_str();
}
void avb200_clear_error(void) { std::lock_guard<std::mutex> lk(g_mu); g_err.clear(); g_err_is_fault = false; }
void avb200_set_log_callback(void (*cb)(int, const char *)) { g_log_cb = cb; }

void *avb200_malloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("avb200_malloc", cudaGetLastError()); return nullptr; }
    return p;
}
void avb200_free(void *p) { if (p) cudaFree(p); }
void *avb200_host_alloc(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) { set_error("avb200_host_alloc", cudaGetLastError()); return nullptr; }
    return p;
}
void avb200_host_free(void *p) { if (p) cudaFreeHost(p); }
int avb200_host_register(void *p, size_t bytes)
{
    AVB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterDefault), "avb200_host_register");
    return 0;
}
int avb200_host_unregister(void *p) { AVB_CUDA(cudaHostUnregister(p), "avb200_host_unregister"); return 0; }

int avb200_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream)
{
    AVB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream), "avb200_memcpy_h2d");
    return 0;
}
int avb200_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream)
{
    AVB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "avb200_memcpy_d2h");
    return 0;
}
int avb200_memcpy2d_h2d(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream)
{
    AVB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyHostToDevice, (cudaStream_t)stream), "avb200_memcpy2d_h2d");
    return 0;
}
int avb200_memcpy2d_d2h(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t height, void *stream)
{
    AVB_CUDA(cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, (cudaStream_t)stream), "avb200_memcpy2d_d2h");
    return 0;
}
int avb200_memset(void *dst, int value, size_t bytes, void *stream)
{
    AVB_CUDA(cudaMemsetAsync(dst, value, bytes, (cudaStream_t)stream), "avb200_memset");
    return 0;
}
void *avb200_stream_create(void)
{
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) { set_error("avb200_stream_create", cudaGetLastError()); return nullptr; }
    return (void *)s;
}
void avb200_stream_destroy(void *s) { if (s) cudaStreamDestroy((cudaStream_t)s); }
int avb200_stream_sync(void *stream)
{
    AVB_CUDA(cudaStreamSynchronize((cudaStream_t)stream), "avb200_stream_sync");
    return 0;
}
int avb200_device_sync(void) { AVB_CUDA(cudaDeviceSynchronize(), "avb200_device_sync"); return 0; }

/* event pair timing on the launching stream (bench.py times kernels launched on the library's own streams
 * with these; torch.cuda.Event only sees torch's current stream) */
void *avb200_event_create(void)
{
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) { set_error("avb200_event_create", cudaGetLastError()); return nullptr; }
    return (void *)e;
}
void avb200_event_destroy(void *e) { if (e) cudaEventDestroy((cudaEvent_t)e); }
int avb200_event_record(void *e, void *stream) { AVB_CUDA(cudaEventRecord((cudaEvent_t)e, (cudaStream_t)stream), "avb200_event_record"); return 0; }
int avb200_event_sync(void *e) { AVB_CUDA(cudaEventSynchronize((cudaEvent_t)e), "avb200_event_sync"); return 0; }
float avb200_event_elapsed_ms(void *a, void *b)
{
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, (cudaEvent_t)a, (cudaEvent_t)b) != cudaSuccess) { set_error("avb200_event_elapsed_ms", cudaGetLastError()); return -1.f; }
    return ms;
}

}  // extern "C"

// libav_b200/csrc/scratch.cu -- see scratch.h
#include "scratch.h"
#include "common.cuh"

namespace avb {

static Scratch g_scratch;
static std::recursive_mutex g_scratch_mu;
Scratch &scratch() { return g_scratch; }
std::recursive_mutex &scratch_mutex() { return g_scratch_mu; }

void *Scratch::dev(int slot, size_t bytes)
{
    if (slot < 0 || slot >= NDEV) return nullptr;
    if (bytes < 256) bytes = 256;
    if (dn_[slot] >= bytes) return d_[slot];
    if (d_[slot]) { cudaFree(d_[slot]); d_[slot] = nullptr; dn_[slot] = 0; }
    size_t want = bytes + bytes / 8;
    if (cudaMalloc(&d_[slot], want) != cudaSuccess) { set_error("scratch.dev", cudaGetLastError()); d_[slot] = nullptr; return nullptr; }
    dn_[slot] = want;
    return d_[slot];
}
static void *grow_pinned(void *&p, size_t &n, size_t bytes)
{
    if (bytes < 4096) bytes = 4096;
    if (n >= bytes) return p;
    if (p) { cudaFreeHost(p); p = nullptr; n = 0; }
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { set_error("scratch.pinned", cudaGetLastError()); p = nullptr; return nullptr; }
    n = bytes;
    return p;
}
void *Scratch::pinned(size_t bytes) { return grow_pinned(h_, hn_, bytes); }
void *Scratch::pinned2(size_t bytes) { return grow_pinned(h2_, h2n_, bytes); }
cudaStream_t *Scratch::streams()
{
    if (!st_ok_) {
        for (int i = 0; i < 3; i++)
            if (cudaStreamCreateWithFlags(&st_[i], cudaStreamNonBlocking) != cudaSuccess) { set_error("scratch.streams", cudaGetLastError()); return nullptr; }
        st_ok_ = true;
    }
    return st_;
}
cudaEvent_t Scratch::event(int i)
{
    if (!ev_ok_) {
        for (int k = 0; k < 4; k++) cudaEventCreateWithFlags(&ev_[k], cudaEventDisableTiming);
        ev_ok_ = true;
    }
    return ev_[i & 3];
}
void Scratch::release()
{
    for (int i = 0; i < NDEV; i++) if (d_[i]) { cudaFree(d_[i]); d_[i] = nullptr; dn_[i] = 0; }
    if (h_) { cudaFreeHost(h_); h_ = nullptr; hn_ = 0; }
    if (h2_) { cudaFreeHost(h2_); h2_ = nullptr; h2n_ = 0; }
}

}  // namespace avb

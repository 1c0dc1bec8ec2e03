// libav_b200/csrc/scratch.h -- grow-only device / pinned staging buffers and the three copy/compute
// streams used by the host-pointer entry points (slot functions and *_host_cuda calls).
// One process drives one GPU (one rank per GPU), so a single process-wide pool is enough; calls that
// use it are serialised by ScratchLock (the batched device-pointer API never touches it).
#pragma once
#include <cuda_runtime.h>
#include <mutex>
#include <stddef.h>

namespace avb {

struct Scratch {
    enum { NDEV = 12 };
    void *dev(int slot, size_t bytes);     // device buffer >= bytes (contents undefined), nullptr on failure
    void *pinned(size_t bytes);            // pinned host buffer >= bytes
    void *pinned2(size_t bytes);           // second pinned buffer
    cudaStream_t *streams();               // 3 non-blocking streams
    cudaEvent_t event(int i);              // small pool of timing-disabled events
    void release();
  private:
    void *d_[NDEV] = {}; size_t dn_[NDEV] = {};
    void *h_ = nullptr; size_t hn_ = 0;
    void *h2_ = nullptr; size_t h2n_ = 0;
    cudaStream_t st_[3] = {}; bool st_ok_ = false;
    cudaEvent_t ev_[4] = {}; bool ev_ok_ = false;
};
Scratch &scratch();
std::recursive_mutex &scratch_mutex();
void enter();
struct ScratchLock { ScratchLock() { enter(); scratch_mutex().lock(); } ~ScratchLock() { scratch_mutex().unlock(); } };

}  // namespace avb

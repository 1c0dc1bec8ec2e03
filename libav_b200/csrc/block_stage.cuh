// libav_b200/csrc/block_stage.cuh -- a warp's 32 consecutive 8x8 int16 blocks (4 KB) staged through shared memory, the layout of the
// 8-bit simple IDCT kernel (idctdsp.cu) as a reusable piece for the thread-per-block transforms (10-bit IDCT, 10-bit forward DCT).
//
//   in:   eight coalesced 512-byte cp.async requests (LDGSTS) per group; 16-byte chunk c of block b lands at b * 128 + ((c ^ (b & 7)) << 4),
//         so that every lane reading row r of ITS block (b = lane) hits a different bank group: conflict-free LDS.128
//   out:  a lane stores its rows back into its own slots, and the tile leaves as eight coalesced 512-byte stores
//   two buffers per warp: the next group is in flight while this one is transformed; no CTA-level barrier anywhere
// A ragged last group is zero-filled on the way in and clipped on the way out.
#pragma once
#include "common.cuh"

namespace avb {

struct WarpBlockStage {
    unsigned tile_s, own, key;
    int lane, l3, lc;
    // tile: the warp's 2 x 4096 bytes of shared memory (128-byte aligned)
    __device__ __forceinline__ WarpBlockStage(void *tile, int lane_) : lane(lane_), l3(lane_ >> 3), lc(lane_ & 7)
    {
        tile_s = (unsigned)__cvta_generic_to_shared(tile);
        own = (unsigned)lane * 128; key = (unsigned)lc << 4;
    }
    __device__ __forceinline__ unsigned buffer(unsigned buf) const { return tile_s + buf; }       // buf = 0 or 4096
    __device__ __forceinline__ unsigned slot(int b) const { return (unsigned)b * 128 + (((unsigned)lc ^ (unsigned)(b & 7)) << 4); }
    __device__ __forceinline__ void issue(const int16_t *blocks, size_t grp, size_t n, unsigned tb) const
    {
        const char *src = reinterpret_cast<const char *>(blocks) + grp * 4096 + lane * 16;
#pragma unroll
        for (int j = 0; j < 8; j++) {                          // request j: blocks 4j .. 4j + 3; this lane: chunk lc of block 4j + l3
            const int b = 4 * j + l3;
            const bool ok = grp * 32 + b < n;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;"
                         :: "r"(tb + slot(b)), "l"(ok ? src + j * 512 : reinterpret_cast<const char *>(blocks)), "r"(ok ? 16 : 0) : "memory");
        }
        cp_async_commit();
    }
    __device__ __forceinline__ uint4 row(unsigned tb, int r) const
    {
        uint4 v;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tb + own + (((unsigned)r << 4) ^ key)));
        return v;
    }
    __device__ __forceinline__ void put_row(unsigned tb, int r, uint4 v) const
    { asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" :: "r"(tb + own + (((unsigned)r << 4) ^ key)), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory"); }
    // tile -> blocks (call after every lane's put_row: it synchronises the warp first)
    __device__ __forceinline__ void flush(int16_t *blocks, size_t grp, size_t n, unsigned tb) const
    {
        __syncwarp();
        char *back = reinterpret_cast<char *>(blocks) + grp * 4096 + lane * 16;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int b = 4 * j + l3;
            uint4 v;
            asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(tb + slot(b)));
            if (grp * 32 + b < n) *reinterpret_cast<uint4 *>(back + j * 512) = v;
        }
    }
};

}  // namespace avb

// libav_b200/csrc/h264intra.cu -- batched H.264 intra reconstruction: what hl_decode_mb() does for
// the intra macroblocks of a picture (libavcodec/h264_mb_template.c:40-, hl_decode_mb_predict_luma h264_mb.c:607-731,
// chroma :hl_decode_mb 158-197): intra prediction (H264PredContext) interleaved with the residual
// (h264_idct_add / dc_add / idct8 / add16intra / add8) block by block, because every block predicts from pixels the
// previous blocks just reconstructed.
//
// Dependencies reach left, up-left, up and up-right, so macroblock rows run as a wavefront: one warp per macroblock row,
// staying two macroblocks behind the row above (the same progress-flag scheme as the deblocking kernel).  A macroblock is
// reconstructed inside a shared-memory tile that also holds its neighbour row / column; prediction of a block runs one
// sample per lane from the edge arrays (h264pred.cuh), the residual reuses the device transforms of h264dsp.cuh.
// Non-intra macroblocks are left as the earlier passes (motion compensation + residual) wrote them; they only feed
// neighbour samples to the intra ones.
#include "h264pred.cuh"
#include "h264dsp.cuh"
#include "../../include/avdsp_b200.h"

namespace avb {

constexpr int LP = 32;        // luma tile: 17 rows x 32; sample (x, y) of the macroblock at [1 + y][4 + x]
constexpr int CP = 16;        // chroma tile: 9 rows x 16; sample (x, y) at [1 + y][4 + x]

__device__ __forceinline__ uint8_t ld_cg8(const uint8_t *p) { return __ldcg(p); }

// raw edges of the n x n block whose top-left sample is tile[r0][c0] (pitch P); `tr_ok`: the top-right samples exist
template <int P>
__device__ __forceinline__ void gather_raw(IntraRaw &r, const uint8_t *tile, int r0, int c0, int n, bool tr_ok)
{
    const uint8_t *top = tile + (r0 - 1) * P + c0;
    r.corner = top[-1];
    for (int i = 0; i < n; i++) { r.top[i] = top[i]; r.left[i] = tile[(r0 + i) * P + c0 - 1]; }
    if (n <= 8) for (int i = n; i < 2 * n; i++) r.top[i] = tr_ok ? top[i] : top[n - 1];
}

__global__ void __launch_bounds__(32)
h264_intra_kernel(const FFH264IntraMB *__restrict__ mbs, int mb_w, int rows_pp, int16_t *__restrict__ coeffs, size_t coeff_stride,
                  const uint8_t *__restrict__ nnzc_all, uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls, int uvls, uint32_t *progress)
{
    __shared__ __align__(16) uint8_t Y[17 * LP];
    __shared__ __align__(16) uint8_t C[2][9 * CP];
    __shared__ FFH264IntraMB M;
    __shared__ uint8_t nnzc[120];
    __shared__ IntraRaw raw;
    __shared__ IntraEdges edges;
    __shared__ IntraBig big;
    const int lane = threadIdx.x, row = blockIdx.x, prow = row % rows_pp;      // prow: row inside its picture
    volatile uint32_t *prog = progress;
    uint8_t *const cplane[2] = { cb, cr };

    for (int x = 0; x < mb_w; x++) {
        if (prow > 0) {
            if (lane == 0) { const uint32_t need = min(x + 2, mb_w); while (prog[row - 1] < need) { } }
            __syncwarp();
        }
        const size_t m = (size_t)row * mb_w + x;
        if (lane < (int)(sizeof(FFH264IntraMB) / 4)) reinterpret_cast<uint32_t *>(&M)[lane] = reinterpret_cast<const uint32_t *>(&mbs[m])[lane];
        // the previous macroblock's last column becomes this one's left column (tile column 3), corner included
        if (x > 0) {
            if (lane < 17) Y[lane * LP + 3] = Y[lane * LP + 19];
            if (lane < 18) { const int p = lane / 9, r = lane % 9; C[p][r * CP + 3] = C[p][r * CP + 11]; }
        }
        __syncwarp();
        // row above: samples x = 0 .. 23 (the last 8 belong to the macroblock up-right)
        if (prow > 0) {
            const uint8_t *g = luma + (size_t)(row * 16 - 1) * ls + x * 16;
            if (lane < 24 && (lane < 16 || x + 1 < mb_w)) Y[4 + lane] = ld_cg8(g + lane);
            if (lane < 16) { const int p = lane >> 3, i = lane & 7; C[p][4 + i] = ld_cg8(cplane[p] + (size_t)(row * 8 - 1) * uvls + x * 8 + i); }
            if (x == 0 && lane == 0) { Y[3] = 0; C[0][3] = C[1][3] = 0; }
        }
        const int kind = M.kind;
        if (kind == 0) {
            // not ours: only keep what its right / lower neighbours will read (last column, via the tile; the row below
            // re-reads from memory)
            for (int i = lane; i < 256; i += 32) Y[(1 + (i >> 4)) * LP + 4 + (i & 15)] = ld_cg8(luma + (size_t)(row * 16 + (i >> 4)) * ls + x * 16 + (i & 15));
            for (int i = lane; i < 128; i += 32) {
                const int p = i >> 6, k = i & 63;
                C[p][(1 + (k >> 3)) * CP + 4 + (k & 7)] = ld_cg8(cplane[p] + (size_t)(row * 8 + (k >> 3)) * uvls + x * 8 + (k & 7));
            }
            __syncwarp();
        } else {
            for (int i = lane; i < 120; i += 32) nnzc[i] = nnzc_all[m * 120 + i];
            __syncwarp();
            int16_t *mb = coeffs + m * coeff_stride;
            if (kind == 1) {                                   // intra 4x4: 16 blocks in coding order
                for (int i = 0; i < 16; i++) {
                    const int bx = (i & 1) + 2 * ((i >> 2) & 1), by = ((i >> 1) & 1) + 2 * (i >> 3);
                    const int mode = M.mode4[i];
                    if (lane == 0) {
                        gather_raw<LP>(raw, Y, 1 + 4 * by, 4 + 4 * bx, 4, (M.topright_samples_available << i) & 0x8000);
                        intra_edges4(edges, raw);
                    }
                    __syncwarp();
                    if (lane < 16) Y[(1 + 4 * by + (lane >> 2)) * LP + 4 + 4 * bx + (lane & 3)] = (uint8_t)intra_directional(edges, 4, mode, lane & 3, lane >> 2);
                    __syncwarp();
                    if (lane == 0) {
                        const int nnz = nnzc[scan8_of(i)];
                        uint8_t *d = &Y[(1 + 4 * by) * LP + 4 + 4 * bx];
                        if (nnz) { if (nnz == 1 && mb[16 * i]) h264_dc_add(d, mb + 16 * i, LP, 4); else h264_idct4_add(d, mb + 16 * i, LP); }
                    }
                    __syncwarp();
                }
            } else if (kind == 2) {                            // intra 8x8
                for (int k = 0; k < 4; k++) {
                    const int i = 4 * k, bx = k & 1, by = k >> 1, mode = M.mode4[i];
                    if (lane == 0) {
                        const bool tl = (M.topleft_samples_available << i) & 0x8000, tr = (M.topright_samples_available << i) & 0x4000;
                        gather_raw<LP>(raw, Y, 1 + 8 * by, 4 + 8 * bx, 8, tr);
                        intra_edges8(edges, raw, tl, tr);
                    }
                    __syncwarp();
                    for (int s = lane; s < 64; s += 32)
                        Y[(1 + 8 * by + (s >> 3)) * LP + 4 + 8 * bx + (s & 7)] = (uint8_t)intra_directional(edges, 8, mode, s & 7, s >> 3);
                    __syncwarp();
                    if (lane == 0) {
                        const int nnz = nnzc[scan8_of(i)];
                        uint8_t *d = &Y[(1 + 8 * by) * LP + 4 + 8 * bx];
                        if (nnz) { if (nnz == 1 && mb[16 * i]) h264_dc_add(d, mb + 16 * i, LP, 8); else h264_idct8_add(d, mb + 16 * i, LP); }
                    }
                    __syncwarp();
                }
            } else {                                           // intra 16x16, then h264_idct_add16intra
                if (lane == 0) { gather_raw<LP>(raw, Y, 1, 4, 16, false); intra_big_prepare(big, raw, 16); }
                __syncwarp();
                for (int s = lane; s < 256; s += 32) Y[(1 + (s >> 4)) * LP + 4 + (s & 15)] = (uint8_t)intra_big_sample(big, raw, 16, M.mode16, s & 15, s >> 4);
                __syncwarp();
                if (lane < 16) {
                    const int bx = (lane & 1) + 2 * ((lane >> 2) & 1), by = ((lane >> 1) & 1) + 2 * (lane >> 3);
                    uint8_t *d = &Y[(1 + 4 * by) * LP + 4 + 4 * bx];
                    if (nnzc[scan8_of(lane)]) h264_idct4_add(d, mb + 16 * lane, LP); else if (mb[16 * lane]) h264_dc_add(d, mb + 16 * lane, LP, 4);
                }
                __syncwarp();
            }
            // chroma: pred8x8 on both planes, then h264_idct_add8
            for (int p = 0; p < 2; p++) {
                if (lane == 0) { gather_raw<CP>(raw, C[p], 1, 4, 8, false); raw.top[8] = 0; intra_big_prepare(big, raw, 8); }
                __syncwarp();
                for (int s = lane; s < 64; s += 32) C[p][(1 + (s >> 3)) * CP + 4 + (s & 7)] = (uint8_t)intra_big_sample(big, raw, 8, M.chroma_mode, s & 7, s >> 3);
                __syncwarp();
            }
            if (M.chroma_residual && lane < 8) {
                const int p = lane >> 2, k = lane & 3, i = 16 + 16 * p + k;
                uint8_t *d = &C[p][(1 + 4 * (k >> 1)) * CP + 4 + 4 * (k & 1)];
                if (nnzc[scan8_of(i)]) h264_idct4_add(d, mb + 16 * i, CP); else if (mb[16 * i]) h264_dc_add(d, mb + 16 * i, CP, 4);
            }
            __syncwarp();
            // write the macroblock back
            for (int i = lane; i < 64; i += 32) {
                const int r = i >> 2, q = i & 3;
                *reinterpret_cast<uint32_t *>(luma + (size_t)(row * 16 + r) * ls + x * 16 + 4 * q) = *reinterpret_cast<const uint32_t *>(&Y[(1 + r) * LP + 4 + 4 * q]);
            }
            {
                const int p = lane >> 4, r = (lane >> 1) & 7, q = lane & 1;
                *reinterpret_cast<uint32_t *>(cplane[p] + (size_t)(row * 8 + r) * uvls + x * 8 + 4 * q) = *reinterpret_cast<const uint32_t *>(&C[p][(1 + r) * CP + 4 + 4 * q]);
            }
        }
        // publish: everything this macroblock wrote must be visible before the row below may read it
        __syncwarp();
        __threadfence();
        if (lane == 0) prog[row] = x + 1;
    }
}

}  // namespace avb

using namespace avb;

extern "C" int ff_h264_intra_mb_batch_cuda(const FFH264IntraMB *mbs, int mb_w, int mb_h, int n_pictures, int16_t *coeffs,
                                           size_t coeff_stride, const uint8_t *nnzc, uint8_t *luma, uint8_t *cb, uint8_t *cr,
                                           int linesize, int uvlinesize, uint32_t *progress, void *stream)
{
    if (mb_w <= 0 || mb_h <= 0 || n_pictures <= 0) return 0;
    if ((linesize & 3) || (uvlinesize & 3) || ((uintptr_t)luma & 3) || ((uintptr_t)cb & 3) || ((uintptr_t)cr & 3)) {
        set_error_msg("ff_h264_intra_mb_batch_cuda", "planes and pitches must be 4-byte aligned"); return -1;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int rows = mb_h * n_pictures;
    AVB_CUDA(cudaMemsetAsync(progress, 0, sizeof(uint32_t) * rows, st), "ff_h264_intra_mb_batch_cuda");
    // a row only ever waits on the row above (a lower block index, dispatched earlier), so any grid size makes progress
    h264_intra_kernel<<<rows, 32, 0, st>>>(mbs, mb_w, mb_h, coeffs, coeff_stride, nnzc, luma, cb, cr, linesize, uvlinesize, progress);
    return check_launch("ff_h264_intra_mb_batch_cuda");
}

// libav_b200/csrc/h264_deblock.cu -- H.264 in-loop deblocking of a batch of pictures (config 3's last stage).
//
// Replaces H264DSPContext.h264_{v,h}_loop_filter_{luma,chroma}{,_intra} (libavcodec/h264dsp_template.c:104-328) applied in the
// reference's order: macroblocks in raster order, inside one the vertical edges 0..3 then the horizontal edges 0..3
// (libavcodec/h264_loopfilter.c:397-415), parameters per edge as filter_mb_edge{v,h,cv,ch} pass them (:104-236; FFH264DeblockMB).
//
// The serial order is a chain along every macroblock row (the vertical edge 0 of macroblock x reads what the horizontal pass of
// macroblock x - 1 left) and a wavefront between rows (the top edge of a macroblock touches the three lines above it).  So the kernel
// spends its parallelism on ROWS, and keeps the step along a row short:
//   * a half-warp (16 lanes) owns one row of one plane kind: 16 luma lines, or 8 cb + 8 cr lines.  A warp carries two rows from
//     different halves of the batch (no dependency between them can exist), so all 32 lanes filter.
//   * vertical edges: a lane holds its line (cols -4 .. 15) unpacked in registers, the four edges are a register-only chain.
//     Through a shared-memory tile the lines become columns; horizontal edges: a lane holds its column (rows -4 .. 15), again a
//     register-only chain; back through the tile the rows are stored.  Two warp barriers per macroblock, no other synchronisation.
//   * everything a macroblock needs from memory is requested one macroblock ahead (its own lines, its parameter record); the lines
//     of the row above -- needed only when the top edge is filtered -- are requested before the vertical pass and used after it.
//   * rows of a picture follow each other as a wavefront: a row publishes how many of its macroblocks are final (release store, issued
//     one macroblock late so that its fence finds the stores it covers already retired),
//     the row below polls that word (acquire load) only for macroblocks whose top edge is filtered -- slice boundaries with
//     disable_deblocking_filter_idc = 2 never wait.  Rows are handed out by an atomic ticket in wavefront order, so the row a warp
//     may wait for is always held by a warp that is already running: no dependence on the order CTAs are dispatched in.
#include "h264dsp.cuh"
#include "../../include/avdsp_b200.h"

namespace avb {

namespace {

__device__ __forceinline__ uint32_t ld_cg(const uint8_t *p) { return __ldcg(reinterpret_cast<const uint32_t *>(p)); }
__device__ __forceinline__ uint32_t ld_acquire(const uint32_t *p)
{ uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t *p)
{ uint32_t v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release(uint32_t *p, uint32_t v)
{ asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ int iabs(int v) { return v < 0 ? -v : v; }
__device__ __forceinline__ int clip3(int v, int lo, int hi) { return min(max(v, lo), hi); }

// bS < 4 luma edge on one line (h264dsp_template.c:104-150)
__device__ __forceinline__ void lf_luma(int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int alpha, int beta, int tc0)
{
    if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta) {
        int tc = tc0, np1 = p1, nq1 = q1;
        const int avg = (p0 + q0 + 1) >> 1;
        if (iabs(p2 - p0) < beta) { if (tc0) np1 = p1 + clip3(((p2 + avg) >> 1) - p1, -tc0, tc0); tc++; }
        if (iabs(q2 - q0) < beta) { if (tc0) nq1 = q1 + clip3(((q2 + avg) >> 1) - q1, -tc0, tc0); tc++; }
        const int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
        p0 = clip_u8(p0 + d); q0 = clip_u8(q0 - d); p1 = np1; q1 = nq1;
    }
}
// bS = 4 luma edge on one line (h264dsp_template.c:166-222)
__device__ __forceinline__ void lf_luma_intra(int p3, int &p2, int &p1, int &p0, int &q0, int &q1, int &q2, int q3, int alpha, int beta)
{
    if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta) {
        int np0, np1 = p1, np2 = p2, nq0, nq1 = q1, nq2 = q2;
        if (iabs(p0 - q0) < ((alpha >> 2) + 2)) {
            if (iabs(p2 - p0) < beta) {
                np0 = (p2 + 2 * p1 + 2 * p0 + 2 * q0 + q1 + 4) >> 3; np1 = (p2 + p1 + p0 + q0 + 2) >> 2; np2 = (2 * p3 + 3 * p2 + p1 + p0 + q0 + 4) >> 3;
            } else np0 = (2 * p1 + p0 + q1 + 2) >> 2;
            if (iabs(q2 - q0) < beta) {
                nq0 = (p1 + 2 * p0 + 2 * q0 + 2 * q1 + q2 + 4) >> 3; nq1 = (p0 + q0 + q1 + q2 + 2) >> 2; nq2 = (2 * q3 + 3 * q2 + q1 + q0 + p0 + 4) >> 3;
            } else nq0 = (2 * q1 + q0 + p1 + 2) >> 2;
        } else {
            np0 = (2 * p1 + p0 + q1 + 2) >> 2; nq0 = (2 * q1 + q0 + p1 + 2) >> 2;
        }
        p0 = np0; p1 = np1; p2 = np2; q0 = nq0; q1 = nq1; q2 = nq2;
    }
}
// chroma edge on one line (h264dsp_template.c:224-328): tc as passed to the slot (tc0 + 1), <= 0 = not filtered; intra = bS 4
__device__ __forceinline__ void lf_chroma(int p1, int &p0, int &q0, int q1, int alpha, int beta, int tc, int intra)
{
    if (iabs(p0 - q0) < alpha && iabs(p1 - p0) < beta && iabs(q1 - q0) < beta) {
        if (intra) { const int a = (2 * p1 + p0 + q1 + 2) >> 2, b = (2 * q1 + q0 + p1 + 2) >> 2; p0 = a; q0 = b; }
        else {
            const int d = clip3((((q0 - p0) << 2) + (p1 - q1) + 4) >> 3, -tc, tc);
            p0 = clip_u8(p0 + d); q0 = clip_u8(q0 - d);
        }
    }
}

__device__ __forceinline__ void unpack4(uint32_t w, int *v)
{ v[0] = (int)(w & 255u); v[1] = (int)((w >> 8) & 255u); v[2] = (int)((w >> 16) & 255u); v[3] = (int)(w >> 24); }
__device__ __forceinline__ uint32_t pack4(const int *v)
{ return (uint32_t)v[0] | ((uint32_t)v[1] << 8) | ((uint32_t)v[2] << 16) | ((uint32_t)v[3] << 24); }

constexpr int DB_WARPS = 4;
constexpr int DB_HALF_WORDS = 112;          // a half-warp's tile: 20 rows x 5 words (luma) / 2 x 10 rows x 3 words (chroma); 112 = 16 (mod 32): the halves' banks interleave
constexpr int DB_PARAM_WORDS = 26;          // sizeof(FFH264DeblockMB) / 4

struct DeblockArgs {
    const FFH264DeblockMB *mbs;
    int mb_w, rows_pp, n_rows, half_rows;   // rows per picture, rows of the batch, rows of the first half of the batch (a multiple of rows_pp)
    uint8_t *luma, *cb, *cr; int ls, uvls;
    uint32_t *prog;                          // [2][n_rows]: luma rows, chroma rows
    uint32_t *ticket;
};

// one pair of rows (rows `pair` and `pair + half_rows`) of one plane kind, from the first to the last macroblock
template <bool CH>
__device__ __forceinline__ void deblock_pair(const DeblockArgs &a, int pair, uint32_t *tile_w, uint32_t *parm_w)
{
    constexpr int W = CH ? 8 : 16;           // macroblock width / height in samples of this plane kind
    constexpr int CTX = CH ? 2 : 4;          // lines of the macroblock above kept above the tile
    constexpr int NE = CH ? 2 : 4;           // edges per direction (chroma edge e = luma edge 2 e)
    constexpr int RW = 1 + W / 4;            // words per tile row: cols -4 .. W - 1
    constexpr int ROWS = CTX + W;            // tile rows per plane
    const int lane = threadIdx.x & 31, half = lane >> 4, hl = lane & 15;
    const int pl = CH ? hl >> 3 : 0, ln = CH ? hl & 7 : hl;          // plane (cb / cr) and line (row pass) or column (column pass) of this lane
    const int row = pair + half * a.half_rows;
    const bool valid = row < a.n_rows;
    const int lrow = valid ? row % a.rows_pp : 0;
    const bool has_above = valid && lrow > 0;
    uint8_t *const plane = CH ? (pl ? a.cr : a.cb) : a.luma;
    const int pitch = CH ? a.uvls : a.ls;
    uint8_t *const grow = plane + (size_t)(valid ? row * W + ln : 0) * pitch;           // this lane's line in the picture (row pass)
    const bool ctx_lane = ln < CTX;                                                      // lanes that fetch / store the lines above
    uint8_t *const gtop = plane + (size_t)(has_above ? row * W - CTX + ln : 0) * pitch;
    uint32_t *const tile = tile_w + half * DB_HALF_WORDS + (CH ? pl * ROWS * RW : 0);   // this lane's plane tile
    uint8_t *const tile_b = reinterpret_cast<uint8_t *>(tile);
    uint32_t *const parm = parm_w + half * 2 * DB_PARAM_WORDS;                            // two parameter buffers per half-warp
    uint32_t *const prog = a.prog + (CH ? a.n_rows : 0);
    const FFH264DeblockMB *const recs = a.mbs + (size_t)(valid ? row : 0) * a.mb_w;

    // prologue: macroblock 0's lines and record
    uint32_t cur[W / 4], left = 0;
#pragma unroll
    for (int k = 0; k < W / 4; k++) cur[k] = valid ? ld_cg(grow + 4 * k) : 0u;
    if (valid) {
        parm[hl] = __ldg(reinterpret_cast<const uint32_t *>(recs) + hl);
        if (hl + 16 < DB_PARAM_WORDS) parm[hl + 16] = __ldg(reinterpret_cast<const uint32_t *>(recs) + hl + 16);
    } else {
        parm[hl] = 0; if (hl + 16 < DB_PARAM_WORDS) parm[hl + 16] = 0;                   // alpha = 0 everywhere: nothing is filtered
    }
    __syncwarp();

    // A release store costs a full memory barrier (it waits for every outstanding access of the warp), so the row only publishes when
    // the row below will look: the leader reads, one macroblock ahead, whether the macroblock below filters its top edge.
    const bool watch = valid && hl == 0 && lrow + 1 < a.rows_pp;
    const uint32_t *const below = reinterpret_cast<const uint32_t *>(recs + a.mb_w);
    bool pub = false;                                                                     // publish at this iteration?
    for (int x = 0; x < a.mb_w; x++) {
        const uint8_t *P = reinterpret_cast<const uint8_t *>(parm + (x & 1) * DB_PARAM_WORDS);
        uint32_t bw[4] = { 0, 0, 0, 0 };                                                  // the top-edge alpha / beta words of the macroblock below
        if (watch) {
            const uint32_t *r = below + (size_t)x * DB_PARAM_WORDS;
            if (!CH) { bw[0] = __ldg(r + 1); bw[1] = __ldg(r + 3); }                      // alpha[1][0..3], beta[1][0..3]
            else { bw[0] = __ldg(r + 13); bw[1] = __ldg(r + 15); bw[2] = __ldg(r + 14); bw[3] = __ldg(r + 16); }   // bytes 52, 60, 56, 64
        }
        // ---- requests for the next macroblock (consumed at the end of this iteration) ----
        uint32_t nxt[W / 4], np0 = 0, np1 = 0;
        const bool more = valid && x + 1 < a.mb_w;
#pragma unroll
        for (int k = 0; k < W / 4; k++) nxt[k] = more ? ld_cg(grow + (x + 1) * W + 4 * k) : 0u;
        if (more) {
            const uint32_t *r = reinterpret_cast<const uint32_t *>(recs + x + 1);
            np0 = __ldg(r + hl);
            if (hl + 16 < DB_PARAM_WORDS) np1 = __ldg(r + hl + 16);
        }
        // ---- does the top edge touch the row above?  then that row's macroblock x must be final ----
        bool top;
        if (!CH) top = has_above && P[4] && P[12];                                       // alpha[1][0], beta[1][0]
        else     top = has_above && ((P[50 + 2] && P[58 + 2]) || (P[50 + 4 + 2] && P[58 + 4 + 2]));   // calpha / cbeta [plane][1][0] of either plane
        {
            // poll with relaxed loads (an acquire load invalidates the SM's L1 every time); one acquire load once the word is there
            // the row above publishes v once its macroblocks < v are final (their last columns included): this top edge needs macroblock x
            const uint32_t need = (uint32_t)min(x + 1, a.mb_w);
            const bool poll = top && hl == 0;
            bool ok = !poll || ld_relaxed(prog + row - 1) >= need;
            while (!__all_sync(0xffffffffu, ok)) ok = !poll || ld_relaxed(prog + row - 1) >= need;
            if (poll) (void)ld_acquire(prog + row - 1);
            __syncwarp();                                     // the leaders' acquire loads are ordered before every lane's loads of the row above
        }
        uint32_t topw[W / 4];
#pragma unroll
        for (int k = 0; k < W / 4; k++) topw[k] = (top && ctx_lane) ? ld_cg(gtop + x * W + 4 * k) : 0u;

        // ---- vertical edges: this lane's line, cols -4 .. W - 1 ----
        int v[4 + W];
        unpack4(left, v);
#pragma unroll
        for (int k = 0; k < W / 4; k++) unpack4(cur[k], v + 4 + 4 * k);
#pragma unroll
        for (int e = 0; e < NE; e++) {
            if (!CH) {
                const int al = P[e], be = P[8 + e];
                if (al && be) {
                    int *c = v + 4 * e;                       // p3 p2 p1 p0 | q0 q1 q2 q3
                    if ((P[48] >> e) & 1) lf_luma_intra(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], al, be);
                    else { const int tc = (int8_t)P[16 + 4 * e + (ln >> 2)]; if (tc >= 0) lf_luma(c[1], c[2], c[3], c[4], c[5], c[6], al, be, tc); }
                }
            } else {
                const int al = P[50 + 4 * pl + e], be = P[58 + 4 * pl + e];
                if (al && be) {
                    int *c = v + 4 * e + 2;                   // p1 p0 | q0 q1
                    const int in = (P[98 + 2 * pl] >> e) & 1, tc = (int8_t)P[66 + 4 * ((pl * 2 + 0) * 2 + e) + (ln >> 1)];
                    if (in || tc > 0) lf_chroma(c[0], c[1], c[2], c[3], al, be, tc, in);
                }
            }
        }
        // the columns left of the macroblock are final now (its left neighbour was filtered horizontally before): store them
        if (valid && x > 0) *reinterpret_cast<uint32_t *>(grow + x * W - 4) = pack4(v);
        uint32_t *trow = tile + (CTX + ln) * RW;
#pragma unroll
        for (int k = 0; k < RW; k++) trow[k] = pack4(v + 4 * k);
        if (ctx_lane) {
#pragma unroll
            for (int k = 0; k < W / 4; k++) tile[ln * RW + 1 + k] = topw[k];
        }
        __syncwarp();

        // ---- horizontal edges: this lane's column, rows -CTX .. W - 1 ----
        int c[ROWS];
#pragma unroll
        for (int k = 0; k < ROWS; k++) c[k] = tile_b[k * RW * 4 + 4 + ln];
#pragma unroll
        for (int e = 0; e < NE; e++) {
            if (!CH) {
                const int al = P[4 + e], be = P[12 + e];
                if (al && be && (e > 0 || top)) {
                    int *q = c + 4 * e;
                    if ((P[49] >> e) & 1) lf_luma_intra(q[0], q[1], q[2], q[3], q[4], q[5], q[6], q[7], al, be);
                    else { const int tc = (int8_t)P[32 + 4 * e + (ln >> 2)]; if (tc >= 0) lf_luma(q[1], q[2], q[3], q[4], q[5], q[6], al, be, tc); }
                }
            } else {
                const int al = P[50 + 4 * pl + 2 + e], be = P[58 + 4 * pl + 2 + e];
                if (al && be && (e > 0 || top)) {
                    int *q = c + 4 * e;                       // rows -2 -1 | 0 1 of the edge
                    const int in = (P[98 + 2 * pl + 1] >> e) & 1, tc = (int8_t)P[66 + 4 * ((pl * 2 + 1) * 2 + e) + (ln >> 1)];
                    if (in || tc > 0) lf_chroma(q[0], q[1], q[2], q[3], al, be, tc, in);
                }
            }
        }
#pragma unroll
        for (int k = 1; k < ROWS; k++) tile_b[k * RW * 4 + 4 + ln] = (uint8_t)c[k];      // (row -CTX is context only)
        __syncwarp();

        // ---- rows out: the macroblock's own lines, and the lines above when the top edge ran ----
        // First publish what is already out: macroblocks < x are final (this iteration's vertical pass has just rewritten the last
        // columns of macroblock x - 1); their stores were issued long ago and are ordered before this point by the barriers above, so
        // the release fence finds them retired instead of waiting for the stores that follow.
        if (pub) st_release(prog + row, (uint32_t)x);
        if (valid) {
#pragma unroll
            for (int k = 0; k < W / 4; k++) {
                const uint32_t w = trow[1 + k];
                *reinterpret_cast<uint32_t *>(grow + x * W + 4 * k) = w;
                if (k == W / 4 - 1) left = w;                 // the next macroblock's left context (its edge 0 will rewrite these columns)
            }
            if (top && ctx_lane && ln >= 1) {
#pragma unroll
                for (int k = 0; k < W / 4; k++) *reinterpret_cast<uint32_t *>(gtop + x * W + 4 * k) = tile[ln * RW + 1 + k];
            }
        }
        // the next macroblock's record into the other parameter buffer
        if (more) {
            uint32_t *pn = parm + ((x + 1) & 1) * DB_PARAM_WORDS;
            pn[hl] = np0;
            if (hl + 16 < DB_PARAM_WORDS) pn[hl + 16] = np1;
        }
#pragma unroll
        for (int k = 0; k < W / 4; k++) cur[k] = nxt[k];
        // macroblock x of the row below waits for the value x + 1, which the next iteration publishes
        if (!CH) pub = watch && (bw[0] & 255u) && (bw[1] & 255u);
        else     pub = watch && (((bw[0] & 255u) && (bw[1] & 255u)) || ((bw[2] & 255u) && (bw[3] & 255u)));
        __syncwarp();                                         // the tile and the parameter buffer may be reused
    }
    if (valid && hl == 0) st_release(prog + row, (uint32_t)a.mb_w);  // (after the barrier that ended the last iteration: every lane's stores are ordered before it)
}

__global__ void __launch_bounds__(DB_WARPS * 32, 8)
h264_deblock_kernel_v3(const DeblockArgs a)
{
    __shared__ __align__(16) uint32_t tile_s[DB_WARPS][2 * DB_HALF_WORDS];
    __shared__ __align__(16) uint32_t parm_s[DB_WARPS][2 * 2 * DB_PARAM_WORDS];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tickets = 2 * a.half_rows;                    // even tickets: luma pairs, odd tickets: chroma pairs, both in wavefront order
    for (;;) {
        int t = 0;
        if (lane == 0) t = (int)atomicAdd(a.ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        if (t >= n_tickets) break;
        if (t & 1) deblock_pair<true>(a, t >> 1, tile_s[warp], parm_s[warp]);
        else       deblock_pair<false>(a, t >> 1, tile_s[warp], parm_s[warp]);
        __syncwarp();
    }
}

}  // namespace

// a small ring of ticket counters: launches on different streams may overlap
static uint32_t *g_tickets = nullptr;
static unsigned g_ticket_next = 0;
constexpr unsigned N_TICKETS = 256;

int launch_h264_deblock_v3(const FFH264DeblockMB *mbs, int mb_w, int mb_h, int n_pictures, uint8_t *luma, uint8_t *cb, uint8_t *cr, int ls,
                           int uvls, uint32_t *progress, cudaStream_t st)
{
    if (!g_tickets) AVB_CUDA(cudaMalloc(&g_tickets, N_TICKETS * sizeof(uint32_t)), "h264_deblock:tickets");
    uint32_t *ticket = g_tickets + (g_ticket_next++ % N_TICKETS);
    const int rows = mb_h * n_pictures;
    AVB_CUDA(cudaMemsetAsync(progress, 0, sizeof(uint32_t) * rows * 2, st), "h264_deblock_picture");
    AVB_CUDA(cudaMemsetAsync(ticket, 0, sizeof(uint32_t), st), "h264_deblock_picture");
    DeblockArgs a;
    a.mbs = mbs; a.mb_w = mb_w; a.rows_pp = mb_h; a.n_rows = rows; a.half_rows = ((n_pictures + 1) / 2) * mb_h;
    a.luma = luma; a.cb = cb; a.cr = cr; a.ls = ls; a.uvls = uvls; a.prog = progress; a.ticket = ticket;
    const int n_tickets = 2 * a.half_rows;
    int grid = (n_tickets + DB_WARPS - 1) / DB_WARPS;
    const int cap = sm_count() * 8;
    if (grid > cap) grid = cap;
    h264_deblock_kernel_v3<<<grid, DB_WARPS * 32, 0, st>>>(a);
    return check_launch("h264_deblock_picture");
}

}  // namespace avb

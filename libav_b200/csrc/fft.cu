// libav_b200/csrc/fft.cu -- float FFT / MDCT filterbank (FFTContext slots, libavcodec/fft.h:73-99) for sm_100a.
//   fft_permute + fft_calc           libavcodec/fft_template.c:186-194, :140-346  (natural-order DFT, no 1/n)
//   imdct_half / imdct_calc / mdct_calc   libavcodec/mdct_template.c:95-214, rotation tables as ff_mdct_init (:86-92)
// Contract for this row: 1e-6 relative to the reference's float result (north star), not bit-exactness.
// One CTA per transform: the sequence is bit-reversed into (bank-swizzled) shared memory, log2(n) radix-2 stages -- four per pass, 16 points per
// thread in registers -- with twiddles read from a per-size table computed in double on the host (so the only float rounding is the
// butterflies'), MDCT pre/post rotations fused around the core.  HBM traffic is one read and one write of the data: 16 B per complex point.
#include "common.cuh"
#include "scratch.h"
#include "../../include/avdsp_b200.h"
#include <math.h>
#include <map>
#include <mutex>
#include <string.h>
#include <utility>
#include <vector>

namespace avb {

__device__ __forceinline__ unsigned bitrev(unsigned v, int bits) { return __brev(v) >> (32 - bits); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) { return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }

// Shared-memory index swizzle: element i lives at i ^ ((i >> 4) & 15).  A thread of the 16-point pass reads elements 16 b + m (one stride of
// 128 bytes per lane: without the swizzle all 32 lanes hit the same two banks), the bit-reversed scatter writes elements that differ in their
// top bits only; with it both spread over the banks, and every later pass (consecutive k per lane) stays a permutation inside 16 elements.
__device__ __forceinline__ int sw(int i) { return i ^ ((i >> 4) & 15); }

// One pass = F radix-2 DIT stages (st .. st + F - 1) on 2^F points carried in registers: the same butterflies in the same order as stage-by-stage
// radix 2, so the result is bit-identical to it.  Stage st + q pairs (m, m + 2^q); element i0 + m * half sits at offset k + (m mod 2^q) * half
// inside its group of that stage, which is its twiddle index (times n >> (stage + 1)).
template <int F>
__device__ __forceinline__ void fft_pass(float2 *s, int n, int st, const float2 *__restrict__ tw)
{
    constexpr int P = 1 << F;
    const int half = 1 << st;
    for (int b = threadIdx.x; b < (n >> F); b += blockDim.x) {
        const int k = b & (half - 1), i0 = ((b >> st) << (st + F)) + k;
        float2 x[P];
#pragma unroll
        for (int m = 0; m < P; m++) x[m] = s[sw(i0 + m * half)];
#pragma unroll
        for (int q = 0; q < F; q++) {
            const int tstep = n >> (st + q + 1);
#pragma unroll
            for (int m = 0; m < P; m++) {
                if (m & (1 << q)) continue;
                const float2 w = __ldg(&tw[(k + (m & ((1 << q) - 1)) * half) * tstep]);
                const float2 t = cmul(x[m + (1 << q)], w);
                x[m + (1 << q)] = make_float2(x[m].x - t.x, x[m].y - t.y);
                x[m] = make_float2(x[m].x + t.x, x[m].y + t.y);
            }
        }
#pragma unroll
        for (int m = 0; m < P; m++) s[sw(i0 + m * half)] = x[m];
    }
}

// in-place DIT on bit-reversed (and swizzled) data in shared memory; tw[k] = exp(sign * 2 pi i k / n), k < n / 2.  Four stages per pass while
// that many remain (16 points per thread: a 1024-point transform is three passes and two barriers instead of five and five), then 3 / 2 / 1.
__device__ __forceinline__ void fft_smem(float2 *s, int nbits, const float2 *__restrict__ tw)
{
    const int n = 1 << nbits;
    int st = 0;
    while (st < nbits) {
        const int rem = nbits - st, f = rem >= 4 ? (rem == 5 ? 3 : 4) : rem;
        __syncthreads();
        if (f == 4) fft_pass<4>(s, n, st, tw); else if (f == 3) fft_pass<3>(s, n, st, tw); else if (f == 2) fft_pass<2>(s, n, st, tw); else fft_pass<1>(s, n, st, tw);
        st += f;
    }
    __syncthreads();
}

// natural-order complex FFT.  revtab != NULL: the input is in the reference's fft_permute order (z_perm[revtab[j]] = z[j])
__global__ void fft_kernel(float2 *__restrict__ z, int nbits, const float2 *__restrict__ tw, const uint16_t *__restrict__ revtab)
{
    extern __shared__ float2 sm[];
    const int n = 1 << nbits;
    float2 *zz = z + (size_t)blockIdx.x * n;
    for (int j = threadIdx.x; j < n; j += blockDim.x) sm[sw(bitrev(j, nbits))] = zz[revtab ? revtab[j] : j];
    fft_smem(sm, nbits, tw);
    for (int j = threadIdx.x; j < n; j += blockDim.x) zz[j] = sm[sw(j)];
}

// op 0 imdct_half (in n/2 -> out n/2), 1 imdct_calc (n/2 -> n), 2 mdct_calc (n -> n/2); FFT of n/4 points inside.
// tcs = tcos[0 .. n/4) followed by tsin[0 .. n/4)
__global__ void mdct_kernel(int op, int nbits, float *__restrict__ out, const float *__restrict__ in, const float2 *__restrict__ tw,
                            const float *__restrict__ tcs)
{
    extern __shared__ float2 sm[];
    const int n = 1 << nbits, n2 = n >> 1, n4 = n >> 2, n8 = n >> 3, n3 = 3 * n4, fb = nbits - 2;
    const float *tcos = tcs, *tsin = tcs + n4;
    const float *x = in + (size_t)blockIdx.x * (op == 2 ? n : n2);
    float *y = out + (size_t)blockIdx.x * (op == 1 ? n : n2);
    if (op < 2) {                                              // pre rotation, mdct_template.c:109-117
        for (int k = threadIdx.x; k < n4; k += blockDim.x) {
            const float a = x[n2 - 1 - 2 * k], b = x[2 * k], c = tcos[k], s = tsin[k];
            sm[sw(bitrev(k, fb))] = make_float2(a * c - b * s, a * s + b * c);
        }
    } else {                                                   // :163-175
        for (int i = threadIdx.x; i < n8; i += blockDim.x) {
            float re = -x[2 * i + n3] - x[n3 - 1 - 2 * i], im = -x[n4 + 2 * i] + x[n4 - 1 - 2 * i];
            float c = -tcos[i], s = tsin[i];
            sm[sw(bitrev(i, fb))] = make_float2(re * c - im * s, re * s + im * c);
            re = x[2 * i] - x[n2 - 1 - 2 * i]; im = -x[n2 + 2 * i] - x[n - 1 - 2 * i];
            c = -tcos[n8 + i]; s = tsin[n8 + i];
            sm[sw(bitrev(n8 + i, fb))] = make_float2(re * c - im * s, re * s + im * c);
        }
    }
    fft_smem(sm, fb, tw);
    float *half = op == 1 ? y + n4 : y;                        // imdct_calc builds its middle half first (:139)
    for (int k = threadIdx.x; k < n8; k += blockDim.x) {
        const float2 za = sm[sw(n8 - k - 1)], zb = sm[sw(n8 + k)];
        float r0, i0, r1, i1;
        if (op < 2) {                                          // :120-130
            float a = za.y, b = za.x, c = tsin[n8 - k - 1], s = tcos[n8 - k - 1];
            r0 = a * c - b * s; i1 = a * s + b * c;
            a = zb.y; b = zb.x; c = tsin[n8 + k]; s = tcos[n8 + k];
            r1 = a * c - b * s; i0 = a * s + b * c;
        } else {                                               // :180-189
            float a = za.x, b = za.y, c = -tsin[n8 - k - 1], s = -tcos[n8 - k - 1];
            i1 = a * c - b * s; r0 = a * s + b * c;
            a = zb.x; b = zb.y; c = -tsin[n8 + k]; s = -tcos[n8 + k];
            i0 = a * c - b * s; r1 = a * s + b * c;
        }
        half[2 * (n8 - k - 1)] = r0; half[2 * (n8 - k - 1) + 1] = i0;
        half[2 * (n8 + k)] = r1; half[2 * (n8 + k) + 1] = i1;
    }
    if (op == 1) {                                             // mirror, :141-144
        __syncthreads();                                       // the CTA's own global writes of `half` are visible after the barrier
        for (int k = threadIdx.x; k < n4; k += blockDim.x) { y[k] = -y[n2 - k - 1]; y[n - k - 1] = y[n2 + k]; }
    }
}

// ---- device tables ------------------------------------------------------------------------------------
static std::mutex g_fft_mu;
static std::map<std::pair<int, int>, float2 *> g_tw;                       // (nbits, inverse) -> n/2 twiddles
static std::map<std::pair<int, double>, float *> g_rot;                    // (nbits, scale)   -> tcos | tsin

static const float2 *twiddles(int nbits, int inverse)
{
    std::lock_guard<std::mutex> lk(g_fft_mu);
    auto key = std::make_pair(nbits, inverse ? 1 : 0);
    auto it = g_tw.find(key);
    if (it != g_tw.end()) return it->second;
    const int n = 1 << nbits, h = n > 1 ? n / 2 : 1;
    std::vector<float2> w(h);
    for (int k = 0; k < h; k++) {
        double a = (inverse ? 2.0 : -2.0) * M_PI * k / n;
        w[k] = make_float2((float)cos(a), (float)sin(a));
    }
    float2 *d = nullptr;
    if (cudaMalloc(&d, sizeof(float2) * h) != cudaSuccess || cudaMemcpy(d, w.data(), sizeof(float2) * h, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("fft tables", cudaGetLastError()); return nullptr;
    }
    g_tw[key] = d;
    return d;
}
static const float *rotations(int nbits, double scale)
{
    std::lock_guard<std::mutex> lk(g_fft_mu);
    auto key = std::make_pair(nbits, scale);
    auto it = g_rot.find(key);
    if (it != g_rot.end()) return it->second;
    const int n = 1 << nbits, n4 = n >> 2;
    std::vector<float> t(2 * n4);
    const double theta = 1.0 / 8.0 + (scale < 0 ? n4 : 0), sc = sqrt(fabs(scale));      // ff_mdct_init, mdct_template.c:86-92
    for (int i = 0; i < n4; i++) {
        double alpha = 2 * M_PI * (i + theta) / n;
        t[i] = (float)(-cos(alpha) * sc);
        t[n4 + i] = (float)(-sin(alpha) * sc);
    }
    float *d = nullptr;
    if (cudaMalloc(&d, sizeof(float) * 2 * n4) != cudaSuccess || cudaMemcpy(d, t.data(), sizeof(float) * 2 * n4, cudaMemcpyHostToDevice) != cudaSuccess) {
        set_error("mdct tables", cudaGetLastError()); return nullptr;
    }
    g_rot[key] = d;
    return d;
}
static int threads_for(int points) { int t = points / 16; return t < 32 ? 32 : t > 256 ? 256 : t; }      // a thread carries 16 points through four stages

int launch_fft(int nbits, int inverse, float *z, size_t n_tr, const uint16_t *d_revtab, cudaStream_t st)
{
    if (!n_tr) return 0;
    if (nbits < 1 || nbits > 12) { set_error_msg("fft_batch", "nbits must be 1..12"); return -1; }
    const float2 *tw = twiddles(nbits, inverse);
    if (!tw) return -1;
    fft_kernel<<<(unsigned)n_tr, threads_for(1 << nbits), sizeof(float2) << nbits, st>>>((float2 *)z, nbits, tw, d_revtab);
    return check_launch("fft_batch");
}
int launch_mdct(int op, int nbits, float *out, const float *in, size_t n_tr, const float *d_tcs, int inverse_core, cudaStream_t st)
{
    if (!n_tr) return 0;
    if (op < 0 || op > 2 || nbits < 4 || nbits > 14) { set_error_msg("mdct_batch", "op 0..2, nbits 4..14"); return -1; }
    const float2 *tw = twiddles(nbits - 2, inverse_core);
    if (!tw) return -1;
    mdct_kernel<<<(unsigned)n_tr, threads_for(1 << (nbits - 2)), sizeof(float2) << (nbits - 2), st>>>(op, nbits, out, in, tw, d_tcs);
    return check_launch("mdct_batch");
}

}  // namespace avb

using namespace avb;

extern "C" {

int ff_fft_batch_cuda(int nbits, int inverse, float *z, size_t n_transforms, void *stream)
{ avb::enter(); return launch_fft(nbits, inverse, z, n_transforms, nullptr, (cudaStream_t)stream); }

int ff_mdct_batch_cuda(int op, int nbits, double scale, float *out, const float *in, size_t n_transforms, void *stream)
{
    avb::enter();
    if (!n_transforms) return 0;
    if (nbits < 4 || nbits > 14) { set_error_msg("mdct_batch", "nbits 4..14"); return -1; }
    const float *rot = rotations(nbits, scale);
    if (!rot) return -1;
    return launch_mdct(op, nbits, out, in, n_transforms, rot, op != 2, (cudaStream_t)stream);   // imdct: inverse core FFT (ff_mdct_init(.., inverse = 1))
}

// ---- table slots: same contract as the C slots, HOST pointers, the context's own revtab / tcos / tsin are honoured ----
static void slot_fft_calc(FFTContext *s, FFTComplex *z)
{
    ScratchLock lk;
    Scratch &S = scratch();
    const int n = 1 << s->nbits;
    cudaStream_t *st = S.streams();
    uint8_t *d = (uint8_t *)S.dev(5, (size_t)n * 8 + (size_t)n * 2 + 64);
    if (!st || !d) return;
    uint16_t *d_rev = (uint16_t *)(d + (size_t)n * 8);
    if (cudaMemcpyAsync(d, z, (size_t)n * 8, cudaMemcpyHostToDevice, st[0]) != cudaSuccess ||
        cudaMemcpyAsync(d_rev, s->revtab, (size_t)n * 2, cudaMemcpyHostToDevice, st[0]) != cudaSuccess) { set_error("fft_calc slot", cudaGetLastError()); return; }
    if (launch_fft(s->nbits, s->inverse, (float *)d, 1, d_rev, st[0])) return;
    if (cudaMemcpyAsync(z, d, (size_t)n * 8, cudaMemcpyDeviceToHost, st[0]) != cudaSuccess || cudaStreamSynchronize(st[0]) != cudaSuccess)
        set_error("fft_calc slot", cudaGetLastError());
}
static void mdct_slot(int op, FFTContext *s, FFTSample *output, const FFTSample *input)
{
    ScratchLock lk;
    Scratch &S = scratch();
    const int nbits = s->mdct_bits, n = 1 << nbits, n2 = n >> 1, nin = op == 2 ? n : n2, nout = op == 1 ? n : n2;
    cudaStream_t *st = S.streams();
    float *d = (float *)S.dev(5, sizeof(float) * (size_t)(nin + nout + n2) + 64);
    if (!st || !d) return;
    float *d_in = d, *d_out = d + nin, *d_tcs = d + nin + nout;
    // tcos and tsin are one allocation in the reference (tsin = tcos + n/4 for FF_MDCT_PERM_NONE, mdct_template.c:76-79)
    if (s->tsin != s->tcos + (n >> 2)) { set_error_msg("mdct slot", "interleaved MDCT tables are not taken over"); return; }
    if (cudaMemcpyAsync(d_in, input, sizeof(float) * nin, cudaMemcpyHostToDevice, st[0]) != cudaSuccess ||
        cudaMemcpyAsync(d_tcs, s->tcos, sizeof(float) * n2, cudaMemcpyHostToDevice, st[0]) != cudaSuccess) { set_error("mdct slot", cudaGetLastError()); return; }
    if (launch_mdct(op, nbits, d_out, d_in, 1, d_tcs, s->inverse, st[0])) return;
    if (cudaMemcpyAsync(output, d_out, sizeof(float) * nout, cudaMemcpyDeviceToHost, st[0]) != cudaSuccess || cudaStreamSynchronize(st[0]) != cudaSuccess)
        set_error("mdct slot", cudaGetLastError());
}
static void slot_imdct_calc(FFTContext *s, FFTSample *o, const FFTSample *i) { mdct_slot(1, s, o, i); }
static void slot_imdct_half(FFTContext *s, FFTSample *o, const FFTSample *i) { mdct_slot(0, s, o, i); }
static void slot_mdct_calc(FFTContext *s, FFTSample *o, const FFTSample *i) { mdct_slot(2, s, o, i); }

void ff_fft_init_cuda(FFTContext *s)
{
    avb::enter();
    if (s->nbits < 1 || s->nbits > 12) return;
    s->fft_calc = slot_fft_calc;                   // fft_permute stays the reference's: fft_calc accepts its revtab order
}
void ff_mdct_init_cuda(FFTContext *s)
{
    avb::enter();
    if (s->mdct_bits < 4 || s->mdct_bits > 14 || s->mdct_permutation != FF_MDCT_PERM_NONE) return;
    s->imdct_calc = slot_imdct_calc; s->imdct_half = slot_imdct_half; s->mdct_calc = slot_mdct_calc;
    s->mdct_calcw = slot_mdct_calc;                // float build: mdct_calcw == mdct_calc (mdct_template.c:68)
}

}  // extern "C"

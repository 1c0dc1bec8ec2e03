// tools/vabsdiff_peak.cu -- measured issue rate of the instructions the SAD search is made of (profiling aid, not product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/vabsdiff_peak tools/vabsdiff_peak.cu && tools/vabsdiff_peak
// Every thread runs 8 independent accumulator chains of `vabsdiff4.u32.u32.u32.add` (4 byte |a-b| + accumulate per
// instruction), respectively of dp4a and of a plain integer add, so the loop is bound by issue rate, not latency.
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(256) burn(unsigned *out, unsigned seed, int iters)
{
    unsigned a[8], acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) { a[k] = seed * (threadIdx.x + 1) + 0x01020304u * k; acc[k] = k; }
    unsigned b = seed ^ 0x9e3779b9u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a[k]), "r"(b));
            else if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(acc[k]) : "r"(a[k]), "r"(b));
            else asm volatile("add.u32 %0, %0, %1;" : "+r"(acc[k]) : "r"(a[k]));
        }
        b += 0x01010101u;
    }
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP> double run(const char *name, int bytes_per_instr)
{
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * 8, iters = 1 << 14;
    unsigned *out; cudaMalloc(&out, grid * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    burn<OP><<<grid, 256>>>(out, 1, 64);
    cudaEventRecord(e0);
    burn<OP><<<grid, 256>>>(out, 7, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)grid * 256 * 8.0 * iters;
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"op\": \"%s\", \"warp_instr_per_clk_per_sm\": %.3f, \"T_instr_per_s\": %.3f, \"T_byte_ops_per_s\": %.3f, \"ms\": %.3f, \"sms\": %d}\n",
           name, instr / 32.0 / (ms * 1e-3) / (clk * 1e3) / sms, instr / (ms * 1e-3) / 1e12, instr * bytes_per_instr / (ms * 1e-3) / 1e12, ms, sms);
    cudaFree(out);
    return ms;
}

int main()
{
    run<2>("add.u32", 0);
    run<1>("dp4a.u32.u32", 4);
    run<0>("vabsdiff4.add", 4);
    return 0;
}

// Micro-benchmark: issue rate of the integer instructions the IB kernels are made of (per SM and clock), one B200.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024) rate_kernel(uint32_t* out, uint32_t seed, int iters)
{
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    const uint32_t k1 = seed | 0x80u, k2 = seed ^ 0x55u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (OP == 1) asm volatile("shf.r.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (OP == 3) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                if (OP == 4) {   // alternating lop3 / mad
                    if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                    else asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                }
                if (OP == 5) {   // alternating lop3 / dp4a
                    if (i & 1) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                    else asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                }
                if (OP == 6) {   // alternating mad / dp4a
                    if (i & 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                    else asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
                }
                if (OP == 7) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(k1), "r"(k2));
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0xdeadbeefu) out[threadIdx.x] = s;
}

template <int OP>
void run(const char* name, uint32_t* d)
{
    const int iters = 4096, sms = 148;
    rate_kernel<OP><<<sms, 1024>>>(d, 3, 16);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    rate_kernel<OP><<<sms, 1024>>>(d, 3, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ops = (double)iters * 64 * 1024;   // per SM (thread-level)
    const double clk = ms * 1e-3 * khz * 1e3;
    printf("%-22s %8.3f ms  %7.1f thread-ops/clk/SM (at the %d MHz attribute clock)\n", name, ms, ops / clk, khz / 1000);
}

int main()
{
    uint32_t* d;
    cudaMalloc(&d, 4096);
    run<0>("lop3", d);
    run<1>("shf", d);
    run<2>("imad", d);
    run<3>("dp4a", d);
    run<7>("prmt", d);
    run<4>("lop3+imad 1:1", d);
    run<5>("lop3+dp4a 1:1", d);
    run<6>("imad+dp4a 1:1", d);
    return 0;
}

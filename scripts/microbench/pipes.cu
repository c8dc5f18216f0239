// Throughput microbenchmark of the integer ops the column pass leans on (sm_100a): per-SM ops/clk for
// dp4a, popc, lop3, iadd3, imad, prmt with 8 independent chains per thread, 1024 threads/SM resident.
#include <cstdio>
#include <cuda_runtime.h>
template <int OP> __global__ void k(unsigned* out, int iters, unsigned seed) {
    unsigned a[8];
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 8 + i;
    unsigned b = seed * 3 + 1, c = seed ^ 0x01010101u;
    for (int it = 0; it < iters; it++) {
        #pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __dp4a(a[i], b, c);
            else if (OP == 1) a[i] = __popc(a[i]) + b;
            else if (OP == 2) a[i] = (a[i] & b) ^ c;
            else if (OP == 3) a[i] = a[i] + b + c;
            else if (OP == 4) a[i] = a[i] * b + c;
            else if (OP == 5) a[i] = __byte_perm(a[i], b, 0x5140);
            else if (OP == 6) a[i] = __funnelshift_r(a[i], b, c);
        }
    }
    unsigned s = 0;
    for (int i = 0; i < 8; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP> void run(const char* name) {
    unsigned* d; cudaMalloc(&d, 148 * 4 * 256 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148 * 4, 256>>>(d, 16, 1);
    cudaEventRecord(e0);
    k<OP><<<148 * 4, 256>>>(d, iters, 7);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double ops = 148.0 * 4 * 256 * 8.0 * iters;
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%-8s %8.3f ms  %7.1f Gop/s  %6.1f thread-ops/clk/SM (at %d MHz)\n", name, ms, ops / ms / 1e6, ops / (ms * 1e-3) / (clk * 1e3) / 148.0, clk / 1000);
    cudaFree(d);
}
int main() { run<0>("dp4a"); run<1>("popc"); run<2>("lop3"); run<3>("iadd3"); run<4>("imad"); run<5>("prmt"); run<6>("shf"); return 0; }

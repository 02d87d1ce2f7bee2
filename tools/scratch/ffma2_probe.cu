// Issue-rate probe: scalar FFMA / FMUL+FADD vs packed FFMA2 on sm_100a.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false -o ffma2_probe ffma2_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0,{%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void unpack(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1},%2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0,%1,%2,%3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
constexpr int ACC = 8, ITERS = 4096;
// MODE 0: scalar fmaf (8 per iter/acc-set)   1: scalar mul + add   2: FFMA2   3: two FFMA2 = exact packed mul then add
template <int MODE>
__global__ void probe(float* out, float k, float nz, float one) {
    float a[ACC * 2];
    for (int i = 0; i < ACC * 2; ++i) a[i] = threadIdx.x * 0.001f + i;
    u64 p[ACC];
    for (int i = 0; i < ACC; ++i) p[i] = pack(a[2 * i], a[2 * i + 1]);
    const u64 K = pack(k, k), NZ = pack(nz, nz), ONE = pack(one, one);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < ACC; ++i) {
            if (MODE == 0) { a[2 * i] = fmaf(a[2 * i], k, nz); a[2 * i + 1] = fmaf(a[2 * i + 1], k, nz); }
            if (MODE == 1) { a[2 * i] = a[2 * i] * k + nz; a[2 * i + 1] = a[2 * i + 1] * k + nz; }
            if (MODE == 2) p[i] = fma2(p[i], K, NZ);
            if (MODE == 3) p[i] = fma2(fma2(p[i], K, NZ), ONE, NZ);
        }
    }
    float s = 0;
    for (int i = 0; i < ACC; ++i) { float x, y; unpack(p[i], x, y); s += x + y + a[2 * i] + a[2 * i + 1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* d, double flops_per_thread_iter) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256;
    probe<MODE><<<blocks, threads>>>(d, 1.0000001f, -0.0f, 1.0f);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; ++r) probe<MODE><<<blocks, threads>>>(d, 1.0000001f, -0.0f, 1.0f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 5;
    const double vals = (double)blocks * threads * ITERS * ACC * 2;   // scalar element-updates per launch
    printf("%-28s %8.3f ms  %7.2f G element-updates/s  (%.1f per clk per SM @1.965GHz)\n", name, ms, vals / ms / 1e6, vals / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA", d, 2); run<1>("scalar FMUL+FADD", d, 2); run<2>("FFMA2", d, 2); run<3>("FFMA2 x2 (exact mul,add)", d, 2);
    return 0;
}

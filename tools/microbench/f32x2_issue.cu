// Microbenchmark: does packed FFMA2 (fma.rn.f32x2) relieve the issue slots of an unfused mul/add stream on sm_100a?
// Variants (per loop iteration, per thread):
//   0: 16 scalar ops   (8 chains x {FMUL, FADD})                       -- what hana_core.cuh's xmul/xadd compile to
//   1:  8 packed ops   (4 chains x {FFMA2(x,a,-0), FFMA2(x,1,b)})      -- the same 16 roundings, exact
//   2: variant 0 + 16 integer ALU ops (LOP3/IADD3 chains)
//   3: variant 1 + 16 integer ALU ops
//   4: 16 integer ALU ops only
//   5: 16 scalar FFMA (3 register operands)
// Prints warp-instructions per clock per SM and ns per iteration.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float& a, float& b) { asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int V>
__global__ void __launch_bounds__(256) k(float* out, float a, float b, float nz, float one, uint32_t ka, int iters, long long* cyc) {
    float x[8];
    u64 X[4];
    uint32_t n[8];
    for (int i = 0; i < 8; i++) { x[i] = 1.f + threadIdx.x * 1e-3f + i; n[i] = threadIdx.x * 7 + i; }
    for (int i = 0; i < 4; i++) X[i] = pk(x[2 * i], x[2 * i + 1]);
    const u64 A = pk(a, a), B = pk(b, b), NZ = pk(nz, nz), ONE = pk(one, one);
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
        if (V == 0 || V == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) { x[i] = __fmul_rn(x[i], a); x[i] = __fadd_rn(x[i], b); }
        }
        if (V == 1 || V == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) { X[i] = fma2(X[i], A, NZ); X[i] = fma2(X[i], ONE, B); }
        }
        if (V == 5) {
#pragma unroll
            for (int i = 0; i < 8; i++) { x[i] = __fmaf_rn(x[i], a, b); x[i] = __fmaf_rn(x[i], a, b); }
        }
        if (V == 2 || V == 3 || V == 4) {
#pragma unroll
            for (int i = 0; i < 8; i++) { n[i] = (n[i] ^ ka) + (n[i] >> 3); n[i] = (n[i] & ka) | (n[(i + 1) & 7] << 1);  }
        }
    }
    long long t1 = clock64();
    float s = 0; uint32_t m = 0;
    for (int i = 0; i < 8; i++) { s += x[i]; m ^= n[i]; }
    for (int i = 0; i < 4; i++) { float p, q; upk(X[i], p, q); s += p + q; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)m;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int V>
void run(const char* name, int fp_inst, int int_inst) {
    int iters = 20000, grid = 148 * 4;
    float* out; long long* cyc;
    cudaMalloc(&out, grid * 256 * 4); cudaMalloc(&cyc, 8);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<V><<<grid, 256>>>(out, 0.999f, 1e-3f, -0.f, 1.f, 0x5a5a5a5au, 100, cyc);
    cudaEventRecord(e0);
    k<V><<<grid, 256>>>(out, 0.999f, 1e-3f, -0.f, 1.f, 0x5a5a5a5au, iters, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    double per_iter_clk = (double)c / iters;             // clocks per iteration with 32 warps per SM resident
    double warps_per_sm = 4 * 8;
    printf("%-34s %7.3f ms  %8.2f clk/iter  fp-inst/clk/SM %5.2f  int-inst/clk/SM %5.2f  total %5.2f\n", name, ms, per_iter_clk,
           fp_inst * warps_per_sm / per_iter_clk, int_inst * warps_per_sm / per_iter_clk, (fp_inst + int_inst) * warps_per_sm / per_iter_clk);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    run<0>("0 scalar FMUL+FADD x16", 16, 0);
    run<1>("1 packed FFMA2 x8 (same roundings)", 8, 0);
    run<2>("2 scalar x16 + int x32", 16, 32);
    run<3>("3 packed x8 + int x32", 8, 32);
    run<4>("4 int x32", 0, 32);
    run<5>("5 scalar FFMA x16", 16, 0);
    return 0;
}

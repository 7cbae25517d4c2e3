// ubench.cu -- per-SM throughput of the instructions the event-detector arithmetic is made of
// (B200, sm_100a).  Prints warp-instructions per clock per SM for each op at full occupancy.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o ubench ubench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITER 2048
template <int OP> __global__ void __launch_bounds__(256) k(float *out, float seed, long long *cyc) {
    float a[8]; double d[8];
    for (int i = 0; i < 8; i++) { a[i] = seed + threadIdx.x * 1e-3f + i; d[i] = (double) a[i] + 0.123456789; }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) a[i] = __fadd_rn(a[i], 1.0001f);
            if (OP == 1) d[i] = __dadd_rn(d[i], 1.0001);
            if (OP == 2) d[i] = __fma_rn(d[i], 0.999, 1.0001);
            if (OP == 3) { d[i] = (double) a[i]; a[i] = __fadd_rn(a[i], (float) (unsigned) (__double2hiint(d[i]) & 1)); }   // F2F f32->f64 (+1 fadd, +lop)
            if (OP == 4) { a[i] = (float) d[i]; d[i] = __dadd_rn(d[i], (double) 0.5) ; }                      // F2F f64->f32 + dadd
            if (OP == 5) a[i] = __fsqrt_rn(a[i] + 2.0f);
            if (OP == 6) a[i] = __fdiv_rn(a[i], 1.0001f + a[(i + 1) & 7]);
            if (OP == 7) d[i] = __ddiv_rn(d[i], 3.0);
            if (OP == 8) d[i] = __ddiv_rn(d[i], d[(i + 1) & 7]);
            if (OP == 9) a[i] = __fdiv_rn(a[i], 3.0f);
            if (OP == 10) { double q = __dmul_rn(d[i], 0.33333333333333331); double r = __fma_rn(-3.0, q, d[i]); d[i] = __fma_rn(r, 0.33333333333333331, q); }
            if (OP == 11) { float q = __fmul_rn(a[i], 0.333333343f); float r = __fmaf_rn(-3.0f, q, a[i]); a[i] = __fmaf_rn(r, 0.333333343f, q); }
            if (OP == 12) { unsigned u = __float_as_uint(a[i]); unsigned hi = ((u >> 3) & 0x0FFFFFFFu) + 0x38000000u | (u & 0x80000000u); d[i] = __hiloint2double(hi, u << 29); a[i] = __fadd_rn(a[i], (float) (hi & 1)); }  // integer f32->f64
            if (OP == 13) a[i] = __fmul_rn(a[i], 1.0001f);
        }
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < 8; i++) s += a[i] + (float) d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP> void run(const char *name, int extra) {
    float *out; long long *cyc;
    int nb = 148 * 8;
    cudaMalloc(&out, nb * 256 * 4); cudaMalloc(&cyc, nb * 8);
    k<OP><<<nb, 256>>>(out, 1.5f, cyc); cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0); k<OP><<<nb, 256>>>(out, 1.5f, cyc); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h[4]; cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
    double winst = (double) nb * 8 * ITER * 8;   // warp-level op groups
    // clock64 cycles of one CTA covers its own run; use event time and 1.965 GHz nominal for a rough rate too
    printf("%-28s %8.3f ms  %7.2f Gop-groups/s (warp)  ~%.3f warp-ops/clk/SM @1.965GHz   cta cycles %lld\n", name, ms, winst / ms / 1e6,
           winst / (ms * 1e-3) / 1.965e9 / 148, h[0]);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FADD", 0); run<13>("FMUL", 0); run<1>("DADD", 0); run<2>("DFMA", 0);
    run<3>("F2F f32->f64 (+FADD,LOP)", 0); run<4>("F2F f64->f32 (+DADD)", 0);
    run<12>("int f32->f64 (+FADD)", 0);
    run<5>("fsqrt_rn (+FADD)", 0); run<6>("fdiv_rn (+FADD)", 0); run<9>("fdiv_rn by 3.0f", 0); run<11>("markstein f32 /3", 0);
    run<7>("ddiv_rn by 3.0", 0); run<8>("ddiv_rn generic", 0); run<10>("markstein f64 /3", 0);
    return 0;
}

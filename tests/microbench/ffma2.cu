// Dev microbenchmark: does fma.rn.f32x2 (Blackwell packed FP32) free issue slots?  nvcc -arch=sm_100a -O3 ffma2.cu -o ffma2
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE> __global__ void k(float *out, int *iout, float a, float b, int m) {
    float x[8];
    float2 y[4];
    int z[8];
    for (int i = 0; i < 8; i++) { x[i] = threadIdx.x * 0.001f + i; z[i] = threadIdx.x + i; }
    for (int i = 0; i < 4; i++) y[i] = make_float2(x[2 * i], x[2 * i + 1]);
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = __fmaf_rn(x[i], a, b);
        }
        if (MODE == 1 || MODE == 3) {
#pragma unroll
            for (int i = 0; i < 4; i++) y[i] = __ffma2_rn(y[i], a2, b2);
        }
        if (MODE >= 2) {
#pragma unroll
            for (int i = 0; i < 8; i++) z[i] = (z[i] ^ m) + it;
        }
    }
    float s = 0; int t = 0;
    for (int i = 0; i < 8; i++) { s += x[i]; t += z[i]; }
    for (int i = 0; i < 4; i++) s += y[i].x + y[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}
template <int MODE> void run(const char *name, float *o, int *io) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<148 * 4, 256>>>(o, io, 1.0001f, 0.5f, 12345);
    cudaEventRecord(e0);
    k<MODE><<<148 * 4, 256>>>(o, io, 1.0001f, 0.5f, 12345);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s %.3f ms\n", name, ms);
}
int main() {
    float *o; int *io; cudaMalloc(&o, 148 * 4 * 256 * 4); cudaMalloc(&io, 148 * 4 * 256 * 4);
    run<0>("8 FFMA", o, io); run<1>("4 FFMA2", o, io); run<2>("8 FFMA + 16 INT", o, io); run<3>("4 FFMA2 + 16 INT", o, io);
    return 0;
}

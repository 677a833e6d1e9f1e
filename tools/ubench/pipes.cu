// Micro-benchmark: issue/pipe cost of scalar vs packed fp32 on sm_100a, and of LDG widths
// when every lane hits the same few L1 lines.  Build: nvcc -arch=sm_100a -O3 -o pipes pipes.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, float a, float b, const float4 *src)
{
    float x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    float2 y0 = make_float2(x0, x1), y1 = make_float2(x2, x3), y2 = make_float2(x4, x5), y3 = make_float2(x6, x7);
    float2 y4 = y0, y5 = y1, y6 = y2, y7 = y3;
    const float2 a2 = make_float2(a, a), b2 = make_float2(b, b);
    int i0 = threadIdx.x, i1 = i0 + 1, i2 = i0 + 2, i3 = i0 + 3;
    float4 acc = make_float4(0, 0, 0, 0);
    const float4 *p = src + (threadIdx.x & 7);
    for (int i = 0; i < ITERS; ++i) {
        if (MODE == 0) {          // 8 scalar FFMA
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        } else if (MODE == 1) {   // 8 packed FFMA2
            y0 = __ffma2_rn(y0, a2, b2); y1 = __ffma2_rn(y1, a2, b2); y2 = __ffma2_rn(y2, a2, b2); y3 = __ffma2_rn(y3, a2, b2);
            y4 = __ffma2_rn(y4, a2, b2); y5 = __ffma2_rn(y5, a2, b2); y6 = __ffma2_rn(y6, a2, b2); y7 = __ffma2_rn(y7, a2, b2);
        } else if (MODE == 2) {   // 4 packed + 4 scalar
            y0 = __ffma2_rn(y0, a2, b2); y1 = __ffma2_rn(y1, a2, b2); y2 = __ffma2_rn(y2, a2, b2); y3 = __ffma2_rn(y3, a2, b2);
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        } else if (MODE == 3) {   // 8 scalar FADD
            x0 += a; x1 += a; x2 += a; x3 += a; x4 += a; x5 += a; x6 += a; x7 += a;
        } else if (MODE == 4) {   // 4 IMAD + 4 FFMA
            i0 = i0 * i1 + i2; i1 = i1 * i2 + i3; i2 = i2 * i3 + i0; i3 = i3 * i0 + i1;
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
        } else if (MODE == 5) {   // 8 IMAD
            i0 = i0 * i1 + i2; i1 = i1 * i2 + i3; i2 = i2 * i3 + i0; i3 = i3 * i0 + i1;
            i0 = i0 * i1 + i3; i1 = i1 * i2 + i0; i2 = i2 * i3 + i1; i3 = i3 * i0 + i2;
        } else if (MODE == 6) {   // 2 LDG.128 (L1-resident, 8 distinct 16B per warp) + 8 FFMA2
            float4 v = __ldg(p + ((i & 3) << 3)), w = __ldg(p + ((i & 3) << 3) + 1);
            acc.x += v.x + w.x; acc.y += v.y + w.y;
            y0 = __ffma2_rn(y0, a2, b2); y1 = __ffma2_rn(y1, a2, b2); y2 = __ffma2_rn(y2, a2, b2); y3 = __ffma2_rn(y3, a2, b2);
        } else if (MODE == 7) {   // 2 LDG.128 only
            float4 v = __ldg(p + ((i & 3) << 3)), w = __ldg(p + ((i & 3) << 3) + 1);
            acc.x += v.x + w.x; acc.y += v.y + w.y; acc.z += v.z + w.z; acc.w += v.w + w.w;
        } else if (MODE == 8) {   // 8 LDG.32 only
            const float *q = reinterpret_cast<const float *>(p) + ((i & 3) << 5);
            acc.x += __ldg(q) + __ldg(q + 1) + __ldg(q + 2) + __ldg(q + 3) + __ldg(q + 4) + __ldg(q + 5) + __ldg(q + 6) + __ldg(q + 7);
        } else if (MODE == 9) {   // 4 FADD2 + 4 FFMA2
            y0 = __fadd2_rn(y0, a2); y1 = __fadd2_rn(y1, a2); y2 = __fadd2_rn(y2, a2); y3 = __fadd2_rn(y3, a2);
            y4 = __ffma2_rn(y4, a2, b2); y5 = __ffma2_rn(y5, a2, b2); y6 = __ffma2_rn(y6, a2, b2); y7 = __ffma2_rn(y7, a2, b2);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7 + y0.x + y0.y + y1.x + y1.y + y2.x + y2.y +
        y3.x + y3.y + y4.x + y4.y + y5.x + y5.y + y6.x + y6.y + y7.x + y7.y + i0 + i1 + i2 + i3 + acc.x + acc.y + acc.z + acc.w;
}

template <int MODE>
void run(const char *name, float ops_per_iter, float *out, const float4 *src)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, src);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, 1.0001f, 0.5f, src);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    // warp-instructions per SM per clock at the nominal max clock
    const double winst = (double)blocks * 8 * ITERS * ops_per_iter;
    const double clks = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-inst/clk/SM (of the %g counted per iteration)\n", name, ms, winst / clks / 148, ops_per_iter);
}

int main()
{
    float *out; float4 *src;
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaMalloc(&src, 4096);
    cudaMemset(src, 0, 4096);
    run<0>("8 FFMA", 8, out, src);
    run<1>("8 FFMA2", 8, out, src);
    run<2>("4 FFMA2 + 4 FFMA", 8, out, src);
    run<3>("8 FADD", 8, out, src);
    run<4>("4 IMAD + 4 FFMA", 8, out, src);
    run<5>("8 IMAD", 8, out, src);
    run<9>("4 FADD2 + 4 FFMA2", 8, out, src);
    run<6>("2 LDG.128 + 4 FFMA2", 6, out, src);
    run<7>("2 LDG.128", 2, out, src);
    run<8>("8 LDG.32", 8, out, src);
    return 0;
}

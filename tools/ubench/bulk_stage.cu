// Micro-benchmark (north_star names "TMA-staged volume bricks"): the fan-plane projector's corner-column traffic
// -- per step four z-runs of 512 contiguous bytes of the resident volume, a new in-plane cell every other step --
// fetched (a) the way the kernel does it, one predicated LDG.128 per lane and corner issued one step ahead, or
// (b) staged into shared memory with cp.async.bulk (the non-tensor TMA path; an mbarrier per slot, three cells
// ahead) and read back with LDS.128.  Same warps per SM (shared memory padded to 14 KB per warp, like the
// kernel), same volume, same walk; prints ns per step and warp.
//   nvcc -arch=sm_100a -O3 -o bulk_stage bulk_stage.cu && ./bulk_stage
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define NW 4
#define STEPS 1024
#define SLOTS 4                       // staging slots of 4 corners x 512 B
#define PAD_BYTES (14 * 1024)

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }

// in-plane cell of step t for this warp: a straight chord at a warp-dependent angle, half a voxel per step
__device__ __forceinline__ unsigned cell_off(int t, float x0, float y0, float dx, float dy, int nxp, int nzp)
{
    const int cx = (int)(x0 + t * dx), cy = (int)(y0 + t * dy);
    return (unsigned)((cy * nxp + cx) * nzp) * 4u;
}

template <int MODE>
__global__ void __launch_bounds__(NW * 32) k(const float *__restrict__ vol, float *out, int nxp, int nyp, int nzp)
{
    extern __shared__ __align__(128) unsigned char smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char *mine = smem + (size_t)warp * PAD_BYTES;
    float4 *stage = reinterpret_cast<float4 *>(mine);                       // [SLOTS][4][32] float4
    uint64_t *bar = reinterpret_cast<uint64_t *>(mine + SLOTS * 4 * 512);
    const int gw = blockIdx.x * NW + warp;
    const float ang = 0.3f + 0.001f * (gw % 1024) + 0.05f * (gw / 1024);
    const float dx = 0.5f * cosf(ang), dy = 0.5f * sinf(ang);
    const float x0 = 0.5f * nxp - 0.5f * STEPS * dx * 0.9f, y0 = 0.5f * nyp - 0.5f * STEPS * dy * 0.9f;
    const size_t dx4 = (size_t)nzp * 4, dy4 = (size_t)nxp * nzp * 4;
    const char *base = reinterpret_cast<const char *>(vol) + lane * 16;
    float4 acc = make_float4(0, 0, 0, 0);
    if (MODE == 0) {
        float4 c00, c01, c10, c11;
        unsigned off = cell_off(0, x0, y0, dx, dy, nxp, nzp);
        c00 = __ldg(reinterpret_cast<const float4 *>(base + off));
        c01 = __ldg(reinterpret_cast<const float4 *>(base + off + dx4));
        c10 = __ldg(reinterpret_cast<const float4 *>(base + off + dy4));
        c11 = __ldg(reinterpret_cast<const float4 *>(base + off + dy4 + dx4));
        for (int t = 0; t < STEPS; ++t) {
            const unsigned offn = cell_off(t + 1, x0, y0, dx, dy, nxp, nzp);
            const float w = 0.25f + 1e-4f * t;
            acc.x = fmaf(w, c00.x + c01.x + c10.x + c11.x, acc.x); acc.y = fmaf(w, c00.y + c01.y + c10.y + c11.y, acc.y);
            acc.z = fmaf(w, c00.z + c01.z + c10.z + c11.z, acc.z); acc.w = fmaf(w, c00.w + c01.w + c10.w + c11.w, acc.w);
            if (offn != off) {                                        // (uniform over the warp in this benchmark)
                c00 = __ldg(reinterpret_cast<const float4 *>(base + offn));
                c01 = __ldg(reinterpret_cast<const float4 *>(base + offn + dx4));
                c10 = __ldg(reinterpret_cast<const float4 *>(base + offn + dy4));
                c11 = __ldg(reinterpret_cast<const float4 *>(base + offn + dy4 + dx4));
            }
            off = offn;
        }
    } else {
        if (lane == 0)
            for (int s = 0; s < SLOTS; ++s)
                asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(smem_u32(bar + s)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        __syncwarp();
        // the walk as a list of distinct cells; fill the slots SLOTS-1 cells ahead
        auto issue = [&](int slot, unsigned off) {
            if (lane == 0)
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar + slot)), "r"(2048));
            __syncwarp();
            if (lane < 4) {
                const char *src = reinterpret_cast<const char *>(vol) + off + (lane & 1) * dx4 + (lane >> 1) * dy4;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             :: "r"(smem_u32(stage + (slot * 4 + lane) * 32)), "l"(src), "r"(512), "r"(smem_u32(bar + slot))
                             : "memory");
            }
        };
        int t_issue = 0, n_issued = 0, n_used = 0;
        unsigned last_issued = 0xffffffffu;
        auto pump = [&]() {                                          // keep SLOTS-1 cells in flight
            while (n_issued - n_used < SLOTS - 1 && t_issue <= STEPS) {
                const unsigned o = cell_off(t_issue, x0, y0, dx, dy, nxp, nzp);
                if (o != last_issued) { issue(n_issued % SLOTS, o); last_issued = o; ++n_issued; }
                ++t_issue;
            }
        };
        pump();
        unsigned off = 0xffffffffu;
        int cur = -1;
        for (int t = 0; t < STEPS; ++t) {
            const unsigned o = cell_off(t, x0, y0, dx, dy, nxp, nzp);
            if (o != off) {                                           // next cell: wait for its slot
                if (cur >= 0) { ++n_used; __syncwarp(); pump(); }
                ++cur;
                const unsigned parity = (cur / SLOTS) & 1;
                unsigned done = 0;
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
                                 "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(bar + cur % SLOTS)), "r"(parity) : "memory");
                off = o;
            }
            const float4 *s = stage + (cur % SLOTS) * 4 * 32 + lane;
            const float4 c00 = s[0], c01 = s[32], c10 = s[64], c11 = s[96];
            const float w = 0.25f + 1e-4f * t;
            acc.x = fmaf(w, c00.x + c01.x + c10.x + c11.x, acc.x); acc.y = fmaf(w, c00.y + c01.y + c10.y + c11.y, acc.y);
            acc.z = fmaf(w, c00.z + c01.z + c10.z + c11.z, acc.z); acc.w = fmaf(w, c00.w + c01.w + c10.w + c11.w, acc.w);
        }
    }
    out[(size_t)gw * 32 + lane] = acc.x + acc.y + acc.z + acc.w;
}

int main()
{
    const int nxp = 514, nyp = 514, nzp = 136;
    const size_t n = (size_t)nxp * nyp * nzp;
    float *vol, *out;
    cudaMalloc(&vol, n * 4);
    cudaMemset(vol, 0, n * 4);
    const int ctas = 148 * 4 * 8;                       // 8 waves of 4 CTAs per SM
    cudaMalloc(&out, (size_t)ctas * NW * 32 * 4);
    const size_t smem = (size_t)NW * PAD_BYTES;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        float best = 1e30f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<ctas, NW * 32, smem>>>(vol, out, nxp, nyp, nzp);
            else k<1><<<ctas, NW * 32, smem>>>(vol, out, nxp, nyp, nzp);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t err = cudaGetLastError();
        const double warp_steps = (double)ctas * NW * STEPS;
        printf("%s: %.3f ms, %.2f ns per step and resident warp slot (16 warps/SM), %s\n",
               mode == 0 ? "LDG.128 per lane, predicated, one step ahead" : "cp.async.bulk staging, 3 cells ahead, LDS.128",
               best, best * 1e6 / (warp_steps / (148.0 * 16.0)), cudaGetErrorString(err));
    }
    return 0;
}

// Ray-driven trilinear cone-beam forward projector (tigre.Ax 'interpolated';
// reference call site pCT_CBCT_Reconstruction.py:86, algorithm SURVEY.md A.3) with the
// scatter-combine and OS-SART residual epilogues fused in.  sm_100a, no tensor cores:
// the work is a gather, bound by L1 wavefronts and issue slots, not by a contraction.
//
// Mapping.  A warp owns 32 consecutive detector rows v of ONE detector column u, so its
// 32 rays share the in-plane trajectory and differ only in z slope.  The resident
// volume is stored z-fastest with a zero rim ((Y+2, X+2, Zp)), so at every step the
// warp's eight corner loads each fall into one or two 128-byte lines, and no load
// needs a bounds check.  Sample positions follow TIGRE's lattice exactly:
// q_i = S + i*(P-S)/ceil(|P-S|/accuracy); only the samples that can be non-zero
// (inside the open box (-1,N)^3) are visited.  Per-ray set-up runs in double so that
// the lattice (the ceil) matches a float64 evaluation; the march is fp32 with exact
// (software) trilinear weights -- hardware texture filtering has 8-bit weights.
#include "pax_internal.h"

struct AxArgs {
    const AngleDev *angles;
    const float *vol0;
    const float *vol1;
    float *out;
    const float *b;
    double *sumsq;
    unsigned long long *nsamples;
    int first, nv, nu;
    int nxp, nzp;                 // resident strides: x -> nzp, y -> nxp*nzp
    float hix, hiy, hiz;          // exclusive upper bounds of valid padded coordinates
    float loz;                    // inclusive lower z bound (PAX_ZOFF-1); x,y lower bound is 0
    double acc;
    float dX, dY, dZ;
    float a0, a1, c;
    float whx, why, whz, wthr;
};

#define AX_MAGIC 8388608.0f          // 2^23: (q + 2^23) rounded down keeps floor(q) in the mantissa
#define AX_MAGIC_I 0x4B000000

template <int NCH>
__device__ __forceinline__ void ax_sample(const float *__restrict__ v0, const float *__restrict__ v1,
                                          float qx, float qy, float qz, int sx, int sy, float &s0,
                                          float &s1)
{
    const float tx = __fadd_rd(qx, AX_MAGIC), ty = __fadd_rd(qy, AX_MAGIC), tz = __fadd_rd(qz, AX_MAGIC);
    const float fx = qx - (tx - AX_MAGIC), fy = qy - (ty - AX_MAGIC), fz = qz - (tz - AX_MAGIC);
    const int ix = __float_as_int(tx) - AX_MAGIC_I, iy = __float_as_int(ty) - AX_MAGIC_I,
              iz = __float_as_int(tz) - AX_MAGIC_I;
    const int lin = iy * sy + ix * sx + iz;
    {
        const float *p = v0 + lin;
        const float c000 = __ldg(p), c001 = __ldg(p + 1);
        const float c010 = __ldg(p + sx), c011 = __ldg(p + sx + 1);
        const float c100 = __ldg(p + sy), c101 = __ldg(p + sy + 1);
        const float c110 = __ldg(p + sy + sx), c111 = __ldg(p + sy + sx + 1);
        const float a00 = fmaf(fz, c001 - c000, c000), a01 = fmaf(fz, c011 - c010, c010);
        const float a10 = fmaf(fz, c101 - c100, c100), a11 = fmaf(fz, c111 - c110, c110);
        const float b0 = fmaf(fx, a01 - a00, a00), b1 = fmaf(fx, a11 - a10, a10);
        s0 += fmaf(fy, b1 - b0, b0);
    }
    if (NCH == 2) {
        const float *p = v1 + lin;
        const float c000 = __ldg(p), c001 = __ldg(p + 1);
        const float c010 = __ldg(p + sx), c011 = __ldg(p + sx + 1);
        const float c100 = __ldg(p + sy), c101 = __ldg(p + sy + 1);
        const float c110 = __ldg(p + sy + sx), c111 = __ldg(p + sy + sx + 1);
        const float a00 = fmaf(fz, c001 - c000, c000), a01 = fmaf(fz, c011 - c010, c010);
        const float a10 = fmaf(fz, c101 - c100, c100), a11 = fmaf(fz, c111 - c110, c110);
        const float b0 = fmaf(fx, a01 - a00, a00), b1 = fmaf(fx, a11 - a10, a10);
        s1 += fmaf(fy, b1 - b0, b0);
    }
}

// LV = detector rows per warp (32/LV columns).  8 warps per CTA, stacked along u.
template <int NCH, int MODE, int LV>
__global__ void __launch_bounds__(256) k_ax(const AxArgs a)
{
    constexpr int LU = 32 / LV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = blockIdx.x * LV + (lane % LV);
    const int u = (blockIdx.y * 8 + warp) * LU + lane / LV;
    const int ia = blockIdx.z;
    const bool live = (u < a.nu) && (v < a.nv);
    const AngleDev &A = a.angles[a.first + ia];

    float s0 = 0.f, s1 = 0.f, steplen = 0.f;
    int ns = 0;
    if (live) {
        // ---- per-ray set-up in double: TIGRE's lattice n = ceil(|P-S|/accuracy)
        const double px = A.ox + u * A.ux, py = A.oy + u * A.uy, pz = A.oz + v * A.vz;
        double dx = px - A.sx, dy = py - A.sy, dz = pz - A.sz;
        const double n = ceil(sqrt(dx * dx + dy * dy + dz * dz) / a.acc);
        const double rn = 1.0 / n;
        dx *= rn; dy *= rn; dz *= rn;
        // ---- samples i in [t0,t1] with q_i inside [0,hix) x [0,hiy) x [loz,hiz)
        double t0 = 0.0, t1 = n;
        bool hit = true;
        {
            const double lo[3] = {0.0, 0.0, (double)a.loz};
            const double hi[3] = {(double)a.hix, (double)a.hiy, (double)a.hiz};
            const double S[3] = {A.sx, A.sy, A.sz}, D[3] = {dx, dy, dz};
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                if (D[k] == 0.0) {
                    hit &= (S[k] >= lo[k] && S[k] < hi[k]);
                } else {
                    const double r = 1.0 / D[k];
                    const double ta = (lo[k] - S[k]) * r, tb = (hi[k] - S[k]) * r;
                    t0 = fmax(t0, fmin(ta, tb));
                    t1 = fmin(t1, fmax(ta, tb));
                }
            }
        }
        if (hit && t0 <= t1) {
            const double i0 = ceil(t0), i1 = floor(t1);
            ns = (int)(i1 - i0) + 1;
            const float ex = (float)(A.sx + i0 * dx), ey = (float)(A.sy + i0 * dy),
                        ez = (float)(A.sz + i0 * dz);
            const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
            // the fp32 positions of the first/last sample must themselves be in range
            // (the march does no bounds checks): trim the ends if rounding pushed them out.
            int j0 = 0;
            auto inside = [&](int j) {
                const float qx = fmaf((float)j, fdx, ex), qy = fmaf((float)j, fdy, ey),
                            qz = fmaf((float)j, fdz, ez);
                return qx >= 0.f && qx < a.hix && qy >= 0.f && qy < a.hiy && qz >= a.loz && qz < a.hiz;
            };
            while (j0 < ns && !inside(j0)) ++j0;
            while (ns > j0 && !inside(ns - 1)) --ns;
            const float *__restrict__ v0 = a.vol0;
            const float *__restrict__ v1 = a.vol1;
            const int sx = a.nzp, sy = a.nxp * a.nzp;
#pragma unroll 2
            for (int j = j0; j < ns; ++j) {
                const float fj = (float)j;
                ax_sample<NCH>(v0, v1, fmaf(fj, fdx, ex), fmaf(fj, fdy, ey), fmaf(fj, fdz, ez), sx, sy,
                               s0, s1);
            }
            ns -= j0;
            const float mx = fdx * a.dX, my = fdy * a.dY, mz = fdz * a.dZ;
            steplen = sqrtf(mx * mx + my * my + mz * mz);
        }
    }
    float r2 = 0.f;
    if (live) {
        const size_t o = ((size_t)ia * a.nv + v) * a.nu + u;
        const float p0 = s0 * steplen;
        if (MODE == PAX_AX_PLAIN) {
            float val = a.a0 * p0;
            if (NCH == 2) val = fmaf(a.a1, s1 * steplen, val);
            if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, u, v, a.whx, a.why, a.whz), val);
            a.out[o] = val;
        } else {
            const float L = pax_ray_chord(A, u, v, a.whx, a.why, a.whz);
            const float w = L > a.wthr ? 1.f / L : 0.f;
            const float r = __ldg(a.b + o) - p0;
            r2 = r * r;
            a.out[o] = w * r;
        }
    }
    if (a.nsamples) {
        unsigned t = live ? (unsigned)max(ns, 0) : 0u;
        t = __reduce_add_sync(0xffffffffu, t);
        if (lane == 0 && t) atomicAdd(a.nsamples, (unsigned long long)t);
    }
    if (MODE == PAX_AX_RESIDUAL && a.sumsq) {
        for (int o = 16; o; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if (lane == 0) atomicAdd(a.sumsq, (double)r2);
    }
}

template <int NCH, int MODE>
static void launch_ax(const AxArgs &a, int count, cudaStream_t st)
{
    // rows-per-warp: tall detectors get the full 32-row column warp; short ones fall back
    if (a.nv >= 32) {
        dim3 grid(pax_cdiv(a.nv, 32), pax_cdiv(a.nu, 8), count);
        k_ax<NCH, MODE, 32><<<grid, 256, 0, st>>>(a);
    } else {
        dim3 grid(pax_cdiv(a.nv, 8), pax_cdiv(a.nu, 32), count);
        k_ax<NCH, MODE, 8><<<grid, 256, 0, st>>>(a);
    }
}

extern "C" int pax_ax(const PaxPlan *p, const float *vol_res0, const float *vol_res1, int first,
                      int count, float *proj_out, const PaxAxEpilogue *epi, void *stream)
{
    PAX_REQUIRE(p && vol_res0 && proj_out && epi, "pax_ax: NULL argument");
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles,
                "pax_ax: angle range [%d,%d) outside plan (%d angles)", first, first + count, p->n_angles);
    PAX_REQUIRE(count <= 65535, "pax_ax: at most 65535 angles per call");
    PAX_REQUIRE(epi->mode == PAX_AX_PLAIN || epi->mode == PAX_AX_RESIDUAL, "pax_ax: bad epilogue mode %d",
                epi->mode);
    PAX_REQUIRE(epi->mode != PAX_AX_RESIDUAL || epi->b, "pax_ax: residual mode needs b");
    PAX_REQUIRE(epi->mode != PAX_AX_RESIDUAL || !vol_res1, "pax_ax: residual mode is single-channel");
    AxArgs a;
    a.angles = p->d_angles;
    a.vol0 = vol_res0; a.vol1 = vol_res1; a.out = proj_out; a.b = epi->b;
    a.sumsq = epi->sumsq; a.nsamples = epi->nsamples;
    a.first = first; a.nv = p->geo.nDetecV; a.nu = p->geo.nDetecU;
    a.nxp = p->nxp; a.nzp = p->nzp;
    a.hix = (float)(p->geo.nVoxelX + 1); a.hiy = (float)(p->geo.nVoxelY + 1);
    a.hiz = (float)(p->geo.nVoxelZ + PAX_ZOFF); a.loz = (float)(PAX_ZOFF - 1);
    a.acc = p->geo.accuracy;
    a.dX = (float)p->geo.dVoxelX; a.dY = (float)p->geo.dVoxelY; a.dZ = (float)p->geo.dVoxelZ;
    a.a0 = epi->a0; a.a1 = epi->a1; a.c = epi->c;
    a.whx = epi->w_half_x; a.why = epi->w_half_y; a.whz = epi->w_half_z; a.wthr = epi->w_thr;
    cudaStream_t st = (cudaStream_t)stream;
    if (epi->mode == PAX_AX_RESIDUAL) launch_ax<1, PAX_AX_RESIDUAL>(a, count, st);
    else if (vol_res1) launch_ax<2, PAX_AX_PLAIN>(a, count, st);
    else launch_ax<1, PAX_AX_PLAIN>(a, count, st);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

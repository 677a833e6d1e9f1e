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
//
// Two gather back-ends share the set-up and the epilogues:
//   QUAD (default)  the caller's workspace holds a re-packed copy of the volume whose entry
//                   (y,x,z) is the float4 {v[y,x], v[y,x+1], v[y+1,x], v[y+1,x+1]} at depth z,
//                   z fastest.  A sample is two 128-bit loads at consecutive addresses and the
//                   seven lerps run as packed FFMA2/FADD2 (sm_100 f32x2): ~26 issue slots per
//                   sample instead of ~60.
//   PLAIN           eight 32-bit loads straight from the resident layout (first version; kept
//                   for A/B measurement, selected when no workspace is given).
// CTAs are ordered u-tile fastest, then angle, then v-tile: all angles of a launch that
// look at the same z slab of the volume run together, so the slab stays in the 126 MB L2.
#include <cstdlib>

#include "pax_internal.h"


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

// ---- packed fp32x2 helpers (sm_100: FADD2 / FFMA2 issue one slot for two lanes of work)
__device__ __forceinline__ float2 f2_add_rm(float2 a, float2 b)
{
    unsigned long long r;
    asm("add.rm.f32x2 %0, %1, %2;" : "=l"(r)
        : "l"(*reinterpret_cast<unsigned long long *>(&a)), "l"(*reinterpret_cast<unsigned long long *>(&b)));
    return *reinterpret_cast<float2 *>(&r);
}
__device__ __forceinline__ float2 f2_sub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 f2_lerp(float f, float2 a, float2 b)
{
    return __ffma2_rn(make_float2(f, f), f2_sub(b, a), a);
}

// trilinear value of one quad-packed channel: lo/hi = the cell's four in-plane corners at z0/z1
__device__ __forceinline__ float quad_trilinear(const float4 lo, const float4 hi, float fx, float fy,
                                                float fz)
{
    const float2 a0 = f2_lerp(fz, make_float2(lo.x, lo.y), make_float2(hi.x, hi.y));   // y0: (x0, x1)
    const float2 a1 = f2_lerp(fz, make_float2(lo.z, lo.w), make_float2(hi.z, hi.w));   // y1: (x0, x1)
    const float2 b = f2_lerp(fy, a0, a1);                                              // (x0, x1)
    return fmaf(fx, b.y - b.x, b.x);
}

// Address arithmetic without integer multiplies (IMAD is half rate on the FMA pipes, which bound
// this kernel): the quad columns are padded to a power-of-two byte stride 2^qshift, the column
// number cy*nxq + cx is formed by ONE FFMA on the (exactly integral) floats, landing biased by 2^23
// in a mantissa, and the biases cancel in 32-bit modular arithmetic.  Shifts and adds go to the
// otherwise idle ALU pipe.
template <int NCH>
__device__ __forceinline__ void ax_sample_quad(const char *__restrict__ b0, const char *__restrict__ b1,
                                               float2 qxy, float qz, float nxq_f, int qshift,
                                               unsigned kbias, float &s0, float &s1)
{
    const float2 mg = make_float2(AX_MAGIC, AX_MAGIC);
    const float2 txy = f2_add_rm(qxy, mg);
    const float tz = __fadd_rd(qz, AX_MAGIC);
    const float2 cxy = f2_sub(txy, mg);                  // (cx, cy): exact small integers
    const float2 fxy = f2_sub(qxy, cxy);
    const float fz = qz - (tz - AX_MAGIC);
    const float col = fmaf(cxy.y, nxq_f, txy.x);         // cy*nxq + cx + 2^23, exact (< 2^24)
    // 32-bit modular arithmetic: the biases cancel mod 2^32, the true byte offset is < 4 GB
    const unsigned off = (__float_as_uint(col) << qshift) + (__float_as_uint(tz) << 4) - kbias;
    {
        const float4 *p = reinterpret_cast<const float4 *>(b0 + off);
        const float4 lo = __ldg(p), hi = __ldg(p + 1);
        s0 += quad_trilinear(lo, hi, fxy.x, fxy.y, fz);
    }
    if (NCH == 2) {
        const float4 *p = reinterpret_cast<const float4 *>(b1 + off);
        const float4 lo = __ldg(p), hi = __ldg(p + 1);
        s1 += quad_trilinear(lo, hi, fxy.x, fxy.y, fz);
    }
}

// LV = detector rows per warp (32/LV columns).  8 warps per CTA, stacked along u.
template <int NCH, int MODE, int LV, bool QUAD, int NW, int MINB, int UNR>
__global__ void __launch_bounds__(NW * 32, MINB) k_ax(const AxArgs a)
{
    constexpr int LU = 32 / LV;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int v = blockIdx.z * LV + (lane % LV);
    const int u = (blockIdx.x * NW + warp) * LU + lane / LV;
    const int ia = blockIdx.y;
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
            const float fdx = (float)dx, fdy = (float)dy, fdz = (float)dz;
            // Four interleaved sub-lattices: sample j sits at E[j&3] + (j & ~3) * d with the origins
            // E[k] = S + (i0+k) d rounded from double.  One counter update per four samples, and no
            // rounding accumulates along the ray.
            float2 exy[4];
            float ezk[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double jk = i0 + (double)k;
                exy[k] = make_float2((float)(A.sx + jk * dx), (float)(A.sy + jk * dy));
                ezk[k] = (float)(A.sz + jk * dz);
            }
            const float ex = exy[0].x, ey = exy[0].y, ez = ezk[0];
            // the fp32 positions of the first/last sample must themselves be in range
            // (the march does no bounds checks): trim the ends if rounding pushed them out.
            int j0 = 0;
            auto inside = [&](int j) {
                const float fb = (float)(j & ~3);
                const int k = j & 3;
                const float bx = k == 0 ? exy[0].x : k == 1 ? exy[1].x : k == 2 ? exy[2].x : exy[3].x;
                const float by = k == 0 ? exy[0].y : k == 1 ? exy[1].y : k == 2 ? exy[2].y : exy[3].y;
                const float bz = k == 0 ? ezk[0] : k == 1 ? ezk[1] : k == 2 ? ezk[2] : ezk[3];
                const float qx = fmaf(fb, fdx, bx), qy = fmaf(fb, fdy, by), qz = fmaf(fb, fdz, bz);
                return qx >= 0.f && qx < a.hix && qy >= 0.f && qy < a.hiy && qz >= a.loz && qz < a.hiz;
            };
            while (j0 < ns && !inside(j0)) ++j0;
            while (ns > j0 && !inside(ns - 1)) --ns;
            if (QUAD) {
                const char *__restrict__ q0 = a.quad0;
                const char *__restrict__ q1 = a.quad1;
                const float sx = a.nxq_f;
                const int sy = a.qshift;
                const unsigned kb = ((unsigned)AX_MAGIC_I << sy) + ((unsigned)AX_MAGIC_I << 4);
                const float2 dxy = make_float2(fdx, fdy);
                int j = j0;
                // head: up to the next multiple of four
                for (; (j & 3) && j < ns; ++j) {
                    const float fb = (float)(j & ~3);
                    const int k = j & 3;
                    const float2 e = k == 1 ? exy[1] : k == 2 ? exy[2] : exy[3];
                    const float z = k == 1 ? ezk[1] : k == 2 ? ezk[2] : ezk[3];
                    ax_sample_quad<NCH>(q0, q1, __ffma2_rn(make_float2(fb, fb), dxy, e), fmaf(fb, fdz, z), sx, sy,
                                        kb, s0, s1);
                }
                float fm = (float)j;
#pragma unroll 1
                for (; j + 4 <= ns; j += 4) {
                    const float2 fm2 = make_float2(fm, fm);
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        ax_sample_quad<NCH>(q0, q1, __ffma2_rn(fm2, dxy, exy[k]), fmaf(fm, fdz, ezk[k]), sx,
                                            sy, kb, s0, s1);
                    fm += 4.f;
                }
                for (int k = 0; j < ns; ++j, ++k) {
                    const float2 e = k == 0 ? exy[0] : k == 1 ? exy[1] : exy[2];
                    const float z = k == 0 ? ezk[0] : k == 1 ? ezk[1] : ezk[2];
                    ax_sample_quad<NCH>(q0, q1, __ffma2_rn(make_float2(fm, fm), dxy, e), fmaf(fm, fdz, z), sx, sy,
                                        kb, s0, s1);
                }
            } else {
                const float *__restrict__ v0 = a.vol0;
                const float *__restrict__ v1 = a.vol1;
                const int sx = a.nzp, sy = a.nxp * a.nzp;
                (void)ex; (void)ey; (void)ez;
#pragma unroll 2
                for (int j = j0; j < ns; ++j) {
                    const float fb = (float)(j & ~3);
                    const int k = j & 3;
                    const float2 e = k == 0 ? exy[0] : k == 1 ? exy[1] : k == 2 ? exy[2] : exy[3];
                    const float z = k == 0 ? ezk[0] : k == 1 ? ezk[1] : k == 2 ? ezk[2] : ezk[3];
                    ax_sample<NCH>(v0, v1, fmaf(fb, fdx, e.x), fmaf(fb, fdy, e.y), fmaf(fb, fdz, z), sx, sy,
                                   s0, s1);
                }
            }
            ns -= j0;
            const float mx = fdx * a.dX, my = fdy * a.dY, mz = fdz * a.dZ;
            steplen = sqrtf(mx * mx + my * my + mz * mz);
        }
    }
    float r2 = 0.f, mymax = -INFINITY;
    if (live) {
        const size_t o = ((size_t)ia * a.nv + v) * a.nu + u;
        const float p0 = s0 * steplen;
        if (MODE == PAX_AX_PLAIN) {
            float val = a.a0 * p0;
            if (NCH == 2) {
                if (a.out1) a.out1[o] = s1 * steplen;
                else val = fmaf(a.a1, s1 * steplen, val);
            }
            if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, u, v, a.whx, a.why, a.whz), val);
            a.out[o] = val;
            mymax = val;
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
    if (MODE == PAX_AX_PLAIN && a.maxout) {
        for (int o = 16; o; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
        if (lane == 0) {
            const unsigned b = __float_as_uint(mymax);
            atomicMax(a.maxout, (b & 0x80000000u) ? ~b : (b | 0x80000000u));
        }
    }
    if (MODE == PAX_AX_RESIDUAL && a.sumsq) {
        for (int o = 16; o; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if (lane == 0) atomicAdd(a.sumsq, (double)r2);
    }
}

template <int NCH, int MODE, bool QUAD, int NW = 8, int MINB = 4, int UNR = 4>
static void launch_ax(const AxArgs &a, int count, cudaStream_t st)
{
    // rows-per-warp: tall detectors get the full 32-row column warp; short ones fall back
    if (a.nv >= 32) {
        dim3 grid(pax_cdiv(a.nu, NW), count, pax_cdiv(a.nv, 32));
        k_ax<NCH, MODE, 32, QUAD, NW, MINB, UNR><<<grid, NW * 32, 0, st>>>(a);
    } else {
        dim3 grid(pax_cdiv(a.nu, NW * 4), count, pax_cdiv(a.nv, 8));
        k_ax<NCH, MODE, 8, QUAD, NW, MINB, UNR><<<grid, NW * 32, 0, st>>>(a);
    }
}

// resident layout -> quad-packed copy (one float4 per cell corner (y,x) and depth z; columns
// padded to zq = 2^k entries so that the march needs no integer multiply)
__global__ void __launch_bounds__(256) k_quad_pack(const float *__restrict__ P, float4 *__restrict__ Q,
                                                   int nyq, int nxq, int nxp, int nzp, int zq)
{
    const size_t n = (size_t)nyq * nxq * nzp;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += (size_t)gridDim.x * blockDim.x) {
        const int z = (int)(t % nzp);
        const size_t c = t / nzp;
        const int x = (int)(c % nxq), y = (int)(c / nxq);
        const float *p = P + ((size_t)y * nxp + x) * nzp + z;
        const size_t sy = (size_t)nxp * nzp;
        Q[c * zq + z] = make_float4(__ldg(p), __ldg(p + nzp), __ldg(p + sy), __ldg(p + sy + nzp));
    }
}

static int quad_zq(const PaxPlan *p)
{
    int zq = 8;
    while (zq < p->nzp) zq <<= 1;
    return zq;
}

extern "C" size_t pax_ax_workspace_bytes(const PaxPlan *p, int n_channels)
{
    if (!p || n_channels < 1) return 0;
    if (pax_plan_projector(p) == PAX_PROJECTOR_FAN) return 0;   // gathers from the resident layout
    if ((size_t)(p->nyp - 1) * (p->nxp - 1) >= ((size_t)1 << 23)) return 0;   // column index must fit a mantissa
    if ((size_t)(p->nyp - 1) * (p->nxp - 1) * quad_zq(p) * sizeof(float4) >= ((size_t)1 << 32)) return 0;
    return (size_t)n_channels * (size_t)(p->nyp - 1) * (p->nxp - 1) * quad_zq(p) * sizeof(float4);
}

extern "C" int pax_ax(const PaxPlan *p, const float *vol_res0, const float *vol_res1, int first,
                      int count, float *proj_out, const PaxAxEpilogue *epi, void *workspace,
                      void *stream)
{
    PAX_REQUIRE(p && vol_res0 && proj_out && epi, "pax_ax: NULL argument");
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles,
                "pax_ax: angle range [%d,%d) outside plan (%d angles)", first, first + count, p->n_angles);
    PAX_REQUIRE(count <= 65535, "pax_ax: at most 65535 angles per call");
    PAX_REQUIRE(epi->mode == PAX_AX_PLAIN || epi->mode == PAX_AX_RESIDUAL, "pax_ax: bad epilogue mode %d",
                epi->mode);
    PAX_REQUIRE(epi->mode != PAX_AX_RESIDUAL || epi->b, "pax_ax: residual mode needs b");
    PAX_REQUIRE(epi->mode != PAX_AX_RESIDUAL || !vol_res1, "pax_ax: residual mode is single-channel");
    PAX_REQUIRE(!epi->out1 || (vol_res1 && epi->mode == PAX_AX_PLAIN), "pax_ax: out1 needs two channels, PLAIN");
    AxArgs a;
    a.angles = p->d_angles;
    a.vol0 = vol_res0; a.vol1 = vol_res1; a.out = proj_out; a.out1 = epi->out1; a.b = epi->b;
    a.sumsq = epi->sumsq; a.nsamples = epi->nsamples; a.maxout = epi->maxout;
    a.first = first; a.nv = p->geo.nDetecV; a.nu = p->geo.nDetecU;
    a.nxp = p->nxp; a.nzp = p->nzp;
    a.hix = (float)(p->geo.nVoxelX + 1); a.hiy = (float)(p->geo.nVoxelY + 1);
    a.hiz = (float)(p->geo.nVoxelZ + PAX_ZOFF); a.loz = (float)(PAX_ZOFF - 1);
    a.acc = p->geo.accuracy;
    a.dX = (float)p->geo.dVoxelX; a.dY = (float)p->geo.dVoxelY; a.dZ = (float)p->geo.dVoxelZ;
    a.a0 = epi->a0; a.a1 = epi->a1; a.c = epi->c;
    a.accumulate = 0;
    a.whx = epi->w_half_x; a.why = epi->w_half_y; a.whz = epi->w_half_z; a.wthr = epi->w_thr;
    a.fan_rows = 0; a.fan_epoch = 0; a.fan_count = 0; a.fan_order = nullptr; a.fan_ntiles = 0; a.fan_split = 1;
    a.fan_queue = nullptr; a.fan_nblock = 0; a.fan_bw = 0; a.fan_nband = 1; a.fan_sms = 0;
    a.fan_scratch = reinterpret_cast<float *>(workspace);      // (fan projector: optional scratch of a split launch)
    cudaStream_t st = (cudaStream_t)stream;
    PAX_REQUIRE(p->fan_share_parts == 1 || pax_plan_projector(p) == PAX_PROJECTOR_FAN,
                "pax_ax: a column share (pax_plan_set_column_share) needs the fan-plane projector");
    if (pax_plan_projector(p) == PAX_PROJECTOR_FAN) {
        // one channel per launch; a second channel is a first pass whose result the second adds to
        if (vol_res1) {
            AxArgs a1 = a;
            a1.vol0 = vol_res1; a1.vol1 = nullptr; a1.nsamples = nullptr; a1.maxout = nullptr;
            a1.c = 0.f; a1.accumulate = 0;
            if (epi->out1) { a1.out = epi->out1; a1.a0 = 1.f; }
            else { a1.a0 = epi->a1; }
            int rc = pax_ax_fan_launch(p, a1, PAX_AX_PLAIN, count, st);
            if (rc != PAX_OK) return rc;
            a.accumulate = epi->out1 ? 0 : 1;
        }
        a.vol1 = nullptr; a.out1 = nullptr;
        int rc = pax_ax_fan_launch(p, a, epi->mode, count, st);
        if (rc != PAX_OK) return rc;
        PAX_CHECK_CUDA(cudaGetLastError());
        return PAX_OK;
    }
    if (workspace) {
        PAX_REQUIRE(((uintptr_t)workspace & 15) == 0, "pax_ax: workspace must be 16-byte aligned");
        const int nyq = p->nyp - 1, nxq = p->nxp - 1, zq = quad_zq(p);
        PAX_REQUIRE((size_t)nyq * nxq < ((size_t)1 << 23),
                    "pax_ax: in-plane size too large for the workspace kernel; pass workspace=NULL");
        const size_t per = (size_t)nyq * nxq * zq;
        float4 *q = reinterpret_cast<float4 *>(workspace);
        const int blocks = 148 * 16;
        k_quad_pack<<<blocks, 256, 0, st>>>(vol_res0, q, nyq, nxq, p->nxp, p->nzp, zq);
        if (vol_res1) k_quad_pack<<<blocks, 256, 0, st>>>(vol_res1, q + per, nyq, nxq, p->nxp, p->nzp, zq);
        int qshift = 4;
        while ((1 << (qshift - 4)) < zq) ++qshift;
        PAX_REQUIRE(per * sizeof(float4) < ((size_t)1 << 32), "pax_ax: quad channel exceeds 4 GB; pass workspace=NULL");
        a.quad0 = reinterpret_cast<const char *>(q);
        a.quad1 = vol_res1 ? reinterpret_cast<const char *>(q + per) : nullptr;
        a.qshift = qshift; a.nxq_f = (float)nxq;
        if (epi->mode == PAX_AX_RESIDUAL) {
            launch_ax<1, PAX_AX_RESIDUAL, true>(a, count, st);
        } else if (vol_res1) launch_ax<2, PAX_AX_PLAIN, true>(a, count, st);
        else launch_ax<1, PAX_AX_PLAIN, true>(a, count, st);
    } else {
        a.quad0 = a.quad1 = nullptr; a.qshift = 0; a.nxq_f = 0.f;
        if (epi->mode == PAX_AX_RESIDUAL) launch_ax<1, PAX_AX_RESIDUAL, false>(a, count, st);
        else if (vol_res1) launch_ax<2, PAX_AX_PLAIN, false>(a, count, st);
        else launch_ax<1, PAX_AX_PLAIN, false>(a, count, st);
    }
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

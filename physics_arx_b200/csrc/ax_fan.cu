// Fan-plane forward projector: tigre.Ax 'interpolated' (reference call site
// pCT_CBCT_Reconstruction.py:86 and, inside ossart, :103; algorithm SURVEY.md A.3) evaluated on
// TIGRE's own sample lattice, but with the work shared between the rays of a detector column.
//
// Why it can be shared.  Pixel (v,u) of one projection is at P = O + u*U + v*V with V = z-hat, so
// all rays of a detector COLUMN u lie in one vertical plane through the source and share the
// in-plane chord S_xy -> P_xy.  TIGRE samples ray (u,v) at q_i = S + i*(P-S)/n, n = ceil(|P-S|/acc):
// the in-plane positions q_i,xy = S_xy + i*(P_xy-S_xy)/n depend on the row v ONLY through the
// integer n(v), which is constant over long runs of rows when the detector is finer than the
// slices (C2: 768 rows, 3..4 distinct n per column).  For a run of rows with equal n
//   sample(i, v) = (1-t) * G_i(k) + t * G_i(k+1),   k = floor(z_i(v)),  t = z_i(v) - k,
//   G_i(k)       = bilinear_xy(volume slice k) at q_i,xy          -- one value per (step, slice),
// i.e. the in-plane half of the trilinear interpolation is done once per (step, slice) instead of
// once per (step, row): ~8x fewer gathers at C2 (768 rows look at ~90 slices).
//
// Rays are nearly horizontal (|dz| per step <= 0.02 slices at C2), so a ray stays in one slice
// pair (k,k+1) for tens of steps, during which t is LINEAR in the step index.  With the running
// sums  P_k(j) = sum_{i<j} G_i(k)  and  Q_k(j) = sum_{i<j} (i-j0) G_i(k)  of a block of steps,
//   sum_{i=a..b} sample(i,v) = dP_k + alpha*(dP_{k+1} - dP_k) + delta*(dQ_{k+1} - dQ_k)
// (alpha, delta = intercept and slope of t over the stretch), so a row costs four table look-ups
// per stretch instead of one trilinear sample per step.  The sums are exactly TIGRE's Riemann sum
// re-associated; only fp32 rounding differs (measured ~1e-7 relative L2 against the float64 oracle).
//
// Kernel shape (sm_100a, no tensor cores: this is a gather plus prefix sums).  Every WARP works on
// its own detector column (one angle, one u, a tile of <= 256 rows) with its own tables in shared memory;
// there is no CTA-wide barrier, so the gather-heavy and the table-heavy phases of different warps
// overlap on an SM.  Per run of equal n the warp walks the chord in chunks of G*8 steps:
//   phase 0  once per chunk: a lane prepares one or two steps -- in-plane cell offset and the four
//            bilinear weights (chunk origin split into integer cell + fraction in double, steps added
//            in fp32, so nothing accumulates along the ray);
//   groups   the rows of the run are cut, for THIS chunk, into groups whose z extent over the chunk fits
//            the 4*LQ slices the lanes of a stream cover (closed form: z is linear in row and step);
//   phase 1  per group: the warp advances G streams of 8 consecutive steps at once; a lane owns one
//            stream and four consecutive slices of the group's z window (the resident volume is
//            z-fastest, so a corner of a step is one 16-byte load per lane), accumulates P and Q down
//            its stream and writes them to the table; a shuffle scan over the streams gives each its
//            carry.  Loads are predicated (same cell: no load) and issued two steps ahead into two
//            alternating register sets;
//   phase 2  per group: a lane owns two detector rows at a time and adds up their stretches from the
//            table (+ carry); the rows' running sums live in shared memory across chunks.
// pax_fan_configure checks from the geometry that at least one row always fits (else the kernel is not
// offered) and estimates how many do (AUTO's criterion).
#include <climits>
#include <cmath>
#include <cstdlib>

#include "pax_internal.h"

#define FAN_NW 8            // warps per CTA (independent; the CTA only groups adjacent columns for L1)
#define FAN_GROUP 128       // most rows summed from one set of tables
#define FAN_MAXTILE 256     // detector rows one warp walks (their sums and run-start masks live in smem)

// LQ lanes (z quads) per step, SB steps per stream
template <int LQ, int SB>
struct FanCfg {
    static constexpr int W = 4 * LQ;            // slices of a z window
    static constexpr int G = 32 / LQ;           // streams a warp advances at once
    static constexpr int B = SB;
    static constexpr int LOGB = SB == 16 ? 4 : 3;
    static constexpr int CW = G * SB;           // steps per chunk
    static constexpr int TS = W + 2;            // table row stride in float2 (rows stay 16-byte aligned)
    static constexpr int TROWS = CW + G;        // one spare row per stream (bank skew)
};

template <int LQ, int SB>
struct __align__(16) FanWarpSmem {
    using Cfg = FanCfg<LQ, SB>;
    float2 T[Cfg::TROWS * Cfg::TS];             // stream-local running (P,Q) per (step, slice)
    float2 carry[Cfg::G * Cfg::TS];             // (P,Q) of all earlier streams of the chunk
    float2 total[Cfg::TS];                      // (P,Q) of the whole chunk
    float4 w[Cfg::TROWS];                       // per step (stream-padded): three bilinear weights (the fourth
                                                // is 1 - their sum) and the in-plane cell's byte offset
    unsigned starts[FAN_MAXTILE / 32];          // bit r of word w: row 32w+r starts a run of equal n
    float rsum[FAN_MAXTILE];                    // running line integral of every row of the tile
};

// first row after `cur` that starts a new run (or nRows)
__device__ __forceinline__ int fan_next_start(const unsigned *starts, int cur, int nRows)
{
    const int nw = (nRows + 31) >> 5;
    int w = (cur + 1) >> 5;
    if (w >= nw) return nRows;
    unsigned bits = starts[w] & (0xffffffffu << ((cur + 1) & 31));
    for (;;) {
        if (bits) return min(nRows, w * 32 + __ffs(bits) - 1);
        if (++w >= nw) return nRows;
        bits = starts[w];
    }
}

// conservative range of real sample indices t with s + t*d inside [lo, hi)
__device__ __forceinline__ void fan_slab(double s, double d, double lo, double hi, double &t0, double &t1,
                                         bool &hit)
{
    if (d == 0.0) {
        hit = hit && (s >= lo && s < hi);
    } else {
        const double r = 1.0 / d;
        const double ta = (lo - s) * r, tb = (hi - s) * r;
        t0 = fmax(t0, fmin(ta, tb));
        t1 = fmin(t1, fmax(ta, tb));
    }
}

// predicated 16-byte read-only load: v keeps its value when `take` is false (no branch, so the load
// can be issued a step ahead of its use)
__device__ __forceinline__ void fan_ld4_if(float4 &v, const void *p, bool take)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                 : "r"((int)take), "l"(p));
}
// packed fp32x2 (sm_100 FMUL2/FFMA2): one issue slot for two slices
__device__ __forceinline__ float2 fan_mul2(float w, float a, float b)
{
    return __fmul2_rn(make_float2(w, w), make_float2(a, b));
}
__device__ __forceinline__ float2 fan_fma2(float w, float a, float b, float2 acc)
{
    return __ffma2_rn(make_float2(w, w), make_float2(a, b), acc);
}
__device__ __forceinline__ float2 fan_shfl_up2(float2 v, int d)
{
    return make_float2(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

// inclusive running (P,Q) of chunk-local step i (i >= 0) at table columns c and c+1
template <int LQ, int SB>
__device__ __forceinline__ void fan_lookup(const float2 *T, const float2 *carry, int i, int c, float2 &lo,
                                           float2 &hi)
{
    constexpr int TS = FanCfg<LQ, SB>::TS;
    const int s = i >> FanCfg<LQ, SB>::LOGB;
    const float2 *t = T + (i + s) * TS + c;
    const float2 *k = carry + s * TS + c;
    const float2 t0 = t[0], t1 = t[1], k0 = k[0], k1 = k[1];
    lo = make_float2(t0.x + k0.x, t0.y + k0.y);
    hi = make_float2(t1.x + k1.x, t1.y + k1.y);
}

// ---- phase 1 of a chunk: stream-local running sums of G, four slices per lane, into the table; then
// the carries of the streams and the chunk totals.  kA = first slice of the window (multiple of 4),
// nq = quads in it, nst = steps of the chunk.
template <int LQ, int SB>
__device__ __forceinline__ void fan_phase1(FanWarpSmem<LQ, SB> &sm, const float *__restrict__ vol, int sxs,
                                           int sys, int kA, int nq, int nst, int lane)
{
    using Cfg = FanCfg<LQ, SB>;
    constexpr int G = Cfg::G, CW = Cfg::CW, TS = Cfg::TS, B = SB;
    const int g = lane / LQ, q = lane % LQ;                       // stream g, z quad q
        const bool on = q < nq;
        const float *__restrict__ base = vol + kA + 4 * q;
        float2 pq0 = make_float2(0.f, 0.f), pq1 = pq0, pq2 = pq0, pq3 = pq0;   // running (P,Q) x 4 slices
        float4 c00 = make_float4(0.f, 0.f, 0.f, 0.f), c01 = c00, c10 = c00, c11 = c00;
        const int row0 = g * (B + 1);             // table row of the stream's first step
        const float4 *wp = sm.w + row0;
        float4 *Tw = reinterpret_cast<float4 *>(sm.T + row0 * TS + 4 * q);
        float jf = (float)(g * B - CW / 2);
        if (g * B < nst) {
            // Two register sets of corner columns, one for the even and one for the odd steps: as soon
            // as step s has consumed its set, the columns of step s+2 are requested into it (predicated:
            // no load when the cell is the same), so a whole step of arithmetic covers the L2 latency
            // and no register moves are needed.
            // byte offsets are unsigned 32-bit (pax_ax_fan_launch checks the volume is < 4 GB):
            // one 64-bit add per step, the other corners are uniform strides away
            const char *bb = reinterpret_cast<const char *>(base);
            const size_t dx4 = (size_t)sxs * 4, dy4 = (size_t)sys * 4;
            // one register set: the columns of step s+1 are requested right after step s consumed them
            float4 w = wp[0];
            unsigned off = __float_as_uint(w.w);
            {
                const char *p00 = bb + off, *p10 = p00 + dy4;
                fan_ld4_if(c00, p00, on); fan_ld4_if(c01, p00 + dx4, on);
                fan_ld4_if(c10, p10, on); fan_ld4_if(c11, p10 + dx4, on);
            }
    #pragma unroll
            for (int s = 0; s < B; ++s) {
                float4 wn = w;
                if (s + 1 < B) wn = wp[s + 1];
                const unsigned offn = __float_as_uint(wn.w);
                const bool chg = on && offn != off;
                const float w11 = 1.f - w.x - w.y - w.z;              // the weights sum to one
                const char *p00 = bb + offn, *p10 = p00 + dy4;
                float2 v0 = fan_mul2(w.x, c00.x, c00.y), v1 = fan_mul2(w.x, c00.z, c00.w);
                v0 = fan_fma2(w.y, c01.x, c01.y, v0); v1 = fan_fma2(w.y, c01.z, c01.w, v1);
                v0 = fan_fma2(w.z, c10.x, c10.y, v0); v1 = fan_fma2(w.z, c10.z, c10.w, v1);
                v0 = fan_fma2(w11, c11.x, c11.y, v0); v1 = fan_fma2(w11, c11.z, c11.w, v1);
                if (s + 1 < B) {
                    fan_ld4_if(c00, p00, chg); fan_ld4_if(c01, p00 + dx4, chg);
                    fan_ld4_if(c10, p10, chg); fan_ld4_if(c11, p10 + dx4, chg);
                }
                pq0.x += v0.x; pq1.x += v0.y; pq2.x += v1.x; pq3.x += v1.y;
                pq0.y = fmaf(jf, v0.x, pq0.y); pq1.y = fmaf(jf, v0.y, pq1.y);
                pq2.y = fmaf(jf, v1.x, pq2.y); pq3.y = fmaf(jf, v1.y, pq3.y);
                jf += 1.f;
                if (on) {
                    Tw[s * (TS / 2)] = make_float4(pq0.x, pq0.y, pq1.x, pq1.y);
                    Tw[s * (TS / 2) + 1] = make_float4(pq2.x, pq2.y, pq3.x, pq3.y);
                }
                off = offn;
                w = wn;
            }
        }
        // carry of a stream = totals of the earlier streams: inclusive scan over g, minus own
        float2 t0 = pq0, t1 = pq1, t2 = pq2, t3 = pq3;
    #pragma unroll
        for (int d = LQ; d < 32; d <<= 1) {
            const float2 o0 = fan_shfl_up2(t0, d), o1 = fan_shfl_up2(t1, d);
            const float2 o2 = fan_shfl_up2(t2, d), o3 = fan_shfl_up2(t3, d);
            if (lane >= d) {
                t0 = __fadd2_rn(t0, o0); t1 = __fadd2_rn(t1, o1);
                t2 = __fadd2_rn(t2, o2); t3 = __fadd2_rn(t3, o3);
            }
        }
        float4 *cw = reinterpret_cast<float4 *>(sm.carry + g * TS + 4 * q);
        cw[0] = make_float4(t0.x - pq0.x, t0.y - pq0.y, t1.x - pq1.x, t1.y - pq1.y);
        cw[1] = make_float4(t2.x - pq2.x, t2.y - pq2.y, t3.x - pq3.x, t3.y - pq3.y);
        if (g == G - 1) {                           // inclusive total of the last stream = chunk total
            float4 *tw = reinterpret_cast<float4 *>(sm.total + 4 * q);
            tw[0] = make_float4(t0.x, t0.y, t1.x, t1.y);
            tw[1] = make_float4(t2.x, t2.y, t3.x, t3.y);
        }
}

// ---- phase 2 state of one detector row within a chunk
struct FanRow {
    float dzf, rinv, zc0, zl, kf, acc;
    int j, hi;
    bool up, bad;

    // the row's own step range is re-derived in fp32 (the double version counted the samples; a boundary
    // sample has zero weight, so an ulp either way is harmless)
    __device__ __forceinline__ void init(int r, bool exists, double zr0, double zvr, double sz, float szf,
                                         float kminf, float kmaxf, float fI0, float fI1, int c0, int nst)
    {
        const double ddz = zr0 + (double)r * zvr;
        dzf = (float)ddz;
        float flo = fI0, fhi = exists ? fI1 : -1.f;
        const float rz = __frcp_rn(dzf);                      // inf when dzf == 0: not used then
        if (dzf != 0.f) {
            const float ta = (kminf - szf) * rz, tb = (kmaxf - szf) * rz;
            flo = fmaxf(flo, ceilf(fminf(ta, tb)));
            fhi = fminf(fhi, floorf(fmaxf(ta, tb)));
        } else if (!(szf >= kminf && szf < kmaxf)) {
            fhi = -1.f;
        }
        j = max((int)flo, c0) - c0;
        hi = min((int)fhi, c0 + nst - 1) - c0;
        if (j > hi) j = max(hi, -1) + 1;                      // nothing to do here: keep the look-up indices in range
        rinv = fabsf(rz);
        up = dzf > 0.f;
        zc0 = (float)(sz + (double)c0 * ddz);
        zl = fmaf((float)j, dzf, zc0);
        // fp32 z may round onto the clip bounds: keep the slice pair inside [kmin, kmax] (tt then
        // extrapolates by an ulp)
        kf = fminf(fmaxf(floorf(zl), kminf), kmaxf - 1.f);
        acc = 0.f;
        bad = false;
    }

    // one stretch (the steps from j on that stay in one slice pair), branch-free; a no-op once j > hi
    template <int LQ, int SB>
    __device__ __forceinline__ void stretch(const float2 *T, const float2 *carry, int kA, int nq, float kminf,
                                            float kmaxf)
    {
        constexpr int W = FanCfg<LQ, SB>::W, CW = FanCfg<LQ, SB>::CW;
        const bool live = j <= hi;
        const float tt = zl - kf;
        // steps that stay in slice pair (k, k+1): tt + m dz stays inside [0,1).  An exact hit of the boundary
        // may go either way (the interpolant is continuous); dz == 0 gives inf or NaN here, which fminf turns
        // into "the rest of the chunk".
        const int ms = (int)fminf(floorf((up ? 1.f - tt : tt) * rinv), 1.0e4f);
        const int je = max(min(hi, j + max(ms, 0)), 0);
        const int kr = (int)kf - kA;
        const int kk = min(max(kr, 0), W - 2);
        bad |= live && (kr < 0 || kr > 4 * nq - 2);
        float2 a0, a1, b0, b1;
        fan_lookup<LQ, SB>(T, carry, max(j - 1, 0), kk, a0, a1);
        fan_lookup<LQ, SB>(T, carry, je, kk, b0, b1);
        if (j <= 0) { a0 = make_float2(0.f, 0.f); a1 = a0; }
        const float dP0 = b0.x - a0.x, dQ0 = b0.y - a0.y;
        const float dP1 = b1.x - a1.x, dQ1 = b1.y - a1.y;
        const float alpha = fmaf((float)(CW / 2 - j), dzf, tt);                // t at index CW/2
        const float val = fmaf(dzf, dQ1 - dQ0, fmaf(alpha, dP1 - dP0, dP0));
        acc += live ? val : 0.f;
        j = live ? je + 1 : j;
        zl = fmaf((float)j, dzf, zc0);
        kf = fminf(fmaxf(floorf(zl), kminf), kmaxf - 1.f);
    }
};

template <int LQ, int SB, int NWARPS, int MODE>
__global__ void __launch_bounds__(NWARPS * 32) k_ax_fan(const AxArgs a)
{
    using Cfg = FanCfg<LQ, SB>;
    constexpr int W = Cfg::W, CW = Cfg::CW, LOGB = Cfg::LOGB;
    static_assert(SB == 8 || SB == 16, "8 or 16 steps per stream");
    extern __shared__ __align__(16) unsigned char fan_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    FanWarpSmem<LQ, SB> &sm = reinterpret_cast<FanWarpSmem<LQ, SB> *>(fan_smem_raw)[warp];

    const int u = blockIdx.x * NWARPS + warp;
    if (u >= a.nu) return;                                        // warps are independent: no barriers below
    const int ia = blockIdx.y;
    const int V0 = blockIdx.z * a.fan_rows;
    const int nRows = min(a.fan_rows, a.nv - V0);
    const AngleDev &A = a.angles[a.first + ia];
    const double sx = A.sx, sy = A.sy, sz = A.sz;
    const double zlo = (double)a.loz, zhi = (double)a.hiz;
    const int kmin = (int)a.loz, kmax = (int)a.hiz;               // slices a sample can touch: [kmin, kmax]
    const float kminf = a.loz, kmaxf = a.hiz;
    const int sxs = a.nzp, sys = a.nxp * a.nzp;
    const int xmaxc = (int)a.hix - 1, ymaxc = (int)a.hiy - 1;    // last in-plane cell whose +1 neighbour exists
    const float szf = (float)sz;
    const double dx = A.ox + u * A.ux - sx, dy = A.oy + u * A.uy - sy;
    const double dxy2 = dx * dx + dy * dy;
    const double racc = 1.0 / a.acc;
    const double dz0 = A.oz + (double)V0 * A.vz - sz;             // P_z - S_z of the tile's row 0; +vz per row
    unsigned my_ns = 0;
    float r2 = 0.f, mymax = -INFINITY;
    bool bad = false;

    // ---- runs of rows with equal n = ceil(|P-S| / accuracy)  (TIGRE's lattice)
    {
        int last = -2;
        for (int r0 = 0; r0 < nRows; r0 += 32) {
            const bool in = r0 + lane < nRows;
            const double dz = dz0 + (double)(r0 + lane) * A.vz;
            const int n = in ? (int)ceil(sqrt(dxy2 + dz * dz) * racc) : -1;
            int prev = __shfl_up_sync(0xffffffffu, n, 1);
            if (lane == 0) prev = last;
            const unsigned m = __ballot_sync(0xffffffffu, in && n != prev);
            if (lane == 0) sm.starts[r0 >> 5] = m;
            last = __shfl_sync(0xffffffffu, n, 31);
        }
        __syncwarp();
    }

    int cur = 0;
    while (cur < nRows) {
        // ---- the run of rows [cur, e) that share n (uniform over the warp)
        const int e = fan_next_start(sm.starts, cur, nRows);
        const double dzc = dz0 + (double)cur * A.vz;
        const int nrun = (int)ceil(sqrt(dxy2 + dzc * dzc) * racc);
        const double rn = 1.0 / (double)nrun;
        const double ddx = dx * rn, ddy = dy * rn;
        double t0 = 0.0, t1 = (double)nrun;
        bool hit = true;
        fan_slab(sx, ddx, 0.0, (double)a.hix, t0, t1, hit);
        fan_slab(sy, ddy, 0.0, (double)a.hiy, t0, t1, hit);
        // ---- rows of the run: clear their sums, count their samples, find the run's step range
        int I0 = INT_MAX, I1 = -1;
        for (int r = cur + lane; r < e; r += 32) {
            sm.rsum[r] = 0.f;
            double tz0 = t0, tz1 = t1;
            bool hz = hit;
            fan_slab(sz, (dz0 + (double)r * A.vz) * rn, zlo, zhi, tz0, tz1, hz);
            if (hz && tz0 <= tz1) {
                const int i0 = (int)ceil(tz0), i1 = (int)floor(tz1);
                if (i0 <= i1) { I0 = min(I0, i0); I1 = max(I1, i1); my_ns += (unsigned)(i1 - i0 + 1); }
            }
        }
        I0 = __reduce_min_sync(0xffffffffu, I0);
        I1 = __reduce_max_sync(0xffffffffu, I1);
        if (I0 <= I1) {
            // z (slices) of row r at step j:  sz + j * (zr0 + r * zvr)
            const double zr0 = dz0 * rn, zvr = A.vz * rn;
            const float fI0 = (float)I0, fI1 = (float)I1;
#pragma unroll 1
            for (int c0 = I0; c0 <= I1; c0 += CW) {
                const int nst = min(CW, I1 - c0 + 1);             // steps of this chunk
                __syncwarp();                                     // the previous chunk is done with w/off
                // ---- phase 0, once per (run, chunk): in-plane cell and bilinear weights of every step.
                // The chunk origin is split in double into an integer cell and a fraction; steps add t*d
                // to the fraction in fp32 (|t| < CW): a position is good to ~2e-6 voxels, nothing accumulates.
                {
                    const double ex = sx + (double)c0 * ddx, ey = sy + (double)c0 * ddy;
                    const double exi = floor(ex), eyi = floor(ey);
                    const float exf = (float)(ex - exi), eyf = (float)(ey - eyi);
                    const int exc = (int)exi, eyc = (int)eyi;
                    const float fddx = (float)ddx, fddy = (float)ddy;
#pragma unroll
                    for (int t = lane; t < CW; t += 32) {
                        // a step past the run's end reads the all-zero rim column (0,0) with weight one
                        float4 w = make_float4(1.f, 0.f, 0.f, __uint_as_float(0u));
                        if (t < nst) {
                            const float px = fmaf((float)t, fddx, exf), py = fmaf((float)t, fddy, eyf);
                            const float flx = floorf(px), fly = floorf(py);
                            const int cx = min(max(exc + (int)flx, 0), xmaxc);
                            const int cy = min(max(eyc + (int)fly, 0), ymaxc);
                            // fraction against the CLAMPED cell: a sample the double clip admitted may sit
                            // an ulp outside in fp32, and must then weigh the rim, not the next column
                            const float fx = fminf(fmaxf(px - (float)(cx - exc), 0.f), 1.f);
                            const float fy = fminf(fmaxf(py - (float)(cy - eyc), 0.f), 1.f);
                            const unsigned off = (unsigned)((cy * a.nxp + cx) * a.nzp) * 4u;
                            const float gx = 1.f - fx, gy = 1.f - fy;
                            w = make_float4(gx * gy, fx * gy, gx * fy, __uint_as_float(off));
                        }
                        sm.w[t + (t >> LOGB)] = w;
                    }
                }
                // z of row r at the first / last step of the chunk: Aa + Ba r, Ab + Bb r  (Ba, Bb >= 0)
                const double ja = (double)c0, jb = (double)(c0 + nst - 1);
                const float Aa = (float)(sz + ja * zr0), Ba = (float)(ja * zvr);
                const float Ab = (float)(sz + jb * zr0), Bb = (float)(jb * zvr);
                // (approximate reciprocals: the row bounds below are re-checked against the window)
                const float rBa = Ba > 0.f ? __frcp_rn(Ba) : 0.f, rBb = Bb > 0.f ? __frcp_rn(Bb) : 0.f;
                // rows that can touch a slice during this chunk
                int ga = cur, ge = e;
                {
                    const float lo_a = Ba > 0.f ? (kminf - 1.f - Aa) * rBa : (Aa >= kminf - 1.f ? -1e9f : 1e9f);
                    const float lo_b = Bb > 0.f ? (kminf - 1.f - Ab) * rBb : (Ab >= kminf - 1.f ? -1e9f : 1e9f);
                    const float hi_a = Ba > 0.f ? (kmaxf + 1.f - Aa) * rBa : (Aa <= kmaxf + 1.f ? 1e9f : -1e9f);
                    const float hi_b = Bb > 0.f ? (kmaxf + 1.f - Ab) * rBb : (Ab <= kmaxf + 1.f ? 1e9f : -1e9f);
                    ga = max(ga, (int)fmaxf(fminf(ceilf(fminf(lo_a, lo_b)), 1e6f), -1e6f));
                    ge = min(ge, (int)fmaxf(fminf(floorf(fmaxf(hi_a, hi_b)), 1e6f), -1e6f) + 1);
                }
                __syncwarp();
                // ---- groups of rows whose z window over this chunk fits the W slices of a stream
#pragma unroll 1
                while (ga < ge) {
                    const float fga = (float)ga;
                    const float zmin = fminf(fmaf(Ba, fga, Aa), fmaf(Bb, fga, Ab));
                    const int kA = max(kmin, (int)floorf(zmin - 2e-3f)) & ~3;
                    int gb = min(ge, ga + a.fan_piece);
                    if (kA + W - 1 < kmax) {
                        // floor(zmax + eps) + 1 <= kA + W - 1  for both ends of the chunk
                        const float lim = (float)(kA + W - 1) - 4e-3f;
                        const float ra = Ba > 0.f ? (lim - Aa) * rBa : 1e9f, rb = Bb > 0.f ? (lim - Ab) * rBb : 1e9f;
                        gb = min(gb, (int)fmaxf(fminf(floorf(fminf(ra, rb) - 1e-2f), 1e6f), -1e6f) + 1);
                    }
                    if (gb <= ga) {                               // a single row does not fit: see pax_fan_configure
                        if (lane == 0 && a.fault) atomicExch(a.fault, 1);
                        gb = ga + 1;
                    }
                    const float fgl = (float)(gb - 1);
                    const float zmax = fmaxf(fmaf(Ba, fgl, Aa), fmaf(Bb, fgl, Ab));
                    const int kB = min(kmax, (int)floorf(zmax + 2e-3f) + 1);
                    const int nq = (kB - kA + 4) >> 2;            // quads in the window
                    if (nq > LQ && lane == 0 && a.fault) atomicExch(a.fault, 3);
                    if (nq > 0) {
                        fan_phase1<LQ, SB>(sm, a.vol0, sxs, sys, kA, nq, nst, lane);
                        __syncwarp();
                        // ---- phase 2: two detector rows per lane and pass, advanced side by side so that their
                        // (serial) address and look-up chains overlap
                        for (int r = ga + lane; r < gb; r += 64) {
                            FanRow ra, rb;
                            ra.init(r, true, zr0, zvr, sz, szf, kminf, kmaxf, fI0, fI1, c0, nst);
                            rb.init(r + 32, r + 32 < gb, zr0, zvr, sz, szf, kminf, kmaxf, fI0, fI1, c0, nst);
                            while (ra.j <= ra.hi || rb.j <= rb.hi) {
                                ra.template stretch<LQ, SB>(sm.T, sm.carry, kA, nq, kminf, kmaxf);
                                rb.template stretch<LQ, SB>(sm.T, sm.carry, kA, nq, kminf, kmaxf);
                            }
                            bad |= ra.bad | rb.bad;
                            sm.rsum[r] += ra.acc;
                            if (r + 32 < gb) sm.rsum[r + 32] += rb.acc;
                        }
                        if (bad && a.fault) atomicExch(a.fault, 2);        // cannot happen: the group was cut to fit
                        __syncwarp();                             // the table is free for the next group
                    }
                    ga = gb;
                }
            }
        }
        __syncwarp();
        // ---- epilogue of the run.  The warps of the CTA own adjacent columns and fill one 32-byte sector
        // of every row between them; L2 merges the partial writes.
        for (int r = cur + lane; r < e; r += 32) {
            const int v = V0 + r;
            const size_t o = ((size_t)ia * a.nv + v) * a.nu + u;
            const double mx = ddx * a.dX, my = ddy * a.dY, mz = (dz0 + (double)r * A.vz) * rn * a.dZ;
            const float p0 = sm.rsum[r] * (float)sqrt(mx * mx + my * my + mz * mz);
            if (MODE == PAX_AX_PLAIN) {
                float val = a.a0 * p0;
                if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, u, v, a.whx, a.why, a.whz), val);
                if (a.accumulate) val += a.out[o];
                a.out[o] = val;
                mymax = fmaxf(mymax, val);
            } else {
                const float L = pax_ray_chord(A, u, v, a.whx, a.why, a.whz);
                const float w = L > a.wthr ? 1.f / L : 0.f;
                const float r_ = __ldg(a.b + o) - p0;
                r2 = fmaf(r_, r_, r2);
                a.out[o] = w * r_;
            }
        }
        cur = e;
    }
    if (a.nsamples) {
        const unsigned t = __reduce_add_sync(0xffffffffu, my_ns);
        if (lane == 0 && t) atomicAdd(a.nsamples, (unsigned long long)t);
    }
    if (MODE == PAX_AX_PLAIN && a.maxout) {
        for (int o = 16; o; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
        if (lane == 0 && mymax > -INFINITY) {
            const unsigned b = __float_as_uint(mymax);
            atomicMax(a.maxout, (b & 0x80000000u) ? ~b : (b | 0x80000000u));
        }
    }
    if (MODE == PAX_AX_RESIDUAL && a.sumsq) {
        for (int o = 16; o; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if (lane == 0) atomicAdd(a.sumsq, (double)r2);
    }
}

// ------------------------------------------------------------------------------ host side
// Piece size (rows) for which every chunk's z window fits the lanes, per kernel configuration, and
// the mean length of the equal-n runs (what the sharing is worth).  All from the plan's geometry.
template <int LQ, int SB>
static int fan_piece_rows(const PaxGeo &g, double vz, double tau, double slope)
{
    // window of a chunk <= rows*vz*tau (spread of the rows) + steps*slope (travel) + 2 (floor, +1)
    //                      + 3 (alignment to a quad) + slack  must fit W slices
    const double room = (double)FanCfg<LQ, SB>::W - 5.3 - FanCfg<LQ, SB>::CW * slope;
    int rows = room > 0 ? (int)std::floor(room / (vz * tau)) : 0;
    rows = std::max(0, std::min(rows, std::min(FAN_GROUP, (int)g.nDetecV)));
    return rows >= 32 ? rows / 32 * 32 : rows;      // whole warps of rows: no idle lanes in phase 2
}

#define FAN_NW8 5            // W=32 tables are twice as large: five warps per CTA, two CTAs per SM

int pax_fan_configure(PaxPlan *p)
{
    const PaxGeo &g = p->geo;
    const double vz = g.dDetecV / g.dVoxelZ;
    const double dmax = std::fmax(g.dVoxelX, g.dVoxelY);
    const double hx = 0.5 * g.nVoxelX * g.dVoxelX, hy = 0.5 * g.nVoxelY * g.dVoxelY;
    // tau: largest fraction of a ray's length at which it can still be inside the volume (far corner);
    // slope: largest |dz| per step of any ray, in slices
    double tau = 0.0, slope = 0.0;
    for (const AngleDev &A : p->h_angles) {
        const double rx = hx + std::fabs(A.offOX) + g.dVoxelX, ry = hy + std::fabs(A.offOY) + g.dVoxelY;
        tau = std::fmax(tau, std::fmin(1.0, ((double)A.DSO + std::sqrt(rx * rx + ry * ry)) / (double)A.DSD));
        const double nmin = std::ceil((double)A.DSD / dmax / g.accuracy);
        const double za = std::fabs(A.oz - A.sz), zb = std::fabs(A.oz + (g.nDetecV - 1) * A.vz - A.sz);
        slope = std::fmax(slope, std::fmax(za, zb) / nmin);
    }
    const int p4 = fan_piece_rows<4, 8>(g, vz, tau, slope);
    const int p8 = fan_piece_rows<8, 16>(g, vz, tau, slope);
    // a W=32 chunk costs twice a W=16 chunk per step: take it when it carries more than twice the rows
    const char *force = getenv("PAX_FAN_LQ");                  // tuning hook
    int lq = (p8 > 3 * p4) ? 8 : 4;                             // (measured at C2: 12.1 ms vs 10.0 ms)
    if (force && (atoi(force) == 4 || atoi(force) == 8)) lq = atoi(force);
    p->fan_lq = lq;
    // worst-case number of rows whose z window fits (AUTO's criterion, pax_plan_fan_info); the kernel
    // itself cuts the groups per chunk from the actual z range, up to FAN_GROUP rows
    p->fan_fit = lq == 8 ? p8 : p4;
    p->fan_piece = p->fan_fit >= 1 ? FAN_GROUP : 0;
    p->fan_warps = lq == 8 ? FAN_NW8 : FAN_NW;
    // rows one warp walks: whole columns when short, else equal tiles of <= FAN_MAXTILE rows
    const int tiles = (g.nDetecV + FAN_MAXTILE - 1) / FAN_MAXTILE;
    p->fan_rows = (g.nDetecV + tiles - 1) / tiles;
    // mean run length of equal n over three columns of the first angle
    {
        const AngleDev &A = p->h_angles[0];
        long runs = 0;
        const int us[3] = {0, g.nDetecU / 2, g.nDetecU - 1};
        for (int c = 0; c < 3; ++c) {
            const double dx = A.ox + us[c] * A.ux - A.sx, dy = A.oy + us[c] * A.uy - A.sy;
            double prev = -1.0;
            for (int v = 0; v < g.nDetecV; ++v) {
                const double dz = A.oz + v * A.vz - A.sz;
                const double n = std::ceil(std::sqrt(dx * dx + dy * dy + dz * dz) / g.accuracy);
                if (n != prev) ++runs;
                prev = n;
            }
        }
        p->fan_run = 3.0 * g.nDetecV / (double)runs;
    }
    return PAX_OK;
}

template <int LQ, int SB, int NWARPS, int MODE>
static int fan_launch(const AxArgs &a, int count, cudaStream_t st)
{
    const size_t smem = sizeof(FanWarpSmem<LQ, SB>) * NWARPS;
    PAX_CHECK_CUDA(cudaFuncSetAttribute(k_ax_fan<LQ, SB, NWARPS, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(pax_cdiv(a.nu, NWARPS), count, pax_cdiv(a.nv, a.fan_rows));
    k_ax_fan<LQ, SB, NWARPS, MODE><<<grid, NWARPS * 32, smem, st>>>(a);
    return PAX_OK;
}

int pax_ax_fan_launch(const PaxPlan *p, const AxArgs &a0, int mode, int count, cudaStream_t st)
{
    AxArgs a = a0;
    a.fan_rows = p->fan_rows;
    a.fan_piece = p->fan_piece;
    a.fault = p->d_fault;
    PAX_REQUIRE(((uintptr_t)a.vol0 & 15) == 0, "pax_ax: resident volume must be 16-byte aligned");
    PAX_REQUIRE(p->vol_elems < ((size_t)1 << 30), "pax_ax: resident volume too large for the fan-plane kernel");
    PAX_REQUIRE(p->fan_piece >= 1, "pax_ax: the fan-plane projector does not fit this geometry "
                                   "(the z window of a single detector row exceeds its lanes)");
    const bool plain = mode == PAX_AX_PLAIN;
    if (p->fan_lq == 8)
        return plain ? fan_launch<8, 16, FAN_NW8, PAX_AX_PLAIN>(a, count, st)
                     : fan_launch<8, 16, FAN_NW8, PAX_AX_RESIDUAL>(a, count, st);
    return plain ? fan_launch<4, 8, FAN_NW, PAX_AX_PLAIN>(a, count, st)
                 : fan_launch<4, 8, FAN_NW, PAX_AX_RESIDUAL>(a, count, st);
}

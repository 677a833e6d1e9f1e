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
// Kernel shape (sm_100a, no tensor cores: this is a gather plus prefix sums).  A CTA owns one angle,
// UC adjacent detector columns and a tile of `fan_rows` rows.  Per column and per run of equal n it
// walks the chord in chunks of NW*B steps:
//   phase 0  every thread prepares one step: in-plane cell offset and the four bilinear weights
//            (positions from double, so nothing accumulates along the ray);
//   phase 1  warp w owns B consecutive steps, lane l owns slice kA_w + l of the block's z window
//            (the resident volume is z-fastest, so the four corner loads of a step are four
//            coalesced 128-byte rows); it accumulates P and Q down the block and writes them to
//            shared memory;
//   phase 2  thread r owns detector row r of the tile and adds up its stretches from the tables.
// The z window of a block (rows of the run x B steps) must fit the 32 lanes; pax_fan_configure
// derives the row tile from the geometry so that it always does.
#include <climits>
#include <cmath>
#include <cstdlib>

#include "pax_internal.h"

#define FAN_LQ 8           // lanes (z quads) per step: a chunk's z window is <= 32 slices
#define FAN_W (4 * FAN_LQ)
#define FAN_G (32 / FAN_LQ) // streams (runs of consecutive steps) a warp advances at once
#define FAN_B 8            // steps per stream
#define FAN_TS (FAN_W + 2) // table row stride in float2 (rows stay 16-byte aligned)
#define FAN_UC 8           // detector columns per CTA (one 32-byte sector of every output row)

template <int NW>
struct FanSmem {
    static constexpr int R = NW * 32;                  // threads; rows per tile <= R
    static constexpr int NS = NW * FAN_G;              // streams per chunk
    static constexpr int C = NS * FAN_B;               // steps per chunk
    static constexpr int TROWS = C + NS;               // one spare row per stream (bank skew)
    float2 T[TROWS * FAN_TS];                          // running (P,Q) per (step, slice), stream-local
    float2 carry[NS * FAN_TS];                         // (P,Q) of all earlier streams of the chunk
    float2 gsum[NW * FAN_W];                           // per-warp totals, input of the carry scan
    float4 w[C + NS];                                  // bilinear weights per step (stream-padded)
    int off[C + NS];                                   // in-plane cell offset per step
    int n[R];
    unsigned starts[NW];                               // bit r of word w: row 32w+r starts a run of equal n
    int red[2];
};

// first row after `cur` that starts a new run (or nRows)
__device__ __forceinline__ int fan_next_start(const unsigned *starts, int cur, int nRows, int nw)
{
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

__device__ __forceinline__ float4 fan_ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// inclusive running (P,Q) of chunk-local step i (i >= 0) at table column c and c+1
__device__ __forceinline__ void fan_lookup(const float2 *T, const float2 *carry, int i, int c, float2 &lo,
                                           float2 &hi)
{
    const int s = i >> 3;                              // FAN_B == 8
    const float2 *t = T + (i + s) * FAN_TS + c;
    const float2 *k = carry + s * FAN_TS + c;
    const float2 t0 = t[0], t1 = t[1], k0 = k[0], k1 = k[1];
    lo = make_float2(t0.x + k0.x, t0.y + k0.y);
    hi = make_float2(t1.x + k1.x, t1.y + k1.y);
}

template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32) k_ax_fan(const AxArgs a)
{
    using SM = FanSmem<NW>;
    constexpr int R = SM::R, C = SM::C, NS = SM::NS, B = FAN_B, TS = FAN_TS, W = FAN_W;
    static_assert(FAN_B == 8, "fan_lookup assumes 8 steps per stream");
    extern __shared__ __align__(16) unsigned char fan_smem_raw[];
    SM &sm = *reinterpret_cast<SM *>(fan_smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ia = blockIdx.y;
    const int V0 = blockIdx.z * a.fan_rows;
    const int nRows = min(a.fan_rows, a.nv - V0);
    const int ucn = min(FAN_UC, a.nu - (int)blockIdx.x * FAN_UC);
    const AngleDev &A = a.angles[a.first + ia];
    const double sx = A.sx, sy = A.sy, sz = A.sz;
    const double dz = A.oz + (double)(V0 + tid) * A.vz - sz;      // this thread's row
    const double zlo = (double)a.loz, zhi = (double)a.hiz;
    const int kmin = (int)a.loz, kmax = (int)a.hiz;               // slices a sample can touch: [kmin, kmax]
    const float kminf = a.loz, kmaxf = a.hiz;
    const int sxs = a.nzp, sys = a.nxp * a.nzp;
    // phase-1 role of this lane: stream g of the warp, z quad q of the window
    const int g = lane / FAN_LQ, q = lane % FAN_LQ;
    const int strm = warp * FAN_G + g;
    unsigned my_ns = 0;
    float r2 = 0.f, mymax = -INFINITY;

#pragma unroll 1
    for (int uc = 0; uc < ucn; ++uc) {
        float sum = 0.f, steplen = 0.f;
        const int u = blockIdx.x * FAN_UC + uc;
        const double dx = A.ox + u * A.ux - sx, dy = A.oy + u * A.uy - sy;
        const double dxy2 = dx * dx + dy * dy;
        // TIGRE's lattice: n = ceil(|P-S| / accuracy), per row
        sm.n[tid] = (tid < nRows) ? (int)ceil(sqrt(dxy2 + dz * dz) / a.acc) : -1;
        __syncthreads();
        {
            const bool st = tid < nRows && (tid == 0 || sm.n[tid] != sm.n[tid - 1]);
            const unsigned bal = __ballot_sync(0xffffffffu, st);
            if (lane == 0) sm.starts[warp] = bal;
        }
        __syncthreads();
        int cur = 0;
        while (cur < nRows) {
            // ---- the run of rows [cur, e) that share n (uniform over the CTA)
            const int nrun = sm.n[cur];
            const int e = fan_next_start(sm.starts, cur, nRows, NW);
            const double rn = 1.0 / (double)nrun;
            const double ddx = dx * rn, ddy = dy * rn;
            double t0 = 0.0, t1 = (double)nrun;
            bool hit = true;
            fan_slab(sx, ddx, 0.0, (double)a.hix, t0, t1, hit);
            fan_slab(sy, ddy, 0.0, (double)a.hiy, t0, t1, hit);
            // ---- this thread's row: its own z clip
            const bool mine = tid >= cur && tid < e;
            const double ddz = dz * rn;
            int i0 = 0, i1 = -1;
            if (mine) {
                double tz0 = t0, tz1 = t1;
                bool hz = hit;
                fan_slab(sz, ddz, zlo, zhi, tz0, tz1, hz);
                if (hz && tz0 <= tz1) { i0 = (int)ceil(tz0); i1 = (int)floor(tz1); }
            }
            const bool valid = i0 <= i1;
            if (tid == 0) { sm.red[0] = INT_MAX; sm.red[1] = -1; }
            __syncthreads();
            {
                const int lo = __reduce_min_sync(0xffffffffu, valid ? i0 : INT_MAX);
                const int hi = __reduce_max_sync(0xffffffffu, valid ? i1 : -1);
                if (lane == 0 && hi >= 0) { atomicMin(&sm.red[0], lo); atomicMax(&sm.red[1], hi); }
            }
            __syncthreads();
            const int I0 = sm.red[0], I1 = sm.red[1];
            if (valid) {
                const double mx = ddx * a.dX, my = ddy * a.dY, mz = ddz * a.dZ;
                steplen = (float)sqrt(mx * mx + my * my + mz * mz);
                my_ns += (unsigned)(i1 - i0 + 1);
            }
            if (I0 <= I1) {
                // z slopes per step of the run's first and last row: the corners of every z window
                const double dzA = (A.oz + (double)(V0 + cur) * A.vz - sz) * rn;
                const double dzB = (A.oz + (double)(V0 + e - 1) * A.vz - sz) * rn;
                const float fdz = (float)ddz;
                const float inv = 1.f / fabsf(fdz);
#pragma unroll 1
                for (int c0 = I0; c0 <= I1; c0 += C) {
                    const int nst = min(C, I1 - c0 + 1);          // steps of this chunk
                    // ---- phase 0: every thread prepares steps (in-plane cell + bilinear weights)
                    for (int t = tid; t < C; t += R) {
                        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                        int off = 0;
                        if (t < nst) {
                            const double j = (double)(c0 + t);
                            const float qx = (float)(sx + j * ddx), qy = (float)(sy + j * ddy);
                            const float cx = fminf(fmaxf(floorf(qx), 0.f), a.hix - 1.f);
                            const float cy = fminf(fmaxf(floorf(qy), 0.f), a.hiy - 1.f);
                            const float fx = fminf(fmaxf(qx - cx, 0.f), 1.f), fy = fminf(fmaxf(qy - cy, 0.f), 1.f);
                            off = ((int)cy * a.nxp + (int)cx) * a.nzp;
                            const float gx = 1.f - fx, gy = 1.f - fy;
                            w = make_float4(gx * gy, fx * gy, gx * fy, fx * fy);
                        }
                        sm.w[t + (t >> 3)] = w;
                        sm.off[t + (t >> 3)] = off;
                    }
                    // ---- z window of the chunk: all rows of the run x all its steps (uniform)
                    int kA, nq;
                    {
                        const double ja = (double)c0, jb = (double)(c0 + nst - 1);
                        const double z0 = sz + ja * dzA, z1 = sz + ja * dzB, z2 = sz + jb * dzA, z3 = sz + jb * dzB;
                        const double zmin = fmin(fmin(z0, z1), fmin(z2, z3));
                        const double zmax = fmax(fmax(z0, z1), fmax(z2, z3));
                        kA = max(kmin, (int)floor(zmin - 1e-3)) & ~3;
                        const int kB = min(kmax, (int)floor(zmax + 1e-3) + 1);
                        nq = (kB - kA + 4) >> 2;                  // quads in the window; <= FAN_LQ by the
                        if (nq > FAN_LQ && tid == 0 && a.fault)   // row tile pax_fan_configure chose
                            atomicExch(a.fault, 1);
                    }
                    __syncthreads();
                    // ---- phase 1: stream-local running sums, four slices per lane
                    {
                        const bool on = q < nq;
                        const float *__restrict__ base = a.vol0 + kA + 4 * q;
                        float4 P = make_float4(0.f, 0.f, 0.f, 0.f), Q = P;
                        float4 c00 = P, c01 = P, c10 = P, c11 = P;
                        int poff = -1;
                        const int i_first = strm * B;
                        const float4 *wp = sm.w + i_first + strm;
                        const int *op = sm.off + i_first + strm;
                        float4 *Tw = reinterpret_cast<float4 *>(sm.T + (i_first + strm) * TS + 4 * q);
                        float jf = (float)(i_first - C / 2);
                        if (i_first < nst) {
#pragma unroll 2
                            for (int s = 0; s < B; ++s) {
                                const float4 w = wp[s];
                                const int off = op[s];
                                if (on && off != poff) {
                                    const float *p = base + off;
                                    c00 = fan_ld4(p); c01 = fan_ld4(p + sxs);
                                    c10 = fan_ld4(p + sys); c11 = fan_ld4(p + sys + sxs);
                                }
                                poff = off;
                                float4 v;
                                v.x = fmaf(w.w, c11.x, fmaf(w.z, c10.x, fmaf(w.y, c01.x, w.x * c00.x)));
                                v.y = fmaf(w.w, c11.y, fmaf(w.z, c10.y, fmaf(w.y, c01.y, w.x * c00.y)));
                                v.z = fmaf(w.w, c11.z, fmaf(w.z, c10.z, fmaf(w.y, c01.z, w.x * c00.z)));
                                v.w = fmaf(w.w, c11.w, fmaf(w.z, c10.w, fmaf(w.y, c01.w, w.x * c00.w)));
                                P.x += v.x; P.y += v.y; P.z += v.z; P.w += v.w;
                                Q.x = fmaf(jf, v.x, Q.x); Q.y = fmaf(jf, v.y, Q.y);
                                Q.z = fmaf(jf, v.z, Q.z); Q.w = fmaf(jf, v.w, Q.w);
                                jf += 1.f;
                                if (on) {
                                    Tw[s * (TS / 2)] = make_float4(P.x, Q.x, P.y, Q.y);
                                    Tw[s * (TS / 2) + 1] = make_float4(P.z, Q.z, P.w, Q.w);
                                }
                            }
                        }
                        // stream totals -> warp totals (exclusive over the warp's streams = carry within warp)
                        // lanes of stream g hold its totals for quad q; scan over g with shuffles
                        float4 cP = make_float4(0.f, 0.f, 0.f, 0.f), cQ = cP;      // sum of earlier streams of this warp
                        float4 tP = P, tQ = Q;                                     // running inclusive
#pragma unroll
                        for (int d = FAN_LQ; d < 32; d <<= 1) {
                            float4 oP, oQ;
                            oP.x = __shfl_up_sync(0xffffffffu, tP.x, d); oP.y = __shfl_up_sync(0xffffffffu, tP.y, d);
                            oP.z = __shfl_up_sync(0xffffffffu, tP.z, d); oP.w = __shfl_up_sync(0xffffffffu, tP.w, d);
                            oQ.x = __shfl_up_sync(0xffffffffu, tQ.x, d); oQ.y = __shfl_up_sync(0xffffffffu, tQ.y, d);
                            oQ.z = __shfl_up_sync(0xffffffffu, tQ.z, d); oQ.w = __shfl_up_sync(0xffffffffu, tQ.w, d);
                            if (lane >= d) {
                                tP.x += oP.x; tP.y += oP.y; tP.z += oP.z; tP.w += oP.w;
                                tQ.x += oQ.x; tQ.y += oQ.y; tQ.z += oQ.z; tQ.w += oQ.w;
                            }
                        }
                        cP = make_float4(tP.x - P.x, tP.y - P.y, tP.z - P.z, tP.w - P.w);
                        cQ = make_float4(tQ.x - Q.x, tQ.y - Q.y, tQ.z - Q.z, tQ.w - Q.w);
                        // park the in-warp carry; the cross-warp part is added after the barrier
                        float4 *cw = reinterpret_cast<float4 *>(sm.carry + strm * TS + 4 * q);
                        cw[0] = make_float4(cP.x, cQ.x, cP.y, cQ.y);
                        cw[1] = make_float4(cP.z, cQ.z, cP.w, cQ.w);
                        if (g == FAN_G - 1) {                      // last stream's inclusive total = warp total
                            float4 *gw = reinterpret_cast<float4 *>(sm.gsum + warp * W + 4 * q);
                            gw[0] = make_float4(tP.x, tQ.x, tP.y, tQ.y);
                            gw[1] = make_float4(tP.z, tQ.z, tP.w, tQ.w);
                        }
                    }
                    __syncthreads();
                    // ---- carry across warps: thread (stream s, column c) adds the totals of earlier warps
                    for (int t = tid; t < NS * W; t += R) {
                        const int s = t / W, c = t % W;
                        const int ws = s / FAN_G;
                        float2 acc = sm.carry[s * TS + c];
                        for (int w2 = 0; w2 < ws; ++w2) {
                            const float2 o = sm.gsum[w2 * W + c];
                            acc.x += o.x; acc.y += o.y;
                        }
                        sm.carry[s * TS + c] = acc;
                    }
                    __syncthreads();
                    // ---- phase 2: one detector row per thread, one table look-up pair per stretch end
                    if (valid) {
                        const int lo = max(i0, c0) - c0, hi = min(i1, c0 + C - 1) - c0;   // chunk-local steps
                        if (lo <= hi) {
                            const float zc0 = (float)(sz + (double)c0 * ddz);
                            const float jc = (float)(C / 2);
                            int j = lo;
                            // running sums just before the first step
                            float zl = fmaf((float)j, fdz, zc0);
                            while (j <= hi) {
                                // fp32 z may round onto the clip bounds the double range excluded: keep the
                                // slice pair inside [kmin, kmax] (tt then extrapolates by an ulp)
                                const float kf = fminf(fmaxf(floorf(zl), kminf), kmaxf - 1.f);
                                const float tt = zl - kf;
                                // steps that stay in slice pair (k, k+1)
                                int m;
                                if (fdz > 0.f) m = (int)fminf(ceilf((1.f - tt) * inv), 1.0e4f) - 1;
                                else if (fdz < 0.f) m = (int)fminf(floorf(tt * inv), 1.0e4f);
                                else m = C;
                                const int je = min(hi, j + max(m, 0));
                                int kk = (int)kf - kA;
                                if (kk < 0 || kk > 4 * nq - 2) {               // cannot happen: see pax_fan_configure
                                    if (a.fault) atomicExch(a.fault, 2);
                                    kk = min(max(kk, 0), W - 2);
                                }
                                float2 a0 = make_float2(0.f, 0.f), a1 = a0, b0, b1;
                                if (j > 0) fan_lookup(sm.T, sm.carry, j - 1, kk, a0, a1);
                                fan_lookup(sm.T, sm.carry, je, kk, b0, b1);
                                const float dP0 = b0.x - a0.x, dQ0 = b0.y - a0.y;
                                const float dP1 = b1.x - a1.x, dQ1 = b1.y - a1.y;
                                const float alpha = fmaf(jc - (float)j, fdz, tt);      // t at index C/2
                                sum += fmaf(fdz, dQ1 - dQ0, fmaf(alpha, dP1 - dP0, dP0));
                                j = je + 1;
                                zl = fmaf((float)j, fdz, zc0);
                            }
                        }
                    }
                    __syncthreads();
                }
            } else {
                __syncthreads();          // sm.red is re-armed by the next run
            }
            cur = e;
        }
        // ---- epilogue of this column for this thread's row.  The eight columns of the CTA fill one
        // 32-byte sector of the row between them; L2 merges the partial writes.
        if (tid < nRows) {
            const int v = V0 + tid;
            const size_t o = ((size_t)ia * a.nv + v) * a.nu + u;
            const float p0 = sum * steplen;
            if (MODE == PAX_AX_PLAIN) {
                float val = a.a0 * p0;
                if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, u, v, a.whx, a.why, a.whz), val);
                if (a.accumulate) val += a.out[o];
                a.out[o] = val;
                mymax = fmaxf(mymax, val);
            } else {
                const float L = pax_ray_chord(A, u, v, a.whx, a.why, a.whz);
                const float w = L > a.wthr ? 1.f / L : 0.f;
                const float r = __ldg(a.b + o) - p0;
                r2 = fmaf(r, r, r2);
                a.out[o] = w * r;
            }
        }
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
// Row tile for which every block's z window (rows of a run x FAN_B steps) fits FAN_K slices, and
// the mean length of the equal-n runs (what the sharing is worth).  All from the plan's geometry.
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
    // window of a chunk <= rows*vz*tau (spread of the rows) + steps*slope (travel) + 2 (floor, +1)
    //                      + 3 (alignment to a quad) must fit FAN_W slices
    p->fan_rows = 0;
    p->fan_warps = 0;
    const int nws[3] = {8, 6, 4};
    const char *force = getenv("PAX_FAN_WARPS");               // tuning hook
    for (int c = 0; c < 3; ++c) {
        const int nw = nws[c], steps = nw * FAN_G * FAN_B;
        if (force && atoi(force) != nw) continue;
        const double room = (double)FAN_W - 5.3 - steps * slope;
        int rows = room > 0 ? (int)std::floor(room / (vz * tau)) : 0;
        rows = std::min(rows, std::min(nw * 32, (int)g.nDetecV));
        if (rows < 1) continue;
        // equal tiles; prefer the configuration that keeps the most threads busy in phase 2
        const int tiles = (g.nDetecV + rows - 1) / rows;
        rows = (g.nDetecV + tiles - 1) / tiles;
        if (p->fan_rows == 0 || rows * p->fan_warps > p->fan_rows * nw) {
            p->fan_rows = rows;
            p->fan_warps = nw;
        }
    }
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

template <int NW, int MODE>
static int fan_launch(const AxArgs &a, int count, cudaStream_t st)
{
    const size_t smem = sizeof(FanSmem<NW>);
    PAX_CHECK_CUDA(cudaFuncSetAttribute(k_ax_fan<NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        (int)smem));
    dim3 grid(pax_cdiv(a.nu, FAN_UC), count, pax_cdiv(a.nv, a.fan_rows));
    k_ax_fan<NW, MODE><<<grid, NW * 32, smem, st>>>(a);
    return PAX_OK;
}

int pax_ax_fan_launch(const PaxPlan *p, const AxArgs &a0, int mode, int count, cudaStream_t st)
{
    AxArgs a = a0;
    a.fan_rows = p->fan_rows;
    a.fault = p->d_fault;
    PAX_REQUIRE(((uintptr_t)a.vol0 & 15) == 0, "pax_ax: resident volume must be 16-byte aligned");
    PAX_REQUIRE(p->fan_rows >= 1, "pax_ax: the fan-plane projector does not fit this geometry "
                                  "(z window of a single detector row exceeds %d slices)", FAN_W);
    const bool plain = mode == PAX_AX_PLAIN;
    switch (p->fan_warps) {
    case 4: return plain ? fan_launch<4, PAX_AX_PLAIN>(a, count, st) : fan_launch<4, PAX_AX_RESIDUAL>(a, count, st);
    case 6: return plain ? fan_launch<6, PAX_AX_PLAIN>(a, count, st) : fan_launch<6, PAX_AX_RESIDUAL>(a, count, st);
    default: return plain ? fan_launch<8, PAX_AX_PLAIN>(a, count, st) : fan_launch<8, PAX_AX_RESIDUAL>(a, count, st);
    }
}

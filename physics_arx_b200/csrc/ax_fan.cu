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

#define FAN_K 32           // slices per block window = lanes
#define FAN_KP 33          // table row stride in float2 (padding against bank conflicts)
#define FAN_B 32           // steps per warp block
#define FAN_UC 8           // detector columns per CTA (one 32-byte sector of every output row)

template <int NW>
struct FanSmem {
    static constexpr int R = NW * 32;
    static constexpr int C = NW * FAN_B;
    float2 T[NW * (FAN_B + 1) * FAN_KP];
    float4 w[C];
    int off[C];
    int n[R];
    unsigned starts[NW];          // bit r of word w: row 32w+r starts a run of equal n
    float out[FAN_UC * R];
    int kA[NW];
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

template <int NW, int MODE>
__global__ void __launch_bounds__(NW * 32) k_ax_fan(const AxArgs a)
{
    constexpr int R = NW * 32, C = NW * FAN_B, B = FAN_B, KP = FAN_KP;
    extern __shared__ __align__(16) unsigned char fan_smem_raw[];
    FanSmem<NW> &sm = *reinterpret_cast<FanSmem<NW> *>(fan_smem_raw);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ia = blockIdx.y;
    const int V0 = blockIdx.z * a.fan_rows;
    const int nRows = min(a.fan_rows, a.nv - V0);
    const int ucn = min(FAN_UC, a.nu - (int)blockIdx.x * FAN_UC);
    const AngleDev &A = a.angles[a.first + ia];
    const double sx = A.sx, sy = A.sy, sz = A.sz;
    const double dz = A.oz + (double)(V0 + tid) * A.vz - sz;      // this thread's row
    const double zlo = (double)a.loz, zhi = (double)a.hiz;
    const int kmin = (int)a.loz, kmax = (int)a.hiz;               // slices that can be non-zero: [kmin+1, kmax-1]
    const int sxs = a.nzp, sys = a.nxp * a.nzp;
    unsigned my_ns = 0;

    for (int uc = 0; uc < ucn; ++uc) {
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
        float sum = 0.f, steplen = 0.f;
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
                for (int c0 = I0; c0 <= I1; c0 += C) {
                    // ---- phase 0: one step per thread
                    for (int t = tid; t < C; t += R) {
                        const int j = c0 + t;
                        float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
                        int off = 0;
                        if (j <= I1) {
                            const float qx = (float)(sx + (double)j * ddx), qy = (float)(sy + (double)j * ddy);
                            const float cx = fminf(fmaxf(floorf(qx), 0.f), a.hix - 1.f);
                            const float cy = fminf(fmaxf(floorf(qy), 0.f), a.hiy - 1.f);
                            const float fx = fminf(fmaxf(qx - cx, 0.f), 1.f), fy = fminf(fmaxf(qy - cy, 0.f), 1.f);
                            off = ((int)cy * a.nxp + (int)cx) * a.nzp;
                            const float gx = 1.f - fx, gy = 1.f - fy;
                            w = make_float4(gx * gy, fx * gy, gx * fy, fx * fy);
                        }
                        sm.w[t] = w;
                        sm.off[t] = off;
                    }
                    // ---- z window of this warp's block of steps
                    const int jb = c0 + warp * B;
                    const int jl = min(jb + B - 1, I1);
                    int kA = 0, kB = -1;
                    if (jb <= I1) {
                        const double z0 = sz + (double)jb * dzA, z1 = sz + (double)jb * dzB;
                        const double z2 = sz + (double)jl * dzA, z3 = sz + (double)jl * dzB;
                        const double zmin = fmin(fmin(z0, z1), fmin(z2, z3));
                        const double zmax = fmax(fmax(z0, z1), fmax(z2, z3));
                        kA = max(kmin, (int)floor(zmin - 1e-3));
                        kB = min(kmax, (int)floor(zmax + 1e-3) + 1);
                    }
                    if (lane == 0) sm.kA[warp] = kA;
                    __syncthreads();
                    // ---- phase 1: running sums of G over the block, one slice per lane
                    if (kB >= kA) {
                        const int k = kA + lane;
                        const bool on = k <= kB;
                        const float *__restrict__ base = a.vol0 + k;
                        float2 *Tw = sm.T + warp * (B + 1) * KP + lane;
                        Tw[0] = make_float2(0.f, 0.f);
                        float P = 0.f, Q = 0.f;
                        const int nst = jl - jb + 1;
                        const float4 *wp = sm.w + warp * B;
                        const int *op = sm.off + warp * B;
#pragma unroll 4
                        for (int s = 0; s < nst; ++s) {
                            const float4 w = wp[s];
                            const int off = op[s];
                            float g = 0.f;
                            if (on) {
                                const float *p = base + off;
                                g = w.x * __ldg(p);
                                g = fmaf(w.y, __ldg(p + sxs), g);
                                g = fmaf(w.z, __ldg(p + sys), g);
                                g = fmaf(w.w, __ldg(p + sys + sxs), g);
                            }
                            P += g;
                            Q = fmaf((float)s, g, Q);
                            Tw[(s + 1) * KP] = make_float2(P, Q);
                        }
                    }
                    __syncthreads();
                    // ---- phase 2: one detector row per thread
                    if (valid) {
                        const int lo = max(i0, c0) - c0, hi = min(i1, c0 + C - 1) - c0;   // chunk-local steps
                        if (lo <= hi) {
                            const float zc0 = (float)(sz + (double)c0 * ddz);
                            for (int w2 = lo / B; w2 <= hi / B; ++w2) {
                                const int jb2 = w2 * B;
                                const int bl = min(hi, jb2 + B - 1);
                                const int kAw = sm.kA[w2];
                                const float2 *Tb = sm.T + w2 * (B + 1) * KP;
                                int j = max(lo, jb2);
                                while (j <= bl) {
                                    const float zl = fmaf((float)j, fdz, zc0);
                                    const float kf = floorf(zl);
                                    const float tt = zl - kf;
                                    // steps that stay in slice pair (k, k+1)
                                    int m;
                                    if (fdz > 0.f) m = (int)fminf(ceilf((1.f - tt) * inv), 1.0e4f) - 1;
                                    else if (fdz < 0.f) m = (int)fminf(floorf(tt * inv), 1.0e4f);
                                    else m = B;
                                    const int je = min(bl, j + max(m, 0));
                                    const int r0 = j - jb2, r1 = je - jb2 + 1;
                                    const int kk = min(max((int)kf - kAw, 0), FAN_K - 2);
                                    const float2 A0 = Tb[r0 * KP + kk], A1 = Tb[r1 * KP + kk];
                                    const float2 B0 = Tb[r0 * KP + kk + 1], B1 = Tb[r1 * KP + kk + 1];
                                    const float dP0 = A1.x - A0.x, dQ0 = A1.y - A0.y;
                                    const float dP1 = B1.x - B0.x, dQ1 = B1.y - B0.y;
                                    const float alpha = fmaf(-(float)r0, fdz, tt);
                                    sum += fmaf(fdz, dQ1 - dQ0, fmaf(alpha, dP1 - dP0, dP0));
                                    j = je + 1;
                                }
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
        sm.out[uc * R + tid] = sum * steplen;
    }
    __syncthreads();

    // ---- epilogue, transposed so that a thread group writes whole 32-byte sectors of a row
    float r2 = 0.f, mymax = -INFINITY;
    for (int el = tid; el < nRows * FAN_UC; el += R) {
        const int row = el / FAN_UC, uc = el % FAN_UC;
        if (uc >= ucn) continue;
        const int v = V0 + row, u = blockIdx.x * FAN_UC + uc;
        const size_t o = ((size_t)ia * a.nv + v) * a.nu + u;
        const float p0 = sm.out[uc * R + row];
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
    double tau = 0.0, slope = 0.0;
    for (const AngleDev &A : p->h_angles) {
        const double r = std::sqrt((hx + std::fabs(A.offOX) + g.dVoxelX) * (hx + std::fabs(A.offOX) + g.dVoxelX) +
                                   (hy + std::fabs(A.offOY) + g.dVoxelY) * (hy + std::fabs(A.offOY) + g.dVoxelY));
        tau = std::fmax(tau, std::fmin(1.0, ((double)A.DSO + r) / (double)A.DSD));
        const double nmin = std::ceil((double)A.DSD / dmax / g.accuracy);
        const double za = std::fabs(A.oz - A.sz), zb = std::fabs(A.oz + (g.nDetecV - 1) * A.vz - A.sz);
        slope = std::fmax(slope, std::fmax(za, zb) / nmin);
    }
    const double room = (double)FAN_K - 3.2 - FAN_B * slope;
    int rows = room > 0 ? (int)std::floor(room / (vz * tau)) : 0;
    rows = std::min(rows, std::min(256, (int)g.nDetecV));
    p->fan_rows = 0;
    p->fan_warps = 0;
    if (rows >= 1) {
        // whole warps of rows when the tile is large: no idle lanes in phase 2
        if (rows >= 32 && rows < (int)g.nDetecV) rows = rows / 32 * 32;
        const int nw = rows > 192 ? 8 : rows > 128 ? 6 : 4;
        p->fan_rows = rows;
        p->fan_warps = nw;
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
    PAX_REQUIRE(p->fan_rows >= 1, "pax_ax: the fan-plane projector does not fit this geometry "
                                  "(z window of a single detector row exceeds %d slices)", FAN_K);
    const bool plain = mode == PAX_AX_PLAIN;
    switch (p->fan_warps) {
    case 4: return plain ? fan_launch<4, PAX_AX_PLAIN>(a, count, st) : fan_launch<4, PAX_AX_RESIDUAL>(a, count, st);
    case 6: return plain ? fan_launch<6, PAX_AX_PLAIN>(a, count, st) : fan_launch<6, PAX_AX_RESIDUAL>(a, count, st);
    default: return plain ? fan_launch<8, PAX_AX_PLAIN>(a, count, st) : fan_launch<8, PAX_AX_RESIDUAL>(a, count, st);
    }
}

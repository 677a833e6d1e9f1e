// Fan-plane forward projector, event-driven: tigre.Ax 'interpolated' (reference call site
// pCT_CBCT_Reconstruction.py:86 and, inside ossart, :103; algorithm SURVEY.md A.3) on TIGRE's own sample
// lattice, with the in-plane half of the trilinear interpolation shared by all rays of a detector column
// AND the rows touching the shared tables only when they cross a slice boundary.
//
// Sharing.  The rays of a detector column u lie in one vertical plane through the source;
// rows with equal n = ceil(|P-S|/accuracy) (a "run") share the in-plane sample positions, so
//   sample(i, v) = (1-t) G_i(k) + t G_i(k+1),  k = floor(z_i(v)), t = z_i(v) - k,  G_i(k) = bilinear_xy(slice k).
//
// Events.  With running sums over the steps of an epoch,  A_k(j) = sum_{i<=j} G_i(k),
// C_k(j) = sum_{i<=j} (i - ref) G_i(k), the sum of a row over the steps it spends in slice pair (k,k+1) is a
// difference of  f_k(j) = A_k + tau_k (A_{k+1} - A_k) + s (C_{k+1} - C_k)   (tau_k = z_ref(v) - k, s = dz per step),
// and the differences telescope: the line integral of row v over an epoch is
//   f_K(last step)  -  sum over crossings of a boundary k at step j of  dir * (tau_k D2A_k(j-1) + s D2C_k(j-1)),
//   D2X_k = X_{k-1} - 2 X_k + X_{k+1},  dir = +1 upwards.
// A row costs one table look-up per slice-boundary crossing (~10 per ray at C2) and one per epoch, instead of
// one trilinear sample per step (~1000).  The sums are
// TIGRE's Riemann sum re-associated; only fp32 rounding differs (tools/fan_event_model.py: 2e-7 relative L2 at
// epochs of 256 steps).
//
// "Row v is above boundary k at step i" is DEFINED as  v >= ceil(fma(a_k, rcp(i), -b)),  a_k = (k - sz) n / vz,
// b = dz0 / vz.  It is monotone in i and in k by construction (correctly rounded operations are monotone),
// so the crossings enumerated per boundary and the pair found at an epoch's end always agree; that it differs
// from the exact predicate by ~1e-5 slices is harmless, the interpolant being continuous across a boundary.
//
// Kernel shape (sm_100a; a gather plus prefix sums, no tensor cores).  One WARP owns one detector column
// (angle, u, up to FAN_MAXROWS rows).  A lane owns four consecutive slices (one 16-byte load per bilinear
// corner: the resident volume is z-fastest) of one run; the lanes of a warp are packed, per epoch, with the
// z ranges the runs of the column look at (greedy; what does not fit 32 lanes goes to a second pass).  Per step
// a lane does 4 predicated LDG.128 (no load while the in-plane cell is unchanged), 8 packed FMAs for G, 4 packed
// accumulates and 2 STS.128 of its running (A, C) into a ring of CW steps.  After every CW steps the lanes
// enumerate the rows that crossed their boundaries (closed form: a row interval per boundary), and the warp
// serves those events 32 at a time from the ring; at the end of an epoch every row reads its pair once.
// The warps are independent (no CTA barrier); by default they are persistent and take their columns from a
// counter (k_ax_fan_queued), the plain grid launch (k_ax_fan) serves column shares and chord-split launches.
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdlib>

#include "pax_internal.h"

#define FAN_MAXSEG 8            // runs (pieces of runs) packed into the lanes of one pass
#define FAN_QCAP 96             // crossing events served per round
#define FAN_MAXROWS 1024        // detector rows one warp walks
#define FAN_EPMAX 256           // most steps per epoch
#define FAN_BW 16               // columns per block of the queued launch

template <int CW>
struct __align__(16) FanSmem {
    float ring[CW * 256 + 8];                  // 4 guard floats, then [step][A of the 128 lane slices | C of them]
    float4 w[FAN_MAXSEG][CW + 1];              // per lattice (distinct n): three bilinear weights + cell offset of the
                                               // steps w0+1 .. w0+CW of the window (slot CW: the epoch's first step)
    float4 c_geo[FAN_MAXSEG];                  // per lattice: {ddx, ddy, fraction x, fraction y} of the epoch's origin
    int4 c_cell[FAN_MAXSEG];                   // per lattice: {origin cell x, y, first step, last step} of the epoch
    float4 s_f[FAN_MAXSEG];                    // per run: {n / vz, dz per step of row 0, its increment per row, z_ref of row 0}
    int4 s_i[FAN_MAXSEG];                      // per run: {first lane, lanes, first quad, last step of the epoch}
    float s_zrefv[FAN_MAXSEG];                 // per run: increment of z_ref per row
    int s_ra[FAN_MAXSEG], s_rb[FAN_MAXSEG], s_cls[FAN_MAXSEG], c_n[FAN_MAXSEG];
    float rtab[36];                            // rcp(step) at the epoch's window boundaries (the crossing predicate)
    unsigned queue[FAN_QCAP];
    unsigned char lane_seg[32];
    unsigned starts[FAN_MAXROWS / 32];         // bit r of word w: row 32w+r starts a run of equal n
    int run_n[32];                             // n of the column's first 32 runs (more: recomputed on the fly)
    int sched[4];                              // k_ax_fan_queued: {row band, chord part, column, angle} of the column at hand
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

// range of the chord parameter lambda in [0,1] with s + lambda*d inside [lo, hi)
__device__ __forceinline__ void fan_slab(double s, double d, double lo, double hi, double &t0, double &t1, bool &hit)
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

__device__ __forceinline__ void fan_ld4_if(float4 &v, const void *p, bool take)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "@p ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%5];\n\t}"
                 : "+f"(v.x), "+f"(v.y), "+f"(v.z), "+f"(v.w)
                 : "r"((int)take), "l"(p));
}
// base + (unsigned) byte offset in one instruction (IMAD.WIDE.U32)
__device__ __forceinline__ const char *fan_at(const char *base, unsigned off)
{
    unsigned long long r;
    asm("mad.wide.u32 %0, %1, 1, %2;" : "=l"(r) : "r"(off), "l"((unsigned long long)base));
    return reinterpret_cast<const char *>(r);
}
__device__ __forceinline__ float2 fan_mul2(float w, float a, float b)
{
    return __fmul2_rn(make_float2(w, w), make_float2(a, b));
}
__device__ __forceinline__ float2 fan_fma2(float w, float a, float b, float2 acc)
{
    return __ffma2_rn(make_float2(w, w), make_float2(a, b), acc);
}

// THE crossing predicate: first row that is above boundary k at the step whose reciprocal is r
// (a = (k - sz) n / vz, b = dz0 / vz).  Every use goes through this function, bit for bit.
__device__ __forceinline__ int fan_cv(float a, float r, float b)
{
    return __float2int_ru(__fmaf_rn(a, r, -b));            // F2I saturates: monotone for any finite input
}
__device__ __forceinline__ float fan_ak(int k, float szf, float nvz)
{
    return __fmul_rn(__fsub_rn((float)k, szf), nvz);
}
__device__ __forceinline__ float fan_rcp_step(int i)
{
    return i > 0 ? __frcp_rn((float)i) : 2.0f;             // decreasing in i, finite at the source
}

// One detector column (angle ia of the launch, column u, row band `by`, chord part `bz` of a.fan_split), walked by
// the calling warp alone: `sm` and `rsum` are the warp's own shared memory.
// (QUEUED: band and part come from the warp's shared memory where they are needed, not through registers)
template <int CW, int MODE, bool QUEUED>
__device__ __forceinline__ void fan_column(const AxArgs &a, FanSmem<CW> &sm, float *rsum, const int lane, const int u,
                                           const int ia)
{
    auto part_of = [&]() { return QUEUED ? ((volatile int *)sm.sched)[1] : (int)blockIdx.z; };
    auto col_of = [&]() { return QUEUED ? ((volatile int *)sm.sched)[2] : u; };
    auto ang_of = [&]() { return QUEUED ? ((volatile int *)sm.sched)[3] : ia; };
    const int V0 = (QUEUED ? ((volatile int *)sm.sched)[0] : (int)blockIdx.y) * a.fan_rows;
    const int nRows = min(a.fan_rows, a.nv - V0);
    const AngleDev &A = a.angles[a.first + ia];
    const double sx = A.sx, sy = A.sy, sz = A.sz;
    const int kmin = (int)a.loz, kmax = (int)a.hiz;               // slices a sample can touch: [kmin, kmax]
    const int xmaxc = (int)a.hix - 1, ymaxc = (int)a.hiy - 1;    // last in-plane cell whose +1 neighbour exists
    const float szf = (float)sz;
    const double dx = A.ox + u * A.ux - sx, dy = A.oy + u * A.uy - sy;
    const double dxy2 = dx * dx + dy * dy;
    const double racc = 1.0 / a.acc;
    const double vz = A.vz;
    const double dz0 = A.oz + (double)V0 * vz - sz;               // P_z - S_z of the band's row 0; +vz per row
    const float bq = (float)(dz0 / vz);                           // the predicate's b
    const double rvz = 1.0 / vz;
    const int EP = a.fan_epoch;
    const size_t dx4 = (size_t)a.nzp * 4, dy4 = (size_t)a.nxp * a.nzp * 4;
    float r2 = 0.f, mymax = -INFINITY;
    unsigned my_ns = 0;

    // ---- runs of rows with equal n (TIGRE's lattice); clear the rows' sums
    int nlo = INT_MAX, nhi = 0;
    {
        int last = -2, nruns = 0;
        for (int r0 = 0; r0 < nRows; r0 += 32) {
            const bool in = r0 + lane < nRows;
            const double dz = dz0 + (double)(r0 + lane) * vz;
            const int n = in ? (int)ceil(sqrt(dxy2 + dz * dz) * racc) : -1;
            if (in) { nlo = min(nlo, n); nhi = max(nhi, n); rsum[r0 + lane] = 0.f; }
            int prev = __shfl_up_sync(0xffffffffu, n, 1);
            if (lane == 0) prev = last;
            const bool st = in && n != prev;
            const unsigned m = __ballot_sync(0xffffffffu, st);
            if (lane == 0) sm.starts[r0 >> 5] = m;
            const int ord = nruns + __popc(m & ((1u << lane) - 1u));
            if (st && ord < 32) sm.run_n[ord] = n;
            nruns += __popc(m);
            last = __shfl_sync(0xffffffffu, n, 31);
        }
        nlo = __reduce_min_sync(0xffffffffu, nlo);
        nhi = __reduce_max_sync(0xffffffffu, nhi);
        __syncwarp();
    }

    // ---- the chord's part inside the volume in x and y, as a range of lambda = i / n (all runs share it)
    double lam0 = 0.0, lam1 = 1.0;
    bool hit = true;
    fan_slab(sx, dx, 0.0, (double)a.hix, lam0, lam1, hit);
    fan_slab(sy, dy, 0.0, (double)a.hiy, lam0, lam1, hit);

    if (hit && lam0 <= lam1) {
        const int Emin = max((int)floor(lam0 * (double)nlo), 0), Emax = (int)ceil(lam1 * (double)nhi);
        // a launch too small to fill the GPU more than a few times over is cut into `fan_split` parts along the
        // chord (whole epochs: their sums are independent), one warp each; k_fan_finish adds the parts up
        const int nep = (Emax - Emin) / EP + 1;
        const int bz = part_of();
        const int mA = nep * bz / a.fan_split, mB = nep * (bz + 1) / a.fan_split;
#pragma unroll 1
        for (int e0 = Emin + mA * EP; e0 < Emin + mB * EP; e0 += EP) {
            const int e1 = min(e0 + EP - 1, Emax);
            const int ref = e0 + (EP >> 1);
            const int nwin = (e1 - e0 + CW) / CW;
            if (lane <= nwin) sm.rtab[lane] = fan_rcp_step(e0 + lane * CW);   // nwin <= 32 (pax_fan_configure)
            int cur = 0, ri = 0;                                  // row, and the ordinal of the run it is in
#pragma unroll 1
            while (cur < nRows) {
                // ================= pack runs (or pieces of runs) into the lanes of one pass (uniform over the warp)
                int nseg = 0, ncls = 0, used = 0;
#pragma unroll 1
                while (cur < nRows && nseg < FAN_MAXSEG && used < 31) {
                    const int e = fan_next_start(sm.starts, cur, nRows);
                    int nrun;
                    if (ri < 32) nrun = sm.run_n[ri];
                    else {
                        const double dzc = dz0 + (double)cur * vz;
                        nrun = (int)ceil(sqrt(dxy2 + dzc * dzc) * racc);
                    }
                    const int I0 = max((int)ceil(lam0 * (double)nrun), 0), I1 = (int)floor(lam1 * (double)nrun);
                    const int ja = max(I0, e0), jb = min(I1, e1);
                    if (ja > jb) { cur = e; ++ri; continue; }     // the run has no sample in this epoch
                    const double rn = 1.0 / (double)nrun;
                    const float s0f = (float)(dz0 * rn), zvrf = (float)(vz * rn);
                    const float fa = (float)ja, fb = (float)jb;
                    const float sA = fmaf((float)cur, zvrf, s0f);
                    const float zlo = szf + fminf(fa * sA, fb * sA);
                    const int ka = max(kmin, (int)floorf(fmaxf(zlo - 4e-3f, -1.0e6f)));
                    if (ka > kmax) { cur = nRows; break; }        // this row and all above it look over the volume
                    const int qa = ka >> 2, avail = 32 - used;
                    // last row whose highest slice (+1 neighbour) still falls into the lanes left
                    const float lim = (float)(4 * (qa + avail) - 1) - 8e-3f - szf;
                    float rmaxf = (float)(e - 1);
                    if (zvrf > 0.f) {
                        const float rz = __frcp_rn(zvrf);
                        if (fb > 0.f) rmaxf = fminf(rmaxf, (__fdividef(lim, fb) - s0f) * rz - 2e-2f);
                        if (fa > 0.f) rmaxf = fminf(rmaxf, (__fdividef(lim, fa) - s0f) * rz - 2e-2f);
                    }
                    int rb = min(max((int)floorf(fmaxf(rmaxf, -1.0f)) + 1, cur + 1), e);
                    int nq;
                    for (;;) {
                        const float sB = fmaf((float)(rb - 1), zvrf, s0f);
                        const float zhi = szf + fmaxf(fa * sB, fb * sB);
                        const int kb = max(min(kmax, (int)floorf(fminf(zhi + 4e-3f, 1.0e6f)) + 1), ka);
                        nq = (kb >> 2) - qa + 1;
                        if (nq <= avail || rb <= cur + 1) break;
                        --rb;
                    }
                    if (nq > avail) {
                        if (used > 0) break;                      // start a new pass with this run
                        if (lane == 0 && a.fault) atomicExch(a.fault, 1);   // a single row does not fit: see pax_fan_configure
                        nq = avail;
                    }
                    // the run's lattice: runs with the same n (below and above the source plane) share their steps
                    int cls = 0;
                    while (cls < ncls && sm.c_n[cls] != nrun) ++cls;
                    if (lane == 0) {
                        const int g = nseg;
                        sm.s_ra[g] = cur; sm.s_rb[g] = rb; sm.s_cls[g] = cls;
                        sm.s_i[g] = make_int4(used, nq, qa, jb);
                        sm.s_f[g] = make_float4((float)((double)nrun * rvz), s0f, zvrf, (float)(sz + (double)ref * dz0 * rn));
                        sm.s_zrefv[g] = (float)((double)ref * vz * rn);
                        if (cls == ncls) {
                            // in-plane origin of the epoch, split in double into an integer cell and a fraction; steps
                            // add t*d to the fraction in fp32 (t < EP): a position is good to a few 1e-6 voxels
                            const double ddx = dx * rn, ddy = dy * rn;
                            const double ex = sx + (double)e0 * ddx, ey = sy + (double)e0 * ddy;
                            const double exi = floor(ex), eyi = floor(ey);
                            sm.c_n[cls] = nrun;
                            sm.c_geo[cls] = make_float4((float)ddx, (float)ddy, (float)(ex - exi), (float)(ey - eyi));
                            sm.c_cell[cls] = make_int4((int)exi, (int)eyi, ja, jb);
                        }
                    }
                    __syncwarp();
                    if (cls == ncls) ++ncls;
                    used += nq;
                    ++nseg;
                    cur = rb;
                    if (rb == e) ++ri;
                }
                if (nseg == 0) continue;

                // ================= the pass
                int myseg = -1;
#pragma unroll 1
                for (int g = 0; g < nseg; ++g) {
                    const int4 si = sm.s_i[g];
                    if (lane >= si.x && lane < si.x + si.y) myseg = g;
                }
                const bool on = myseg >= 0;
                const int sg = on ? myseg : 0;
                const int4 mysi = sm.s_i[sg];
                const int quad = mysi.z + lane - mysi.x;
                const int segra = sm.s_ra[sg], segrb = sm.s_rb[sg];
                const int segjb = mysi.w;                         // last step of my run in this epoch: no sample, hence
                                                                  // no crossing to serve, after it
                sm.lane_seg[lane] = (unsigned char)sg;
                float ak0, ak1, ak2, ak3;
                {
                    const float nvz = sm.s_f[sg].x;
                    ak0 = fan_ak(4 * quad + 0, szf, nvz); ak1 = fan_ak(4 * quad + 1, szf, nvz);
                    ak2 = fan_ak(4 * quad + 2, szf, nvz); ak3 = fan_ak(4 * quad + 3, szf, nvz);
                }
                int cs0, cs1, cs2, cs3;                           // rows above my boundaries so far: [cs, segrb)
                {
                    const float r0 = sm.rtab[0];
                    cs0 = min(max(fan_cv(ak0, r0, bq), segra), segrb); cs1 = min(max(fan_cv(ak1, r0, bq), segra), segrb);
                    cs2 = min(max(fan_cv(ak2, r0, bq), segra), segrb); cs3 = min(max(fan_cv(ak3, r0, bq), segra), segrb);
                }
                // the four corner columns of in-plane cell (0,0), at my quad
                const char *b00 = reinterpret_cast<const char *>(a.vol0 + 4 * (on ? quad : 0));
                const char *b01 = b00 + dx4, *b10 = b00 + dy4, *b11 = b10 + dx4;
                const float4 *wp = sm.w[sm.s_cls[sg]];
                float *const ring = sm.ring + 4;
                float4 Aq = make_float4(0.f, 0.f, 0.f, 0.f), Cq = Aq;     // running (A, C) of my four slices
                float4 c00 = make_float4(0.f, 0.f, 0.f, 0.f), c01 = c00, c10 = c00, c11 = c00;
                // event words of my four slices: row | slot << 16 | edge flags (no neighbour slice below / above
                // inside the run's lanes: it is a zero slice) | direction
                const unsigned evw = ((unsigned)(lane * 4) << 16);
                const unsigned elo = (lane == mysi.x) ? (1u << 23) : 0u, ehi = (lane == mysi.x + mysi.y - 1) ? (1u << 24) : 0u;

                // weights and cells of `count` steps from `first` on into slots slot0.., all lattices of the pass
                auto fill_w = [&](int first, int count, int slot0) {
                    for (int q = lane; q < ncls * count; q += 32) {
                        const int g = q / count, t = q - g * count, i = first + t;
                        const int4 cc = sm.c_cell[g];
                        // a step outside the run's range reads the all-zero rim column (0,0) with weight one
                        float4 w = make_float4(1.f, 0.f, 0.f, __uint_as_float(0u));
                        if (i >= cc.z && i <= cc.w) {
                            const float4 cg = sm.c_geo[g];
                            const float tf = (float)(i - e0);
                            const float px = fmaf(tf, cg.x, cg.z), py = fmaf(tf, cg.y, cg.w);
                            const int ix = __float2int_rd(px), iy = __float2int_rd(py);
                            const int cx = min(max(cc.x + ix, 0), xmaxc);
                            const int cy = min(max(cc.y + iy, 0), ymaxc);
                            // fraction against the CLAMPED cell: a sample the double clip admitted may sit an ulp
                            // outside in fp32, and must then weigh the rim, not the next column
                            const float fx = fminf(fmaxf(px - (float)(cx - cc.x), 0.f), 1.f);
                            const float fy = fminf(fmaxf(py - (float)(cy - cc.y), 0.f), 1.f);
                            const unsigned off = (unsigned)((cy * a.nxp + cx) * a.nzp) * 4u;
                            const float gx = 1.f - fx, gy = 1.f - fy;
                            w = make_float4(gx * gy, fx * gy, gx * fy, __uint_as_float(off));
                        }
                        sm.w[g][slot0 + t] = w;
                    }
                };
                fill_w(e0, 1, CW);
                fill_w(e0 + 1, CW, 0);
                __syncwarp();
                float4 wv = wp[CW];
                unsigned off = __float_as_uint(wv.w);
                fan_ld4_if(c00, fan_at(b00, off), on); fan_ld4_if(c01, fan_at(b01, off), on);
                fan_ld4_if(c10, fan_at(b10, off), on); fan_ld4_if(c11, fan_at(b11, off), on);
                float jf = (float)(e0 - ref);
#pragma unroll 1
                for (int m = 0; m < nwin; ++m) {
                    const int w0 = e0 + m * CW;
                    if (m) { fill_w(w0 + 1, CW, 0); __syncwarp(); }
                    // ---- CW steps: running (A, C) of my four slices into the ring
                    const int iend = min(w0 + CW, e1);
                    int lo0 = 0, lo1 = 0, lo2 = 0, lo3 = 0, n0 = 0, n1 = 0, n2 = 0, n3 = 0, mine = 0, incl = 0;
                    unsigned up = 0;
#pragma unroll
                    for (int t = 0; t < CW; ++t) {
                        const float4 wn = wp[t];                   // the next step
                        const unsigned offn = __float_as_uint(wn.w);
                        const bool chg = on && offn != off;
                        const float w11 = 1.f - wv.x - wv.y - wv.z;              // the weights sum to one
                        float2 v0 = fan_mul2(wv.x, c00.x, c00.y), v1 = fan_mul2(wv.x, c00.z, c00.w);
                        v0 = fan_fma2(wv.y, c01.x, c01.y, v0); v1 = fan_fma2(wv.y, c01.z, c01.w, v1);
                        v0 = fan_fma2(wv.z, c10.x, c10.y, v0); v1 = fan_fma2(wv.z, c10.z, c10.w, v1);
                        v0 = fan_fma2(w11, c11.x, c11.y, v0); v1 = fan_fma2(w11, c11.z, c11.w, v1);
                        // the columns of the next step, as soon as this one has consumed the registers
                        fan_ld4_if(c00, fan_at(b00, offn), chg); fan_ld4_if(c01, fan_at(b01, offn), chg);
                        fan_ld4_if(c10, fan_at(b10, offn), chg); fan_ld4_if(c11, fan_at(b11, offn), chg);
                        // (scalar accumulates: packed ones would pin the sums to register pairs that the 16-byte
                        // stores then have to gather with eight moves)
                        Aq.x += v0.x; Aq.y += v0.y; Aq.z += v1.x; Aq.w += v1.y;
                        Cq.x = fmaf(jf, v0.x, Cq.x); Cq.y = fmaf(jf, v0.y, Cq.y);
                        Cq.z = fmaf(jf, v1.x, Cq.z); Cq.w = fmaf(jf, v1.y, Cq.w);
                        jf += 1.f;
                        // (idle lanes store too: predicating the stores measured 4 % slower)
                        *reinterpret_cast<float4 *>(ring + t * 256 + 4 * lane) = Aq;
                        *reinterpret_cast<float4 *>(ring + t * 256 + 128 + 4 * lane) = Cq;
                        off = offn;
                        wv = wn;
                    }
                    __syncwarp();
                    // ---- crossings at steps w0+1 .. iend (prefix through the step before, i.e. ring row t-1)
                    {
                        const int il = min(iend, segjb);
                        int c0 = cs0, c1 = cs1, c2 = cs2, c3 = cs3;
                        if (on && il > w0) {
                            // (the predicate's reciprocal: tabulated at the window boundaries, else -- the run's
                            // last window -- computed by the same function)
                            const float re = (il == w0 + CW) ? sm.rtab[m + 1] : fan_rcp_step(il);
                            c0 = min(max(fan_cv(ak0, re, bq), segra), segrb); c1 = min(max(fan_cv(ak1, re, bq), segra), segrb);
                            c2 = min(max(fan_cv(ak2, re, bq), segra), segrb); c3 = min(max(fan_cv(ak3, re, bq), segra), segrb);
                        }
                        lo0 = min(c0, cs0); n0 = abs(c0 - cs0); up |= (c0 < cs0) ? 0x80000000u : 0u;
                        lo1 = min(c1, cs1); n1 = abs(c1 - cs1); up |= (c1 < cs1) ? 0x40000000u : 0u;
                        lo2 = min(c2, cs2); n2 = abs(c2 - cs2); up |= (c2 < cs2) ? 0x20000000u : 0u;
                        lo3 = min(c3, cs3); n3 = abs(c3 - cs3); up |= (c3 < cs3) ? 0x10000000u : 0u;
                        cs0 = c0; cs1 = c1; cs2 = c2; cs3 = c3;
                        mine = n0 + n1 + n2 + n3;
                        incl = mine;
                    }
#pragma unroll
                    for (int d = 1; d < 32; d <<= 1) {
                        const int o = __shfl_up_sync(0xffffffffu, incl, d);
                        if (lane >= d) incl += o;
                    }
                    const int total = __shfl_sync(0xffffffffu, incl, 31);
#pragma unroll 1
                    for (int qb = 0; qb < total; qb += FAN_QCAP) {
                        // my events into the queue
                        {
                            const int i0 = incl - mine - qb, i1 = i0 + n0, i2 = i1 + n1, i3 = i2 + n2;
                            const unsigned e0w = (unsigned)lo0 | evw | elo | (up & 0x80000000u);
                            const unsigned e1w = (unsigned)lo1 | (evw + (1u << 16)) | ((up << 1) & 0x80000000u);
                            const unsigned e2w = (unsigned)lo2 | (evw + (2u << 16)) | ((up << 2) & 0x80000000u);
                            const unsigned e3w = (unsigned)lo3 | (evw + (3u << 16)) | ehi | ((up << 3) & 0x80000000u);
                            const int nmax = max(max(n0, n1), max(n2, n3));
                            for (int r = 0; r < nmax; ++r) {
                                if (r < n0 && (unsigned)(i0 + r) < FAN_QCAP) sm.queue[i0 + r] = e0w + r;
                                if (r < n1 && (unsigned)(i1 + r) < FAN_QCAP) sm.queue[i1 + r] = e1w + r;
                                if (r < n2 && (unsigned)(i2 + r) < FAN_QCAP) sm.queue[i2 + r] = e2w + r;
                                if (r < n3 && (unsigned)(i3 + r) < FAN_QCAP) sm.queue[i3 + r] = e3w + r;
                            }
                        }
                        __syncwarp();
                        const int nev = min(FAN_QCAP, total - qb);
                        for (int q = lane; q < nev; q += 32) {
                            const unsigned ev = sm.queue[q];
                            const int v = (int)(ev & 0xffffu), slot = (int)((ev >> 16) & 0x7fu);
                            const int g = sm.lane_seg[slot >> 2];
                            const int4 si = sm.s_i[g];
                            const float4 sf = sm.s_f[g];
                            const float fk = (float)(4 * (si.z - si.x) + slot), fv = (float)v;
                            const float ak = __fmul_rn(__fsub_rn(fk, szf), sf.x);
                            // first step of the window at which the row is on the new side of the boundary.  The
                            // closed form may differ from the predicate's by a step when z hits the boundary to
                            // ~1e-5: harmless, the interpolant is continuous there (only WHICH window serves the
                            // crossing has to agree with the pair found at the epoch's end, and that is the predicate's)
                            const int th = min(iend, si.w) - w0;
                            const int tl = min(max(__float2int_ru(__fdividef(ak, fv + bq)) - w0, 1), th);
                            const float *row = ring + (tl - 1) * 256 + slot;
                            float am = row[-1], a0_ = row[0], ap = row[1], cm = row[127], c0_ = row[128], cp = row[129];
                            if (ev & (1u << 23)) { am = 0.f; cm = 0.f; }
                            if (ev & (1u << 24)) { ap = 0.f; cp = 0.f; }
                            const float d2a = fmaf(-2.f, a0_, am + ap), d2c = fmaf(-2.f, c0_, cm + cp);
                            const float sv = fmaf(fv, sf.z, sf.y);
                            const float tau = fmaf(fv, sm.s_zrefv[g], sf.w) - fk;
                            const float val = fmaf(tau, d2a, sv * d2c);
                            atomicAdd(rsum + v, (int)ev < 0 ? -val : val);
                        }
                        __syncwarp();
                    }
                }
                // ---- end of the epoch: every row reads its slice pair once
                {
                    const float *row = ring + (e1 - (e0 + (nwin - 1) * CW)) * 256;
#pragma unroll 1
                    for (int g = 0; g < nseg; ++g) {
                        const int4 si = sm.s_i[g];
                        const float4 sf = sm.s_f[g];
                        const float zrv = sm.s_zrefv[g];
                        // the run's last step of the epoch; the sums do not change after it (G = 0)
                        const float re = fan_rcp_step(si.w), fe = (float)si.w;
                        const int k0 = 4 * si.z;
                        for (int v = sm.s_ra[g] + lane; v < sm.s_rb[g]; v += 32) {
                            const float fv = (float)v;
                            const float sv = fmaf(fv, sf.z, sf.y);
                            int K = __float2int_rd(fmaf(fe, sv, szf));
                            // the pair the PREDICATE puts the row in (an ulp of z may disagree with it)
                            if (v >= fan_cv(fan_ak(K + 1, szf, sf.x), re, bq)) ++K;
                            else if (v < fan_cv(fan_ak(K, szf, sf.x), re, bq)) --K;
                            const int kap = K - k0;
                            if (kap >= 0 && kap + 1 < 4 * si.y) {
                                const float *pr = row + 4 * si.x + kap;
                                const float aK = pr[0], aK1 = pr[1], cK = pr[128], cK1 = pr[129];
                                const float tau = fmaf(fv, zrv, sf.w) - (float)K;
                                rsum[v] += fmaf(sv, cK1 - cK, fmaf(tau, aK1 - aK, aK));
                            }
                        }
                    }
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();

    if (a.fan_split > 1) {                                        // raw sums of my part; k_fan_finish does the rest
        float *dst = a.fan_scratch + (((size_t)part_of() * a.fan_count + ang_of()) * a.nv + V0) * a.nu + col_of();
        for (int r = lane; r < nRows; r += 32) dst[(size_t)r * a.nu] = rsum[r];
        return;
    }
    // ---- epilogue, run by run (the step length in mm depends on n).  The warps of the CTA own adjacent columns and
    // fill one 32-byte sector of every row between them; L2 merges the partial writes.
    const int ue = col_of(), iae = ang_of();
    int cur = 0;
    while (cur < nRows) {
        const int e = fan_next_start(sm.starts, cur, nRows);
        const double dzc = dz0 + (double)cur * vz;
        const int nrun = (int)ceil(sqrt(dxy2 + dzc * dzc) * racc);
        const double rn = 1.0 / (double)nrun;
        const double ddx = dx * rn, ddy = dy * rn;
        double t0 = 0.0, t1 = (double)nrun;
        bool hitr = true;
        if (a.nsamples) {
            fan_slab(sx, ddx, 0.0, (double)a.hix, t0, t1, hitr);
            fan_slab(sy, ddy, 0.0, (double)a.hiy, t0, t1, hitr);
        }
        for (int r = cur + lane; r < e; r += 32) {
            const int v = V0 + r;
            const size_t o = ((size_t)iae * a.nv + v) * a.nu + ue;
            const double mx = ddx * a.dX, my = ddy * a.dY, mz = (dz0 + (double)r * vz) * rn * a.dZ;
            const float p0 = rsum[r] * (float)sqrt(mx * mx + my * my + mz * mz);
            if (a.nsamples) {                                     // exact count of the samples inside the open box
                double tz0 = t0, tz1 = t1;
                bool hz = hitr;
                fan_slab(sz, (dz0 + (double)r * vz) * rn, (double)a.loz, (double)a.hiz, tz0, tz1, hz);
                if (hz && tz0 <= tz1) {
                    const int i0 = (int)ceil(tz0), i1 = (int)floor(tz1);
                    if (i0 <= i1) my_ns += (unsigned)(i1 - i0 + 1);
                }
            }
            if (MODE == PAX_AX_PLAIN) {
                float val = a.a0 * p0;
                if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, ue, v, a.whx, a.why, a.whz), val);
                if (a.accumulate) val += a.out[o];
                a.out[o] = val;
                mymax = fmaxf(mymax, val);
            } else {
                const float L = pax_ray_chord(A, ue, v, a.whx, a.why, a.whz);
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

// Launch order of the (column tile, angle) pairs of a launch: angle fastest, then the column tile.  The angles of a
// launch are a few degrees apart, so at any time the resident warps walk chords through one narrow bow-tie of the
// volume, which stays in L2 while the bow-tie sweeps across (column-tile-fastest order re-read the volume from
// DRAM once per angle; the 656 angles of C2 in one launch take 146 ms, 20 at a time 153 ms).  Column tiles in
// order of decreasing chord length (a.fan_order): a warp walks its column alone, so a long chord that starts late
// would keep the launch waiting with most of the SMs idle.
__device__ __forceinline__ void fan_decode(int p, int count, int &ti, int &ia)
{
    ti = p / count;
    ia = p - ti * count;
}

template <int CW, int NWARPS, int MINB, int MODE>
__global__ void __launch_bounds__(NWARPS * 32, MINB) k_ax_fan(const AxArgs a)
{
    extern __shared__ __align__(16) unsigned char fan_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_pad = (a.fan_rows + 3) & ~3;
    const size_t per_warp = sizeof(FanSmem<CW>) + (size_t)rows_pad * sizeof(float);
    FanSmem<CW> &sm = *reinterpret_cast<FanSmem<CW> *>(fan_smem_raw + (size_t)warp * per_warp);
    float *rsum = reinterpret_cast<float *>(&sm + 1);             // running line integral of every row of the band

    int ti, ia;
    fan_decode((int)blockIdx.x, a.fan_count, ti, ia);
    const int u = a.fan_order[ti] * NWARPS + warp;
    if (u >= a.nu) return;                                        // warps are independent: no CTA barriers
    fan_column<CW, MODE, false>(a, sm, rsum, lane, u, ia);
}

// The same walk with the columns handed out at run time, one per WARP, from one counter over the launch's columns
// in the order above (blocks of FAN_BW adjacent columns of one angle): a warp that is done takes the next column,
// so no warp slot waits for the other warps of a CTA and the tail of the launch is balanced column by column.
// (Per-SM counters with stealing, meant to keep the 16 warps of an SM on neighbouring chords for L1, measured
// slower: the L1 left beside the tables is smaller than one step's corner columns of 16 warps.)
template <int CW, int NWARPS, int MINB, int MODE>
__global__ void __launch_bounds__(NWARPS * 32, MINB) k_ax_fan_queued(const AxArgs a)
{
    extern __shared__ __align__(16) unsigned char fan_smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int rows_pad = (a.fan_rows + 3) & ~3;
    const size_t per_warp = sizeof(FanSmem<CW>) + (size_t)rows_pad * sizeof(float);
    FanSmem<CW> &sm = *reinterpret_cast<FanSmem<CW> *>(fan_smem_raw + (size_t)warp * per_warp);
    float *rsum = reinterpret_cast<float *>(&sm + 1);
    constexpr int bw = FAN_BW;
    const int total = a.fan_nblock * bw;
#pragma unroll 1
    for (;;) {
        int idx = 0;
        if (lane == 0) idx = atomicAdd(a.fan_queue, 1);
        idx = __shfl_sync(0xffffffffu, idx, 0);
        if (idx >= total) break;
        const int blk = idx / bw;                                 // the block, in launch order
        const int per = a.fan_ntiles * a.fan_count, band = blk / per;
        int ti, ia;
        fan_decode(blk - band * per, a.fan_count, ti, ia);
        const int u = a.fan_order[ti] * bw + (idx - blk * bw);
        if (lane == 0) {                                          // row band fastest, then the part of the chord
            sm.sched[0] = band % a.fan_nband; sm.sched[1] = band / a.fan_nband;
            sm.sched[2] = u; sm.sched[3] = ia;
        }
        __syncwarp();
        if (u < a.nu) fan_column<CW, MODE, true>(a, sm, rsum, lane, u, ia);
        __syncwarp();
    }
}

// The epilogue of a split launch: add the parts of every pixel up, scale by the step length of the pixel's ray
// (same arithmetic as above) and apply the PLAIN / RESIDUAL epilogue.
template <int MODE>
__global__ void __launch_bounds__(256) k_fan_finish(const AxArgs a, int nwarps)
{
    // one thread per pixel of the launch's column tiles (all of the detector, or this plan's share of it)
    const size_t npix = (size_t)a.nv * a.nu, nall = npix * a.fan_count;
    const size_t per_angle = (size_t)a.fan_ntiles * nwarps * a.nv, n = per_angle * a.fan_count;
    const double racc = 1.0 / a.acc;
    float r2 = 0.f, mymax = -INFINITY;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (size_t)gridDim.x * blockDim.x) {
        const int ia = (int)(t / per_angle);
        const size_t w = t % per_angle;
        const int j = (int)(w % nwarps), v = (int)((w / nwarps) % a.nv), ti = (int)(w / ((size_t)nwarps * a.nv));
        const int u = a.fan_order[ti] * nwarps + j;
        if (u >= a.nu) continue;
        const size_t e = ((size_t)ia * a.nv + v) * a.nu + u;
        float sum = 0.f;
        for (int k = 0; k < a.fan_split; ++k) sum += a.fan_scratch[(size_t)k * nall + e];
        const AngleDev &A = a.angles[a.first + ia];
        const double dx = A.ox + u * A.ux - A.sx, dy = A.oy + u * A.uy - A.sy, dz = A.oz + (double)v * A.vz - A.sz;
        const int nrun = (int)ceil(sqrt(dx * dx + dy * dy + dz * dz) * racc);
        const double rn = 1.0 / (double)nrun;
        const double mx = dx * rn * a.dX, my = dy * rn * a.dY, mz = dz * rn * a.dZ;
        const float p0 = sum * (float)sqrt(mx * mx + my * my + mz * mz);
        if (MODE == PAX_AX_PLAIN) {
            float val = a.a0 * p0;
            if (a.c != 0.f) val = fmaf(a.c, pax_ray_chord(A, u, v, a.whx, a.why, a.whz), val);
            if (a.accumulate) val += a.out[e];
            a.out[e] = val;
            mymax = fmaxf(mymax, val);
        } else {
            const float L = pax_ray_chord(A, u, v, a.whx, a.why, a.whz);
            const float wt = L > a.wthr ? 1.f / L : 0.f;
            const float r_ = __ldg(a.b + e) - p0;
            r2 = fmaf(r_, r_, r2);
            a.out[e] = wt * r_;
        }
    }
    if (MODE == PAX_AX_PLAIN && a.maxout) {
        for (int o = 16; o; o >>= 1) mymax = fmaxf(mymax, __shfl_xor_sync(0xffffffffu, mymax, o));
        if ((threadIdx.x & 31) == 0 && mymax > -INFINITY) {
            const unsigned b = __float_as_uint(mymax);
            atomicMax(a.maxout, (b & 0x80000000u) ? ~b : (b | 0x80000000u));
        }
    }
    if (MODE == PAX_AX_RESIDUAL && a.sumsq) {
        for (int o = 16; o; o >>= 1) r2 += __shfl_xor_sync(0xffffffffu, r2, o);
        if ((threadIdx.x & 31) == 0) atomicAdd(a.sumsq, (double)r2);
    }
}

// ------------------------------------------------------------------------------ host side
// Parts a launch of `count` angles is cut into along the chords: none while its warps fill the GPU's 16 warp slots
// per SM three times over, else as many as that takes (at most 4; a chord has 6-9 epochs at C2).
static int fan_parts(const PaxPlan *p, int count)
{
    const char *e = getenv("PAX_FAN_SPLIT");                    // tuning hook / tests: force 1..4 parts
    if (e && atoi(e) >= 1 && atoi(e) <= 4) return atoi(e);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const long warps = (long)p->fan_ntiles * p->fan_warps * count * pax_cdiv(p->geo.nDetecV, p->fan_rows);
    const long want = 3L * sms * 16;
    if (warps <= 0 || warps >= want) return 1;
    return (int)std::min(4L, (want + warps - 1) / warps);
}

extern "C" size_t pax_ax_fan_scratch_bytes(const PaxPlan *p, int count)
{
    if (!p || count < 1 || pax_plan_projector(p) != PAX_PROJECTOR_FAN || p->fan_epoch < 1) return 0;
    const int parts = fan_parts(p, count);
    return parts > 1 ? (size_t)parts * count * p->geo.nDetecV * p->geo.nDetecU * sizeof(float) : 0;
}

template <int CW, int NWARPS, int MINB, int MODE>
static int fan_launch(const AxArgs &a, int count, cudaStream_t st)
{
    const size_t per_warp = sizeof(FanSmem<CW>) + (size_t)((a.fan_rows + 3) & ~3) * sizeof(float);
    const size_t smem = per_warp * NWARPS;
    PAX_REQUIRE(smem <= 227 * 1024, "pax_ax: fan-plane tables do not fit shared memory");
    PAX_CHECK_CUDA(cudaFuncSetAttribute(k_ax_fan<CW, NWARPS, MINB, MODE>,
                                        cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    AxArgs b = a;
    b.fan_count = count;
    dim3 grid(a.fan_ntiles * count, pax_cdiv(a.nv, a.fan_rows), a.fan_split);
    if (a.fan_ntiles == 0) return PAX_OK;
    if (a.fan_queue) {
        PAX_CHECK_CUDA(cudaFuncSetAttribute(k_ax_fan_queued<CW, NWARPS, MINB, MODE>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        b.fan_nband = (int)grid.y;
        b.fan_nblock = a.fan_ntiles * count * (int)grid.y * a.fan_split;
        PAX_CHECK_CUDA(cudaMemsetAsync(a.fan_queue, 0, sizeof(int), st));
        k_ax_fan_queued<CW, NWARPS, MINB, MODE><<<a.fan_sms * MINB, NWARPS * 32, smem, st>>>(b);
        if (a.fan_split > 1) {
            const size_t n = (size_t)count * a.nv * a.fan_ntiles * a.fan_bw;
            const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
            k_fan_finish<MODE><<<blocks, 256, 0, st>>>(b, a.fan_bw);
        }
        return PAX_OK;
    }
    k_ax_fan<CW, NWARPS, MINB, MODE><<<grid, NWARPS * 32, smem, st>>>(b);
    if (a.fan_split > 1) {
        const size_t n = (size_t)count * a.nv * a.fan_ntiles * NWARPS;
        const int blocks = (int)std::min<size_t>((n + 255) / 256, 148 * 16);
        k_fan_finish<MODE><<<blocks, 256, 0, st>>>(b, NWARPS);
    }
    return PAX_OK;
}

template <int CW, int NWARPS, int MINB>
static int fan_launch_mode(const AxArgs &a, int mode, int count, cudaStream_t st)
{
    return mode == PAX_AX_PLAIN ? fan_launch<CW, NWARPS, MINB, PAX_AX_PLAIN>(a, count, st)
                                : fan_launch<CW, NWARPS, MINB, PAX_AX_RESIDUAL>(a, count, st);
}

int pax_ax_fan_launch(const PaxPlan *p, const AxArgs &a0, int mode, int count, cudaStream_t st)
{
    AxArgs a = a0;
    a.fan_rows = p->fan_rows;
    a.fan_epoch = p->fan_epoch;
    a.fault = p->d_fault;
    a.fan_order = p->d_fan_order;
    a.fan_ntiles = p->fan_ntiles;
    // split along the chords when the launch is small and the caller gave the scratch for it (a0.fan_scratch);
    // not when samples are counted
    a.fan_split = (a0.fan_scratch && !a0.nsamples) ? fan_parts(p, count) : 1;
    if (a.fan_split == 1) a.fan_scratch = nullptr;
    // (the small launches that are split along the chords keep the plain grid: 2 / 3 / 5 angles of C2 take
    // 0.680 / 0.937 / 1.399 ms on the grid and 0.688 / 0.930 / 1.408 ms from the queue; PAX_FAN_QUEUED=2 queues them too)
    if (p->fan_queued && (a.fan_split == 1 || p->fan_queued == 2) && p->fan_share_parts == 1 && !p->h_fan_order_q.empty()) {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        a.fan_queue = p->d_fan_queue;
        a.fan_sms = sms;
        a.fan_order = p->d_fan_order_q;
        a.fan_bw = p->fan_bw;
        a.fan_ntiles = (int)p->h_fan_order_q.size();
    }
    PAX_REQUIRE(a.fan_order, "pax_ax: the fan-plane projector was not configured for this plan");
    PAX_REQUIRE(((uintptr_t)a.vol0 & 15) == 0, "pax_ax: resident volume must be 16-byte aligned");
    PAX_REQUIRE(p->vol_elems < ((size_t)1 << 30), "pax_ax: resident volume too large for the fan-plane kernel");
    PAX_REQUIRE(p->fan_epoch >= 1, "pax_ax: the fan-plane projector does not fit this geometry "
                                   "(the z travel of a single detector row exceeds its lanes)");
    switch (p->fan_cw * 10000 + p->fan_warps * 100 + p->fan_minb) {
    case 80404: return fan_launch_mode<8, 4, 4>(a, mode, count, st);          // the default
    case 80802: return fan_launch_mode<8, 8, 2>(a, mode, count, st);
    default: break;
    }
    return pax_set_error(PAX_ERR_INVALID, "pax_ax: unsupported fan configuration CW=%d warps=%d minb=%d", p->fan_cw,
                         p->fan_warps, p->fan_minb);
}

// Epoch length, rows per warp and launch shape from the plan's geometry.
int pax_fan_configure(PaxPlan *p)
{
    const PaxGeo &g = p->geo;
    const double dmax = std::fmax(g.dVoxelX, g.dVoxelY);
    // largest |dz| per step of any ray, in slices
    double slope = 0.0;
    for (const AngleDev &A : p->h_angles) {
        const double nmin = std::ceil((double)A.DSD / dmax / g.accuracy);
        const double za = std::fabs(A.oz - A.sz), zb = std::fabs(A.oz + (g.nDetecV - 1) * A.vz - A.sz);
        slope = std::fmax(slope, std::fmax(za, zb) / nmin);
    }
    const char *e;
    int cw = 8, nw = 4;
    if ((e = getenv("PAX_FAN_CW")) && atoi(e) == 8) cw = atoi(e);                                   // tuning hooks
    if ((e = getenv("PAX_FAN_NW")) && (atoi(e) == 4 || atoi(e) == 8)) nw = atoi(e);
    int minb = 16 / nw;
    if ((e = getenv("PAX_FAN_MINB")) && atoi(e) >= 1) minb = atoi(e);
    p->fan_minb = minb;
    // a single row must fit the 32 lanes (128 slices) of a pass over one epoch, with room to spare for its
    // neighbours: halve the epoch until its z travel is at most 64 slices
    int ep = 160;
    if ((e = getenv("PAX_FAN_EPOCH")) && atoi(e) >= cw && atoi(e) <= FAN_EPMAX) ep = atoi(e) / cw * cw;
    if (ep > 32 * cw) ep = 32 * cw;                             // at most 32 windows per epoch (rtab)
    while (ep >= cw && ep * slope > 64.0) ep /= 2;
    if (ep < cw) ep = 0;
    else ep = ep / cw * cw;
    p->fan_epoch = ep;
    p->fan_cw = cw;
    p->fan_warps = nw;
    const int tiles = (g.nDetecV + FAN_MAXROWS - 1) / FAN_MAXROWS;
    p->fan_rows = (g.nDetecV + tiles - 1) / tiles;
    // launch order of the column tiles: longest chords first (mean over a few of the plan's angles of the in-plane
    // chord through the volume's box, summed over the tile's columns)
    auto order_of = [&](int width, std::vector<int> &order) {
        const int ntile = (g.nDetecU + width - 1) / width;
        std::vector<double> cost(ntile, 0.0);
        const int na = (int)p->h_angles.size(), nsamp = std::min(na, 8);
        for (int s = 0; s < nsamp; ++s) {
            const AngleDev &A = p->h_angles[(size_t)s * na / nsamp];
            for (int u = 0; u < g.nDetecU; ++u) {
                const double dx = A.ox + u * A.ux - A.sx, dy = A.oy + u * A.uy - A.sy;
                double t0 = 0.0, t1 = 1.0;
                bool hit = true;
                const double lo = 0.0, hx = g.nVoxelX + 1.0, hy = g.nVoxelY + 1.0;
                if (dx == 0.0) hit = hit && A.sx >= lo && A.sx < hx;
                else { const double a0 = (lo - A.sx) / dx, a1 = (hx - A.sx) / dx;
                       t0 = std::fmax(t0, std::fmin(a0, a1)); t1 = std::fmin(t1, std::fmax(a0, a1)); }
                if (dy == 0.0) hit = hit && A.sy >= lo && A.sy < hy;
                else { const double a0 = (lo - A.sy) / dy, a1 = (hy - A.sy) / dy;
                       t0 = std::fmax(t0, std::fmin(a0, a1)); t1 = std::fmin(t1, std::fmax(a0, a1)); }
                if (hit && t1 > t0) cost[u / width] += (t1 - t0) * std::sqrt(dx * dx + dy * dy);
            }
        }
        order.resize(ntile);
        for (int t = 0; t < ntile; ++t) order[t] = t;
        std::stable_sort(order.begin(), order.end(), [&](int a_, int b_) { return cost[a_] > cost[b_]; });
    };
    order_of(nw, p->h_fan_order);
    p->fan_bw = FAN_BW;         // (4 .. 32 columns per block measured within 1.5 % of each other)
    order_of(p->fan_bw, p->h_fan_order_q);
    p->fan_queued = (e = getenv("PAX_FAN_QUEUED")) ? atoi(e) : 1;                      // tuning hook
    // AUTO's criteria (pax_plan_projector, pax_plan_fan_info): how many detector rows fit the 32 lanes of a warp
    // over one epoch in the worst case, and how long the runs of equal n are (what the sharing is worth)
    {
        const double vz = g.dDetecV / g.dVoxelZ;
        const double hx = 0.5 * g.nVoxelX * g.dVoxelX, hy = 0.5 * g.nVoxelY * g.dVoxelY;
        double tau = 0.0;       // largest fraction of a ray's length at which it can still be inside the volume
        for (const AngleDev &A : p->h_angles) {
            const double rx = hx + std::fabs(A.offOX) + g.dVoxelX, ry = hy + std::fabs(A.offOY) + g.dVoxelY;
            tau = std::fmax(tau, std::fmin(1.0, ((double)A.DSO + std::sqrt(rx * rx + ry * ry)) / (double)A.DSD));
        }
        const double room = 128.0 - 6.0 - ep * slope;
        int fit = (ep > 0 && room > 0) ? (int)std::floor(room / (vz * tau)) : 0;
        p->fan_fit = std::max(ep > 0 ? 1 : 0, std::min(fit, (int)g.nDetecV));
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

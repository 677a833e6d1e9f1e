// Internal declarations shared by the libpax translation units (not part of the ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <vector>

#include "pax.h"

#define PAX_ZOFF 4   // z index of voxel 0 inside a resident column (keeps float4 alignment)

// Per-angle constants, computed in double on the host (plan.cu).
struct AngleDev {
    // ray frame in the padded index space of the resident volume: voxel (k,j,i) centre
    // sits at (x,y,z) = (i+1, j+1, k+PAX_ZOFF).   [SURVEY.md A.2/A.3]
    double sx, sy, sz;     // source
    double ox, oy, oz;     // centre of pixel (v=0,u=0)
    double ux, uy;         // step per +1 in u (no z component: u-hat is in-plane)
    double vz;             // step per +1 in v (v-hat = z)
    // the same frame in mm relative to the centre of the W box (= offOrigin)   [A.5]
    float wsx, wsy, wsz;
    float wox, woy, woz;
    float wux, wuy, wvz;
    // voxel-driven backprojection   [A.4]
    float cosv, sinv, DSD, DSO;
    float offOZ, offOY, offOX, offDV, offDU;
    float pad_;
};

struct PaxPlan {
    PaxGeo geo;
    int n_angles;
    int device;
    int nyp, nxp, nzp;          // resident (padded) dims
    size_t vol_elems;
    AngleDev *d_angles;
    std::vector<AngleDev> h_angles;
    int projector;              // PAX_PROJECTOR_* requested (AUTO resolves per call)
    // ax_fan.cu (fan-plane projector): launch shape and AUTO's criteria, from the geometry (pax_fan_configure)
    int fan_rows;               // detector rows one warp walks
    int fan_epoch;              // steps per epoch (0: the kernel does not fit the geometry)
    int fan_cw;                 // steps per window (ring depth)
    int fan_warps;              // warps per CTA
    int fan_minb;               // CTAs per SM the kernel is compiled for
    int fan_fit;                // detector rows whose z range fits a warp's lanes over an epoch, worst case
    double fan_run;             // mean length (rows) of the equal-n runs of a detector column
    int *d_fault;               // device word the fan kernel raises if a single row's z range overflowed its lanes
    std::vector<int> h_fan_order;   // column tiles (fan_warps columns each) sorted by decreasing chord length
    int *d_fan_order;               // the tiles this plan's launches walk (all of them, or its share)
    int fan_ntiles;                 // how many
    int fan_share_part, fan_share_parts;   // pax_plan_set_column_share
    // the queued launch (k_ax_fan_queued): blocks of 16 columns in the same order, and the counter the warps take from
    std::vector<int> h_fan_order_q;
    int *d_fan_order_q;             // inside d_fan_order's allocation
    int *d_fan_queue;               // the counter (PAX_FAN_MAXQ ints reserved), cleared before every queued launch
    int fan_bw;                     // columns per block
    int fan_queued;                 // 1: launches that own all columns and are not split go through the queues
};
#define PAX_FAN_MAXQ 4


// arguments shared by the two forward-projector kernels (ax.cu ray march, ax_fan.cu fan planes)
struct AxArgs {
    const AngleDev *angles;
    const float *vol0;
    const float *vol1;
    const char *quad0;            // quad-packed channels
    const char *quad1;
    int qshift;                   // log2 of the quad column stride in bytes
    float nxq_f;                  // quad columns per y row
    float *out;
    float *out1;
    const float *b;
    double *sumsq;
    unsigned long long *nsamples;
    unsigned *maxout;
    int first, nv, nu;
    int nxp, nzp;                 // resident strides: x -> nzp, y -> nxp*nzp
    float hix, hiy, hiz;          // exclusive upper bounds of valid padded coordinates
    float loz;                    // inclusive lower z bound (PAX_ZOFF-1); x,y lower bound is 0
    double acc;
    float dX, dY, dZ;
    float a0, a1, c;
    float whx, why, whz, wthr;
    int fan_rows;                 // ax_fan.cu: detector rows one warp walks
    int fan_epoch;                // ax_fan.cu: steps per epoch
    int fan_count;                // ax_fan.cu: angles of the launch
    const int *fan_order;         // ax_fan.cu: column tiles in launch order (longest chords first)
    int fan_ntiles;               // ax_fan.cu: how many of them
    int fan_split;                // ax_fan.cu: parts the launch is cut into along the chords (1: none)
    float *fan_scratch;           // ax_fan.cu: [fan_split][count][V][U] raw sums of a split launch
    int accumulate;               // ax_fan.cu, PLAIN: out += instead of out = (second channel pass)
    int *fan_queue;               // ax_fan.cu, queued launch: next column of the launch
    int fan_bw;                   // ax_fan.cu, queued launch: columns per block
    int fan_nblock;               // ax_fan.cu, queued launch: blocks of columns in the launch
    int fan_sms;                  // ax_fan.cu, queued launch: SMs of the device (the grid is fan_minb CTAs on each)
    int fan_nband;                // ax_fan.cu, queued launch: row bands (grid.y of the plain launch)
    int *fault;                   // ax_fan.cu: set to 1 if a single row's z range did not fit the lanes (never, by construction)
};

// ax_fan.cu
int pax_fan_configure(PaxPlan *p);
int pax_ax_fan_launch(const PaxPlan *p, const AxArgs &a, int mode, int count, cudaStream_t st);


int pax_set_error(int code, const char *fmt, ...);

#define PAX_CHECK_CUDA(expr)                                                              \
    do {                                                                                  \
        cudaError_t e__ = (expr);                                                         \
        if (e__ != cudaSuccess)                                                           \
            return pax_set_error(PAX_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,            \
                                 cudaGetErrorString(e__), __FILE__, __LINE__);            \
    } while (0)

#define PAX_REQUIRE(cond, ...)                                                            \
    do {                                                                                  \
        if (!(cond)) return pax_set_error(PAX_ERR_INVALID, __VA_ARGS__);                  \
    } while (0)

static inline int pax_cdiv(long a, long b) { return (int)((a + b - 1) / b); }

#ifdef __CUDACC__
// chord length (mm) of the segment source->pixel through the box |x|<=hx, |y|<=hy, |z|<=hz
__device__ __forceinline__ float box_chord(float sx, float sy, float sz, float dx, float dy, float dz,
                                           float hx, float hy, float hz)
{
    float t0 = 0.f, t1 = 1.f;
    bool miss = false;
    {
        if (dx == 0.f) miss |= (sx < -hx || sx > hx);
        else { const float r = 1.f / dx; const float a = (-hx - sx) * r, b = (hx - sx) * r;
               t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
        if (dy == 0.f) miss |= (sy < -hy || sy > hy);
        else { const float r = 1.f / dy; const float a = (-hy - sy) * r, b = (hy - sy) * r;
               t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
        if (dz == 0.f) miss |= (sz < -hz || sz > hz);
        else { const float r = 1.f / dz; const float a = (-hz - sz) * r, b = (hz - sz) * r;
               t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b)); }
    }
    if (miss || !(t1 > t0)) return 0.f;
    return (t1 - t0) * sqrtf(dx * dx + dy * dy + dz * dz);
}

__device__ __forceinline__ float pax_ray_chord(const AngleDev &A, int u, int v, float hx, float hy, float hz)
{
    const float px = A.wox + u * A.wux, py = A.woy + u * A.wuy, pz = A.woz + v * A.wvz;
    return box_chord(A.wsx, A.wsy, A.wsz, px - A.wsx, py - A.wsy, pz - A.wsz, hx, hy, hz);
}
#endif  // __CUDACC__

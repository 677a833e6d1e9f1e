// Plan (per-geometry device constants), error reporting, layout conversion, and the
// small elementwise / reduction kernels of the path.  sm_100a only.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <new>

#include "pax_internal.h"

// ----------------------------------------------------------------------------- errors
static thread_local char g_err[512] = "";

int pax_set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

extern "C" const char *pax_last_error(void) { return g_err; }
extern "C" int pax_version(void) { return PAX_VERSION; }

// ------------------------------------------------------------------------------- plan
extern "C" int pax_plan_create(const PaxGeo *geo, const double *rows, int n, PaxPlan **out)
{
    PAX_REQUIRE(geo && rows && out, "pax_plan_create: NULL argument");
    PAX_REQUIRE(n >= 1, "pax_plan_create: n_angles=%d", n);
    PAX_REQUIRE(geo->nVoxelX >= 1 && geo->nVoxelY >= 1 && geo->nVoxelZ >= 1 &&
                    geo->nDetecU >= 1 && geo->nDetecV >= 1,
                "pax_plan_create: non-positive dimensions");
    PAX_REQUIRE(geo->dVoxelX > 0 && geo->dVoxelY > 0 && geo->dVoxelZ > 0 && geo->dDetecU > 0 &&
                    geo->dDetecV > 0 && geo->accuracy > 0,
                "pax_plan_create: non-positive spacing/accuracy");
    PAX_REQUIRE((long)geo->nVoxelX < 8192 && (long)geo->nVoxelY < 8192 && (long)geo->nVoxelZ < 8192,
                "pax_plan_create: volume dimension too large");
    PaxPlan *p = new (std::nothrow) PaxPlan();
    if (!p) return pax_set_error(PAX_ERR_NOMEM, "pax_plan_create: out of host memory");
    p->geo = *geo;
    p->n_angles = n;
    p->nyp = geo->nVoxelY + 2;
    p->nxp = geo->nVoxelX + 2;
    p->nzp = ((geo->nVoxelZ + PAX_ZOFF + 1) + 3) / 4 * 4;
    p->vol_elems = (size_t)p->nyp * p->nxp * p->nzp;
    if (p->vol_elems >= (size_t)1 << 31) {
        delete p;
        return pax_set_error(PAX_ERR_INVALID, "pax_plan_create: resident volume exceeds 2^31 elements");
    }
    p->h_angles.resize(n);
    const double nX = geo->nVoxelX, nY = geo->nVoxelY, nZ = geo->nVoxelZ;
    const double dX = geo->dVoxelX, dY = geo->dVoxelY, dZ = geo->dVoxelZ;
    for (int i = 0; i < n; ++i) {
        const double *a = rows + (size_t)i * PAX_ANGLE_STRIDE;
        const double th = a[0], DSD = a[1], DSO = a[2];
        const double oz = a[3], oy = a[4], ox = a[5], ov = a[6], ou = a[7];
        if (!(DSO > 0 && DSD > DSO)) {
            delete p;
            return pax_set_error(PAX_ERR_INVALID, "pax_plan_create: need DSD > DSO > 0 (angle %d)", i);
        }
        const double c = cos(th), s = sin(th);
        AngleDev &A = p->h_angles[i];
        // world -> index: (r - off + sVoxel/2 - dVoxel/2)/dVoxel, then the padding shift
        const double bx = -ox + 0.5 * nX * dX - 0.5 * dX;
        const double by = -oy + 0.5 * nY * dY - 0.5 * dY;
        const double bz = -oz + 0.5 * nZ * dZ - 0.5 * dZ;
        const double u0 = (0.5 - 0.5 * geo->nDetecU) * geo->dDetecU + ou;
        const double v0 = (0.5 - 0.5 * geo->nDetecV) * geo->dDetecV + ov;
        const double Sx = DSO * c, Sy = DSO * s, Sz = 0.0;
        const double Px = -(DSD - DSO) * c - u0 * s, Py = -(DSD - DSO) * s + u0 * c, Pz = v0;
        A.sx = (Sx + bx) / dX + 1.0;
        A.sy = (Sy + by) / dY + 1.0;
        A.sz = (Sz + bz) / dZ + PAX_ZOFF;
        A.ox = (Px + bx) / dX + 1.0;
        A.oy = (Py + by) / dY + 1.0;
        A.oz = (Pz + bz) / dZ + PAX_ZOFF;
        A.ux = -s * geo->dDetecU / dX;
        A.uy = c * geo->dDetecU / dY;
        A.vz = geo->dDetecV / dZ;
        A.wsx = (float)(Sx - ox); A.wsy = (float)(Sy - oy); A.wsz = (float)(Sz - oz);
        A.wox = (float)(Px - ox); A.woy = (float)(Py - oy); A.woz = (float)(Pz - oz);
        A.wux = (float)(-s * geo->dDetecU); A.wuy = (float)(c * geo->dDetecU);
        A.wvz = (float)geo->dDetecV;
        A.cosv = (float)c; A.sinv = (float)s; A.DSD = (float)DSD; A.DSO = (float)DSO;
        A.offOZ = (float)oz; A.offOY = (float)oy; A.offOX = (float)ox;
        A.offDV = (float)ov; A.offDU = (float)ou;
        A.pad_ = 0.f;
    }
    cudaError_t e = cudaGetDevice(&p->device);
    if (e == cudaSuccess) e = cudaMalloc(&p->d_angles, sizeof(AngleDev) * n + sizeof(int));
    if (e == cudaSuccess)
        e = cudaMemcpy(p->d_angles, p->h_angles.data(), sizeof(AngleDev) * n, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (p->d_angles) cudaFree(p->d_angles);
        delete p;
        return pax_set_error(PAX_ERR_CUDA, "pax_plan_create: %s", cudaGetErrorString(e));
    }
    p->d_fault = reinterpret_cast<int *>(p->d_angles + n);       // same allocation, after the angles
    cudaMemset(p->d_fault, 0, sizeof(int));
    p->projector = PAX_PROJECTOR_AUTO;
    p->d_fan_order = nullptr;
    p->d_fan_order_q = nullptr;
    p->d_fan_queue = nullptr;
    p->fan_share_part = 0;
    p->fan_share_parts = 1;
    pax_fan_configure(p);
    p->fan_ntiles = (int)p->h_fan_order.size();
    if (!p->h_fan_order.empty()) {
        // the fan-plane kernel's column tiles, longest chords first (ax_fan.cu: tail of the last wave)
        const size_t n4 = p->h_fan_order.size(), nq = p->h_fan_order_q.size();
        e = cudaMalloc(&p->d_fan_order, sizeof(int) * (n4 + nq + PAX_FAN_MAXQ));
        if (e == cudaSuccess)
            e = cudaMemcpy(p->d_fan_order, p->h_fan_order.data(), sizeof(int) * n4, cudaMemcpyHostToDevice);
        p->d_fan_order_q = p->d_fan_order + n4;
        p->d_fan_queue = p->d_fan_order_q + nq;
        if (e == cudaSuccess && nq)
            e = cudaMemcpy(p->d_fan_order_q, p->h_fan_order_q.data(), sizeof(int) * nq, cudaMemcpyHostToDevice);
        if (e != cudaSuccess) {
            if (p->d_fan_order) cudaFree(p->d_fan_order);
            cudaFree(p->d_angles);
            delete p;
            return pax_set_error(PAX_ERR_CUDA, "pax_plan_create: %s", cudaGetErrorString(e));
        }
    }
    *out = p;
    return PAX_OK;
}

extern "C" int pax_plan_set_projector(PaxPlan *p, int kind)
{
    PAX_REQUIRE(p, "pax_plan_set_projector: NULL plan");
    PAX_REQUIRE(kind == PAX_PROJECTOR_AUTO || kind == PAX_PROJECTOR_RAY || kind == PAX_PROJECTOR_FAN,
                "pax_plan_set_projector: unknown projector %d", kind);
    PAX_REQUIRE(kind != PAX_PROJECTOR_FAN || p->fan_epoch >= 1,
                "pax_plan_set_projector: the fan-plane projector does not fit this geometry");
    p->projector = kind;
    return PAX_OK;
}

// which kernel pax_ax will run for this plan: AUTO picks the fan-plane projector when the
// equal-n runs of a detector column are long enough to pay for the shared in-plane pass
extern "C" int pax_plan_projector(const PaxPlan *p)
{
    if (!p) return PAX_PROJECTOR_RAY;
    if (p->projector != PAX_PROJECTOR_AUTO) return p->projector;
    // FAN pays when many rows share one lattice (long runs of equal n) and the rows fit a warp's lanes.  Measured
    // with C2's volume and coarser and coarser detector rows (tools/bench_crossover.py, r2, fan / ray in ms per 20
    // angles): mean run 154 rows 5.3 / 18.6; 51 rows 4.4 / 6.7; 38 rows 4.2 / 5.2; 26 rows 4.1 / 3.7
    return (p->fan_fit >= 32 && p->fan_run >= 32.0) ? PAX_PROJECTOR_FAN : PAX_PROJECTOR_RAY;
}

// The fan-plane projector of this plan works on every `parts`-th column tile only (tiles dealt out in order of
// decreasing cost, back and forth, so that the shares are balanced): how the ranks of a multi-GPU run split the
// detector columns of an angle between them.  Columns outside the share are neither read nor written by pax_ax.
extern "C" int pax_plan_set_column_share(PaxPlan *p, int part, int parts)
{
    PAX_REQUIRE(p, "pax_plan_set_column_share: NULL plan");
    PAX_REQUIRE(parts >= 1 && part >= 0 && part < parts, "pax_plan_set_column_share: share %d of %d", part, parts);
    PAX_REQUIRE(!p->h_fan_order.empty() && p->fan_epoch >= 1,
                "pax_plan_set_column_share: the fan-plane projector does not fit this geometry");
    std::vector<int> mine;
    const int n = (int)p->h_fan_order.size();
    for (int i = 0; i < n; ++i) {
        const int round = i / parts, pos = i % parts;
        if (((round & 1) ? parts - 1 - pos : pos) == part) mine.push_back(p->h_fan_order[i]);
    }
    p->fan_share_part = part;
    p->fan_share_parts = parts;
    p->fan_ntiles = (int)mine.size();
    if (!mine.empty())
        PAX_CHECK_CUDA(cudaMemcpy(p->d_fan_order, mine.data(), sizeof(int) * mine.size(), cudaMemcpyHostToDevice));
    return PAX_OK;
}

extern "C" int pax_plan_fan_info(const PaxPlan *p, int *rows_per_tile, int *warps, double *mean_run)
{
    PAX_REQUIRE(p, "pax_plan_fan_info: NULL plan");
    if (rows_per_tile) *rows_per_tile = p->fan_fit;
    if (warps) *warps = p->fan_warps;
    if (mean_run) *mean_run = p->fan_run;
    return PAX_OK;
}

// 1 if a launch of the FAN kernel met a z window wider than its lanes (a bug in the tile choice;
// results of that launch are wrong).  Synchronises the device.  Test hook.
extern "C" int pax_plan_fault(const PaxPlan *p)
{
    int f = 0;
    if (!p || cudaMemcpy(&f, p->d_fault, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return f;
}

extern "C" int pax_plan_destroy(PaxPlan *p)
{
    if (!p) return PAX_OK;
    if (p->d_angles) cudaFree(p->d_angles);
    if (p->d_fan_order) cudaFree(p->d_fan_order);
    delete p;
    return PAX_OK;
}

extern "C" int pax_plan_n_angles(const PaxPlan *p) { return p ? p->n_angles : 0; }
extern "C" size_t pax_vol_elems(const PaxPlan *p) { return p ? p->vol_elems : 0; }

// ------------------------------------------------------------------ layout conversion
// (Z,Y,X) x-fastest  <->  resident (Y+2, X+2, Zp) z-fastest: a 32x32 smem transpose per y.
template <bool PACK>
__global__ void __launch_bounds__(256) k_transpose(const float *__restrict__ src,
                                                   float *__restrict__ dst, int nz, int ny, int nx,
                                                   int nxp, int nzp)
{
    __shared__ float tile[32][33];
    const int y = blockIdx.z;
    const int x0 = blockIdx.x * 32, z0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    if (PACK) {
        for (int r = ty; r < 32; r += 8) {
            const int z = z0 + r, x = x0 + tx;
            tile[r][tx] = (z < nz && x < nx) ? src[((size_t)z * ny + y) * nx + x] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int x = x0 + r, z = z0 + tx;
            if (x < nx && z < nz)
                dst[((size_t)(y + 1) * nxp + (x + 1)) * nzp + PAX_ZOFF + z] = tile[tx][r];
        }
    } else {
        for (int r = ty; r < 32; r += 8) {
            const int x = x0 + r, z = z0 + tx;
            tile[r][tx] = (x < nx && z < nz)
                              ? src[((size_t)(y + 1) * nxp + (x + 1)) * nzp + PAX_ZOFF + z] : 0.f;
        }
        __syncthreads();
        for (int r = ty; r < 32; r += 8) {
            const int z = z0 + r, x = x0 + tx;
            if (z < nz && x < nx) dst[((size_t)z * ny + y) * nx + x] = tile[tx][r];
        }
    }
}

extern "C" int pax_vol_pack(const PaxPlan *p, const float *vol_zyx, float *vol_res, void *stream)
{
    PAX_REQUIRE(p && vol_zyx && vol_res, "pax_vol_pack: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    PAX_CHECK_CUDA(cudaMemsetAsync(vol_res, 0, p->vol_elems * sizeof(float), st));
    dim3 grid(pax_cdiv(p->geo.nVoxelX, 32), pax_cdiv(p->geo.nVoxelZ, 32), p->geo.nVoxelY);
    k_transpose<true><<<grid, 256, 0, st>>>(vol_zyx, vol_res, p->geo.nVoxelZ, p->geo.nVoxelY,
                                            p->geo.nVoxelX, p->nxp, p->nzp);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

extern "C" int pax_vol_unpack(const PaxPlan *p, const float *vol_res, float *vol_zyx, void *stream)
{
    PAX_REQUIRE(p && vol_zyx && vol_res, "pax_vol_unpack: NULL argument");
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid(pax_cdiv(p->geo.nVoxelX, 32), pax_cdiv(p->geo.nVoxelZ, 32), p->geo.nVoxelY);
    k_transpose<false><<<grid, 256, 0, st>>>(vol_res, vol_zyx, p->geo.nVoxelZ, p->geo.nVoxelY,
                                             p->geo.nVoxelX, p->nxp, p->nzp);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------------------------------- SART update
// x = max(0, x + lambda*inv_v[y,x]*upd) over the interior of the resident layout.
__global__ void __launch_bounds__(256) k_sart_update(float *__restrict__ x,
                                                     const float *__restrict__ upd,
                                                     const float *__restrict__ inv_v, int nz, int ny,
                                                     int nx, int nxp, int nzp, float lambda, int noneg)
{
    // one thread per (column, 4 z): consecutive threads walk z fastest -> coalesced float4
    const int nz4 = (nz + 3) / 4;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long ncol = (long)ny * nx;
    if (t >= ncol * nz4) return;
    const int q = (int)(t % nz4);
    const long col = t / nz4;
    const int yy = (int)(col / nx), xx = (int)(col % nx);
    const float s = lambda * inv_v[col];
    const size_t base = ((size_t)(yy + 1) * nxp + (xx + 1)) * nzp + PAX_ZOFF + 4 * q;
    float4 a = *reinterpret_cast<const float4 *>(x + base);
    const float4 b = *reinterpret_cast<const float4 *>(upd + base);
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
    if (noneg) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    const int zr = nz - 4 * q;   // valid z in this quad; never touch the zero rim
    if (zr >= 4) {
        *reinterpret_cast<float4 *>(x + base) = a;
    } else {
        x[base] = a.x;
        if (zr > 1) x[base + 1] = a.y;
        if (zr > 2) x[base + 2] = a.z;
    }
}

extern "C" int pax_sart_update(const PaxPlan *p, float *x_res, const float *upd_res,
                               const float *inv_v, float lambda, int noneg, void *stream)
{
    PAX_REQUIRE(p && x_res && upd_res && inv_v, "pax_sart_update: NULL argument");
    const long n = (long)p->geo.nVoxelY * p->geo.nVoxelX * ((p->geo.nVoxelZ + 3) / 4);
    k_sart_update<<<pax_cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(
        x_res, upd_res, inv_v, p->geo.nVoxelZ, p->geo.nVoxelY, p->geo.nVoxelX, p->nxp, p->nzp, lambda,
        noneg);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------- SART update fused with its collective
// Multi-GPU OS-SART: every rank has backprojected its share of the subset into a partial
// update; the iterate on every rank must become  x = max(0, x + lambda * invV * sum_ranks(upd)).  Instead of
// an NCCL all-reduce of the 144 MB update followed by a local update kernel, ONE kernel does the reduction,
// the update and the broadcast over NVLink peer memory: rank r owns the r-th slice of the volume's float4
// units; for those it
//   * multicast path (NVSwitch / NVLS):  sum = multimem.ld_reduce.add over all ranks' update buffers (the
//     switch adds), computes the new x and writes it to ALL ranks with one multimem.st;
//   * peer-pointer path:  loads the W partial updates through the peers' mapped pointers, adds them in rank
//     order, and stores the new x into every peer's iterate.
// The buffers are symmetric allocations (torch.distributed._symmetric_memory on the host side, which also
// provides the barriers before and after).  Every rank ends up with bit-identical x.
// Measured on 8 x B200 at C2 (tools/bench_collective.py): all-reduce + update kernel 0.472 ms, this kernel
// with its two barriers 0.381 ms (multicast) / 0.448 ms (peer pointers); four units per thread: 0.413 ms;
// a __threadfence_system() per thread: +0.2 ms.  On 2 GPUs it is 0.05 ms slower than NCCL.
struct PeerPtrs {
    float *x[PAX_MAX_PEERS];
    const float *upd[PAX_MAX_PEERS];
};

__device__ __forceinline__ float4 mc_ld_reduce_add(const float *p)
{
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void mc_st4(float *p, float4 v)
{
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_red4(float *p, float4 v)
{
    asm volatile("multimem.red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void mc_st1(float *p, float v)
{
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

template <bool MC>
__global__ void __launch_bounds__(256) k_sart_update_fused(float *__restrict__ x_local, float *x_mc,
                                                           const float *upd_mc, const PeerPtrs peers, int world,
                                                           const float *__restrict__ inv_v, int nz, int ny, int nx,
                                                           int nxp, int nzp, float lambda, int noneg, int rank,
                                                           long n_local)
{
    // rank r owns the volume rows y = r, r + world, ...: whatever part of the plane a subset sees (a half disc
    // whose orientation turns with the angles) is then spread evenly over the ranks
    const int nz4 = (nz + 3) / 4;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_local) return;
    const int q = (int)(t % nz4);
    const long cl = t / nz4;
    const int yy = (int)(cl / nx) * world + rank, xx = (int)(cl % nx);
    const long col = (long)yy * nx + xx;
    const float s = lambda * inv_v[col];
    // a column no ray of this subset sees (V_s = 0: outside the half-fan's footprint, about half of the plane)
    // gets no update: x stays what it is on every rank, and nothing has to cross the fabric for it
    if (s == 0.f) return;
    const size_t base = ((size_t)(yy + 1) * nxp + (xx + 1)) * nzp + PAX_ZOFF + 4 * q;
    float4 b;
    if (MC) {
        b = mc_ld_reduce_add(upd_mc + base);
    } else {
        b = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < world; ++r) {                      // fixed order: identical sums on every rank
            const float4 u = *reinterpret_cast<const float4 *>(peers.upd[r] + base);
            b.x += u.x; b.y += u.y; b.z += u.z; b.w += u.w;
        }
    }
    float4 a = *reinterpret_cast<const float4 *>(x_local + base);
    a.x = fmaf(s, b.x, a.x); a.y = fmaf(s, b.y, a.y); a.z = fmaf(s, b.z, a.z); a.w = fmaf(s, b.w, a.w);
    if (noneg) { a.x = fmaxf(a.x, 0.f); a.y = fmaxf(a.y, 0.f); a.z = fmaxf(a.z, 0.f); a.w = fmaxf(a.w, 0.f); }
    const int zr = nz - 4 * q;   // valid z in this quad; never touch the zero rim
    if (MC) {
        if (zr >= 4) {
            mc_st4(x_mc + base, a);
        } else {
            mc_st1(x_mc + base, a.x);
            if (zr > 1) mc_st1(x_mc + base + 1, a.y);
            if (zr > 2) mc_st1(x_mc + base + 2, a.z);
        }
    } else {
        for (int r = 0; r < world; ++r) {
            float *dst = peers.x[r] + base;
            if (zr >= 4) {
                *reinterpret_cast<float4 *>(dst) = a;
            } else {
                dst[0] = a.x;
                if (zr > 1) dst[1] = a.y;
                if (zr > 2) dst[2] = a.z;
            }
        }
    }
    // no fence here: the kernel's completion orders its (peer) stores before the barrier kernel that the
    // host side enqueues next on the same stream
}

// ---------------------------------------------------------------- broadcast of chunks of a symmetric buffer
struct PeerDst {
    char *p[PAX_MAX_PEERS];
};

template <bool MC>
__global__ void __launch_bounds__(256) k_bcast_chunks(const char *__restrict__ src, long src_first, char *dst_mc,
                                                      const PeerDst peers, int world, size_t chunk16, long first,
                                                      long stride, int add)
{
    // blockIdx.y = chunk, grid-stride over its 16-byte units: a warp moves 512 contiguous bytes per instruction
    const size_t off = ((size_t)first + (size_t)blockIdx.y * (size_t)stride) * chunk16 * 16;
    const size_t soff = src_first < 0 ? off : ((size_t)src_first + (size_t)blockIdx.y * (size_t)stride) * chunk16 * 16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < chunk16; i += (size_t)gridDim.x * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4 *>(src + soff + i * 16);
        if (MC) {
            if (add) mc_red4(reinterpret_cast<float *>(dst_mc + off + i * 16), v);
            else mc_st4(reinterpret_cast<float *>(dst_mc + off + i * 16), v);
        } else {
            for (int r = 0; r < world; ++r) {
                float *d = reinterpret_cast<float *>(peers.p[r] + off + i * 16);
                if (add) { atomicAdd(d, v.x); atomicAdd(d + 1, v.y); atomicAdd(d + 2, v.z); atomicAdd(d + 3, v.w); }
                else *reinterpret_cast<float4 *>(d) = v;
            }
        }
    }
}

extern "C" int pax_bcast_chunks2(const void *src, long src_first_chunk, void *dst_mc, const uint64_t *dst_peers,
                                 int world, size_t chunk_bytes, long first_chunk, long chunk_stride, long n_chunks,
                                 int add, void *stream);

extern "C" int pax_bcast_chunks(const void *src_local, void *dst_mc, const uint64_t *dst_peers, int world,
                                size_t chunk_bytes, long first_chunk, long chunk_stride, long n_chunks, void *stream)
{
    return pax_bcast_chunks2(src_local, -1, dst_mc, dst_peers, world, chunk_bytes, first_chunk, chunk_stride, n_chunks,
                             0, stream);
}

// The general form: the chunks are read from `src` starting at chunk src_first_chunk (-1: at the same place as in
// the destination, i.e. src is this rank's copy of the symmetric buffer), and with add != 0 they are ADDED to what
// every rank's buffer holds (multimem.red: the switch does the adds; ranks that each hold some columns of one
// projection merge them this way into a zeroed slot).
extern "C" int pax_bcast_chunks2(const void *src_local, long src_first_chunk, void *dst_mc, const uint64_t *dst_peers,
                                 int world, size_t chunk_bytes, long first_chunk, long chunk_stride, long n_chunks,
                                 int add, void *stream)
{
    PAX_REQUIRE(src_local && (dst_mc || dst_peers), "pax_bcast_chunks: NULL argument");
    PAX_REQUIRE(world >= 1 && world <= PAX_MAX_PEERS, "pax_bcast_chunks: world %d", world);
    PAX_REQUIRE(chunk_bytes % 16 == 0 && ((uintptr_t)src_local & 15) == 0, "pax_bcast_chunks: 16-byte units");
    PAX_REQUIRE(first_chunk >= 0 && chunk_stride >= 1 && n_chunks >= 0 && n_chunks < 65536, "pax_bcast_chunks: range");
    if (n_chunks == 0 || chunk_bytes == 0) return PAX_OK;
    PeerDst peers;
    for (int r = 0; r < PAX_MAX_PEERS; ++r)
        peers.p[r] = (!dst_mc && r < world) ? reinterpret_cast<char *>(dst_peers[r]) : nullptr;
    const size_t c16 = chunk_bytes / 16;
    const int bx = (int)((c16 + 255) / 256 < 64 ? (c16 + 255) / 256 : 64);
    dim3 grid(bx, (unsigned)n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    if (dst_mc)
        k_bcast_chunks<true><<<grid, 256, 0, st>>>(reinterpret_cast<const char *>(src_local), src_first_chunk,
                                                   reinterpret_cast<char *>(dst_mc), peers, world, c16, first_chunk,
                                                   chunk_stride, add);
    else
        k_bcast_chunks<false><<<grid, 256, 0, st>>>(reinterpret_cast<const char *>(src_local), src_first_chunk, nullptr,
                                                    peers, world, c16, first_chunk, chunk_stride, add);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// The volume rows a rank owns (y = part mod parts), to every rank -- but only the columns the subset's rays see
// (inv_v != 0): the others were not updated, so every rank holds them already.  About half of the plane at C2.
template <bool MC>
__global__ void __launch_bounds__(256) k_bcast_xrows(const float *__restrict__ x_local, float *x_mc, const PeerDst peers,
                                                     int world, const float *__restrict__ inv_v, int nz, int nx,
                                                     int nxp, int nzp, int part, int parts, long n_local)
{
    const int nz4 = (nz + 3) / 4;
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_local) return;
    const int q = (int)(t % nz4);
    const long cl = t / nz4;
    const int yy = (int)(cl / nx) * parts + part, xx = (int)(cl % nx);
    if (inv_v[(long)yy * nx + xx] == 0.f) return;
    const size_t base = ((size_t)(yy + 1) * nxp + (xx + 1)) * nzp + PAX_ZOFF + 4 * q;
    const float4 v = *reinterpret_cast<const float4 *>(x_local + base);      // (the rim beyond nz holds zeros)
    if (MC) mc_st4(x_mc + base, v);
    else
        for (int r = 0; r < world; ++r) *reinterpret_cast<float4 *>(reinterpret_cast<float *>(peers.p[r]) + base) = v;
}

extern "C" int pax_bcast_xrows(const PaxPlan *p, const float *x_local, float *x_mc, const uint64_t *x_peers, int world,
                               const float *inv_v, int part, int parts, void *stream)
{
    PAX_REQUIRE(p && x_local && inv_v && (x_mc || x_peers), "pax_bcast_xrows: NULL argument");
    PAX_REQUIRE(world >= 1 && world <= PAX_MAX_PEERS && parts >= 1 && part >= 0 && part < parts,
                "pax_bcast_xrows: share %d of %d, world %d", part, parts, world);
    const long rows = (p->geo.nVoxelY - part + parts - 1) / parts;
    const long n_local = rows * p->geo.nVoxelX * ((p->geo.nVoxelZ + 3) / 4);
    if (n_local <= 0) return PAX_OK;
    PeerDst peers;
    for (int r = 0; r < PAX_MAX_PEERS; ++r)
        peers.p[r] = (!x_mc && r < world) ? reinterpret_cast<char *>(x_peers[r]) : nullptr;
    const int blocks = pax_cdiv(n_local, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_mc)
        k_bcast_xrows<true><<<blocks, 256, 0, st>>>(x_local, x_mc, peers, world, inv_v, p->geo.nVoxelZ, p->geo.nVoxelX,
                                                    p->nxp, p->nzp, part, parts, n_local);
    else
        k_bcast_xrows<false><<<blocks, 256, 0, st>>>(x_local, nullptr, peers, world, inv_v, p->geo.nVoxelZ,
                                                     p->geo.nVoxelX, p->nxp, p->nzp, part, parts, n_local);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

extern "C" int pax_sart_update_fused(const PaxPlan *p, float *x_local, float *x_mc, const float *upd_mc,
                                     const uint64_t *x_peers, const uint64_t *upd_peers, int rank, int world,
                                     const float *inv_v, float lambda, int noneg, void *stream)
{
    PAX_REQUIRE(p && x_local && inv_v, "pax_sart_update_fused: NULL argument");
    PAX_REQUIRE(world >= 1 && world <= PAX_MAX_PEERS && rank >= 0 && rank < world,
                "pax_sart_update_fused: rank %d of %d", rank, world);
    const bool mc = x_mc && upd_mc;
    PAX_REQUIRE(mc || (x_peers && upd_peers), "pax_sart_update_fused: neither multicast nor peer pointers given");
    const long rows = (p->geo.nVoxelY - rank + world - 1) / world;          // rows y = rank (mod world)
    const long n_local = rows * p->geo.nVoxelX * ((p->geo.nVoxelZ + 3) / 4);
    if (n_local <= 0) return PAX_OK;
    PeerPtrs peers;
    for (int r = 0; r < PAX_MAX_PEERS; ++r) {
        peers.x[r] = (!mc && r < world) ? reinterpret_cast<float *>(x_peers[r]) : nullptr;
        peers.upd[r] = (!mc && r < world) ? reinterpret_cast<const float *>(upd_peers[r]) : nullptr;
    }
    const int blocks = pax_cdiv(n_local, 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (mc)
        k_sart_update_fused<true><<<blocks, 256, 0, st>>>(x_local, x_mc, upd_mc, peers, world, inv_v,
                                                         p->geo.nVoxelZ, p->geo.nVoxelY, p->geo.nVoxelX, p->nxp,
                                                         p->nzp, lambda, noneg, rank, n_local);
    else
        k_sart_update_fused<false><<<blocks, 256, 0, st>>>(x_local, x_mc, upd_mc, peers, world, inv_v,
                                                          p->geo.nVoxelZ, p->geo.nVoxelY, p->geo.nVoxelX, p->nxp,
                                                          p->nzp, lambda, noneg, rank, n_local);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// -------------------------------------------------------------------------------- max
__device__ __forceinline__ unsigned enc_ordered(float f)
{
    const unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ordered(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void __launch_bounds__(256) k_max(const float *__restrict__ x, size_t n, unsigned *out)
{
    float m = -INFINITY;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const size_t n4 = n / 4;
    const float4 *x4 = reinterpret_cast<const float4 *>(x);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        const float4 v = x4[i];
        m = fmaxf(fmaxf(m, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
    }
    for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        m = fmaxf(m, x[i]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    __shared__ float sm[8];
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < 8 ? sm[threadIdx.x] : -INFINITY;
        for (int o = 4; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        if (threadIdx.x == 0) atomicMax(out, enc_ordered(m));
    }
}
__global__ void k_max_decode(unsigned *out) { *reinterpret_cast<float *>(out) = dec_ordered(*out); }

extern "C" int pax_reduce_max(const float *x, size_t n, float *out_dev, void *stream)
{
    PAX_REQUIRE(x && out_dev && n > 0, "pax_reduce_max: bad argument");
    PAX_REQUIRE(((uintptr_t)x & 15) == 0, "pax_reduce_max: input must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    PAX_CHECK_CUDA(cudaMemsetAsync(out_dev, 0, sizeof(float), st));
    const int blocks = (int)std::min<size_t>(148 * 8, (n / 4 + 255) / 256 + 1);
    k_max<<<blocks, 256, 0, st>>>(x, n, reinterpret_cast<unsigned *>(out_dev));
    k_max_decode<<<1, 1, 0, st>>>(reinterpret_cast<unsigned *>(out_dev));
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

extern "C" int pax_max_decode(uint32_t *word_dev, void *stream)
{
    PAX_REQUIRE(word_dev, "pax_max_decode: NULL argument");
    k_max_decode<<<1, 1, 0, (cudaStream_t)stream>>>(word_dev);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

__global__ void __launch_bounds__(256) k_axpy(float4 *__restrict__ y, const float4 *__restrict__ x, float a,
                                              size_t n4)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 v = y[i];
        const float4 w = x[i];
        v.x = fmaf(a, w.x, v.x); v.y = fmaf(a, w.y, v.y); v.z = fmaf(a, w.z, v.z); v.w = fmaf(a, w.w, v.w);
        y[i] = v;
    }
}

__global__ void __launch_bounds__(256) k_lincomb(float *out, const float *x, const float *y, float a,
                                                 float b, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        out[i] = fmaf(a, x[i], b * y[i]);
}

extern "C" int pax_lincomb(float *out, const float *x, const float *y, float a, float b, size_t n,
                           void *stream)
{
    PAX_REQUIRE(out && x && y && n > 0, "pax_lincomb: bad argument");
    k_lincomb<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(out, x, y, a, b, n);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// the rim of both operands is zero, so running over the whole padded array keeps it zero
extern "C" int pax_vol_axpy(const PaxPlan *p, float *y_res, const float *x_res, float a, void *stream)
{
    PAX_REQUIRE(p && y_res && x_res, "pax_vol_axpy: NULL argument");
    PAX_REQUIRE((((uintptr_t)y_res | (uintptr_t)x_res) & 15) == 0, "pax_vol_axpy: volumes must be 16-byte aligned");
    const size_t n4 = p->vol_elems / 4;      // nzp is a multiple of 4
    k_axpy<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<float4 *>(y_res),
                                                      reinterpret_cast<const float4 *>(x_res), a, n4);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------------------------------------ noise
// Philox4x32-10 (Salmon et al. 2011), one block of 4 words per projection pixel keyed by
// (seed, global pixel index); same word assignment as oracle_ctnoise.
__device__ __forceinline__ void philox4x32_10(unsigned c[4], unsigned k0, unsigned k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const unsigned n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// CTnoise.add (reference :88) in double: I0 = 1e8 counts do not fit fp32.
__global__ void __launch_bounds__(256) k_ctnoise(float *__restrict__ p, size_t n, uint64_t index0,
                                                 size_t inner, uint64_t outer_stride, double I0,
                                                 double mean, double sd,
                                                 const float *__restrict__ max_dev, uint64_t seed)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const double m = (double)*max_dev;
    // local element e = (outer, in) of an [n/inner][inner] block cut out of the global array
    const uint64_t idx = index0 + (e / inner) * outer_stride + (e % inner);
    unsigned c[4] = {(unsigned)idx, (unsigned)(idx >> 32), 0u, 0u};
    philox4x32_10(c, (unsigned)seed, (unsigned)(seed >> 32));
    const double inv32 = 1.0 / 4294967296.0;
    const double u1 = ((double)c[0] + 0.5) * inv32, u2 = ((double)c[1] + 0.5) * inv32;
    const double u3 = ((double)c[2] + 0.5) * inv32, u4 = ((double)c[3] + 0.5) * inv32;
    const double lam = I0 * exp(-(double)p[e] / m);
    double k;
    if (lam < 512.0) {   // exact sequential inversion
        double pk = exp(-lam), F = pk;
        k = 0.0;
        while (u1 > F && k < 4096.0) { k += 1.0; pk *= lam / k; F += pk; }
    } else {             // rounded normal approximation
        const double zp = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
        k = floor(lam + sqrt(lam) * zp + 0.5);
        if (k < 0.0) k = 0.0;
    }
    const double zg = sqrt(-2.0 * log(u3)) * cospi(2.0 * u4);
    double I = k + mean + sd * zg;
    if (I <= 0.0) I = 1e-6;
    p[e] = (float)(-log(I / I0) * m);
}

extern "C" int pax_ctnoise(float *proj, size_t n, uint64_t index0, size_t inner, uint64_t outer_stride,
                           double I0, double mean, double sd, const float *max_dev, uint64_t seed,
                           void *stream)
{
    PAX_REQUIRE(proj && max_dev && n > 0, "pax_ctnoise: bad argument");
    PAX_REQUIRE(I0 > 0 && sd >= 0, "pax_ctnoise: need I0 > 0 and std >= 0");
    if (inner == 0) { inner = n; outer_stride = n; }
    PAX_REQUIRE(n % inner == 0 && outer_stride >= inner, "pax_ctnoise: bad block shape");
    k_ctnoise<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        proj, n, index0, inner, outer_stride, I0, mean, sd, max_dev, seed);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ---------------------------------------------------------------------------------- W
__global__ void __launch_bounds__(256) k_compute_w(const AngleDev *__restrict__ angles, int first,
                                                   int nv, int nu, float hz, float hy, float hx,
                                                   float thr, float *__restrict__ w)
{
    const int u = blockIdx.x * 32 + (threadIdx.x & 31);
    const int v = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int ia = blockIdx.z;
    if (u >= nu || v >= nv) return;
    const AngleDev A = angles[first + ia];
    const float L = pax_ray_chord(A, u, v, hx, hy, hz);
    w[((size_t)ia * nv + v) * nu + u] = L > thr ? 1.f / L : 0.f;
}

extern "C" int pax_compute_w(const PaxPlan *p, int first, int count, float hz, float hy, float hx,
                             float thr, float *w_out, void *stream)
{
    PAX_REQUIRE(p && w_out, "pax_compute_w: NULL argument");
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles,
                "pax_compute_w: angle range [%d,%d) outside plan (%d)", first, first + count, p->n_angles);
    dim3 grid(pax_cdiv(p->geo.nDetecU, 32), pax_cdiv(p->geo.nDetecV, 8), count);
    k_compute_w<<<grid, 256, 0, (cudaStream_t)stream>>>(p->d_angles, first, p->geo.nDetecV,
                                                        p->geo.nDetecU, hz, hy, hx, thr, w_out);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

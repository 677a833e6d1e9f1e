// "TIGRE-style" baseline kernels: the launch shapes of the toolbox the reference
// calls -- recalled, TIGRE is not vendored -- written for sm_100a as THE KERNELS TO BEAT, not as product:
//   Ax   one thread per detector pixel and projection, blocks of 9 x 9 pixels x 9 projections, per-projection
//        parameters in constant memory, the volume in a 3-D texture sampled with the HARDWARE trilinear filter
//        (8-bit interpolation weights: that is why the product path does not use it);
//   Atb  blocks of 16 x 32 voxel columns, 8 z voxels per thread, up to 32 projections per launch with their
//        parameters in constant memory, the projections in a layered 2-D texture with hardware bilinear filter.
// Same geometry frame as the product (PaxPlan / AngleDev), same lattice (n = ceil(|P-S|/accuracy) samples, the part
// of the ray inside the volume's box only -- a favour to the baseline).  bench.py --impl tigre_style times them
// beside the product kernels and reports their error against the float64 oracle.
#include <cmath>

#include "pax_internal.h"

#define TS_PPB 9                 // pixels per block edge / projections per block (Ax)
#define TS_ATB_PROJ 32           // projections per Atb launch
#define TS_ATB_ZT 8

struct TsAxProj {                // in (Z,Y,X) index space (voxel centre i at i), per projection
    float sx, sy, sz;
    float ox, oy, oz;            // pixel (v=0,u=0)
    float ux, uy, vz;
    float pad;
};
struct TsAtbProj {
    float cosv, sinv, DSD, DSO;
    float offOX, offOY, offOZ, offDU, offDV;
    float pad[3];
};
__constant__ TsAxProj c_ts_ax[TS_PPB * 8];
__constant__ TsAtbProj c_ts_atb[TS_ATB_PROJ];

struct PaxTs {
    const PaxPlan *plan;
    cudaArray_t vol_arr, proj_arr;
    cudaTextureObject_t vol_tex, proj_tex;
    int proj_layers;
};

__global__ void __launch_bounds__(TS_PPB * TS_PPB * TS_PPB)
k_ts_ax(cudaTextureObject_t tex, float *__restrict__ out, int nu, int nv, int nproj, float acc, float hix, float hiy,
        float hiz, float dX, float dY, float dZ)
{
    const int u = blockIdx.x * TS_PPB + threadIdx.x, v = blockIdx.y * TS_PPB + threadIdx.y;
    const int ip = blockIdx.z * TS_PPB + threadIdx.z;
    if (u >= nu || v >= nv || ip >= nproj) return;
    const TsAxProj P = c_ts_ax[ip];
    const float px = P.ox + u * P.ux, py = P.oy + u * P.uy, pz = P.oz + v * P.vz;
    float dx = px - P.sx, dy = py - P.sy, dz = pz - P.sz;
    const float len = sqrtf(dx * dx + dy * dy + dz * dz);
    const float n = ceilf(len / acc);
    dx /= n; dy /= n; dz /= n;
    // the part of the ray inside the open box (-1, N)^3
    float t0 = 0.f, t1 = n;
    bool hit = true;
    const float s[3] = {P.sx, P.sy, P.sz}, d[3] = {dx, dy, dz}, hi[3] = {hix, hiy, hiz};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        if (d[k] == 0.f) hit = hit && s[k] > -1.f && s[k] < hi[k];
        else {
            const float a = (-1.f - s[k]) / d[k], b = (hi[k] - s[k]) / d[k];
            t0 = fmaxf(t0, fminf(a, b)); t1 = fminf(t1, fmaxf(a, b));
        }
    }
    float sum = 0.f;
    if (hit && t0 <= t1) {
        const int i0 = (int)ceilf(t0), i1 = (int)floorf(t1);
        for (int i = i0; i <= i1; ++i) {
            const float fi = (float)i;
            sum += tex3D<float>(tex, fmaf(fi, dx, P.sx) + 0.5f, fmaf(fi, dy, P.sy) + 0.5f, fmaf(fi, dz, P.sz) + 0.5f);
        }
    }
    const float mx = dx * dX, my = dy * dY, mz = dz * dZ;
    out[((size_t)ip * nv + v) * nu + u] = sum * sqrtf(mx * mx + my * my + mz * mz);
}

__global__ void __launch_bounds__(16 * 32)
k_ts_atb(cudaTextureObject_t tex, float *__restrict__ vol, int nx, int ny, int nz, int nu, int nv, int nproj,
         float dX, float dY, float dZ, float x0, float y0, float z0, float rdU, float rdV, float cu, float cv,
         int accumulate)
{
    const int ix = blockIdx.x * 16 + threadIdx.x, iy = blockIdx.y * 32 + threadIdx.y, kz = blockIdx.z * TS_ATB_ZT;
    if (ix >= nx || iy >= ny) return;
    float acc[TS_ATB_ZT];
#pragma unroll
    for (int k = 0; k < TS_ATB_ZT; ++k) acc[k] = 0.f;
    const float xb = x0 + ix * dX, yb = y0 + iy * dY, zb = z0 + kz * dZ;
    for (int ip = 0; ip < nproj; ++ip) {
        const TsAtbProj A = c_ts_atb[ip];
        const float x = xb + A.offOX, y = yb + A.offOY, z = zb + A.offOZ;
        const float xr = fmaf(A.cosv, x, A.sinv * y), yr = fmaf(A.cosv, y, -A.sinv * x);
        const float rden = 1.f / (A.DSO - xr), t = A.DSD * rden;
        const float u = fmaf(fmaf(yr, t, -A.offDU), rdU, cu);
        const float wq = A.DSO * rden, w = wq * wq;
        const float vstep = dZ * t * rdV, vbase = fmaf(fmaf(z, t, -A.offDV), rdV, cv);
#pragma unroll
        for (int k = 0; k < TS_ATB_ZT; ++k)
            acc[k] = fmaf(w, tex2DLayered<float>(tex, u + 0.5f, fmaf((float)k, vstep, vbase) + 0.5f, ip), acc[k]);
    }
#pragma unroll
    for (int k = 0; k < TS_ATB_ZT; ++k)
        if (kz + k < nz) {
            float *p = vol + ((size_t)(kz + k) * ny + iy) * nx + ix;
            *p = accumulate ? *p + acc[k] : acc[k];
        }
}

extern "C" int pax_ts_create(const PaxPlan *p, PaxTs **out)
{
    PAX_REQUIRE(p && out, "pax_ts_create: NULL argument");
    PaxTs *t = new PaxTs();
    t->plan = p;
    t->vol_arr = t->proj_arr = nullptr;
    t->vol_tex = t->proj_tex = 0;
    t->proj_layers = 0;
    const PaxGeo &g = p->geo;
    cudaChannelFormatDesc ch = cudaCreateChannelDesc<float>();
    cudaError_t e = cudaMalloc3DArray(&t->vol_arr, &ch, make_cudaExtent(g.nVoxelX, g.nVoxelY, g.nVoxelZ));
    if (e == cudaSuccess)
        e = cudaMalloc3DArray(&t->proj_arr, &ch, make_cudaExtent(g.nDetecU, g.nDetecV, TS_ATB_PROJ), cudaArrayLayered);
    if (e == cudaSuccess) {
        cudaResourceDesc rd = {};
        rd.resType = cudaResourceTypeArray;
        cudaTextureDesc td = {};
        td.filterMode = cudaFilterModeLinear;
        td.readMode = cudaReadModeElementType;
        td.normalizedCoords = 0;
        td.addressMode[0] = td.addressMode[1] = td.addressMode[2] = cudaAddressModeBorder;
        rd.res.array.array = t->vol_arr;
        e = cudaCreateTextureObject(&t->vol_tex, &rd, &td, nullptr);
        if (e == cudaSuccess) {
            rd.res.array.array = t->proj_arr;
            e = cudaCreateTextureObject(&t->proj_tex, &rd, &td, nullptr);
        }
    }
    if (e != cudaSuccess) {
        if (t->vol_tex) cudaDestroyTextureObject(t->vol_tex);
        if (t->vol_arr) cudaFreeArray(t->vol_arr);
        if (t->proj_arr) cudaFreeArray(t->proj_arr);
        delete t;
        return pax_set_error(PAX_ERR_CUDA, "pax_ts_create: %s", cudaGetErrorString(e));
    }
    *out = t;
    return PAX_OK;
}

extern "C" int pax_ts_destroy(PaxTs *t)
{
    if (!t) return PAX_OK;
    if (t->vol_tex) cudaDestroyTextureObject(t->vol_tex);
    if (t->proj_tex) cudaDestroyTextureObject(t->proj_tex);
    if (t->vol_arr) cudaFreeArray(t->vol_arr);
    if (t->proj_arr) cudaFreeArray(t->proj_arr);
    delete t;
    return PAX_OK;
}

// (Z,Y,X) device volume -> the 3-D texture (TIGRE uploads the volume into a cudaArray the same way)
extern "C" int pax_ts_set_volume(PaxTs *t, const float *vol_zyx, void *stream)
{
    PAX_REQUIRE(t && vol_zyx, "pax_ts_set_volume: NULL argument");
    const PaxGeo &g = t->plan->geo;
    cudaMemcpy3DParms c = {};
    c.srcPtr = make_cudaPitchedPtr(const_cast<float *>(vol_zyx), g.nVoxelX * sizeof(float), g.nVoxelX, g.nVoxelY);
    c.dstArray = t->vol_arr;
    c.extent = make_cudaExtent(g.nVoxelX, g.nVoxelY, g.nVoxelZ);
    c.kind = cudaMemcpyDeviceToDevice;
    PAX_CHECK_CUDA(cudaMemcpy3DAsync(&c, (cudaStream_t)stream));
    return PAX_OK;
}

extern "C" int pax_ts_ax(PaxTs *t, int first, int count, float *proj_out, void *stream)
{
    PAX_REQUIRE(t && proj_out, "pax_ts_ax: NULL argument");
    const PaxPlan *p = t->plan;
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles, "pax_ts_ax: angle range");
    const PaxGeo &g = p->geo;
    cudaStream_t st = (cudaStream_t)stream;
    // constant memory holds 8 blocks' worth of projections: launch in chunks of 72
    for (int off = 0; off < count; off += TS_PPB * 8) {
        const int n = count - off < TS_PPB * 8 ? count - off : TS_PPB * 8;
        TsAxProj h[TS_PPB * 8];
        for (int i = 0; i < n; ++i) {
            const AngleDev &A = p->h_angles[first + off + i];
            // padded frame -> plain (Z,Y,X) index space
            h[i].sx = (float)(A.sx - 1.0); h[i].sy = (float)(A.sy - 1.0); h[i].sz = (float)(A.sz - PAX_ZOFF);
            h[i].ox = (float)(A.ox - 1.0); h[i].oy = (float)(A.oy - 1.0); h[i].oz = (float)(A.oz - PAX_ZOFF);
            h[i].ux = (float)A.ux; h[i].uy = (float)A.uy; h[i].vz = (float)A.vz; h[i].pad = 0.f;
        }
        PAX_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_ts_ax, h, sizeof(TsAxProj) * n, 0, cudaMemcpyHostToDevice, st));
        dim3 grid(pax_cdiv(g.nDetecU, TS_PPB), pax_cdiv(g.nDetecV, TS_PPB), pax_cdiv(n, TS_PPB));
        k_ts_ax<<<grid, dim3(TS_PPB, TS_PPB, TS_PPB), 0, st>>>(
            t->vol_tex, proj_out + (size_t)off * g.nDetecV * g.nDetecU, g.nDetecU, g.nDetecV, n, (float)g.accuracy,
            (float)g.nVoxelX, (float)g.nVoxelY, (float)g.nVoxelZ, (float)g.dVoxelX, (float)g.dVoxelY,
            (float)g.dVoxelZ);
        PAX_CHECK_CUDA(cudaStreamSynchronize(st));              // the host buffer h is reused by the next chunk
    }
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// Atb 'FDK' of projections [first, first+count) (device, (count,V,U)) into a (Z,Y,X) device volume
extern "C" int pax_ts_atb(PaxTs *t, const float *proj, int first, int count, float *vol_zyx, void *stream)
{
    PAX_REQUIRE(t && proj && vol_zyx, "pax_ts_atb: NULL argument");
    const PaxPlan *p = t->plan;
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles, "pax_ts_atb: angle range");
    const PaxGeo &g = p->geo;
    cudaStream_t st = (cudaStream_t)stream;
    for (int off = 0; off < count; off += TS_ATB_PROJ) {
        const int n = count - off < TS_ATB_PROJ ? count - off : TS_ATB_PROJ;
        cudaMemcpy3DParms c = {};
        c.srcPtr = make_cudaPitchedPtr(const_cast<float *>(proj + (size_t)off * g.nDetecV * g.nDetecU),
                                       g.nDetecU * sizeof(float), g.nDetecU, g.nDetecV);
        c.dstArray = t->proj_arr;
        c.extent = make_cudaExtent(g.nDetecU, g.nDetecV, n);
        c.kind = cudaMemcpyDeviceToDevice;
        PAX_CHECK_CUDA(cudaMemcpy3DAsync(&c, st));
        TsAtbProj h[TS_ATB_PROJ];
        for (int i = 0; i < n; ++i) {
            const AngleDev &A = p->h_angles[first + off + i];
            h[i].cosv = A.cosv; h[i].sinv = A.sinv; h[i].DSD = A.DSD; h[i].DSO = A.DSO;
            h[i].offOX = A.offOX; h[i].offOY = A.offOY; h[i].offOZ = A.offOZ; h[i].offDU = A.offDU; h[i].offDV = A.offDV;
        }
        PAX_CHECK_CUDA(cudaMemcpyToSymbolAsync(c_ts_atb, h, sizeof(TsAtbProj) * n, 0, cudaMemcpyHostToDevice, st));
        dim3 grid(pax_cdiv(g.nVoxelX, 16), pax_cdiv(g.nVoxelY, 32), pax_cdiv(g.nVoxelZ, TS_ATB_ZT));
        k_ts_atb<<<grid, dim3(16, 32), 0, st>>>(
            t->proj_tex, vol_zyx, g.nVoxelX, g.nVoxelY, g.nVoxelZ, g.nDetecU, g.nDetecV, n, (float)g.dVoxelX,
            (float)g.dVoxelY, (float)g.dVoxelZ, (float)(-0.5 * g.nVoxelX * g.dVoxelX + 0.5 * g.dVoxelX),
            (float)(-0.5 * g.nVoxelY * g.dVoxelY + 0.5 * g.dVoxelY),
            (float)(-0.5 * g.nVoxelZ * g.dVoxelZ + 0.5 * g.dVoxelZ), (float)(1.0 / g.dDetecU),
            (float)(1.0 / g.dDetecV), (float)(0.5 * g.nDetecU - 0.5), (float)(0.5 * g.nDetecV - 0.5), off > 0);
        PAX_CHECK_CUDA(cudaStreamSynchronize(st));
    }
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

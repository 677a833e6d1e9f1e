// Voxel-driven FDK-weighted backprojector (tigre.Atb 'FDK', reached through ossart at
// pCT_CBCT_Reconstruction.py:103; algorithm SURVEY.md A.4) with the OS-SART image update
// x = max(0, x + lambda * V^-1 * acc) fused into its epilogue, and set_v (A.5).
//
// Mapping.  One thread owns ZT consecutive z voxels of one (y,x) column: the detector
// column u, the magnification t and the distance weight w are computed once per angle
// and shared by the whole z run, whose row coordinate v advances by a constant.  The
// per-angle constants of the launch are staged in shared memory.  Accumulators stay in
// registers over all angles of the call; the resident volume is z-fastest, so every
// thread reads/writes one aligned run (float4 x ZT/4).
#include "pax_internal.h"

#define ATB_ZT 8
#define ATB_MAX_ANGLES PAX_ATB_SART_MAX

struct AtbAngle {
    float cosv, sinv, DSD, DSO;
    float offOX, offOY, offOZ, offDU, offDV;
};

struct AtbArgs {
    const AngleDev *angles;
    const float *proj;
    float *vol;
    const float *inv_v;
    int first, count;
    int nz, ny, nx, nv, nu;
    int nxp, nzp;
    float dX, dY, dZ;          // voxel pitch (already scaled for set_v)
    float x0, y0, z0;          // centre of voxel 0 before offOrigin: -s/2 + d/2
    float rdU, rdV, cu, cv;    // 1/dU, 1/dV, nU/2-0.5, nV/2-0.5
    float lambda;
    int mode, noneg;
    int row_part, row_parts;   // only the rows y = row_part (mod row_parts)
};

__device__ __forceinline__ void atb_load_angles(AtbAngle *sm, const AngleDev *angles, int first,
                                                int count)
{
    for (int i = threadIdx.x; i < count; i += blockDim.x) {
        const AngleDev &A = angles[first + i];
        AtbAngle t;
        t.cosv = A.cosv; t.sinv = A.sinv; t.DSD = A.DSD; t.DSO = A.DSO;
        t.offOX = A.offOX; t.offOY = A.offOY; t.offOZ = A.offOZ; t.offDU = A.offDU; t.offDV = A.offDV;
        sm[i] = t;
    }
    __syncthreads();
}

// bilinear fetch with zero border; iu in [-1,nu-1], iv in [-1,nv-1] guaranteed by the caller
__device__ __forceinline__ float atb_bilinear(const float *__restrict__ p, int nu, int nv, int iu,
                                              int iv, float fu, float fv)
{
    const bool u0 = iu >= 0, u1 = iu + 1 < nu, r0 = iv >= 0, r1 = iv + 1 < nv;
    const float *q = p + (long)iv * nu + iu;
    const float p00 = (r0 && u0) ? __ldg(q) : 0.f;
    const float p01 = (r0 && u1) ? __ldg(q + 1) : 0.f;
    const float p10 = (r1 && u0) ? __ldg(q + nu) : 0.f;
    const float p11 = (r1 && u1) ? __ldg(q + nu + 1) : 0.f;
    const float a = fmaf(fu, p01 - p00, p00), b = fmaf(fu, p11 - p10, p10);
    return fmaf(fv, b - a, a);
}

__global__ void __launch_bounds__(128) k_atb(const AtbArgs a)
{
    __shared__ AtbAngle sm[ATB_MAX_ANGLES];
    atb_load_angles(sm, a.angles, a.first, a.count);
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31);
    const int iy = (blockIdx.y * 4 + (threadIdx.x >> 5)) * a.row_parts + a.row_part;
    const int kz = blockIdx.z * ATB_ZT;
    if (ix >= a.nx || iy >= a.ny) return;

    float acc[ATB_ZT];
#pragma unroll
    for (int k = 0; k < ATB_ZT; ++k) acc[k] = 0.f;
    const float xb = a.x0 + ix * a.dX, yb = a.y0 + iy * a.dY, zb = a.z0 + kz * a.dZ;
    const size_t pstride = (size_t)a.nv * a.nu;
    const float unu = (float)a.nu, vnv = (float)a.nv;

    for (int ia = 0; ia < a.count; ++ia) {
        const AtbAngle A = sm[ia];
        const float x = xb + A.offOX, y = yb + A.offOY, z = zb + A.offOZ;
        const float xr = fmaf(A.cosv, x, A.sinv * y);
        const float yr = fmaf(A.cosv, y, -A.sinv * x);
        const float rden = 1.f / (A.DSO - xr);
        const float t = A.DSD * rden;
        const float u = fmaf(fmaf(yr, t, -A.offDU), a.rdU, a.cu);
        if (!(u > -1.f && u < unu)) continue;
        const float wq = A.DSO * rden;
        const float w = wq * wq;
        const float fu0 = floorf(u);
        const int iu = (int)fu0;
        const float fu = u - fu0;
        const float vstep = a.dZ * t * a.rdV;
        const float vbase = fmaf(fmaf(z, t, -A.offDV), a.rdV, a.cv);
        const float *__restrict__ p = a.proj + (size_t)ia * pstride;
#pragma unroll
        for (int k = 0; k < ATB_ZT; ++k) {
            const float v = fmaf((float)k, vstep, vbase);
            if (v > -1.f && v < vnv) {
                const float fv0 = floorf(v);
                acc[k] = fmaf(w, atb_bilinear(p, a.nu, a.nv, iu, (int)fv0, fu, v - fv0), acc[k]);
            }
        }
    }

    float *dst = a.vol + ((size_t)(iy + 1) * a.nxp + (ix + 1)) * a.nzp + PAX_ZOFF + kz;
    const int zr = a.nz - kz;     // valid z in this run
    float s = 1.f;
    if (a.mode == PAX_ATB_SART) s = a.lambda * __ldg(a.inv_v + (size_t)iy * a.nx + ix);
#pragma unroll
    for (int q = 0; q < ATB_ZT / 4; ++q) {
        float4 o = make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]);
        if (a.mode != PAX_ATB_STORE) {
            const float4 old = *reinterpret_cast<const float4 *>(dst + 4 * q);
            o.x = fmaf(s, o.x, old.x); o.y = fmaf(s, o.y, old.y);
            o.z = fmaf(s, o.z, old.z); o.w = fmaf(s, o.w, old.w);
            if (a.mode == PAX_ATB_SART && a.noneg) {
                o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
            }
        }
        const int left = zr - 4 * q;
        if (left >= 4) {
            *reinterpret_cast<float4 *>(dst + 4 * q) = o;
        } else if (left > 0) {     // never write into the zero rim
            dst[4 * q] = o.x;
            if (left > 1) dst[4 * q + 1] = o.y;
            if (left > 2) dst[4 * q + 2] = o.z;
        }
    }
}

extern "C" int pax_atb(const PaxPlan *p, const float *proj, int first, int count, float *vol_res,
                       const PaxAtbEpilogue *epi, void *stream)
{
    PAX_REQUIRE(p && proj && vol_res && epi, "pax_atb: NULL argument");
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles,
                "pax_atb: angle range [%d,%d) outside plan (%d angles)", first, first + count, p->n_angles);
    PAX_REQUIRE(epi->mode >= PAX_ATB_STORE && epi->mode <= PAX_ATB_SART, "pax_atb: bad epilogue mode %d",
                epi->mode);
    PAX_REQUIRE(epi->mode != PAX_ATB_SART || epi->inv_v, "pax_atb: SART mode needs inv_v");
    PAX_REQUIRE(epi->mode != PAX_ATB_SART || count <= ATB_MAX_ANGLES,
                "pax_atb: SART mode takes at most %d angles per call", ATB_MAX_ANGLES);
    const PaxGeo &g = p->geo;
    AtbArgs a;
    a.angles = p->d_angles; a.vol = vol_res; a.inv_v = epi->inv_v;
    a.nz = g.nVoxelZ; a.ny = g.nVoxelY; a.nx = g.nVoxelX; a.nv = g.nDetecV; a.nu = g.nDetecU;
    a.nxp = p->nxp; a.nzp = p->nzp;
    a.dX = (float)g.dVoxelX; a.dY = (float)g.dVoxelY; a.dZ = (float)g.dVoxelZ;
    a.x0 = (float)(-0.5 * g.nVoxelX * g.dVoxelX + 0.5 * g.dVoxelX);
    a.y0 = (float)(-0.5 * g.nVoxelY * g.dVoxelY + 0.5 * g.dVoxelY);
    a.z0 = (float)(-0.5 * g.nVoxelZ * g.dVoxelZ + 0.5 * g.dVoxelZ);
    a.rdU = (float)(1.0 / g.dDetecU); a.rdV = (float)(1.0 / g.dDetecV);
    a.cu = (float)(0.5 * g.nDetecU - 0.5); a.cv = (float)(0.5 * g.nDetecV - 0.5);
    a.lambda = epi->lambda; a.noneg = epi->noneg;
    a.row_parts = epi->row_parts > 1 ? epi->row_parts : 1;
    a.row_part = a.row_parts > 1 ? epi->row_part : 0;
    PAX_REQUIRE(a.row_part >= 0 && a.row_part < a.row_parts, "pax_atb: row share %d of %d", a.row_part, a.row_parts);
    const int rows = (g.nVoxelY - a.row_part + a.row_parts - 1) / a.row_parts;
    if (rows <= 0) return PAX_OK;
    dim3 grid(pax_cdiv(g.nVoxelX, 32), pax_cdiv(rows, 4), pax_cdiv(g.nVoxelZ, ATB_ZT));
    cudaStream_t st = (cudaStream_t)stream;
    for (int off = 0; off < count; off += ATB_MAX_ANGLES) {
        a.first = first + off;
        a.count = count - off < ATB_MAX_ANGLES ? count - off : ATB_MAX_ANGLES;
        a.proj = proj + (size_t)off * g.nDetecV * g.nDetecU;
        a.mode = (off == 0) ? epi->mode : PAX_ATB_ADD;
        k_atb<<<grid, 128, 0, st>>>(a);
    }
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ---------------------------------------------------------------------------------- V
// set_v (A.5 variant A): V(y,x) = mean_z Atb_FDK(ones) on the geometry whose dVoxel is
// scaled by vscale.  Bilinear interpolation of an all-ones image with a zero border is the
// product of two 1-D ramps, so no projection data is read.
__device__ __forceinline__ float ones_cov(float c, float n)
{
    if (!(c > -1.f && c < n)) return 0.f;
    if (c < 0.f) return c + 1.f;
    if (c > n - 1.f) return n - c;
    return 1.f;
}

__global__ void __launch_bounds__(128) k_compute_v(const AtbArgs a, float *__restrict__ v_out)
{
    const int ix = blockIdx.x * 32 + (threadIdx.x & 31);
    const int iy = blockIdx.y * 4 + (threadIdx.x >> 5);
    if (ix >= a.nx || iy >= a.ny) return;
    const float xb = a.x0 + ix * a.dX, yb = a.y0 + iy * a.dY;
    const float unu = (float)a.nu, vnv = (float)a.nv;
    float total = 0.f;
    for (int ia = 0; ia < a.count; ++ia) {
        const AngleDev &A = a.angles[a.first + ia];
        const float x = xb + A.offOX, y = yb + A.offOY;
        const float xr = fmaf(A.cosv, x, A.sinv * y);
        const float yr = fmaf(A.cosv, y, -A.sinv * x);
        const float rden = 1.f / (A.DSO - xr);
        const float t = A.DSD * rden;
        const float u = fmaf(fmaf(yr, t, -A.offDU), a.rdU, a.cu);
        const float cu = ones_cov(u, unu);
        if (cu == 0.f) continue;
        const float wq = A.DSO * rden;
        float sumv = 0.f;
        for (int k = 0; k < a.nz; ++k) {
            const float z = a.z0 + k * a.dZ + A.offOZ;
            const float v = fmaf(fmaf(z, t, -A.offDV), a.rdV, a.cv);
            sumv += ones_cov(v, vnv);
        }
        total = fmaf(wq * wq * cu, sumv, total);
    }
    v_out[(size_t)iy * a.nx + ix] = total / (float)a.nz;
}

extern "C" int pax_compute_v(const PaxPlan *p, int first, int count, double vscale, float *v_out,
                             void *stream)
{
    PAX_REQUIRE(p && v_out, "pax_compute_v: NULL argument");
    PAX_REQUIRE(first >= 0 && count >= 1 && first + count <= p->n_angles,
                "pax_compute_v: angle range [%d,%d) outside plan (%d)", first, first + count, p->n_angles);
    PAX_REQUIRE(vscale > 0, "pax_compute_v: vscale must be > 0");
    const PaxGeo &g = p->geo;
    AtbArgs a = {};
    a.angles = p->d_angles; a.first = first; a.count = count;
    a.nz = g.nVoxelZ; a.ny = g.nVoxelY; a.nx = g.nVoxelX; a.nv = g.nDetecV; a.nu = g.nDetecU;
    const double dX = g.dVoxelX * vscale, dY = g.dVoxelY * vscale, dZ = g.dVoxelZ * vscale;
    a.dX = (float)dX; a.dY = (float)dY; a.dZ = (float)dZ;
    a.x0 = (float)(-0.5 * g.nVoxelX * dX + 0.5 * dX);
    a.y0 = (float)(-0.5 * g.nVoxelY * dY + 0.5 * dY);
    a.z0 = (float)(-0.5 * g.nVoxelZ * dZ + 0.5 * dZ);
    a.rdU = (float)(1.0 / g.dDetecU); a.rdV = (float)(1.0 / g.dDetecV);
    a.cu = (float)(0.5 * g.nDetecU - 0.5); a.cv = (float)(0.5 * g.nDetecV - 0.5);
    dim3 grid(pax_cdiv(g.nVoxelX, 32), pax_cdiv(g.nVoxelY, 4));
    k_compute_v<<<grid, 128, 0, (cudaStream_t)stream>>>(a, v_out);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

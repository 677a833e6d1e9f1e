// Artifact-Induction image operations on the device: what the reference
// does with five ITK command-line tools and NRRD files between them
// (Physics-based-Augmentation/Artifact-Induction/Cpp-Codes/*, driven by CBCT_pCT_Artifact_adding.sh:45-145):
//   ResampleImages / ResampleImagesFloat / the resampling inside AddImages / the iso-z branch of
//   RescaleImage   -> pax_resample_linear   (itk::ResampleImageFilter, linear, identity transform)
//   RescaleImage                            -> pax_minmax + pax_rescale_intensity
//   AdaptiveHistogramMatch                  -> pax_minmax + pax_plahe  (power-law adaptive histogram
//                                              equalisation, itk::AdaptiveHistogramEqualizationImageFilter)
// Arrays are (Z,Y,X) C-order float32 (X fastest), like the NRRD files.  All three are HBM- or
// shared-memory-bound elementwise / stencil kernels; no tensor cores.
#include "pax_internal.h"

struct ResampleMap { double m[12]; };

// ((a0 x0 + a1 x1) + a2 x2) + c and acc + ((wx wy) wz) v, each operation rounded on its own
__device__ __forceinline__ double rsum3(double a0, double x0, double a1, double x1, double a2, double x2, double c)
{
    return __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(a0, x0), __dmul_rn(a1, x1)), __dmul_rn(a2, x2)), c);
}
__device__ __forceinline__ double rterm(double acc, double wx, double wy, double wz, float v)
{
    return __dadd_rn(acc, __dmul_rn(__dmul_rn(__dmul_rn(wx, wy), wz), (double)v));
}

// ---------------------------------------------------------------------------- resample
// One thread per destination voxel; coordinates and weights in double, like ITK (the short-pixel
// variant truncates the result, so an fp32 rounding at an integer would show).
__global__ void __launch_bounds__(256) k_resample_linear(const float *__restrict__ src, int snz, int sny,
                                                         int snx, const ResampleMap M,
                                                         float *__restrict__ dst, int dnz, int dny, int dnx,
                                                         float defval, int cast)
{
    const long n = (long)dnz * dny * dnx;
    for (long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const int i = (int)(e % dnx);
        const long r = e / dnx;
        const int j = (int)(r % dny), k = (int)(r / dny);
        // every product and sum rounded separately and in the oracle's order (no FMA contraction): the
        // short-pixel variant truncates the result, so a last-bit difference at an exact integer would show
        const double fi = (double)i, fj = (double)j, fk = (double)k;
        const double c0 = rsum3(M.m[0], fi, M.m[1], fj, M.m[2], fk, M.m[3]);
        const double c1 = rsum3(M.m[4], fi, M.m[5], fj, M.m[6], fk, M.m[7]);
        const double c2 = rsum3(M.m[8], fi, M.m[9], fj, M.m[10], fk, M.m[11]);
        double val = (double)defval;
        // itk::ImageFunction::IsInsideBuffer: [-0.5, n-0.5) per axis
        if (c0 >= -0.5 && c0 < snx - 0.5 && c1 >= -0.5 && c1 < sny - 0.5 && c2 >= -0.5 && c2 < snz - 0.5) {
            // LinearInterpolateImageFunction: base index raised to the first voxel (the then non-positive
            // distance means "no interpolation along this axis"), neighbour past the last voxel not read
            const int b0 = max((int)floor(c0), 0), b1 = max((int)floor(c1), 0), b2 = max((int)floor(c2), 0);
            const double d0 = fmax(__dsub_rn(c0, (double)b0), 0.0), d1 = fmax(__dsub_rn(c1, (double)b1), 0.0),
                         d2 = fmax(__dsub_rn(c2, (double)b2), 0.0);
            const double g0 = __dsub_rn(1.0, d0), g1 = __dsub_rn(1.0, d1), g2 = __dsub_rn(1.0, d2);
            const int e0 = min(b0 + 1, snx - 1), e1 = min(b1 + 1, sny - 1), e2 = min(b2 + 1, snz - 1);
            const float *p00 = src + ((size_t)b2 * sny + b1) * snx, *p01 = src + ((size_t)b2 * sny + e1) * snx;
            const float *p10 = src + ((size_t)e2 * sny + b1) * snx, *p11 = src + ((size_t)e2 * sny + e1) * snx;
            double acc = 0.0;
            acc = rterm(acc, g0, g1, g2, __ldg(p00 + b0));
            acc = rterm(acc, d0, g1, g2, __ldg(p00 + e0));
            acc = rterm(acc, g0, d1, g2, __ldg(p01 + b0));
            acc = rterm(acc, d0, d1, g2, __ldg(p01 + e0));
            acc = rterm(acc, g0, g1, d2, __ldg(p10 + b0));
            acc = rterm(acc, d0, g1, d2, __ldg(p10 + e0));
            acc = rterm(acc, g0, d1, d2, __ldg(p11 + b0));
            acc = rterm(acc, d0, d1, d2, __ldg(p11 + e0));
            val = acc;
        }
        if (cast == 1) val = trunc(fmin(fmax(val, -32768.0), 32767.0));    // static_cast<short> with bounds check
        dst[e] = (float)val;
    }
}

extern "C" int pax_resample_linear(const float *src, int snz, int sny, int snx, const double *index_map,
                                   float *dst, int dnz, int dny, int dnx, float default_value,
                                   int cast_short, void *stream)
{
    PAX_REQUIRE(src && dst && index_map, "pax_resample_linear: NULL argument");
    PAX_REQUIRE(snz >= 1 && sny >= 1 && snx >= 1 && dnz >= 1 && dny >= 1 && dnx >= 1,
                "pax_resample_linear: non-positive dimensions");
    ResampleMap M;
    for (int t = 0; t < 12; ++t) M.m[t] = index_map[t];
    const long n = (long)dnz * dny * dnx;
    const int blocks = (int)min((long)148 * 16, (n + 255) / 256);
    k_resample_linear<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, snz, sny, snx, M, dst, dnz, dny, dnx,
                                                               default_value, cast_short);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------------------------------------ min / max
__device__ __forceinline__ unsigned enc_f(float f)
{
    const unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void k_minmax_init(unsigned *w) { w[0] = 0xffffffffu; w[1] = 0u; }

__global__ void __launch_bounds__(256) k_minmax(const float *__restrict__ x, size_t n, unsigned *w)
{
    float lo = INFINITY, hi = -INFINITY;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const float v = __ldg(x + e);
        lo = fminf(lo, v);
        hi = fmaxf(hi, v);
    }
    for (int o = 16; o; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(w, enc_f(lo));
        atomicMax(w + 1, enc_f(hi));
    }
}

__global__ void k_minmax_decode(unsigned *w)
{
    const float lo = dec_f(w[0]), hi = dec_f(w[1]);
    reinterpret_cast<float *>(w)[0] = lo;
    reinterpret_cast<float *>(w)[1] = hi;
}

extern "C" int pax_minmax(const float *x, size_t n, float *out2_dev, void *stream)
{
    PAX_REQUIRE(x && out2_dev && n >= 1, "pax_minmax: bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    unsigned *w = reinterpret_cast<unsigned *>(out2_dev);
    k_minmax_init<<<1, 1, 0, st>>>(w);
    const int blocks = (int)min((size_t)148 * 8, (n + 255) / 256);
    k_minmax<<<blocks, 256, 0, st>>>(x, n, w);
    k_minmax_decode<<<1, 1, 0, st>>>(w);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// -------------------------------------------------------------------------------- rescale
// itk::RescaleIntensityImageFilter: out = (in + shift) * scale, clamped; arithmetic in double (ITK's
// RealType for float pixels).  RescaleImage.cxx:99-104.
__global__ void __launch_bounds__(256) k_rescale(float *__restrict__ x, size_t n, const float *__restrict__ mm,
                                                 double out_min, double out_max)
{
    const double lo = (double)mm[0], hi = (double)mm[1];
    const double scale = (hi != lo) ? (out_max - out_min) / (hi - lo) : 0.0;
    const double shift = (scale != 0.0) ? out_min / scale - lo : 0.0;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        double v = (scale != 0.0) ? ((double)x[e] + shift) * scale : out_min;
        v = fmin(fmax(v, out_min), out_max);
        x[e] = (float)v;
    }
}

extern "C" int pax_rescale_intensity(float *x, size_t n, const float *minmax_dev, double out_min,
                                     double out_max, void *stream)
{
    PAX_REQUIRE(x && minmax_dev && n >= 1, "pax_rescale_intensity: bad argument");
    const int blocks = (int)min((size_t)148 * 16, (n + 255) / 256);
    k_rescale<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, n, minmax_dev, out_min, out_max);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// --------------------------------------------------------------------------------- PL-AHE
// out = iscale * (0.5 + (1/N) sum_window F(u,v)) + min,  F = 0.5 s |2(u-v)|^alpha - beta (u-v) + beta u
//     = iscale * (0.5 + (0.5 S1 + beta S2) / N) + min,   S1 = sum s |2(u-v)|^alpha,  S2 = sum v
// (AdaptiveHistogramMatch.cxx:30-42 -> itk::AdaptiveHistogramEqualizationImageFilter).  A CTA computes
// an 8 x 8 x 16 block of outputs from a shared-memory tile with the radius-wide halo; a thread owns one
// output voxel and walks its (border-clipped) window.  Per term: one LDS and 3-4 flops, so the kernel is
// bound by shared-memory bandwidth; the x rows of the window are summed in fp32, the rows in double.
#define AHE_TX 16
#define AHE_TY 8
#define AHE_TZ 8
#define AHE_MAXR 5

template <int AMODE>   // 0: alpha == 0, 1: alpha == 0.5, 2: alpha == 1, 3: general
__global__ void __launch_bounds__(AHE_TX *AHE_TY *AHE_TZ)
    k_plahe(const float *__restrict__ src, int nz, int ny, int nx, int radius, float alpha, float beta,
            const float *__restrict__ mm, float *__restrict__ dst)
{
    constexpr int SX = AHE_TX + 2 * AHE_MAXR, SY = AHE_TY + 2 * AHE_MAXR, SZ = AHE_TZ + 2 * AHE_MAXR;
    __shared__ float tile[SZ][SY][SX + 1];
    const float lo = mm[0], hi = mm[1];
    const float iscale = hi - lo, rs = 1.f / iscale;
    const int x0 = blockIdx.x * AHE_TX, y0 = blockIdx.y * AHE_TY, z0 = blockIdx.z * AHE_TZ;
    const int tid = threadIdx.x;
    const int ex = AHE_TX + 2 * radius, ey = AHE_TY + 2 * radius, ez = AHE_TZ + 2 * radius;
    // normalised intensities of the tile + halo (voxels outside the image are never read back)
    for (int t = tid; t < ex * ey * ez; t += AHE_TX * AHE_TY * AHE_TZ) {
        const int lx = t % ex, ly = (t / ex) % ey, lz = t / (ex * ey);
        const int gx = x0 + lx - radius, gy = y0 + ly - radius, gz = z0 + lz - radius;
        float v = 0.f;
        if (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz)
            v = (__ldg(src + ((size_t)gz * ny + gy) * nx + gx) - lo) * rs - 0.5f;
        tile[lz][ly][lx] = v;
    }
    __syncthreads();
    const int tx = tid % AHE_TX, ty = (tid / AHE_TX) % AHE_TY, tz = tid / (AHE_TX * AHE_TY);
    const int gx = x0 + tx, gy = y0 + ty, gz = z0 + tz;
    if (gx >= nx || gy >= ny || gz >= nz) return;
    const float u = tile[tz + radius][ty + radius][tx + radius];
    // window clipped at the image border, in tile coordinates
    const int ax = max(gx - radius, 0) - x0 + radius, bx = min(gx + radius, nx - 1) - x0 + radius;
    const int ay = max(gy - radius, 0) - y0 + radius, by = min(gy + radius, ny - 1) - y0 + radius;
    const int az = max(gz - radius, 0) - z0 + radius, bz = min(gz + radius, nz - 1) - z0 + radius;
    double S1 = 0.0, S2 = 0.0;
    for (int lz = az; lz <= bz; ++lz)
        for (int ly = ay; ly <= by; ++ly) {
            float s1 = 0.f, s2 = 0.f;
            const float *row = &tile[lz][ly][0];
#pragma unroll 11
            for (int lx = ax; lx <= bx; ++lx) {
                const float v = row[lx];
                const float d = u - v;
                s2 += v;
                if (AMODE == 2) {
                    s1 += 2.f * d;
                } else {
                    const float sg = d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
                    if (AMODE == 0) s1 += sg;
                    else if (AMODE == 1) s1 += sg * sqrtf(fabsf(2.f * d));
                    else s1 += sg * powf(fabsf(2.f * d), alpha);
                }
            }
            S1 += (double)s1;
            S2 += (double)s2;
        }
    const double cnt = (double)((bx - ax + 1) * (by - ay + 1) * (bz - az + 1));
    dst[((size_t)gz * ny + gy) * nx + gx] =
        (float)((double)iscale * (0.5 + (0.5 * S1 + (double)beta * S2) / cnt) + (double)lo);
}

extern "C" int pax_plahe(const float *src, int nz, int ny, int nx, int radius, float alpha, float beta,
                         const float *minmax_dev, float *dst, void *stream)
{
    PAX_REQUIRE(src && dst && minmax_dev, "pax_plahe: NULL argument");
    PAX_REQUIRE(nz >= 1 && ny >= 1 && nx >= 1, "pax_plahe: non-positive dimensions");
    PAX_REQUIRE(radius >= 0 && radius <= AHE_MAXR, "pax_plahe: radius %d outside [0,%d]", radius, AHE_MAXR);
    PAX_REQUIRE(src != dst, "pax_plahe: in-place operation is not supported");
    dim3 grid(pax_cdiv(nx, AHE_TX), pax_cdiv(ny, AHE_TY), pax_cdiv(nz, AHE_TZ));
    const int threads = AHE_TX * AHE_TY * AHE_TZ;
    cudaStream_t st = (cudaStream_t)stream;
    if (alpha == 0.f) k_plahe<0><<<grid, threads, 0, st>>>(src, nz, ny, nx, radius, alpha, beta, minmax_dev, dst);
    else if (alpha == 0.5f) k_plahe<1><<<grid, threads, 0, st>>>(src, nz, ny, nx, radius, alpha, beta, minmax_dev, dst);
    else if (alpha == 1.f) k_plahe<2><<<grid, threads, 0, st>>>(src, nz, ny, nx, radius, alpha, beta, minmax_dev, dst);
    else k_plahe<3><<<grid, threads, 0, st>>>(src, nz, ny, nx, radius, alpha, beta, minmax_dev, dst);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------------------------ resize (output side)
// skimage.transform.resize(img, shape, order=0|1, preserve_range=True, anti_aliasing=False) as the reference's
// packing script calls it (preprocess/prep_test_data_pseudo_cbct_3D.py:23-36; scikit-image 0.17.2 per
// environment.yaml:124): for a 3-D array that is scipy.ndimage.map_coordinates at the pixel-centre aligned
// coordinates  c_d = (in_d / out_d) (o_d + 1/2) - 1/2  with mode 'mirror' (d c b | a b c d | c b a) -- order 0
// takes the sample at floor(c + 1/2), order 1 interpolates linearly.  Coordinates in double, like scipy.
__device__ __forceinline__ double rs_mirror(double c, int n)
{
    if (n <= 1) return 0.0;
    const double len = (double)(n - 1);
    if (c < 0.0) {                             // scipy's map_coordinate(), NI_EXTEND_MIRROR
        const double sz2 = 2.0 * len;
        c = sz2 * (double)(long long)(-c / sz2) + c;
        c = c <= -len ? c + sz2 : -c;
    } else if (c > len) {
        const double sz2 = 2.0 * len;
        c -= sz2 * (double)(long long)(c / sz2);
        if (c > len) c = sz2 - c;
    }
    return c;
}

__global__ void __launch_bounds__(256) k_resize(const float *__restrict__ src, int s0, int s1, int s2,
                                                float *__restrict__ dst, int d0, int d1, int d2, int order)
{
    const size_t n = (size_t)d0 * d1 * d2;
    const double f0 = (double)s0 / d0, f1 = (double)s1 / d1, f2 = (double)s2 / d2;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (size_t)gridDim.x * blockDim.x) {
        const int o2 = (int)(e % d2), o1 = (int)((e / d2) % d1), o0 = (int)(e / ((size_t)d1 * d2));
        const double c0 = rs_mirror(f0 * (o0 + 0.5) - 0.5, s0), c1 = rs_mirror(f1 * (o1 + 0.5) - 0.5, s1),
                     c2 = rs_mirror(f2 * (o2 + 0.5) - 0.5, s2);
        if (order == 0) {
            const int i0 = min(max((int)floor(c0 + 0.5), 0), s0 - 1), i1 = min(max((int)floor(c1 + 0.5), 0), s1 - 1),
                      i2 = min(max((int)floor(c2 + 0.5), 0), s2 - 1);
            dst[e] = __ldg(src + ((size_t)i0 * s1 + i1) * s2 + i2);
        } else {
            const int i0 = min((int)floor(c0), max(s0 - 2, 0)), i1 = min((int)floor(c1), max(s1 - 2, 0)),
                      i2 = min((int)floor(c2), max(s2 - 2, 0));
            const double t0 = c0 - i0, t1 = c1 - i1, t2 = c2 - i2;
            const int j0 = min(i0 + 1, s0 - 1), j1 = min(i1 + 1, s1 - 1), j2 = min(i2 + 1, s2 - 1);
            double acc = 0.0;                  // scipy sums the 2^3 taps with the products of the 1-D weights
            for (int a = 0; a < 2; ++a)
                for (int b = 0; b < 2; ++b)
                    for (int c = 0; c < 2; ++c) {
                        const double w = (a ? t0 : 1.0 - t0) * (b ? t1 : 1.0 - t1) * (c ? t2 : 1.0 - t2);
                        acc += w * (double)__ldg(src + ((size_t)(a ? j0 : i0) * s1 + (b ? j1 : i1)) * s2 + (c ? j2 : i2));
                    }
            dst[e] = (float)acc;
        }
    }
}

extern "C" int pax_resize(const float *src, int s0, int s1, int s2, float *dst, int d0, int d1, int d2, int order,
                          void *stream)
{
    PAX_REQUIRE(src && dst && src != dst, "pax_resize: NULL or aliased argument");
    PAX_REQUIRE(s0 >= 1 && s1 >= 1 && s2 >= 1 && d0 >= 1 && d1 >= 1 && d2 >= 1, "pax_resize: empty array");
    PAX_REQUIRE(order == 0 || order == 1, "pax_resize: order %d (0 = nearest, 1 = linear)", order);
    const size_t n = (size_t)d0 * d1 * d2;
    const int blocks = (int)((n + 255) / 256 < 148 * 32 ? (n + 255) / 256 : 148 * 32);
    k_resize<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, s0, s1, s2, dst, d0, d1, d2, order);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

// ------------------------------------------------------ image quality measures (Quameasopts)
// The sums TIGRE's Measure_Quality needs for RMSE, CC and UQI of two volumes, in one pass, in double:
// out[0..5] += { n, sum a, sum b, sum a^2, sum b^2, sum a b }.  (MSSIM is a windowed measure: quality.py.)
__global__ void __launch_bounds__(256) k_pair_sums(const float *__restrict__ a, const float *__restrict__ b,
                                                   size_t n, double *__restrict__ out)
{
    double s[5] = {0, 0, 0, 0, 0};
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const double x = a[i], y = b[i];
        s[0] += x; s[1] += y; s[2] += x * x; s[3] += y * y; s[4] += x * y;
    }
    __shared__ double sm[5][8];
    for (int k = 0; k < 5; ++k) {
        for (int o = 16; o; o >>= 1) s[k] += __shfl_xor_sync(0xffffffffu, s[k], o);
        if ((threadIdx.x & 31) == 0) sm[k][threadIdx.x >> 5] = s[k];
    }
    __syncthreads();
    if (threadIdx.x < 5) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += sm[threadIdx.x][w];
        atomicAdd(out + 1 + threadIdx.x, t);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(out, (double)n);
}

extern "C" int pax_pair_sums(const float *a, const float *b, size_t n, double *out6_dev, void *stream)
{
    PAX_REQUIRE(a && b && out6_dev && n > 0, "pax_pair_sums: bad argument");
    PAX_CHECK_CUDA(cudaMemsetAsync(out6_dev, 0, 6 * sizeof(double), (cudaStream_t)stream));
    const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
    k_pair_sums<<<blocks, 256, 0, (cudaStream_t)stream>>>(a, b, n, out6_dev);
    PAX_CHECK_CUDA(cudaGetLastError());
    return PAX_OK;
}

/*
 * oracle.c -- float64 CPU restatement of the sCBCT physics hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the checker for the CUDA product in
 * physics_arx_b200/csrc; it is imported by tests/, __graft_entry__.smoke() and
 * the cpu_baseline / --impl reference legs of bench.py, and by nothing else.
 * The product never routes through it.
 *
 * PARITY UNPINNED: the reference (nadeemlab/Physics-ArX) holds no tests, no
 * golden vectors and no CPU implementation of this path; its arithmetic lives
 * in the un-vendored, un-pinned TIGRE toolbox (CERN/TIGRE, "git clone" HEAD,
 * README.md:151-155 of the reference).  What follows restates TIGRE's published
 * algorithm as summarised in SURVEY.md Appendix A and is anchored on
 *   - the reference's call sites:
 *       Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py
 *       :67-83 (geometry), :86 (Ax 'interpolated'), :88 (CTnoise.add),
 *       :103 (ossart, blocksize=20)
 *   - analytic known-answer tests in tests/test_oracle_known_answers.py.
 *
 * Conventions: volume (Z,Y,X) C-order, projections (N,V,U),
 * mm and radians, world origin at the isocentre, rotation about z.
 *
 * Per-angle parameter row (8 doubles):
 *   [theta, DSD, DSO, offOrigZ, offOrigY, offOrigX, offDetV, offDetU]
 * Geometry block (11 doubles):
 *   [nZ, nY, nX, nV, nU, dZ, dY, dX, dV, dU, accuracy]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif
/* Thread count of the OpenMP loops below (bench.py's CPU legs set it explicitly: a launcher such as
 * torch.distributed.run exports OMP_NUM_THREADS=1).  Returns the count in effect. */
int oracle_set_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define G_NZ 0
#define G_NY 1
#define G_NX 2
#define G_NV 3
#define G_NU 4
#define G_DZ 5
#define G_DY 6
#define G_DX 7
#define G_DV 8
#define G_DU 9
#define G_ACC 10

#define A_TH 0
#define A_DSD 1
#define A_DSO 2
#define A_OZ 3
#define A_OY 4
#define A_OX 5
#define A_OV 6
#define A_OU 7
#define A_STRIDE 8

typedef struct {
    double sx, sy, sz;       /* source, index space */
    double ox, oy, oz;       /* pixel (v=0,u=0) centre, index space */
    double ux, uy, uz;       /* step per +1 in u */
    double vx, vy, vz;       /* step per +1 in v */
} RayFrame;

/* World (mm) -> index space where voxel centres sit at integers (A.3). */
static void ray_frame(const double *g, const double *a, RayFrame *f)
{
    const double th = a[A_TH], DSD = a[A_DSD], DSO = a[A_DSO];
    const double c = cos(th), s = sin(th);
    const double nZ = g[G_NZ], nY = g[G_NY], nX = g[G_NX];
    const double dZ = g[G_DZ], dY = g[G_DY], dX = g[G_DX];
    /* index = (world - off + sVoxel/2 - dVoxel/2) / dVoxel */
    const double bx = -a[A_OX] + 0.5 * nX * dX - 0.5 * dX;
    const double by = -a[A_OY] + 0.5 * nY * dY - 0.5 * dY;
    const double bz = -a[A_OZ] + 0.5 * nZ * dZ - 0.5 * dZ;
    /* source S = Rz(th) (DSO,0,0) */
    f->sx = (DSO * c + bx) / dX;
    f->sy = (DSO * s + by) / dY;
    f->sz = (0.0 + bz) / dZ;
    /* detector centre D = Rz(th) (-(DSD-DSO),0,0); u-hat = Rz(th)(0,1,0); v-hat = z */
    const double dcx = -(DSD - DSO) * c, dcy = -(DSD - DSO) * s;
    const double uhx = -s, uhy = c;
    const double u0 = (0.0 - 0.5 * g[G_NU] + 0.5) * g[G_DU] + a[A_OU];
    const double v0 = (0.0 - 0.5 * g[G_NV] + 0.5) * g[G_DV] + a[A_OV];
    f->ox = (dcx + u0 * uhx + bx) / dX;
    f->oy = (dcy + u0 * uhy + by) / dY;
    f->oz = (v0 + bz) / dZ;
    f->ux = g[G_DU] * uhx / dX;
    f->uy = g[G_DU] * uhy / dY;
    f->uz = 0.0;
    f->vx = 0.0;
    f->vy = 0.0;
    f->vz = g[G_DV] / dZ;
}

/* trilinear sample, zero outside (CUDA "border" addressing; A.3) */
static inline double tri(const double *vol, long nz, long ny, long nx,
                         double x, double y, double z)
{
    if (!(x > -1.0 && y > -1.0 && z > -1.0 && x < nx && y < ny && z < nz)) return 0.0;
    const double fx0 = floor(x), fy0 = floor(y), fz0 = floor(z);
    const long ix = (long)fx0, iy = (long)fy0, iz = (long)fz0;
    const double fx = x - fx0, fy = y - fy0, fz = z - fz0;
    double acc = 0.0;
    for (int dz = 0; dz < 2; ++dz) {
        const long kz = iz + dz;
        if (kz < 0 || kz >= nz) continue;
        const double wz = dz ? fz : 1.0 - fz;
        for (int dy = 0; dy < 2; ++dy) {
            const long ky = iy + dy;
            if (ky < 0 || ky >= ny) continue;
            const double wy = dy ? fy : 1.0 - fy;
            const double *row = vol + (kz * ny + ky) * nx;
            double r = 0.0;
            if (ix >= 0 && ix < nx) r += (1.0 - fx) * row[ix];
            if (ix + 1 >= 0 && ix + 1 < nx) r += fx * row[ix + 1];
            acc += wz * wy * r;
        }
    }
    return acc;
}

/* conservative range of sample indices i (real-valued) with S+i*d inside (-1,N) on one axis */
static inline void slab(double s, double d, double lo, double hi, double *t0, double *t1)
{
    if (d == 0.0) {
        if (s <= lo || s >= hi) { *t0 = 1.0; *t1 = 0.0; }
        return;
    }
    double a = (lo - s) / d, b = (hi - s) / d;
    if (a > b) { double t = a; a = b; b = t; }
    if (a > *t0) *t0 = a;
    if (b < *t1) *t1 = b;
}

/*
 * Ax 'interpolated' (A.3; reference call site pCT_CBCT_Reconstruction.py:86).
 * nsamples (optional, may be NULL) receives per-angle counts of lattice samples
 * that lie inside the open box (-1,N)^3: the "Ax update" unit of SURVEY.md 8(d).
 */
void oracle_ax_interp(const double *g, const double *ang, long nang,
                      const double *vol, double *proj, double *nsamples)
{
    const long nz = (long)g[G_NZ], ny = (long)g[G_NY], nx = (long)g[G_NX];
    const long nv = (long)g[G_NV], nu = (long)g[G_NU];
    const double acc = g[G_ACC];
    for (long ia = 0; ia < nang; ++ia) {
        RayFrame f;
        ray_frame(g, ang + ia * A_STRIDE, &f);
        double count = 0.0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : count)
        for (long v = 0; v < nv; ++v) {
            for (long u = 0; u < nu; ++u) {
                const double px = f.ox + u * f.ux + v * f.vx;
                const double py = f.oy + u * f.uy + v * f.vy;
                const double pz = f.oz + u * f.uz + v * f.vz;
                double dx = px - f.sx, dy = py - f.sy, dz = pz - f.sz;
                const double len = sqrt(dx * dx + dy * dy + dz * dz);
                const double n = ceil(len / acc);
                dx /= n; dy /= n; dz /= n;
                double t0 = 0.0, t1 = n;
                slab(f.sx, dx, -1.0, (double)nx, &t0, &t1);
                slab(f.sy, dy, -1.0, (double)ny, &t0, &t1);
                slab(f.sz, dz, -1.0, (double)nz, &t0, &t1);
                double sum = 0.0;
                if (t0 <= t1) {
                    long i0 = (long)floor(t0) - 1, i1 = (long)ceil(t1) + 1;
                    if (i0 < 0) i0 = 0;
                    if (i1 > (long)n) i1 = (long)n;
                    for (long i = i0; i <= i1; ++i) {
                        const double x = f.sx + i * dx, y = f.sy + i * dy, z = f.sz + i * dz;
                        if (x > -1.0 && y > -1.0 && z > -1.0 && x < nx && y < ny && z < nz) {
                            sum += tri(vol, nz, ny, nx, x, y, z);
                            count += 1.0;
                        }
                    }
                }
                const double mx = dx * g[G_DX], my = dy * g[G_DY], mz = dz * g[G_DZ];
                proj[(ia * nv + v) * nu + u] = sum * sqrt(mx * mx + my * my + mz * mz);
            }
        }
        if (nsamples) nsamples[ia] = count;
    }
}

/* bilinear sample of one projection, zero outside (A.4) */
static inline double bil(const double *p, long nv, long nu, double v, double u)
{
    if (!(u > -1.0 && v > -1.0 && u < nu && v < nv)) return 0.0;
    const double fu0 = floor(u), fv0 = floor(v);
    const long iu = (long)fu0, iv = (long)fv0;
    const double fu = u - fu0, fv = v - fv0;
    double acc = 0.0;
    for (int dv = 0; dv < 2; ++dv) {
        const long kv = iv + dv;
        if (kv < 0 || kv >= nv) continue;
        const double wv = dv ? fv : 1.0 - fv;
        const double *row = p + kv * nu;
        double r = 0.0;
        if (iu >= 0 && iu < nu) r += (1.0 - fu) * row[iu];
        if (iu + 1 >= 0 && iu + 1 < nu) r += fu * row[iu + 1];
        acc += wv * r;
    }
    return acc;
}

/*
 * Atb 'FDK' (A.4): voxel-driven, bilinear, weight (DSO/(DSO-x'))^2, plain sum
 * over angles.  proj==NULL means "all ones with zero border" (used for V).
 * vscale multiplies dVoxel (the shrunken geometry of set_v, A.5); 1.0 otherwise.
 */
void oracle_atb_fdk(const double *g, const double *ang, long nang,
                    const double *proj, double *vol, double vscale)
{
    const long nz = (long)g[G_NZ], ny = (long)g[G_NY], nx = (long)g[G_NX];
    const long nv = (long)g[G_NV], nu = (long)g[G_NU];
    const double dZ = g[G_DZ] * vscale, dY = g[G_DY] * vscale, dX = g[G_DX] * vscale;
    double *cs = (double *)malloc(sizeof(double) * 2 * nang);
    for (long ia = 0; ia < nang; ++ia) {
        cs[2 * ia] = cos(ang[ia * A_STRIDE + A_TH]);
        cs[2 * ia + 1] = sin(ang[ia * A_STRIDE + A_TH]);
    }
#pragma omp parallel for collapse(2) schedule(static)
    for (long k = 0; k < nz; ++k) {
        for (long j = 0; j < ny; ++j) {
            for (long i = 0; i < nx; ++i) {
                double accum = 0.0;
                for (long ia = 0; ia < nang; ++ia) {
                    const double *a = ang + ia * A_STRIDE;
                    const double x = -0.5 * nx * dX + (i + 0.5) * dX + a[A_OX];
                    const double y = -0.5 * ny * dY + (j + 0.5) * dY + a[A_OY];
                    const double z = -0.5 * nz * dZ + (k + 0.5) * dZ + a[A_OZ];
                    const double c = cs[2 * ia], s = cs[2 * ia + 1];
                    const double xr = c * x + s * y;      /* Rz(-theta) r */
                    const double yr = -s * x + c * y;
                    const double den = a[A_DSO] - xr;
                    const double t = a[A_DSD] / den;
                    const double u = (yr * t - a[A_OU]) / g[G_DU] + 0.5 * nu - 0.5;
                    const double v = (z * t - a[A_OV]) / g[G_DV] + 0.5 * nv - 0.5;
                    const double w = (a[A_DSO] / den) * (a[A_DSO] / den);
                    double val;
                    if (proj) {
                        val = bil(proj + ia * nv * nu, nv, nu, v, u);
                    } else {
                        /* bilinear of an all-ones image with zero border */
                        double cu = 1.0, cv = 1.0;
                        if (!(u > -1.0 && u < nu)) cu = 0.0;
                        else if (u < 0.0) cu = u + 1.0;
                        else if (u > nu - 1.0) cu = nu - u;
                        if (!(v > -1.0 && v < nv)) cv = 0.0;
                        else if (v < 0.0) cv = v + 1.0;
                        else if (v > nv - 1.0) cv = nv - v;
                        val = cu * cv;
                    }
                    accum += w * val;
                }
                vol[(k * ny + j) * nx + i] = accum;
            }
        }
    }
    free(cs);
}

/*
 * W (A.5): chord length in mm of the segment source->pixel through an
 * axis-aligned box of half-sizes (hz,hy,hx) centred at offOrigin.  TIGRE gets
 * this as a Siddon projection of a 2x2x2 volume of ones.  out = 1/len, or 0
 * where len <= thr.
 */
void oracle_raylen_w(const double *g, const double *ang, long nang,
                     double hz, double hy, double hx, double thr, double *w)
{
    const long nv = (long)g[G_NV], nu = (long)g[G_NU];
    for (long ia = 0; ia < nang; ++ia) {
        const double *a = ang + ia * A_STRIDE;
        const double c = cos(a[A_TH]), s = sin(a[A_TH]);
        const double DSD = a[A_DSD], DSO = a[A_DSO];
        const double sx = DSO * c - a[A_OX], sy = DSO * s - a[A_OY], sz = -a[A_OZ];
#pragma omp parallel for schedule(static)
        for (long v = 0; v < nv; ++v) {
            for (long u = 0; u < nu; ++u) {
                const double uu = (u - 0.5 * nu + 0.5) * g[G_DU] + a[A_OU];
                const double vv = (v - 0.5 * nv + 0.5) * g[G_DV] + a[A_OV];
                const double px = -(DSD - DSO) * c - uu * s - a[A_OX];
                const double py = -(DSD - DSO) * s + uu * c - a[A_OY];
                const double pz = vv - a[A_OZ];
                const double dx = px - sx, dy = py - sy, dz = pz - sz;
                double t0 = 0.0, t1 = 1.0;
                const double lo[3] = {-hx, -hy, -hz}, hi[3] = {hx, hy, hz};
                const double S[3] = {sx, sy, sz}, D[3] = {dx, dy, dz};
                int miss = 0;
                for (int k = 0; k < 3; ++k) {
                    if (D[k] == 0.0) {
                        if (S[k] < lo[k] || S[k] > hi[k]) miss = 1;
                        continue;
                    }
                    double ta = (lo[k] - S[k]) / D[k], tb = (hi[k] - S[k]) / D[k];
                    if (ta > tb) { double t = ta; ta = tb; tb = t; }
                    if (ta > t0) t0 = ta;
                    if (tb < t1) t1 = tb;
                }
                double len = 0.0;
                if (!miss && t1 > t0) len = (t1 - t0) * sqrt(dx * dx + dy * dy + dz * dz);
                w[(ia * nv + v) * nu + u] = (len > thr) ? 1.0 / len : 0.0;
            }
        }
    }
}

/* ---------------- counter-based RNG: Philox4x32-10 (Salmon et al., SC'11) ---------------- */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1)
{
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

/* raw block function, for the Random123 known-answer vectors */
void oracle_philox_raw(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
    uint32_t c[4] = {ctr[0], ctr[1], ctr[2], ctr[3]};
    philox4x32_10(c, key[0], key[1]);
    memcpy(out, c, sizeof(uint32_t) * 4);
}

void oracle_philox(uint64_t seed, uint64_t index, uint32_t out[4])
{
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), 0u, 0u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(out, c, sizeof(uint32_t) * 4);
}

#define POISSON_INVERSION_MAX 512.0
static const double TWO_PI = 6.283185307179586476925286766559;

/*
 * CTnoise.add (A.7; reference call site pCT_CBCT_Reconstruction.py:88):
 *   I = I0 exp(-p/m);  I <- Poisson(I) + N(mean, std);  I<=0 -> 1e-6;  p' = -log(I/I0) m
 * with m = max(p) supplied by the caller.  The reference draws from NumPy's
 * unseeded global RNG, so the stream here is this build's own specification:
 * one Philox4x32-10 block per pixel keyed by (seed, global pixel index);
 * words 0,1 -> Box-Muller normal / inversion uniform for the Poisson draw,
 * words 2,3 -> Box-Muller normal for the electronic noise.  Poisson: exact
 * sequential inversion for lambda < 512, rounded normal approximation above.
 */
void oracle_ctnoise(const double *p, long n, uint64_t index0, double I0, double mean,
                    double std, double m, uint64_t seed, double *out)
{
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) {
        uint32_t r[4];
        oracle_philox(seed, index0 + (uint64_t)e, r);
        const double u1 = ((double)r[0] + 0.5) * (1.0 / 4294967296.0);
        const double u2 = ((double)r[1] + 0.5) * (1.0 / 4294967296.0);
        const double u3 = ((double)r[2] + 0.5) * (1.0 / 4294967296.0);
        const double u4 = ((double)r[3] + 0.5) * (1.0 / 4294967296.0);
        const double lam = I0 * exp(-p[e] / m);
        double k;
        if (lam < POISSON_INVERSION_MAX) {
            double pk = exp(-lam), F = pk;
            k = 0.0;
            while (u1 > F && k < 4096.0) { k += 1.0; pk *= lam / k; F += pk; }
        } else {
            const double zp = sqrt(-2.0 * log(u1)) * cos(TWO_PI * u2);
            k = floor(lam + sqrt(lam) * zp + 0.5);
            if (k < 0.0) k = 0.0;
        }
        const double zg = sqrt(-2.0 * log(u3)) * cos(TWO_PI * u4);
        double I = k + mean + std * zg;
        if (I <= 0.0) I = 1e-6;
        out[e] = -log(I / I0) * m;
    }
}

/* x <- max(0, x + lambda * invV(y,x) * upd)   (A.6 update_image + clip) */
void oracle_sart_update(double *x, const double *upd, const double *invv, long nz, long nyx,
                        double lambda, int noneg)
{
#pragma omp parallel for schedule(static)
    for (long k = 0; k < nz; ++k)
        for (long j = 0; j < nyx; ++j) {
            double val = x[k * nyx + j] + lambda * invv[j] * upd[k * nyx + j];
            if (noneg && val < 0.0) val = 0.0;
            x[k * nyx + j] = val;
        }
}

/* ======================================================================================
 * Artifact-Induction image operations: the ITK command-line
 * tools of Physics-based-Augmentation/Artifact-Induction/Cpp-Codes, restated.  ITK itself
 * is not vendored under /root/reference (the tools only #include it), so these follow the
 * published behaviour of the ITK classes the tools instantiate: PARITY UNPINNED, anchored
 * on the tools' own call sequences and on the known answers in
 * tests/test_artifact_induction.py.
 * Arrays are (Z,Y,X) C-order (X fastest), like the NRRD files the tools exchange.
 * ====================================================================================== */

/*
 * itk::ResampleImageFilter + itk::LinearInterpolateImageFunction with the identity transform
 * and default pixel value 0, as instantiated by ResampleImages.cxx:87-102,
 * ResampleImagesFloat.cxx (same, float pixels), AddImages.cxx:40-53 and the iso-z branch of
 * RescaleImage.cxx:44-92.
 *   m[12]: continuous source index (x,y,z) of destination index (i,j,k):
 *          cx = m[0] i + m[1] j + m[2] k + m[3], cy = m[4..7], cz = m[8..11]
 *          (= Dsrc^-1 (Odst + Ddst diag(sdst) idx - Osrc) / ssrc, formed by the caller in double)
 * Inside test: ITK's IsInsideBuffer on the continuous index, [-0.5, n-0.5) per axis.  Inside,
 * LinearInterpolateImageFunction clamps: a base index below the first voxel is raised to it and
 * the (then non-positive) distance means "no interpolation along this axis"; a neighbour past
 * the last voxel is not read (the lower value is returned).  That is trilinear interpolation
 * with edge clamping.  cast: 0 = float, 1 = short (ResampleImageFilter casts the double result
 * with a bounds check and static_cast, i.e. truncation toward zero).
 */
void oracle_resample_linear(const double *src, long snz, long sny, long snx, const double *m,
                            double *dst, long dnz, long dny, long dnx, double defval, int cast)
{
#pragma omp parallel for schedule(static) collapse(2)
    for (long k = 0; k < dnz; ++k)
        for (long j = 0; j < dny; ++j)
            for (long i = 0; i < dnx; ++i) {
                const double c[3] = {m[0] * i + m[1] * j + m[2] * k + m[3],
                                     m[4] * i + m[5] * j + m[6] * k + m[7],
                                     m[8] * i + m[9] * j + m[10] * k + m[11]};
                const long n[3] = {snx, sny, snz};
                double val = defval;
                int inside = 1;
                for (int a = 0; a < 3; ++a)
                    if (!(c[a] >= -0.5 && c[a] < (double)n[a] - 0.5)) inside = 0;
                if (inside) {
                    long b[3], b1[3];
                    double d[3];
                    for (int a = 0; a < 3; ++a) {
                        b[a] = (long)floor(c[a]);
                        if (b[a] < 0) b[a] = 0;
                        d[a] = c[a] - (double)b[a];
                        if (d[a] < 0.0) d[a] = 0.0;
                        b1[a] = b[a] + 1;
                        if (b1[a] > n[a] - 1) { b1[a] = b[a]; }
                    }
                    double acc = 0.0;
                    for (int dz = 0; dz < 2; ++dz)
                        for (int dy = 0; dy < 2; ++dy)
                            for (int dx = 0; dx < 2; ++dx) {
                                const double w = (dx ? d[0] : 1.0 - d[0]) * (dy ? d[1] : 1.0 - d[1]) *
                                                 (dz ? d[2] : 1.0 - d[2]);
                                acc += w * src[((dz ? b1[2] : b[2]) * sny + (dy ? b1[1] : b[1])) * snx +
                                               (dx ? b1[0] : b[0])];
                            }
                    val = acc;
                }
                if (cast == 1) {
                    if (val < -32768.0) val = -32768.0;
                    if (val > 32767.0) val = 32767.0;
                    val = (double)(short)val;             /* static_cast: toward zero */
                }
                dst[(k * dny + j) * dnx + i] = val;
            }
}

/*
 * itk::RescaleIntensityImageFilter (RescaleImage.cxx:99-104): out = (in + shift) * scale with
 * scale = (outMax-outMin)/(inMax-inMin), shift = outMin/scale - inMin, clamped to [outMin,outMax];
 * a constant image maps to outMin... ITK sets scale = 0 then, giving outMin.
 */
void oracle_rescale_intensity(const double *in, long n, double out_min, double out_max, double *out)
{
    double lo = in[0], hi = in[0];
    for (long e = 1; e < n; ++e) { if (in[e] < lo) lo = in[e]; if (in[e] > hi) hi = in[e]; }
    const double scale = (hi != lo) ? (out_max - out_min) / (hi - lo) : 0.0;
    const double shift = (scale != 0.0) ? out_min / scale - lo : 0.0;
#pragma omp parallel for schedule(static)
    for (long e = 0; e < n; ++e) {
        double v = (scale != 0.0) ? (in[e] + shift) * scale : out_min;
        if (v < out_min) v = out_min;
        if (v > out_max) v = out_max;
        out[e] = v;
    }
}

/*
 * itk::AdaptiveHistogramEqualizationImageFilter (AdaptiveHistogramMatch.cxx:30-42; radius 5 and
 * (alpha,beta) in {0,0.5,1}^2 per CBCT_pCT_Artifact_adding.sh:51-87): Stark's power-law adaptive
 * histogram equalisation.  With u, v = intensities normalised to [-0.5,0.5] by the image min/max,
 *   out = iscale * (0.5 + (1/N) sum_{v in window} F(u,v)) + min,
 *   F(u,v) = 0.5 s |2(u-v)|^alpha - beta 0.5 s |2(u-v)| + beta u,   s = sgn(u-v) (0 at u == v),
 * N = number of window voxels inside the image (the window is clipped at the border).
 */
void oracle_plahe(const double *in, long nz, long ny, long nx, long radius, double alpha, double beta,
                  double *out)
{
    const long n = nz * ny * nx;
    double lo = in[0], hi = in[0];
    for (long e = 1; e < n; ++e) { if (in[e] < lo) lo = in[e]; if (in[e] > hi) hi = in[e]; }
    const double iscale = hi - lo;
#pragma omp parallel for schedule(dynamic, 1) collapse(2)
    for (long k = 0; k < nz; ++k)
        for (long j = 0; j < ny; ++j)
            for (long i = 0; i < nx; ++i) {
                const double u = (in[(k * ny + j) * nx + i] - lo) / iscale - 0.5;
                const long k0 = k - radius < 0 ? 0 : k - radius, k1 = k + radius >= nz ? nz - 1 : k + radius;
                const long j0 = j - radius < 0 ? 0 : j - radius, j1 = j + radius >= ny ? ny - 1 : j + radius;
                const long i0 = i - radius < 0 ? 0 : i - radius, i1 = i + radius >= nx ? nx - 1 : i + radius;
                double sum = 0.0;
                for (long kk = k0; kk <= k1; ++kk)
                    for (long jj = j0; jj <= j1; ++jj)
                        for (long ii = i0; ii <= i1; ++ii) {
                            const double v = (in[(kk * ny + jj) * nx + ii] - lo) / iscale - 0.5;
                            const double d = u - v;
                            const double s = d > 0.0 ? 1.0 : (d < 0.0 ? -1.0 : 0.0);
                            const double ad = fabs(2.0 * d);
                            sum += 0.5 * s * pow(ad, alpha) - beta * 0.5 * s * ad + beta * u;
                        }
                const double cnt = (double)((k1 - k0 + 1) * (j1 - j0 + 1) * (i1 - i0 + 1));
                out[(k * ny + j) * nx + i] = iscale * (sum / cnt + 0.5) + lo;
            }
}

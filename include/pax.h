/*
 * pax.h -- C ABI of libpax.so: the B200 (sm_100a) sCBCT physics path.
 *
 * Drop-in boundary for the one hot path of nadeemlab/Physics-ArX: cone-beam
 * forward projection (Ax), scatter/noise injection, FDK-weighted backprojection
 * (Atb) and the OS-SART update with its W / V weights.  The reference reaches
 * this arithmetic through five TIGRE calls in
 *   Physics-based-Augmentation/OSSART-Reconstruction/pCT_CBCT_Reconstruction.py
 * (:67 geometry_default, :86 Ax, :88 CTnoise.add, :103 ossart).  TIGRE's own
 * native boundary (Cython _Ax_ext/_Atb_ext -> interpolation_projection(),
 * voxel_backprojection(); SURVEY.md 8(b), recalled -- TIGRE is not vendored) is
 * host-pointer in / host-pointer out and blocking.  This boundary differs on
 * purpose: DEVICE pointers owned by the caller, asynchronous on the caller's
 * stream, nothing allocated per call, so that a whole OS-SART run stays resident
 * in HBM.  The ctypes binding a maintainer would add is in INTEGRATION.md.
 *
 * Conventions: volumes (Z,Y,X) C-order float32, projections (N,V,U) C-order
 * float32, mm and radians.  Every function returns
 * PAX_OK (0) or a negative PaxStatus; the message for the calling thread's last
 * failure is pax_last_error().  No C++ exception crosses this boundary.
 */
#ifndef PAX_H_
#define PAX_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAX_VERSION 100

typedef enum PaxStatus {
    PAX_OK = 0,
    PAX_ERR_INVALID = -1,   /* bad argument (NULL, range, unsupported geometry) */
    PAX_ERR_CUDA = -2,      /* a CUDA runtime call or launch failed */
    PAX_ERR_NOMEM = -3
} PaxStatus;

/* geo.nVoxel / dVoxel / nDetector / dDetector / accuracy of the reference's
 * geometry bag (pCT_CBCT_Reconstruction.py:67-83).  sVoxel = nVoxel*dVoxel. */
typedef struct PaxGeo {
    int32_t nVoxelZ, nVoxelY, nVoxelX;
    int32_t nDetecV, nDetecU;
    double dVoxelZ, dVoxelY, dVoxelX;   /* mm */
    double dDetecV, dDetecU;            /* mm */
    double accuracy;                    /* ray step in voxel units (TIGRE default 0.5) */
} PaxGeo;

/* One row per projection angle, the per-angle broadcast TIGRE's check_geo makes. */
#define PAX_ANGLE_STRIDE 8
/* [theta, DSD, DSO, offOrigZ, offOrigY, offOrigX, offDetV, offDetU] (doubles) */

typedef struct PaxPlan PaxPlan;   /* per-geometry device constants; opaque */

/* Epilogue of the forward projector.  P0/P1 = line integrals of volume channel 0/1,
 * L = chord length (mm) of the ray through the W box.
 *   PAX_AX_PLAIN    out = a0*P0 + a1*P1 + c*L      (a1 ignored without channel 1)
 *   PAX_AX_RESIDUAL out = W * (b - P0), W = 1/L (0 where L <= w_thr)   [OS-SART, A.6]
 * The scatter transfer of the reference (AddImages.cxx:57-62 adds the artifact
 * volume to the planning CT before projecting) is PLAIN with a0 = 1, a1 = scale. */
#define PAX_AX_PLAIN 0
#define PAX_AX_RESIDUAL 1
typedef struct PaxAxEpilogue {
    int32_t mode;
    float a0, a1, c;
    const float *b;                 /* device, (count,V,U): measured projections of these angles */
    float w_half_z, w_half_y, w_half_x;   /* W box half sizes, mm (geo.w_box) */
    float w_thr;                    /* chord lengths <= w_thr give W = 0 */
    double *sumsq;                  /* optional device scalar: += sum (b-P0)^2 (computel2) */
    unsigned long long *nsamples;   /* optional device scalar: += trilinear samples taken */
    float *out1;                    /* optional (PLAIN, two channels): P1 goes here, (count,V,U), and
                                       out = a0*P0 + c*L -- both line integrals from ONE pass, to fan
                                       several scatter scales out of it afterwards (pax_lincomb) */
    uint32_t *maxout;               /* optional device word (PLAIN): running max of out, order-encoded;
                                       zero it first, finish with pax_max_decode -- the normalising
                                       max of CTnoise.add fused into the projection epilogue */
} PaxAxEpilogue;

/* What the backprojector does with the accumulated sum  acc = sum_angles w * bilinear(proj):
 *   PAX_ATB_STORE   vol  = acc
 *   PAX_ATB_ADD     vol += acc
 *   PAX_ATB_SART    vol  = max(0?, vol + lambda * inv_v[y,x] * acc)     [A.6 update + clip] */
#define PAX_ATB_STORE 0
#define PAX_ATB_ADD 1
#define PAX_ATB_SART 2
#define PAX_ATB_SART_MAX 64         /* most angles of one PAX_ATB_SART call (the kernel's angle table);
                                       STORE / ADD take any count */
#define PAX_MAX_PEERS 16            /* most ranks pax_sart_update_fused serves */
typedef struct PaxAtbEpilogue {
    int32_t mode;
    int32_t noneg;
    float lambda;
    const float *inv_v;             /* device (Y,X): 1/V_s, 0 where V_s == 0 */
    int32_t row_part, row_parts;    /* row_parts > 1: only the volume rows y = row_part (mod row_parts) are
                                       computed and written (multi-GPU: every rank backprojects ALL angles of the
                                       subset into its own rows); 0 / 0 or 0 / 1: every row */
} PaxAtbEpilogue;

int pax_version(void);
const char *pax_last_error(void);

/* replaces tigre's per-call geometry marshalling (Ax.py/Atb.py -> _Ax_ext/_Atb_ext) */
int pax_plan_create(const PaxGeo *geo, const double *angle_rows, int n_angles, PaxPlan **out);
int pax_plan_destroy(PaxPlan *plan);
int pax_plan_n_angles(const PaxPlan *plan);

/* Which forward-projector kernel pax_ax runs.  Both evaluate the same sums on the same sample
 * lattice (tigre.Ax 'interpolated') and agree to fp32 rounding:
 *   RAY  one thread per ray, one trilinear gather per sample (csrc/ax.cu)
 *   FAN  the rays of a detector column share their in-plane interpolation; a row touches the shared
 *        prefix sums only where it crosses a slice boundary (csrc/ax_fan_ev.cu); pays off when the
 *        detector rows are finer than the slices
 *   AUTO (default) FAN when the geometry makes the sharing worthwhile, else RAY.
 * pax_plan_projector returns the kernel AUTO resolves to (or the forced one). */
#define PAX_PROJECTOR_AUTO 0
#define PAX_PROJECTOR_RAY 1
#define PAX_PROJECTOR_FAN 2
int pax_plan_set_projector(PaxPlan *plan, int kind);
int pax_plan_projector(const PaxPlan *plan);
/* tuning facts of the FAN kernel for this geometry: detector rows whose z range fits the 32 lanes of a
 * warp over one epoch in the worst case (0 = the kernel does not fit), warps per CTA, mean length in rows
 * of the runs that share one lattice */
int pax_plan_fan_info(const PaxPlan *plan, int *rows_per_tile, int *warps, double *mean_run);
/* Multi-GPU: restrict the FAN kernel of this plan to share `part` of `parts` of the detector's column tiles
 * (dealt out by decreasing chord length, so the shares cost the same).  pax_ax then reads / writes only the
 * columns of its share; the others keep what the caller put there (zeros, for the backprojection of a
 * partial residual to be exact).  parts = 1 restores the whole detector. */
int pax_plan_set_column_share(PaxPlan *plan, int part, int parts);
/* self-check of the FAN kernel: 1 if any launch so far met a single detector row whose z range over an
 * epoch is wider than its 32 lanes (the epoch length is derived so that this cannot happen; the results
 * of such a launch are wrong); synchronises the device */
int pax_plan_fault(const PaxPlan *plan);

/* Resident volume layout ("z-column": (Y+2, X+2, Zp) floats, z fastest, zero rim).
 * pax_vol_elems = floats per channel.  pack/unpack convert from/to (Z,Y,X). */
size_t pax_vol_elems(const PaxPlan *plan);
int pax_vol_pack(const PaxPlan *plan, const float *vol_zyx, float *vol_res, void *stream);
int pax_vol_unpack(const PaxPlan *plan, const float *vol_res, float *vol_zyx, void *stream);

/* y += a*x on two volumes in resident layout: the reference's image-domain scatter transfer
 * (AddImages.cxx:57-62, pCT + artifact) when only one variation is wanted. */
int pax_vol_axpy(const PaxPlan *plan, float *y_res, const float *x_res, float a, void *stream);

/* tigre.Ax(img, geo, angles, 'interpolated')  -- reference :86; angles [first, first+count)
 * of the plan; vol_res1 may be NULL (single channel). proj_out is (count,V,U).
 * workspace, RAY kernel (pax_ax_workspace_bytes is 0 when the plan resolves to FAN): device
 * scratch of pax_ax_workspace_bytes(plan, channels) bytes, 16-byte aligned, that receives the
 * gather-friendly re-packing of the volume(s) for this call; NULL selects the slower ray kernel
 * that gathers from the resident layout directly.
 * workspace, FAN kernel: optional device scratch of pax_ax_fan_scratch_bytes(plan, count) bytes (0: none
 * wanted).  A launch too small to fill the GPU a few times over -- a rank's few angles in a multi-GPU
 * run -- is then cut into up to four parts along the chords, whose raw sums go to the scratch and are
 * added up by a second, elementwise kernel; NULL keeps the single launch.
 * Calls on ONE plan must be ordered (one stream, or events between streams): the FAN kernel hands the
 * detector columns to its warps through counters that belong to the plan. */
size_t pax_ax_workspace_bytes(const PaxPlan *plan, int n_channels);
size_t pax_ax_fan_scratch_bytes(const PaxPlan *plan, int count);
int pax_ax(const PaxPlan *plan, const float *vol_res0, const float *vol_res1, int first, int count,
           float *proj_out, const PaxAxEpilogue *epi, void *workspace, void *stream);

/* tigre.Atb(proj, geo, angles, 'FDK')  -- reached via ossart, reference :103.
 * proj is (count,V,U) for angles [first, first+count); vol_res is in resident layout. */
int pax_atb(const PaxPlan *plan, const float *proj, int first, int count, float *vol_res,
            const PaxAtbEpilogue *epi, void *stream);

/* x = max(0?, x + lambda*inv_v*upd): the update when upd was summed across GPUs first */
int pax_sart_update(const PaxPlan *plan, float *x_res, const float *upd_res, const float *inv_v,
                    float lambda, int noneg, void *stream);

/* The same update FUSED with its multi-GPU collective: every rank holds a partial update
 * in a symmetric (peer-mapped) buffer; rank r of `world` reduces the r-th slice of the volume over all ranks,
 * applies the update and writes the new x slice into every rank's iterate -- one kernel over NVLink peer
 * memory instead of an all-reduce plus a local update.  Give either the multicast addresses of the two
 * symmetric buffers (x_mc, upd_mc: NVSwitch multimem.ld_reduce / multimem.st) or the `world` peer pointers of
 * each (x_peers, upd_peers: host arrays of device addresses).  The caller synchronises the ranks before (all
 * partial updates complete) and after (all slices landed); every rank ends with bit-identical x. */
/* Multi-GPU plumbing over symmetric (peer-mapped) buffers: copy `n_chunks` chunks of `chunk_bytes` (a multiple of
 * 16) -- chunk numbers first_chunk, first_chunk + chunk_stride, ... -- from this rank's buffer to the same place
 * in EVERY rank's buffer, either through the buffer's multicast address (dst_mc: one multimem.st per 16 bytes, the
 * NVSwitch replicates) or through the `world` peer pointers (dst_peers).  Used to all-gather the residual
 * projections of a subset (chunk = one projection) and to broadcast the volume rows a rank has updated (chunk =
 * one row of the resident volume, stride = world).  The caller synchronises the ranks around it. */
int pax_bcast_chunks(const void *src_local, void *dst_mc, const uint64_t *dst_peers, int world, size_t chunk_bytes,
                     long first_chunk, long chunk_stride, long n_chunks, void *stream);
/* The general form: the chunks are read from `src` starting at chunk src_first_chunk (-1: from the same place
 * as in the destination), and with add != 0 they are ADDED to what every rank's buffer holds (multimem.red --
 * ranks that each computed some columns of a projection, zeros elsewhere, merge them into a zeroed slot). */
int pax_bcast_chunks2(const void *src, long src_first_chunk, void *dst_mc, const uint64_t *dst_peers, int world,
                      size_t chunk_bytes, long first_chunk, long chunk_stride, long n_chunks, int add, void *stream);

/* The same for the iterate: the rows y = part (mod parts) of this rank's resident volume to every rank, skipping
 * the columns with inv_v == 0 (not seen by the subset, hence not updated).  After pax_atb(..., PAX_ATB_SART,
 * row_part/row_parts) has updated exactly those rows. */
int pax_bcast_xrows(const PaxPlan *plan, const float *x_local, float *x_mc, const uint64_t *x_peers, int world,
                    const float *inv_v, int part, int parts, void *stream);

int pax_sart_update_fused(const PaxPlan *plan, float *x_local, float *x_mc, const float *upd_mc,
                          const uint64_t *x_peers, const uint64_t *upd_peers, int rank, int world,
                          const float *inv_v, float lambda, int noneg, void *stream);

/* IterativeReconAlg.set_w / set_v (A.5).  w_out (count,V,U); v_out (Y,X) = mean_z Atb_FDK(ones)
 * on the geometry with dVoxel scaled by vscale. */
int pax_compute_w(const PaxPlan *plan, int first, int count, float half_z, float half_y,
                  float half_x, float thr, float *w_out, void *stream);
int pax_compute_v(const PaxPlan *plan, int first, int count, double vscale, float *v_out,
                  void *stream);

/* out = a*x + b*y over n floats (out may alias x or y): the projection-domain scatter combine
 * P = P_ct + s*P_art for one variation of a batch */
int pax_lincomb(float *out, const float *x, const float *y, float a, float b, size_t n, void *stream);

/* max over n floats -> *out_dev (device float) */
int pax_reduce_max(const float *x, size_t n, float *out_dev, void *stream);
/* turn the order-encoded running max of PaxAxEpilogue.maxout into a float, in place */
int pax_max_decode(uint32_t *word_dev, void *stream);

/* tigre.utilities.CTnoise.add -- reference :88.  In place on n floats.  Noise is a function of
 * (seed, global element index) only, so it does not depend on how projections are sharded:
 * local element e has global index  index0 + (e / inner) * outer_stride + (e % inner)
 * (an [n/inner][inner] block cut out of the global array; inner = 0 means contiguous).
 * *max_dev = max(proj) over the WHOLE projection set. */
int pax_ctnoise(float *proj, size_t n, uint64_t index0, size_t inner, uint64_t outer_stride, double I0,
                double mean, double std, const float *max_dev, uint64_t seed, void *stream);

/* ---- Artifact-Induction image operations (the reference's ITK command-line tools,
 * Physics-based-Augmentation/Artifact-Induction/Cpp-Codes, driven by CBCT_pCT_Artifact_adding.sh:45-145).
 * Arrays are (Z,Y,X) C-order float32 device pointers, X fastest, like the NRRD files the tools exchange. */

/* itk::ResampleImageFilter with LinearInterpolateImageFunction, identity transform, default pixel value:
 * ResampleImages.cxx:87-102 (cast_short = 1: short pixels, the result is truncated like ITK's static_cast),
 * ResampleImagesFloat.cxx, AddImages.cxx:40-53, RescaleImage.cxx:44-92 (iso-z branch).
 * index_map: 12 host doubles, continuous SOURCE index (x,y,z) of DESTINATION index (i,j,k):
 * cx = m0 i + m1 j + m2 k + m3, cy = m4.., cz = m8..  (from both images' origin/spacing/direction). */
int pax_resample_linear(const float *src, int snz, int sny, int snx, const double *index_map, float *dst,
                        int dnz, int dny, int dnx, float default_value, int cast_short, void *stream);
/* out2_dev[0] = min, out2_dev[1] = max over n floats (device result) */
int pax_minmax(const float *x, size_t n, float *out2_dev, void *stream);
/* itk::RescaleIntensityImageFilter to [out_min, out_max], in place; RescaleImage.cxx:99-104 */
int pax_rescale_intensity(float *x, size_t n, const float *minmax_dev, double out_min, double out_max,
                          void *stream);
/* itk::AdaptiveHistogramEqualizationImageFilter(radius, alpha, beta): AdaptiveHistogramMatch.cxx:30-42.
 * minmax_dev = pax_minmax of src; radius <= 5 (the reference uses 5); dst != src. */
int pax_plahe(const float *src, int nz, int ny, int nx, int radius, float alpha, float beta,
              const float *minmax_dev, float *dst, void *stream);

/* Output side.
 * pax_resize: skimage.transform.resize(order 0 | 1, preserve_range, no anti-aliasing) of a 3-D array, as the
 * reference's packing script calls it (preprocess/prep_test_data_pseudo_cbct_3D.py:23-36): pixel-centre aligned
 * coordinates, 'mirror' boundary, double arithmetic.  src (s0,s1,s2) -> dst (d0,d1,d2), both device, C order.
 * pax_pair_sums: out6 = { n, sum a, sum b, sum a^2, sum b^2, sum a b } in double -- what RMSE / CC / UQI of two
 * volumes need (the Quameasopts of ossart, pCT_CBCT_Reconstruction.py:99-103). */
int pax_resize(const float *src, int s0, int s1, int s2, float *dst, int d0, int d1, int d2, int order, void *stream);
int pax_pair_sums(const float *a, const float *b, size_t n, double *out6_dev, void *stream);

/* "TIGRE-style" baseline kernels -- benchmark reference only, not on the product path (csrc/tigre_style.cu): the
 * launch shapes of the toolbox the reference calls (9x9x9 hardware-trilinear texture Ax, 16x32x8z constant-memory
 * texture Atb), compiled for sm_100a as the kernels to beat.  Plain (Z,Y,X) device volumes in and out.  The handle
 * owns a 3-D and a layered CUDA array (these entry points DO allocate, unlike the product's). */
typedef struct PaxTs PaxTs;
int pax_ts_create(const PaxPlan *plan, PaxTs **out);
int pax_ts_destroy(PaxTs *ts);
int pax_ts_set_volume(PaxTs *ts, const float *vol_zyx, void *stream);
int pax_ts_ax(PaxTs *ts, int first, int count, float *proj_out, void *stream);
int pax_ts_atb(PaxTs *ts, const float *proj, int first, int count, float *vol_zyx, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* PAX_H_ */

/* icebin_b200.h -- C ABI of the B200-native IceBin regridding backend (libicebin_b200.so).
 *
 * Every entry point replaces one interface of the reference (citibeth/icebin); the reference file:line each
 * one stands in for is cited beside it.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * Conventions
 *   - Every function returns 0 on success and a negative code on error; ib_last_error() gives the message
 *     (the reference calls (*icebin_error)(-1, fmt, ...) which throws; slib/icebin/error.hpp:22-32).
 *   - "host" pointers are ordinary CPU memory; functions ending in _dev take CUDA device pointers and a
 *     cudaStream_t (passed as void*; NULL = the library's own stream) and do not synchronise.
 *   - Sparse indices are int64 (eigen_types.hpp: sparse_index_type = long), dense indices int32, values fp64.
 *   - Fields ("coupling variables") are stored CELL-MAJOR: X[cell * ld + field], ld >= nvar.  This is the
 *     transpose of the reference's [nvar x ncell] Blitz view (IceCoupler.cpp:231-233); IB_LAYOUT_FIELD_MAJOR
 *     selects the reference layout (costs one transposition on the device).
 *   - There is NO CPU fallback: without a CUDA device every compute entry point fails with IB_ERR_CUDA.
 */
#ifndef ICEBIN_B200_H
#define ICEBIN_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IB_OK 0
#define IB_ERR_ARG (-1)      /* bad argument / unknown spec (std::map::at throw, RegridMatrices_Dynamic.cpp:419) */
#define IB_ERR_CUDA (-2)     /* CUDA runtime failure or no device */
#define IB_ERR_ELEV (-3)     /* "Elevation %g out of bounds" (IceRegridder_L0.cpp:84-85) */
#define IB_ERR_AREA (-4)     /* "Area of cell %ld must be non-zero" (smoother.cpp:88-90) */
#define IB_ERR_DIM (-5)      /* dimension mismatch (hntr.hpp:399-406) */

#define IB_LAYOUT_CELL_MAJOR 0
#define IB_LAYOUT_FIELD_MAJOR 1

#define IB_Z_INTERP 0            /* InterpStyle::Z_INTERP          (IceRegridder.hpp:42-48) */
#define IB_ELEV_CLASS_INTERP 1   /* InterpStyle::ELEV_CLASS_INTERP */

#define IB_SRC_F64 0
#define IB_SRC_I16 1

typedef struct ib_gcm ib_gcm;             /* GCMRegridder_Standard        (GCMRegridder.hpp:312-400) */
typedef struct ib_rm ib_rm;               /* RegridMatrices_Dynamic       (RegridMatrices_Dynamic.hpp:20-61) */
typedef struct ib_dim ib_dim;             /* SparseSetT = spsparse::SparseSet<long,int> (eigen_types.hpp:24) */
typedef struct ib_weighted ib_weighted;   /* ibmisc::linear::Weighted_Eigen (RegridMatrices_Dynamic.cpp:63,167,258) */
typedef struct ib_hntr ib_hntr;           /* icebin::modele::Hntr         (modele/hntr.hpp:63-184) */

const char *ib_last_error(void);
/* number of visible CUDA devices (0 => every compute call will fail) and device selection for this thread */
int ib_device_count(void);
int ib_set_device(int device);
/* number of kernels this library has launched so far in this process (bench.py's gpu_launches) */
int64_t ib_launch_count(void);
/* Performance knobs; results never depend on them.  Keys: "stream_bps", "tma_bps" (CTAs per SM of the one-pass / ring
 * SpMM kernels, 1..8), "stream_spare" (CTA slots the one-pass kernel leaves free -- default 24: faster than a grid that needs every slot, and room for a concurrent small kernel -- 0..64), "gc_nq" (64-field groups per lane of the distinct-column kernel: 0 = by nvar, 1, 2, 4),
 * "use_stream", "use_gc", "use_group" (0 = route around the one-pass / distinct-column / tensor-core group kernel), "hntr_generic" (1 = Hntr
 * always takes its generic kernels), "xchg_blocks" (128-thread blocks of the two ib_xchg_* kernels, default 96: set it to the CTA slots a concurrent apply leaves
 * free -- "stream_spare" -- or the exchange waits for that apply's CTAs to exit). */
int ib_tune(const char *key, int value);
/* Return the device memory the library's stream-ordered allocations cached (cudaMemPoolTrimTo); synchronises the device. */
int ib_trim(void);

/* Device buffers for hosts that do not link the CUDA runtime themselves (include/icebin/coupler_b200.hpp keeps the
 * coupling fields in HBM between the variable transformation and the regrid applies).  Stream-ordered on the library's
 * own stream; upload/download return when the copy is complete. */
int ib_dev_alloc(void **out, int64_t bytes);
void ib_dev_free(void *p);
int ib_dev_upload(void *dst_dev, const void *src_host, int64_t bytes);
int ib_dev_download(void *dst_host, const void *src_dev, int64_t bytes);

/* ---- GCMRegridder_Standard ------------------------------------------------------------------------- */
/* GCMRegridder_Standard::init(AbbrGrid&&, hcdefs, indexingHC, correctA)  GCMRegridder.hpp:346-352,
 * with indexingHC fixed to {A,HC} / HC slowest, iE = ihc*nA + iA          pylib/icebin_cython.cpp:56-73.
 * dimA[n_real] = realised A cells (dense->sparse), native_area[n_real]    AbbrGrid.hpp:103-109 */
int ib_gcm_create(ib_gcm **out, int64_t nA, int32_t n_real, const int64_t *dimA, const double *native_area,
                  int32_t nhc, const double *hcdefs, int correctA);
void ib_gcm_free(ib_gcm *);
/* add_sheet(name, IceRegridder) + IceRegridder::init(...)   GCMRegridder.hpp:357-370, IceRegridder.cpp:81-119.
 * ex_ij[2*nX] interleaved (iA, iI), ex_area[nX]             AbbrGrid.hpp:40-78
 * gridA_proj_area[n_real] is an input (PROJ is upstream of the hot path).  centroid_xy may be NULL (no smoothing).
 * Returns the sheet index (>= 0). */
int ib_gcm_add_sheet(ib_gcm *, const char *name, int64_t nI, int32_t n_real_I, const int64_t *dimI,
                     const double *native_areaI, const double *centroid_xy, int64_t nX, const int32_t *ex_ij,
                     const double *ex_area, const double *gridA_proj_area, int interp_style);
int64_t ib_gcm_nA(const ib_gcm *);                 /* GCMRegridder.hpp:268 */
int64_t ib_gcm_nE(const ib_gcm *);                 /* :269 */
int32_t ib_gcm_nhc(const ib_gcm *);                /* :265-266 */
int64_t ib_gcm_nI(const ib_gcm *, int sheet);      /* :272 */
int64_t ib_gcm_nX(const ib_gcm *, int sheet);      /* IceRegridder.hpp nX() */
int ib_gcm_nsheets(const ib_gcm *);
int ib_gcm_sheet_index(const ib_gcm *, const char *name);   /* ice_regridders().index.at(name) */
int ib_gcm_correctA(const ib_gcm *);
/* GCMRegridder_Standard::filter_cellsA(keepA)  GCMRegridder.hpp:380-389, GCMRegridder.cpp:154-164 (domain decomposition:
 * every MPI rank / GPU keeps its own GCM cells).  keepA[nA]: nonzero = keep.  Exchange cells of dropped GCM cells are
 * removed from every sheet (the survivors keep their order and are renumbered, IceRegridder.cpp:170-176) and the GCM
 * grid's realised set shrinks.  Matrices already built are unaffected; regrid_matrices() built afterwards see the
 * restricted grids. */
int ib_gcm_filter_cellsA(ib_gcm *, const uint8_t *keepA);
int64_t ib_gcm_nA_realized(const ib_gcm *);        /* agridA.dim.dense_extent() */
/* GCMRegridder::wA(accum, sheet, native)  GCMRegridder.hpp:304-314: index[n_real], area[n_real] */
int ib_gcm_wA(const ib_gcm *, int sheet, int native, int64_t *index, double *area);

/* ---- RegridMatrices -------------------------------------------------------------------------------- */
/* GCMRegridder_Standard::regrid_matrices(sheet_index, elevmaskI, RegridParams)  RegridMatrices_Dynamic.cpp:334-402.
 * elevmaskI[nI] (NaN = masked) is COPIED (:352).  sigma may be NULL (= {0,0,0}).   RegridMatrices.hpp:17-36 */
int ib_regrid_matrices(ib_gcm *, int sheet, const double *elevmaskI, int scale, int correctA, const double *sigma,
                       ib_rm **out);
int ib_regrid_matrices_dev(ib_gcm *, int sheet, const double *d_elevmaskI, int scale, int correctA,
                           const double *sigma, void *stream, ib_rm **out);
void ib_rm_free(ib_rm *);

/* SparseSetT: ordered sparse<->dense bijection.  ids = sparse ids in dense order. */
int ib_dim_create(ib_dim **out);                                             /* SparseSetT() */
int ib_dim_create_ids(ib_dim **out, int64_t sparse_extent, int64_t n, const int64_t *ids);
int ib_dim_create_identity(ib_dim **out, int64_t n);                         /* id_sparse_set(n), IceCoupler.cpp:366 */
int64_t ib_dim_dense_extent(const ib_dim *);
int64_t ib_dim_sparse_extent(const ib_dim *);
int ib_dim_get_ids(const ib_dim *, int64_t *ids);                            /* to_sparse(0..dense_extent) */
void ib_dim_free(ib_dim *);

/* RegridMatrices::matrix(spec): fresh dims owned by the result          RegridMatrices_Dynamic.cpp:425-437
 * spec in {AvI IvA AvX XvA EvI IvE EvX XvE EvA AvE}                     :363-390 */
int ib_rm_matrix(ib_rm *, const char *spec, ib_weighted **out);
/* RegridMatrices_Dynamic::matrix_d(spec, {dim0,dim1}, params): borrows the caller's dims and APPENDS   :412-423 */
int ib_rm_matrix_d(ib_rm *, const char *spec, ib_dim *dim0, ib_dim *dim1, int scale, int correctA,
                   const double *sigma, ib_weighted **out);
void ib_w_free(ib_weighted *);

/* Extension (no upstream counterpart): wrap an existing dense-indexed CSR matrix + weights (e.g. one read back from
 * a file written by Weighted_Eigen::ncio, matrix_formats.rst:9-63) so it can be applied on the GPU.  Identity dims. */
int ib_w_create_csr(ib_weighted **out, int64_t nrow, int64_t ncol, const int32_t *rowptr, const int32_t *col,
                    const double *val, const double *wM, const double *Mw, int conservative);

/* The device step of GCMRegridder_ModelE's mismatched regridding (modele/GCMRegridder_ModelE.cpp:285-307 crop_mvp, :367-386,
 * :418-432; icebin_b200/modele.py holds the small GCM-side factors): re-index one side of a matrix -- side 0: row r becomes
 * map[r], side 1: column c becomes map[c]; map < 0 drops the entry -- scale every entry by factor[old index] (NULL: 1) and by
 * other_scale[index on the other side] (NULL: 1; applied first, as the reference scales the large matrix before the product),
 * and add up the entries that now coincide, in storage order.  n_new: dense extent of the re-indexed side.  The result
 * has identity dims and zero weights: ib_w_set_dim / ib_w_set_weights complete it. */
int ib_w_aggregate(const ib_weighted *, int side, const int32_t *map, const double *factor, int64_t n_new, const double *other_scale,
                   ib_weighted **out);
int ib_w_set_weights(ib_weighted *, const double *wM, const double *Mw, int conservative, int scaled);   /* NULL: keep */
int ib_w_set_dim(ib_weighted *, int which, int64_t sparse_extent, int64_t n, const int64_t *ids);        /* dense -> sparse ids */

/* info[0..7] = nrow, ncol, nnz, conservative, scaled, sparse_extent(dim0), sparse_extent(dim1), reserved */
int ib_w_info(const ib_weighted *, int64_t *info);
/* matrix in CSR (rows = output space, columns ascending inside a row) and as sorted COO (to_coo()) */
int ib_w_get_csr(const ib_weighted *, int32_t *rowptr, int32_t *col, double *val);
int ib_w_get_coo(const ib_weighted *, int32_t *row, int32_t *col, double *val);
int ib_w_get_weights(const ib_weighted *, double *wM, double *Mw);            /* get_weights(0), get_weights(1) */
int ib_w_get_dim(const ib_weighted *, int which, int64_t *ids);               /* dims[which] dense->sparse */
/* device views (valid until ib_w_free): rowptr[nrow+1], col[nnz], val[nnz], wM[nrow], Mw[ncol] */
int ib_w_dev_ptrs(const ib_weighted *, const int32_t **rowptr, const int32_t **col, const double **val,
                  const double **wM, const double **Mw);

/* Weighted_Eigen::apply(A_d, fill, force_conservation)  (ibmisc; call sites merge_topo.cpp:65, icebin22m.cpp:153,
 * explicit products IceCoupler.cpp:237-238, 445-446).  Dense indexing.  X: ncol cells, Y: nrow cells. */
int ib_w_apply(const ib_weighted *, const double *X, int64_t ldx, int32_t nvar, double *Y, int64_t ldy, int layout,
               int force_conservation);
int ib_w_apply_dev(const ib_weighted *, const double *dX, int64_t ldx, int32_t nvar, double *dY, int64_t ldy,
                   int force_conservation, void *stream);          /* cell-major only */
/* Several matrices applied to the SAME input fields in one call -- the coupler's pattern: EvI and AvI both act on
 * ice_ovalsI (IceCoupler.cpp:445-446).  The applies are independent, so they run concurrently on internal streams
 * (forked from / joined to `stream`), and the host version uploads X once.  dY[i] / Y[i]: output of W[i]. */
int ib_w_apply_many_dev(int32_t nmat, const ib_weighted *const *W, const double *dX, int64_t ldx, int32_t nvar,
                        double *const *dY, const int64_t *ldy, int force_conservation, void *stream);
int ib_w_apply_many(int32_t nmat, const ib_weighted *const *W, const double *X, int64_t ldx, int32_t nvar, double *const *Y,
                    const int64_t *ldy, int layout, int force_conservation);

/* Row-block sharding of an apply across GPUs (north star: "output row blocks of the I/E grids are split across GPUs with
 * input fields replicated" -- no collective): computes only output rows [row_begin, row_end) of dY = M dX.
 * row_begin and row_end must be multiples of 32 (row_end may also be nrow).  No conservation correction. */
int ib_w_apply_rows_dev(const ib_weighted *, const double *dX, int64_t ldx, int32_t nvar, double *dY, int64_t ldy,
                        int64_t row_begin, int64_t row_end, void *stream);
/* Python apply_M(A_s, fill, force_conservation): sparse-indexed in and out   matrix_formats.rst:150-165
 * A_s: [nvar][sparse_extent(dim1)] field-major, B_s: [nvar][sparse_extent(dim0)] */
int ib_w_apply_M(const ib_weighted *, const double *A_s, int32_t nvar, double *B_s, double fill,
                 int force_conservation);
/* apply_weight(dim, A_s): out[n] = sum_d w_dim[d] * A_s[n][to_sparse(d)]    matrix_formats.rst:168-185 */
int ib_w_apply_weight(const ib_weighted *, int dim, const double *A_s, int32_t nvar, double *out);

/* coo_matvec(yy, xx, ignore_nan, nrow, ncol, row, col, data)  pylib/icebin_cython.cpp:159-192.
 * yy entries of rows with no contribution are left untouched.  The ONE entry point computed on the host: a single-vector
 * convenience of the Python binding over host arrays, not on the hot path (batched fields go through ib_w_apply*). */
int ib_coo_matvec(double *yy, const double *xx, int ignore_nan, int64_t nrow, int64_t ncol, int64_t nnz,
                  const int32_t *row, const int32_t *col, const double *data);

/* VectorMultivec::to_dense_scale + to_dense  slib/icebin/multivec.cpp:35-81: merge n sparse contributions
 * (index[n], weights[n], vals[n][nvar]) from several sheets into dense[nvar][nE]; scaleE[nE] = 1/sum(weights).
 * Host buffers in and out, computed on the device (stable sort by cell, sums in input order: the reference's bits). */
int ib_multivec_to_dense(int64_t n, const int64_t *index, const double *weights, const double *vals, int32_t nvar,
                         int64_t nE, double fill, double *scaleE, double *dense);

/* Device-resident form of the same merge, one ice sheet per GPU (see icebin_b200/multi.py): pack the rows of the
 * cells shared with other sheets (loc[k] = local dense row of shared cell k, or -1), all-reduce `buf` across ranks
 * (NCCL, by the caller), then unpack: shared rows <- buf / wtot, other rows <- Y / wlocal. */
int ib_merge_pack_dev(const double *dY, int64_t ldy, int32_t nvar, const int32_t *d_loc, int64_t nshared, double *d_buf,
                      void *stream);
int ib_merge_unpack_dev(double *dY, int64_t ldy, int32_t nvar, int64_t nrow, const int32_t *d_slot_of_row,
                        const double *d_buf, const double *d_wtot, const double *d_wlocal, void *stream);

/* ---- The same merge across GPUs over NVLink peer memory (one rank = one process = one GPU = one ice sheet) ----------
 * Replaces the MPI gather to the GCM root + VectorMultivec merge (modele/GCMCoupler_ModelE.cpp:761-793, multivec.cpp:35-81)
 * for results that stay in HBM: a rank publishes the rows of its shared cells by storing them into the inboxes of the cells'
 * other owners (their memory is mapped with CUDA IPC; posted stores over NVLink, four inbox slots per peer) and gathers, for
 * every shared cell it holds, its own row and the rows in its inbox, added in ascending rank order (the same bits on every
 * rank), times 1 / (summed weights); rows of cells no other sheet touches are divided by their own weight.  No host round
 * trip, capturable in a CUDA graph; step counters in device memory (one per kernel: ib_xchg_publish_dev of step s + 1 may
 * run on another stream while ib_xchg_gather_dev of step s is still adding) and ready / done counters in the RECEIVER's
 * memory order the ranks: every wait polls local memory (waits give up after 4 s: ib_xchg_status).
 *
 *   host (all ranks):  gather every rank's dims[0] ids  ->  ib_xchg_plan  ->  ib_xchg_create  ->  all-gather the handles
 *                      ->  ib_xchg_open  ->  ib_xchg_set_part (per grid: E, A)  ->  ib_xchg_merge_dev every step.
 * Up to 64 ranks, up to 4 parts (grids) per context.  Ranks of ONE process use ib_xchg_open_local instead of IPC handles.
 * Kernels of different ranks wait on each other, so they must run on different GPUs; on one GPU drive the ranks one after
 * the other with ib_xchg_publish_dev (all ranks) and then ib_xchg_gather_dev (all ranks). */
typedef struct ib_xchg ib_xchg;
/* shared cells of two or more ranks, ascending: ids[r][n[r]] = rank r's dims[0]; shared == NULL returns the count */
int64_t ib_xchg_plan(int nranks, const int64_t *const *ids, const int64_t *n, int64_t n_sparse, int rank, int64_t *shared,
                     int32_t *loc, uint64_t *owners);
int ib_xchg_create(ib_xchg **out, int rank, int nranks, int nparts, const int64_t *nshared, int32_t nvar_max);
void ib_xchg_free(ib_xchg *);
int ib_xchg_handle_bytes(void);                                  /* sizeof(cudaIpcMemHandle_t) */
int ib_xchg_get_handle(const ib_xchg *, void *handle);
int ib_xchg_open(ib_xchg *, const void *handles);                /* [nranks][ib_xchg_handle_bytes()], all-gathered by the host */
int ib_xchg_open_local(ib_xchg *, ib_xchg *const *peers);        /* peers[r] = rank r's context in this process */
/* scaled = 0: dY holds unscaled sums (the coupler's E1vI_unscaled) and every row is divided by its weight in the merge;
 * scaled = 1: dY comes from the scaled matrix, only the shared cells are touched (Y * wlocal is what travels) */
int ib_xchg_set_part(ib_xchg *, int part, int64_t nrow, const int32_t *loc, const uint64_t *owners, const double *wtot,
                     const double *wlocal, int scaled);
/* dY[p]: [nrow_p][ldy_p] unscaled sums of part p in, merged and scaled values out */
int ib_xchg_merge_dev(ib_xchg *, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream);
int ib_xchg_publish_dev(ib_xchg *, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream);
int ib_xchg_gather_dev(ib_xchg *, double *const *dY, const int64_t *ldy, int32_t nvar, void *stream);
int ib_xchg_status(ib_xchg *);                                   /* IB_ERR_CUDA if a wait on a peer timed out; synchronises */
int64_t ib_xchg_step(ib_xchg *);                                 /* steps completed so far; synchronises */

/* ---- Coupling step, regridding side (callers of the path: SURVEY 8f rank 2) ---------------------------- */
/* Variable/unit transformation in front of a regrid apply:  Z = O * M + b.replicate(ncell, 1)
 *   gcm -> ice: ice_ivalsI = IvE0 * (gcm_ovalsE0 * ivg.M + ivg.b)        IceCoupler.cpp:203-252
 *   ice -> gcm: gcm_ivalsX = X1vI * (ice_ovalsI * gvi.M + gvi.b)         IceCoupler.cpp:404-446
 * dO: [ncell][ldo] cell-major device fields (nin variables), dZ: [ncell][ldz] (nout variables).  M is the transposed
 * VarTransformer matrix (apply_scalars(scalars, 'T')), column-major sparse [nin x nout] on the HOST: colptr[nout+1],
 * row[nnz], val[nnz]; b[nout].  Structural zeros are never multiplied (a NaN only reaches the outputs that name it).
 * Synchronises `stream` (the small host arrays are staged). */
int ib_vartrans_dev(const double *dO, int64_t ldo, int64_t ncell, int32_t nin, int32_t nout, const int32_t *colptr,
                    const int32_t *row, const double *val, const double *b, double *dZ, int64_t ldz, void *stream);

/* e1ve0::compute_E1vE0c(XuE1s, XuE0s, nE, areaX)   e1ve0.cpp:55-106, call site GCMCoupler.cpp:488-515.
 * XuE1[i], XuE0[i]: the sheet's "XvE" matrices (scale = 0, correctA = 1, identity X dim) of this and of the previous
 * coupling step; the sparsify() of GCMCoupler.cpp:388-449 is done inside.  Result: tuples (iE1, iE0, value) in SPARSE
 * E indices, sorted by (iE1, iE0), duplicates summed (consolidate, e1ve0.cpp:32-52).
 * Call with rows == NULL to get the entry count (the result is kept until the next call on this thread). */
int64_t ib_e1ve0c(int32_t nsheet, const ib_weighted *const *XuE1, const ib_weighted *const *XuE0, int64_t nE, int64_t *rows,
                  int64_t *cols, double *vals);

/* combine_chunks(ifnames, ofname, sgrids, scale)   modele/combine_global_ec.cpp:95-262: the reference computes global
 * matrices in chunks (ranges of GCM cells, one file each) and concatenates them; here a chunk is an UNSCALED ib_weighted
 * (one per GCM-cell domain, GPU or ice sheet).  Entries are relabelled with sparse indices and, if `scale`, multiplied by
 * 1 / (sum over all chunks of wM) of their row; wM and Mw are returned as sparse (index, value) lists.  Everything is
 * concatenated in chunk order, nothing is merged.  Buffer sizes: sum of nnz, of nrow and of ncol over the chunks (ib_w_info). */
int ib_w_combine_chunks(int32_t nchunk, const ib_weighted *const *W, int scale, int64_t *M_row, int64_t *M_col, double *M_val,
                        int64_t *wM_index, double *wM_val, int64_t *Mw_index, double *Mw_val);

/* ---- Exchange-grid generation for a rectilinear ice grid (the step in front of the path: SURVEY 8f rank 4) --- */
/* make_exchange_grid(gridA, gridI, sproj)   gridgen/GridGen_Exchange.cpp:208-316 (+ overlap_callback :175-204), for GCM
 * cells given as polygons ALREADY PROJECTED into the ice grid's plane (pptr[npoly+1] vertex ranges into pxy[nv][2];
 * the reference subdivides the sides of lon-lat cells before projecting, so 4..92 vertices, not necessarily convex) and
 * an ice grid of rectangles with edges xe[nx+1], ye[ny+1].  Every (polygon, ice cell) overlap of positive area becomes
 * an exchange cell (iA[p], iI, area): Sutherland-Hodgman clipping + shoelace area instead of CGAL + Cell::proj_area.
 * iI = ix*ny + iy if iI_major_x (PISM order, gridgen/searise_grid.cpp:104-108) else iy*nx + ix.
 * Order: polygons in input order, ice cells ascending inside one.  Call with ex_ij == NULL for the count (the result is
 * kept on this thread), then with ex_ij[2*n] (interleaved iA, iI: ExchangeGrid::indices) and ex_area[n]. */
int64_t ib_exgrid_rectilinear(int32_t npoly, const int32_t *pptr, const double *pxy, const int64_t *iA, int32_t nx,
                              const double *xe, int32_t ny, const double *ye, int iI_major_x, int32_t *ex_ij, double *ex_area);

/* ---- Hntr ------------------------------------------------------------------------------------------ */
/* Hntr(yp17, Bspec, Aspec, DATMIS); HntrSpec(im, jm, offi, dlat[minutes])   hntr.cpp:63-82, GridSpec.hpp:143-170 */
int ib_hntr_create(ib_hntr **out, int imB, int jmB, double offiB, double dlatB, int imA, int jmA, double offiA,
                   double dlatA, double datmis);
void ib_hntr_free(ib_hntr *);
/* which: 0 SINA[jmA+1] 1 SINB[jmB+1] 2 FMIN[imB] 3 FMAX[imB] 4 GMIN[jmB] 5 GMAX[jmB] (double)
 *        6 IMIN[imB] 7 IMAX[imB] 8 JMIN[jmB] 9 JMAX[jmB] (int, 1-based values as in the reference)
 *        10 dxyp(A)[jmA] 11 dxyp(B)[jmB]  (HntrGrid::dxyp, hntr.cpp:34-52) */
int ib_hntr_get(const ib_hntr *, int which, double *dout, int32_t *iout);
/* Hntr::regrid(WTA, A, B, mean_polar, wtm, wtb)  hntr.hpp:385-435.  A: [jmA*imA] (i fastest) of src_dtype;
 * wta NULL = constant 1 (ibmisc::const_array, etopo1_ice.cpp:24-26); B: [jmB*imB] fp64. */
int ib_hntr_regrid(const ib_hntr *, const double *wta, double wtm, double wtb, const void *src, int src_dtype,
                   double *dst, int mean_polar);
int ib_hntr_regrid_dev(const ib_hntr *, const double *d_wta, double wtm, double wtb, const void *d_src,
                       int src_dtype, double *d_dst, int mean_polar, void *stream);
/* Hntr::overlap(accum, R) (R >= 0) / scaled_regrid_matrix(accum) (R < 0)  hntr.hpp:249-339.
 * Call with rows == NULL to get the entry count, then again with buffers.  Entries in the reference's emission
 * order; indices 0-based (IJB-1, IJA-1). */
int64_t ib_hntr_matrix(const ib_hntr *, double R, int64_t *rows, int64_t *cols, double *vals);

#ifdef __cplusplus
}
#endif
#endif

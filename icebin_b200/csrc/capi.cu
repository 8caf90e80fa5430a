// capi.cu -- the extern "C" boundary declared in include/icebin_b200.h.  Thin: argument checks, host<->device
// staging for the host-pointer entry points, exception -> status code translation.
#include <cmath>
#include <cstring>
#include <mutex>

#include "types.cuh"

namespace ib {

std::atomic<int64_t> g_launches{0};
Tune g_tune;
static thread_local std::string g_err;
void set_last_error(std::string const &msg) { g_err = msg; }

void ensure_device()
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        fail(IB_ERR_CUDA, "icebin_b200: no CUDA device available (%s); there is no CPU fallback",
             e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
}

int device_index()
{
    int dev = 0;
    IB_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= IB_MAX_DEVICES) fail(IB_ERR_CUDA, "icebin_b200: device index %d out of range", dev);
    return dev;
}

// Freed blocks stay in the device's default memory pool up to 1/4 of the device's memory (steady-state matrix builds --
// ~10 GB of temporaries for a 36 M-cell sheet -- then never reach the driver allocator); what exceeds it goes back to the
// driver at the next synchronisation, so a co-resident allocator (the host application's) is not starved.  ib_trim()
// releases everything.
static uint64_t pool_keep_bytes()
{
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); return 4ull << 30; }
    return std::max<uint64_t>(4ull << 30, tot / 4);
}

cudaStream_t default_stream()
{
    static std::mutex mu;
    static cudaStream_t streams[IB_MAX_DEVICES] = {};
    const int dev = device_index();
    std::lock_guard<std::mutex> lk(mu);
    if (!streams[dev]) {
        IB_CUDA(cudaStreamCreateWithFlags(&streams[dev], cudaStreamNonBlocking));
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            uint64_t thr = 0;
            cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            const uint64_t keep = pool_keep_bytes();
            if (thr < keep) {                     // never lower a threshold the host application has raised
                thr = keep;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            }
        }
    }
    return streams[dev];
}

int sm_count()
{
    static int n[IB_MAX_DEVICES] = {};
    const int dev = device_index();
    if (!n[dev]) cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev);
    return n[dev] > 0 ? n[dev] : 148;
}

// hntr.cu
struct Hntr;
Hntr *hntr_create(int, int, double, double, int, int, double, double, double);
void hntr_free(Hntr *);
void hntr_get(Hntr const &, int, double *, int32_t *);
void hntr_dims(Hntr const &, int *);
void hntr_regrid_dev(Hntr &, const double *, double, double, const void *, int, double *, bool, cudaStream_t);
int64_t hntr_matrix(Hntr &, double, int64_t *, int64_t *, double *, cudaStream_t);

// couple.cu
void vartrans_dev(const double *, int64_t, int64_t, int, int, const int32_t *, const int32_t *, const double *, const double *, double *,
                  int64_t, cudaStream_t);
void e1ve0c(int, Weighted *const *, Weighted *const *, int64_t, std::vector<int64_t> &, std::vector<int64_t> &, std::vector<double> &,
            cudaStream_t);

void combine_chunks(int, Weighted *const *, bool, int64_t *, int64_t *, double *, int64_t *, double *, int64_t *, double *, cudaStream_t);

// exgrid.cu
int64_t exgrid_rectilinear(int, const int32_t *, const double *, const int64_t *, int, const double *, int, const double *, int, int32_t *,
                           double *, cudaStream_t);

// modele.cu
std::unique_ptr<Weighted> aggregate(Weighted &, int, const int32_t *, const double *, int64_t, const double *, cudaStream_t);

// merge.cu
void multivec_to_dense(int64_t, const int64_t *, const double *, const double *, int, int64_t, double, double *, double *, cudaStream_t);
void merge_pack(const double *, int64_t, int, const int32_t *, int64_t, double *, cudaStream_t);
void merge_unpack(double *, int64_t, int, int64_t, const int32_t *, const double *, const double *, const double *, cudaStream_t);

// cell-major <-> field-major transposition: out[c*ldo + r] = in[r*ldi + c]
__global__ void transpose_kernel(const double *__restrict__ in, int64_t ldi, int64_t nr, int64_t nc, double *__restrict__ out, int64_t ldo)
{
    __shared__ double tile[32][33];
    int64_t c0 = (int64_t)blockIdx.x * 32, r0 = (int64_t)blockIdx.y * 32;
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t r = r0 + k, c = c0 + threadIdx.x;
        if (r < nr && c < nc) tile[k][threadIdx.x] = in[r * ldi + c];
    }
    __syncthreads();
    for (int k = threadIdx.y; k < 32; k += blockDim.y) {
        int64_t c = c0 + k, r = r0 + threadIdx.x;
        if (r < nr && c < nc) out[c * ldo + r] = tile[threadIdx.x][k];
    }
}
static void transpose(const double *in, int64_t ldi, int64_t nr, int64_t nc, double *out, int64_t ldo, cudaStream_t s)
{
    if (nr == 0 || nc == 0) return;
    dim3 grid((unsigned)cdiv(nc, 32), (unsigned)cdiv(nr, 32)), block(32, 8);
    transpose_kernel<<<grid, block, 0, s>>>(in, ldi, nr, nc, out, ldo);
    IB_LAUNCHED();
}

// gather/scatter between sparse-indexed field-major host layout and dense cell-major device layout (apply_M)
__global__ void densify_kernel(const double *__restrict__ As, int64_t nsparse, const int64_t *__restrict__ sparse_of, int identity,
                               int64_t ndense, int nvar, double *__restrict__ Xd)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t d = t % ndense;
    int64_t f = t / ndense;
    if (f >= nvar) return;
    int64_t sidx = identity ? d : sparse_of[d];
    Xd[d * nvar + f] = As[f * nsparse + sidx];
}
__global__ void sparsify_kernel(const double *__restrict__ Yd, const int64_t *__restrict__ sparse_of, int identity, int64_t ndense,
                                int nvar, int64_t nsparse, double *__restrict__ Bs)
{
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t d = t % ndense;
    int64_t f = t / ndense;
    if (f >= nvar) return;
    int64_t sidx = identity ? d : sparse_of[d];
    Bs[f * nsparse + sidx] = Yd[d * nvar + f];
}
__global__ void fill_f64_kernel(double *p, int64_t n, double v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}   // namespace ib

using namespace ib;

#define IB_TRY try {
#define IB_CATCH                                                           \
    }                                                                      \
    catch (ib::Error const &e) { g_err = e.what(); return e.code; }        \
    catch (std::exception const &e) { g_err = e.what(); return IB_ERR_ARG; }

struct ib_gcm { Gcm g; };
struct ib_rm { RegridMatrices r; };
struct ib_dim { Dim d; };
struct ib_weighted { std::unique_ptr<Weighted> w; };
struct ib_hntr { ib::Hntr *h; };

extern "C" {

const char *ib_last_error(void) { return g_err.c_str(); }

int ib_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}
int ib_set_device(int device)
{
    IB_TRY
    IB_CUDA(cudaSetDevice(device));
    return IB_OK;
    IB_CATCH
}
int64_t ib_launch_count(void) { return g_launches.load(); }

/* performance knobs (results never depend on them) */
int ib_tune(const char *key, int value)
{
    IB_TRY
    if (!key) fail(IB_ERR_ARG, "ib_tune: null key");
    const std::string k(key);
    if (k == "stream_bps") { if (value < 1 || value > 8) fail(IB_ERR_ARG, "ib_tune: stream_bps in 1..8"); g_tune.stream_bps = value; }
    else if (k == "stream_spare") { if (value < 0 || value > 64) fail(IB_ERR_ARG, "ib_tune: stream_spare in 0..64"); g_tune.stream_spare = value; }
    else if (k == "tma_bps") { if (value < 1 || value > 8) fail(IB_ERR_ARG, "ib_tune: tma_bps in 1..8"); g_tune.tma_bps = value; }
    else if (k == "gc_nq") { if (value != 0 && value != 1 && value != 2 && value != 4) fail(IB_ERR_ARG, "ib_tune: gc_nq in {0,1,2,4}"); g_tune.gc_nq = value; }
    else if (k == "use_stream") g_tune.use_stream = value != 0;
    else if (k == "use_gc") g_tune.use_gc = value != 0;
    else if (k == "use_group") g_tune.use_group = value != 0;
    else if (k == "xchg_blocks") g_tune.xchg_blocks = std::max(1, std::min(value, 1024));
    else if (k == "hntr_generic") g_tune.hntr_generic = value != 0;
    else fail(IB_ERR_ARG, "ib_tune: unknown key '%s'", key);
    return IB_OK;
    IB_CATCH
}
int ib_trim(void)
{
    IB_TRY
    ensure_device();
    cudaMemPool_t pool;
    IB_CUDA(cudaDeviceGetDefaultMemPool(&pool, device_index()));
    IB_CUDA(cudaDeviceSynchronize());
    IB_CUDA(cudaMemPoolTrimTo(pool, 0));
    return IB_OK;
    IB_CATCH
}

// ---------------------------------------------------------------------------------------- device buffers
int ib_dev_alloc(void **out, int64_t bytes)
{
    IB_TRY
    if (!out || bytes < 0) fail(IB_ERR_ARG, "ib_dev_alloc: bad argument");
    ensure_device();
    *out = nullptr;
    cudaStream_t s = default_stream();
    if (bytes > 0) IB_CUDA(cudaMallocAsync(out, (size_t)bytes, s));
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}
void ib_dev_free(void *p)
{
    if (p) cudaFreeAsync(p, default_stream());
}
int ib_dev_upload(void *dst_dev, const void *src_host, int64_t bytes)
{
    IB_TRY
    if (bytes < 0 || (bytes > 0 && (!dst_dev || !src_host))) fail(IB_ERR_ARG, "ib_dev_upload: bad argument");
    ensure_device();
    cudaStream_t s = default_stream();
    if (bytes > 0) IB_CUDA(cudaMemcpyAsync(dst_dev, src_host, (size_t)bytes, cudaMemcpyHostToDevice, s));
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}
int ib_dev_download(void *dst_host, const void *src_dev, int64_t bytes)
{
    IB_TRY
    if (bytes < 0 || (bytes > 0 && (!dst_host || !src_dev))) fail(IB_ERR_ARG, "ib_dev_download: bad argument");
    ensure_device();
    cudaStream_t s = default_stream();
    if (bytes > 0) IB_CUDA(cudaMemcpyAsync(dst_host, src_dev, (size_t)bytes, cudaMemcpyDeviceToHost, s));
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}

// ---------------------------------------------------------------------------------------- GCMRegridder
int ib_gcm_create(ib_gcm **out, int64_t nA, int32_t n_real, const int64_t *dimA, const double *native_area, int32_t nhc,
                  const double *hcdefs, int correctA)
{
    IB_TRY
    if (!out || nA <= 0 || n_real < 0 || nhc < 0 || (n_real && (!dimA || !native_area)) || (nhc && !hcdefs))
        fail(IB_ERR_ARG, "ib_gcm_create: bad argument");
    ensure_device();
    std::unique_ptr<ib_gcm> G(new ib_gcm);
    Gcm &g = G->g;
    g.nA = nA; g.correctA = correctA;
    g.dimA_h.assign(dimA, dimA + n_real);
    for (int64_t v : g.dimA_h) if (v < 0 || v >= nA) fail(IB_ERR_ARG, "ib_gcm_create: A cell index %ld outside [0,%ld)", (long)v, (long)nA);
    g.native_area_h.assign(native_area, native_area + n_real);
    g.hcdefs_h.assign(hcdefs, hcdefs + nhc);
    for (int k = 1; k < nhc; ++k) if (!(hcdefs[k] > hcdefs[k - 1])) fail(IB_ERR_ARG, "ib_gcm_create: hcdefs must be strictly increasing");
    cudaStream_t s = default_stream();
    g.hcdefs.alloc((size_t)nhc, s);
    g.hcdefs.upload(g.hcdefs_h.data(), (size_t)nhc);
    IB_CUDA(cudaStreamSynchronize(s));
    *out = G.release();
    return IB_OK;
    IB_CATCH
}
void ib_gcm_free(ib_gcm *g) { delete g; }

int ib_gcm_filter_cellsA(ib_gcm *G, const uint8_t *keepA)
{
    IB_TRY
    if (!G || !keepA) fail(IB_ERR_ARG, "ib_gcm_filter_cellsA: bad argument");
    filter_cellsA(G->g, keepA, default_stream());
    return IB_OK;
    IB_CATCH
}

int ib_gcm_add_sheet(ib_gcm *G, const char *name, int64_t nI, int32_t n_real_I, const int64_t *dimI, const double *native_areaI,
                     const double *centroid_xy, int64_t nX, const int32_t *ex_ij, const double *ex_area,
                     const double *gridA_proj_area, int interp_style)
{
    IB_TRY
    if (!G || nI <= 0 || n_real_I < 0 || nX < 0 || (nX && (!ex_ij || !ex_area)) || !gridA_proj_area)
        fail(IB_ERR_ARG, "ib_gcm_add_sheet: bad argument");
    if (interp_style != IB_Z_INTERP && interp_style != IB_ELEV_CLASS_INTERP) fail(IB_ERR_ARG, "ib_gcm_add_sheet: unknown interp_style");
    Gcm &g = G->g;
    for (int64_t x = 0; x < nX; ++x) {
        if (ex_ij[2 * x] < 0 || ex_ij[2 * x] >= g.nA || ex_ij[2 * x + 1] < 0 || ex_ij[2 * x + 1] >= nI)
            fail(IB_ERR_ARG, "ib_gcm_add_sheet: exchange cell %ld references (iA=%d, iI=%d) outside the grids", (long)x, ex_ij[2 * x], ex_ij[2 * x + 1]);
    }
    cudaStream_t s = default_stream();
    std::unique_ptr<Sheet> sh(new Sheet);
    sh->name = name ? name : "icesheet";
    sh->nI = nI; sh->nX = nX; sh->interp_style = interp_style;
    sh->dimI_h.assign(dimI, dimI + n_real_I);
    sh->dimI_identity = (n_real_I == nI);
    for (int64_t i = 0; i < n_real_I && sh->dimI_identity; ++i) sh->dimI_identity = (dimI[i] == i);
    sh->dimI.alloc((size_t)n_real_I, s); sh->dimI.upload(dimI, (size_t)n_real_I);
    sh->native_areaI.alloc((size_t)n_real_I, s); sh->native_areaI.upload(native_areaI, (size_t)n_real_I);
    if (centroid_xy && n_real_I > 0) {
        sh->centroid_xy.alloc(2 * (size_t)n_real_I, s); sh->centroid_xy.upload(centroid_xy, 2 * (size_t)n_real_I);
        sh->bbox[0] = sh->bbox[1] = centroid_xy[0]; sh->bbox[2] = sh->bbox[3] = centroid_xy[1];
        for (int64_t i = 0; i < n_real_I; ++i) {
            sh->bbox[0] = std::min(sh->bbox[0], centroid_xy[2 * i]); sh->bbox[1] = std::max(sh->bbox[1], centroid_xy[2 * i]);
            sh->bbox[2] = std::min(sh->bbox[2], centroid_xy[2 * i + 1]); sh->bbox[3] = std::max(sh->bbox[3], centroid_xy[2 * i + 1]);
        }
    }
    sh->ex_ij.alloc((size_t)nX, s); sh->ex_ij.upload(reinterpret_cast<const int2 *>(ex_ij), (size_t)nX);
    sh->ex_area.alloc((size_t)nX, s); sh->ex_area.upload(ex_area, (size_t)nX);
    // sApvA / sEpvE diagonal (IceRegridder.cpp:41-65): native/proj per realised A cell, 0 elsewhere ("missing")
    std::vector<double> ratio((size_t)g.nA, 0.0);
    sh->gridA_proj_area_h.assign(gridA_proj_area, gridA_proj_area + g.dimA_h.size());
    for (size_t d = 0; d < g.dimA_h.size(); ++d) ratio[(size_t)g.dimA_h[d]] = g.native_area_h[d] / gridA_proj_area[d];
    sh->ratioA.alloc((size_t)g.nA, s); sh->ratioA.upload(ratio.data(), ratio.size());
    IB_CUDA(cudaStreamSynchronize(s));
    g.sheets.push_back(std::move(sh));
    return (int)g.sheets.size() - 1;
    IB_CATCH
}

int64_t ib_gcm_nA_realized(const ib_gcm *G) { return G ? (int64_t)G->g.dimA_h.size() : -1; }
int64_t ib_gcm_nA(const ib_gcm *G) { return G->g.nA; }
int64_t ib_gcm_nE(const ib_gcm *G) { return G->g.nE(); }
int32_t ib_gcm_nhc(const ib_gcm *G) { return (int32_t)G->g.hcdefs_h.size(); }
int ib_gcm_nsheets(const ib_gcm *G) { return (int)G->g.sheets.size(); }
int ib_gcm_correctA(const ib_gcm *G) { return G->g.correctA; }
int64_t ib_gcm_nI(const ib_gcm *G, int sheet) { return (sheet < 0 || sheet >= (int)G->g.sheets.size()) ? -1 : G->g.sheets[sheet]->nI; }
int64_t ib_gcm_nX(const ib_gcm *G, int sheet) { return (sheet < 0 || sheet >= (int)G->g.sheets.size()) ? -1 : G->g.sheets[sheet]->nX; }
int ib_gcm_sheet_index(const ib_gcm *G, const char *name)
{
    for (size_t i = 0; i < G->g.sheets.size(); ++i) if (G->g.sheets[i]->name == name) return (int)i;
    g_err = std::string("no ice sheet named '") + name + "'";
    return IB_ERR_ARG;
}
int ib_gcm_wA(const ib_gcm *G, int sheet, int native, int64_t *index, double *area)
{
    IB_TRY
    Gcm const &g = G->g;
    if (sheet < 0 || sheet >= (int)g.sheets.size()) fail(IB_ERR_ARG, "ib_gcm_wA: bad sheet index %d", sheet);
    auto const &a = native ? g.native_area_h : g.sheets[sheet]->gridA_proj_area_h;
    for (size_t d = 0; d < g.dimA_h.size(); ++d) { index[d] = g.dimA_h[d]; area[d] = a[d]; }
    return IB_OK;
    IB_CATCH
}

// ---------------------------------------------------------------------------------------- RegridMatrices
static int regrid_matrices_impl(ib_gcm *G, int sheet, const double *em, bool on_device, int scale, int correctA, const double *sigma,
                                cudaStream_t s, ib_rm **out)
{
    IB_TRY
    if (!G || !em || !out) fail(IB_ERR_ARG, "ib_regrid_matrices: bad argument");
    Gcm &g = G->g;
    if (sheet < 0 || sheet >= (int)g.sheets.size()) fail(IB_ERR_ARG, "ib_regrid_matrices: bad sheet index %d", sheet);
    ensure_device();
    std::unique_ptr<ib_rm> R(new ib_rm);
    R->r.gcm = &g; R->r.sheet = sheet;
    R->r.params.scale = scale != 0; R->r.params.correctA = correctA != 0;
    if (sigma) R->r.params.sigma = {sigma[0], sigma[1], sigma[2]};
    R->r.stream = s;
    size_t n = (size_t)g.sheets[sheet]->nI;
    R->r.elevmaskI.alloc(n, s);
    IB_CUDA(cudaMemcpyAsync(R->r.elevmaskI.p, em, n * sizeof(double), on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s));
    if (!on_device) IB_CUDA(cudaStreamSynchronize(s));
    *out = R.release();
    return IB_OK;
    IB_CATCH
}
int ib_regrid_matrices(ib_gcm *G, int sheet, const double *elevmaskI, int scale, int correctA, const double *sigma, ib_rm **out)
{
    cudaStream_t s;
    IB_TRY s = default_stream(); IB_CATCH
    return regrid_matrices_impl(G, sheet, elevmaskI, false, scale, correctA, sigma, s, out);
}
int ib_regrid_matrices_dev(ib_gcm *G, int sheet, const double *d_elevmaskI, int scale, int correctA, const double *sigma, void *stream,
                           ib_rm **out)
{
    cudaStream_t s;
    IB_TRY s = stream ? (cudaStream_t)stream : default_stream(); IB_CATCH
    return regrid_matrices_impl(G, sheet, d_elevmaskI, true, scale, correctA, sigma, s, out);
}
void ib_rm_free(ib_rm *r) { delete r; }

int ib_dim_create(ib_dim **out)
{
    IB_TRY
    *out = new ib_dim;
    return IB_OK;
    IB_CATCH
}
int ib_dim_create_ids(ib_dim **out, int64_t sparse_extent, int64_t n, const int64_t *ids)
{
    IB_TRY
    if (n < 0 || (n && !ids)) fail(IB_ERR_ARG, "ib_dim_create_ids: bad argument");
    std::unique_ptr<ib_dim> D(new ib_dim);
    D->d.pending.assign(ids, ids + n);
    {   // ids must be distinct (a SparseSet is a bijection)
        std::vector<int64_t> t(D->d.pending);
        std::sort(t.begin(), t.end());
        if (std::adjacent_find(t.begin(), t.end()) != t.end()) fail(IB_ERR_ARG, "ib_dim_create_ids: duplicate sparse id");
    }
    D->d.dense_extent = n;
    D->d.sparse_extent = sparse_extent;
    if (n == 0) D->d.pending.clear();
    *out = D.release();
    return IB_OK;
    IB_CATCH
}
int ib_dim_create_identity(ib_dim **out, int64_t n)
{
    IB_TRY
    if (n < 0) fail(IB_ERR_ARG, "ib_dim_create_identity: bad argument");
    std::unique_ptr<ib_dim> D(new ib_dim);
    D->d.identity = true; D->d.sparse_extent = n; D->d.dense_extent = n;
    *out = D.release();
    return IB_OK;
    IB_CATCH
}
int64_t ib_dim_dense_extent(const ib_dim *d) { return d->d.dense_extent; }
int64_t ib_dim_sparse_extent(const ib_dim *d) { return d->d.sparse_extent; }
int ib_dim_get_ids(const ib_dim *d, int64_t *ids)
{
    IB_TRY
    auto v = d->d.ids(d->d.sparse_of.st ? d->d.sparse_of.st : default_stream());
    std::copy(v.begin(), v.end(), ids);
    return IB_OK;
    IB_CATCH
}
void ib_dim_free(ib_dim *d) { delete d; }

int ib_rm_matrix(ib_rm *R, const char *spec, ib_weighted **out)
{
    IB_TRY
    if (!R || !spec || !out) fail(IB_ERR_ARG, "ib_rm_matrix: bad argument");
    std::unique_ptr<ib_weighted> W(new ib_weighted);
    W->w = build_matrix(R->r, spec, new Dim, new Dim, R->r.params, true);
    *out = W.release();
    return IB_OK;
    IB_CATCH
}
int ib_rm_matrix_d(ib_rm *R, const char *spec, ib_dim *dim0, ib_dim *dim1, int scale, int correctA, const double *sigma,
                   ib_weighted **out)
{
    IB_TRY
    if (!R || !spec || !out || !dim0 || !dim1) fail(IB_ERR_ARG, "ib_rm_matrix_d: bad argument");
    Params p;
    p.scale = scale != 0; p.correctA = correctA != 0;
    if (sigma) p.sigma = {sigma[0], sigma[1], sigma[2]};
    std::unique_ptr<ib_weighted> W(new ib_weighted);
    W->w = build_matrix(R->r, spec, &dim0->d, &dim1->d, p, false);
    *out = W.release();
    return IB_OK;
    IB_CATCH
}
void ib_w_free(ib_weighted *w) { delete w; }

int ib_w_create_csr(ib_weighted **out, int64_t nrow, int64_t ncol, const int32_t *rowptr, const int32_t *col, const double *val,
                    const double *wM, const double *Mw, int conservative)
{
    IB_TRY
    if (!out || nrow < 0 || ncol < 0 || !rowptr) fail(IB_ERR_ARG, "ib_w_create_csr: bad argument");
    ensure_device();
    const int64_t nnz = rowptr[nrow];
    if (rowptr[0] != 0 || nnz < 0 || (nnz && (!col || !val))) fail(IB_ERR_ARG, "ib_w_create_csr: bad rowptr");
    for (int64_t r = 0; r < nrow; ++r) if (rowptr[r + 1] < rowptr[r]) fail(IB_ERR_ARG, "ib_w_create_csr: rowptr not monotone");
    for (int64_t j = 0; j < nnz; ++j) if (col[j] < 0 || col[j] >= ncol) fail(IB_ERR_ARG, "ib_w_create_csr: column index out of range");
    cudaStream_t s = default_stream();
    std::unique_ptr<ib_weighted> W(new ib_weighted);
    W->w.reset(new Weighted);
    Weighted &w = *W->w;
    w.stream = s; w.nrow = nrow; w.ncol = ncol; w.nnz = nnz; w.conservative = conservative != 0;
    for (int k = 0; k < 2; ++k) {
        w.owned[k].reset(new Dim);
        w.owned[k]->identity = true;
        w.owned[k]->sparse_extent = w.owned[k]->dense_extent = k ? ncol : nrow;
        w.dims[k] = w.owned[k].get();
    }
    w.rowptr.alloc((size_t)nrow + 1, s); w.rowptr.upload(rowptr, (size_t)nrow + 1);
    w.col.alloc((size_t)nnz, s); w.col.upload(col, (size_t)nnz);
    w.val.alloc((size_t)nnz, s); w.val.upload(val, (size_t)nnz);
    w.wM.alloc((size_t)nrow, s); w.Mw.alloc((size_t)ncol, s);
    if (wM) w.wM.upload(wM, (size_t)nrow); else w.wM.zero();
    if (Mw) w.Mw.upload(Mw, (size_t)ncol); else w.Mw.zero();
    IB_CUDA(cudaStreamSynchronize(s));
    *out = W.release();
    return IB_OK;
    IB_CATCH
}

int ib_w_aggregate(const ib_weighted *W, int side, const int32_t *map, const double *factor, int64_t n_new, const double *other_scale,
                   ib_weighted **out)
{
    IB_TRY
    if (!W || !out || (side != 0 && side != 1) || n_new < 0 || (!map && (side == 0 ? W->w->nrow : W->w->ncol) > 0))
        fail(IB_ERR_ARG, "ib_w_aggregate: bad argument");
    Weighted &w = *W->w;
    const int64_t n_old = side == 0 ? w.nrow : w.ncol;
    for (int64_t i = 0; i < n_old; ++i)
        if (map[i] >= n_new) fail(IB_ERR_ARG, "ib_w_aggregate: map[%ld] = %d outside the new extent %ld", (long)i, map[i], (long)n_new);
    std::unique_ptr<ib_weighted> N(new ib_weighted);
    N->w = aggregate(w, side, map, factor, n_new, other_scale, w.stream);
    *out = N.release();
    return IB_OK;
    IB_CATCH
}

int ib_w_set_weights(ib_weighted *W, const double *wM, const double *Mw, int conservative, int scaled)
{
    IB_TRY
    if (!W) fail(IB_ERR_ARG, "ib_w_set_weights: bad argument");
    Weighted &w = *W->w;
    if (wM) w.wM.upload(wM, (size_t)w.nrow);
    if (Mw) w.Mw.upload(Mw, (size_t)w.ncol);
    w.conservative = conservative != 0;
    w.scaled = scaled != 0;
    w.cons_v.release();
    IB_CUDA(cudaStreamSynchronize(w.stream));
    return IB_OK;
    IB_CATCH
}

int ib_w_set_dim(ib_weighted *W, int which, int64_t sparse_extent, int64_t n, const int64_t *ids)
{
    IB_TRY
    if (!W || which < 0 || which > 1 || n < 0 || (n && !ids)) fail(IB_ERR_ARG, "ib_w_set_dim: bad argument");
    Weighted &w = *W->w;
    if (n != (which ? w.ncol : w.nrow)) fail(IB_ERR_DIM, "ib_w_set_dim: %ld ids for a dimension of dense extent %ld", (long)n, (long)(which ? w.ncol : w.nrow));
    for (int64_t i = 0; i < n; ++i)
        if (ids[i] < 0 || ids[i] >= sparse_extent) fail(IB_ERR_ARG, "ib_w_set_dim: id %ld outside the sparse extent %ld", (long)ids[i], (long)sparse_extent);
    std::unique_ptr<Dim> d(new Dim);
    d->pending.assign(ids, ids + n);
    d->dense_extent = n;
    d->set_sparse_extent(sparse_extent, w.stream);
    if (n == 0) d->sparse_extent = sparse_extent;
    IB_CUDA(cudaStreamSynchronize(w.stream));
    w.owned[which] = std::move(d);
    w.dims[which] = w.owned[which].get();
    return IB_OK;
    IB_CATCH
}

int ib_w_info(const ib_weighted *W, int64_t *info)
{
    Weighted const &w = *W->w;
    info[0] = w.nrow; info[1] = w.ncol; info[2] = w.nnz; info[3] = w.conservative; info[4] = w.scaled;
    info[5] = w.dims[0]->sparse_extent; info[6] = w.dims[1]->sparse_extent; info[7] = 0;
    return IB_OK;
}
int ib_w_get_csr(const ib_weighted *W, int32_t *rowptr, int32_t *col, double *val)
{
    IB_TRY
    Weighted const &w = *W->w;
    if (rowptr) w.rowptr.download(rowptr, (size_t)w.nrow + 1);
    if (col) w.col.download(col, (size_t)w.nnz);
    if (val) w.val.download(val, (size_t)w.nnz);
    IB_CUDA(cudaStreamSynchronize(w.stream));
    return IB_OK;
    IB_CATCH
}
int ib_w_get_coo(const ib_weighted *W, int32_t *row, int32_t *col, double *val)
{
    IB_TRY
    Weighted const &w = *W->w;
    std::vector<int32_t> rp((size_t)w.nrow + 1);
    w.rowptr.download(rp.data(), rp.size());
    if (col) w.col.download(col, (size_t)w.nnz);
    if (val) w.val.download(val, (size_t)w.nnz);
    IB_CUDA(cudaStreamSynchronize(w.stream));
    if (row)
        for (int64_t r = 0; r < w.nrow; ++r)
            for (int32_t j = rp[r]; j < rp[r + 1]; ++j) row[j] = (int32_t)r;
    return IB_OK;
    IB_CATCH
}
int ib_w_get_weights(const ib_weighted *W, double *wM, double *Mw)
{
    IB_TRY
    Weighted const &w = *W->w;
    if (wM) w.wM.download(wM, (size_t)w.nrow);
    if (Mw) w.Mw.download(Mw, (size_t)w.ncol);
    IB_CUDA(cudaStreamSynchronize(w.stream));
    return IB_OK;
    IB_CATCH
}
int ib_w_get_dim(const ib_weighted *W, int which, int64_t *ids)
{
    IB_TRY
    if (which < 0 || which > 1) fail(IB_ERR_ARG, "ib_w_get_dim: which must be 0 or 1");
    auto v = W->w->dims[which]->ids(W->w->stream);
    std::copy(v.begin(), v.end(), ids);
    return IB_OK;
    IB_CATCH
}
int ib_w_dev_ptrs(const ib_weighted *W, const int32_t **rowptr, const int32_t **col, const double **val, const double **wM,
                  const double **Mw)
{
    Weighted const &w = *W->w;
    if (rowptr) *rowptr = w.rowptr.p;
    if (col) *col = w.col.p;
    if (val) *val = w.val.p;
    if (wM) *wM = w.wM.p;
    if (Mw) *Mw = w.Mw.p;
    return IB_OK;
}

int ib_w_apply_dev(const ib_weighted *W, const double *dX, int64_t ldx, int32_t nvar, double *dY, int64_t ldy, int force_conservation,
                   void *stream)
{
    IB_TRY
    if (!W || !dX || !dY) fail(IB_ERR_ARG, "ib_w_apply_dev: bad argument");
    apply_dev(*W->w, dX, ldx, nvar, dY, ldy, force_conservation != 0, stream ? (cudaStream_t)stream : W->w->stream);
    return IB_OK;
    IB_CATCH
}

// Several matrices applied to the same fields.  Two wide matrices (the coupler's EvI + AvI) are stacked and applied in
// ONE pass over X (apply_stream.cu); anything else forks onto per-call auxiliary streams and joins back on `s`.
struct ForkJoin {            // per device; used by one host thread at a time (mutex held for the whole fork/join)
    cudaStream_t aux[7] = {};
    cudaEvent_t fork_ev = nullptr, join_ev[7] = {};
};
static void apply_many(int nmat, const ib_weighted *const *W, const double *dX, int64_t ldx, int nvar, double *const *dY,
                       const int64_t *ldy, bool force, cudaStream_t s)
{
    if (nmat > 8) fail(IB_ERR_ARG, "ib_w_apply_many: at most 8 matrices per call");
    if (nmat == 2 && g_tune.use_stream) {
        Weighted &a = *W[0]->w, &b = *W[1]->w;
        const bool plain = !force || (a.conservative && b.conservative);     // no conservation correction involved
        if (plain && a.nrow > 0 && b.nrow > 0) {
            const bool a_first = a.nnz >= b.nnz;       // M1 holds the pairs (EvI), M2 the single entries (AvI)
            Weighted &m1 = a_first ? a : b, &m2 = a_first ? b : a;
            const int i1 = a_first ? 0 : 1, i2 = 1 - i1;
            if (apply_stream(m1, &m2, dX, ldx, nvar, dY[i1], ldy[i1], dY[i2], ldy[i2], s)) return;
        }
    }
    if (nmat == 1) { apply_dev(*W[0]->w, dX, ldx, nvar, dY[0], ldy[0], force, s); return; }
    static std::mutex mu;
    static ForkJoin fjs[IB_MAX_DEVICES];
    const int dev = device_index();
    std::lock_guard<std::mutex> lk(mu);
    ForkJoin &fj = fjs[dev];
    if (!fj.fork_ev) {
        IB_CUDA(cudaEventCreateWithFlags(&fj.fork_ev, cudaEventDisableTiming));
        for (int i = 0; i < 7; ++i) {
            IB_CUDA(cudaStreamCreateWithFlags(&fj.aux[i], cudaStreamNonBlocking));
            IB_CUDA(cudaEventCreateWithFlags(&fj.join_ev[i], cudaEventDisableTiming));
        }
    }
    IB_CUDA(cudaEventRecord(fj.fork_ev, s));
    int forked = 0;
    // whatever happens, `s` waits for every auxiliary stream that was given work before this function returns
    auto join = [&]() {
        for (int i = 0; i < forked; ++i) {
            cudaEventRecord(fj.join_ev[i], fj.aux[i]);
            cudaStreamWaitEvent(s, fj.join_ev[i], 0);
        }
    };
    try {
        for (int i = 1; i < nmat; ++i) {
            IB_CUDA(cudaStreamWaitEvent(fj.aux[i - 1], fj.fork_ev, 0));
            forked = i;
            apply_dev(*W[i]->w, dX, ldx, nvar, dY[i], ldy[i], force, fj.aux[i - 1]);
        }
        apply_dev(*W[0]->w, dX, ldx, nvar, dY[0], ldy[0], force, s);
    } catch (...) {
        join();
        throw;
    }
    join();
}

int ib_w_apply_many_dev(int32_t nmat, const ib_weighted *const *W, const double *dX, int64_t ldx, int32_t nvar, double *const *dY,
                        const int64_t *ldy, int force_conservation, void *stream)
{
    IB_TRY
    if (nmat <= 0 || !W || !dX || !dY || !ldy) fail(IB_ERR_ARG, "ib_w_apply_many_dev: bad argument");
    for (int i = 0; i < nmat; ++i) {
        if (!W[i] || !dY[i]) fail(IB_ERR_ARG, "ib_w_apply_many_dev: null matrix or output");
        if (W[i]->w->ncol != W[0]->w->ncol) fail(IB_ERR_DIM, "ib_w_apply_many_dev: the matrices must share their input space");
        for (int k = 0; k < i; ++k) if (W[k] == W[i]) fail(IB_ERR_ARG, "ib_w_apply_many_dev: the same matrix appears twice");
    }
    apply_many(nmat, W, dX, ldx, nvar, dY, ldy, force_conservation != 0, stream ? (cudaStream_t)stream : W[0]->w->stream);
    return IB_OK;
    IB_CATCH
}

int ib_w_apply_many(int32_t nmat, const ib_weighted *const *W, const double *X, int64_t ldx, int32_t nvar, double *const *Y,
                    const int64_t *ldy, int layout, int force_conservation)
{
    IB_TRY
    if (nmat <= 0 || nmat > 8 || !W || !X || !Y || !ldy || nvar <= 0) fail(IB_ERR_ARG, "ib_w_apply_many: bad argument");
    for (int i = 0; i < nmat; ++i) {
        if (!W[i] || !Y[i]) fail(IB_ERR_ARG, "ib_w_apply_many: null matrix or output");
        if (W[i]->w->ncol != W[0]->w->ncol) fail(IB_ERR_DIM, "ib_w_apply_many: the matrices must share their input space");
        for (int k = 0; k < i; ++k) if (W[k] == W[i]) fail(IB_ERR_ARG, "ib_w_apply_many: the same matrix appears twice");
    }
    cudaStream_t s = W[0]->w->stream;
    const int64_t ncol = W[0]->w->ncol;
    const int64_t ldd = (nvar + 1) & ~1;
    DBuf<double> dX((size_t)ncol * ldd, s), tX;
    std::vector<DBuf<double>> dY(nmat), tY(nmat);
    std::vector<double *> dYp(nmat);
    std::vector<int64_t> ldys(nmat, ldd);
    for (int i = 0; i < nmat; ++i) { dY[i].alloc((size_t)W[i]->w->nrow * ldd, s); dYp[i] = dY[i].p; }
    if (layout == IB_LAYOUT_CELL_MAJOR) {
        if (ldx < nvar) fail(IB_ERR_ARG, "ib_w_apply_many: leading dimension smaller than nvar");
        IB_CUDA(cudaMemcpy2DAsync(dX.p, ldd * 8, X, ldx * 8, (size_t)nvar * 8, (size_t)ncol, cudaMemcpyHostToDevice, s));
    } else if (layout == IB_LAYOUT_FIELD_MAJOR) {
        if (ldx < ncol) fail(IB_ERR_ARG, "ib_w_apply_many: leading dimension smaller than the cell count");
        tX.alloc((size_t)nvar * ncol, s);
        IB_CUDA(cudaMemcpy2DAsync(tX.p, ncol * 8, X, ldx * 8, (size_t)ncol * 8, (size_t)nvar, cudaMemcpyHostToDevice, s));
        transpose(tX.p, ncol, nvar, ncol, dX.p, ldd, s);
    } else {
        fail(IB_ERR_ARG, "ib_w_apply_many: unknown layout %d", layout);
    }
    apply_many(nmat, W, dX.p, ldd, nvar, dYp.data(), ldys.data(), force_conservation != 0, s);
    for (int i = 0; i < nmat; ++i) {
        const int64_t nrow = W[i]->w->nrow;
        if (layout == IB_LAYOUT_CELL_MAJOR) {
            if (ldy[i] < nvar) fail(IB_ERR_ARG, "ib_w_apply_many: leading dimension smaller than nvar");
            IB_CUDA(cudaMemcpy2DAsync(Y[i], ldy[i] * 8, dY[i].p, ldd * 8, (size_t)nvar * 8, (size_t)nrow, cudaMemcpyDeviceToHost, s));
        } else {
            if (ldy[i] < nrow) fail(IB_ERR_ARG, "ib_w_apply_many: leading dimension smaller than the cell count");
            tY[i].alloc((size_t)nvar * nrow, s);
            transpose(dY[i].p, ldd, nrow, nvar, tY[i].p, nrow, s);
            IB_CUDA(cudaMemcpy2DAsync(Y[i], ldy[i] * 8, tY[i].p, nrow * 8, (size_t)nrow * 8, (size_t)nvar, cudaMemcpyDeviceToHost, s));
        }
    }
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}

int ib_w_apply_rows_dev(const ib_weighted *W, const double *dX, int64_t ldx, int32_t nvar, double *dY, int64_t ldy, int64_t row_begin,
                        int64_t row_end, void *stream)
{
    IB_TRY
    if (!W || !dX || !dY) fail(IB_ERR_ARG, "ib_w_apply_rows_dev: bad argument");
    apply_dev(*W->w, dX, ldx, nvar, dY, ldy, false, stream ? (cudaStream_t)stream : W->w->stream, row_begin, row_end);
    return IB_OK;
    IB_CATCH
}

int ib_w_apply(const ib_weighted *W, const double *X, int64_t ldx, int32_t nvar, double *Y, int64_t ldy, int layout,
               int force_conservation)
{
    IB_TRY
    if (!W || !X || !Y || nvar <= 0) fail(IB_ERR_ARG, "ib_w_apply: bad argument");
    Weighted &w = *W->w;
    cudaStream_t s = w.stream;
    const int64_t ldd = (nvar + 1) & ~1;    // even leading dimension => 16-byte vector path
    DBuf<double> dX((size_t)w.ncol * ldd, s), dY((size_t)w.nrow * ldd, s);
    if (layout == IB_LAYOUT_CELL_MAJOR) {
        if (ldx < nvar || ldy < nvar) fail(IB_ERR_ARG, "ib_w_apply: leading dimension smaller than nvar");
        IB_CUDA(cudaMemcpy2DAsync(dX.p, ldd * 8, X, ldx * 8, (size_t)nvar * 8, (size_t)w.ncol, cudaMemcpyHostToDevice, s));
        apply_dev(w, dX.p, ldd, nvar, dY.p, ldd, force_conservation != 0, s);
        IB_CUDA(cudaMemcpy2DAsync(Y, ldy * 8, dY.p, ldd * 8, (size_t)nvar * 8, (size_t)w.nrow, cudaMemcpyDeviceToHost, s));
    } else if (layout == IB_LAYOUT_FIELD_MAJOR) {
        if (ldx < w.ncol || ldy < w.nrow) fail(IB_ERR_ARG, "ib_w_apply: leading dimension smaller than the cell count");
        DBuf<double> tX((size_t)nvar * w.ncol, s), tY((size_t)nvar * w.nrow, s);
        IB_CUDA(cudaMemcpy2DAsync(tX.p, w.ncol * 8, X, ldx * 8, (size_t)w.ncol * 8, (size_t)nvar, cudaMemcpyHostToDevice, s));
        transpose(tX.p, w.ncol, nvar, w.ncol, dX.p, ldd, s);
        apply_dev(w, dX.p, ldd, nvar, dY.p, ldd, force_conservation != 0, s);
        transpose(dY.p, ldd, w.nrow, nvar, tY.p, w.nrow, s);
        IB_CUDA(cudaMemcpy2DAsync(Y, ldy * 8, tY.p, w.nrow * 8, (size_t)w.nrow * 8, (size_t)nvar, cudaMemcpyDeviceToHost, s));
    } else {
        fail(IB_ERR_ARG, "ib_w_apply: unknown layout %d", layout);
    }
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}

int ib_w_apply_M(const ib_weighted *W, const double *A_s, int32_t nvar, double *B_s, double fill, int force_conservation)
{
    IB_TRY
    if (!W || !A_s || !B_s || nvar <= 0) fail(IB_ERR_ARG, "ib_w_apply_M: bad argument");
    Weighted &w = *W->w;
    cudaStream_t s = w.stream;
    Dim &db = *w.dims[0], &da = *w.dims[1];
    const int64_t nsa = da.sparse_extent, nsb = db.sparse_extent;
    DBuf<double> dAs((size_t)nvar * nsa, s), dBs((size_t)nvar * nsb, s), dX((size_t)w.ncol * nvar, s), dY((size_t)w.nrow * nvar, s);
    dAs.upload(A_s, (size_t)nvar * nsa);
    if (w.ncol) { densify_kernel<<<(unsigned)cdiv(w.ncol * nvar, 256), 256, 0, s>>>(dAs.p, nsa, da.sparse_of.p, da.identity, w.ncol, nvar, dX.p); IB_LAUNCHED(); }
    apply_dev(w, dX.p, nvar, nvar, dY.p, nvar, force_conservation != 0, s);
    fill_f64_kernel<<<(unsigned)cdiv((int64_t)nvar * nsb, 256), 256, 0, s>>>(dBs.p, (int64_t)nvar * nsb, fill);
    IB_LAUNCHED();
    if (w.nrow) { sparsify_kernel<<<(unsigned)cdiv(w.nrow * nvar, 256), 256, 0, s>>>(dY.p, db.sparse_of.p, db.identity, w.nrow, nvar, nsb, dBs.p); IB_LAUNCHED(); }
    dBs.download(B_s, (size_t)nvar * nsb);
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}

int ib_w_apply_weight(const ib_weighted *W, int dim, const double *A_s, int32_t nvar, double *out)
{
    IB_TRY
    if (!W || !A_s || !out || dim < 0 || dim > 1) fail(IB_ERR_ARG, "ib_w_apply_weight: bad argument");
    Weighted const &w = *W->w;
    // O(n) dot product on gathered entries: metadata-sized, evaluated on the host from downloaded weights
    int64_t n = dim == 0 ? w.nrow : w.ncol;
    std::vector<double> wt((size_t)n);
    (dim == 0 ? w.wM : w.Mw).download(wt.data(), (size_t)n);
    IB_CUDA(cudaStreamSynchronize(w.stream));
    auto ids = w.dims[dim]->ids(w.stream);
    int64_t ns = w.dims[dim]->sparse_extent;
    for (int f = 0; f < nvar; ++f) {
        double s = 0;
        for (int64_t d = 0; d < n; ++d) s += wt[d] * A_s[(size_t)f * ns + ids[d]];
        out[f] = s;
    }
    return IB_OK;
    IB_CATCH
}

int ib_coo_matvec(double *yy, const double *xx, int ignore_nan, int64_t nrow, int64_t ncol, int64_t nnz, const int32_t *row,
                  const int32_t *col, const double *data)
{
    IB_TRY
    // pylib/icebin_cython.cpp:159-192: a Python-side convenience on host arrays (single vector); kept on the host.
    if (!yy || !xx) fail(IB_ERR_ARG, "ib_coo_matvec: bad argument");
    std::vector<char> written((size_t)nrow, 0);
    for (int64_t n = 0; n < nnz; ++n) {
        int64_t r = row[n], c = col[n];
        if (r < 0 || r >= nrow || c < 0 || c >= ncol) fail(IB_ERR_ARG, "ib_coo_matvec: index out of range");
        if (ignore_nan && std::isnan(xx[c])) continue;
        double old = written[r] ? yy[r] : 0.0;
        written[r] = 1;
        yy[r] = old + data[n] * xx[c];
    }
    return IB_OK;
    IB_CATCH
}

int ib_multivec_to_dense(int64_t n, const int64_t *index, const double *weights, const double *vals, int32_t nvar, int64_t nE,
                         double fill, double *scaleE, double *dense)
{
    IB_TRY
    if (n < 0 || nvar < 0 || nE < 0 || (n && (!index || !weights || !vals)) || (nE && (!scaleE || (nvar && !dense))))
        fail(IB_ERR_ARG, "ib_multivec_to_dense: bad argument");
    multivec_to_dense(n, index, weights, vals, nvar, nE, fill, scaleE, dense, default_stream());
    return IB_OK;
    IB_CATCH
}

// ---------------------------------------------------------------------------------------- coupling step
int ib_vartrans_dev(const double *dO, int64_t ldo, int64_t ncell, int32_t nin, int32_t nout, const int32_t *colptr, const int32_t *row,
                    const double *val, const double *b, double *dZ, int64_t ldz, void *stream)
{
    IB_TRY
    if (!colptr || !b || (ncell > 0 && (!dO || !dZ)) || nout <= 0 || (colptr[nout] > 0 && (!row || !val)))
        fail(IB_ERR_ARG, "ib_vartrans_dev: bad argument");
    vartrans_dev(dO, ldo, ncell, nin, nout, colptr, row, val, b, dZ, ldz, stream ? (cudaStream_t)stream : default_stream());
    return IB_OK;
    IB_CATCH
}

int64_t ib_e1ve0c(int32_t nsheet, const ib_weighted *const *XuE1, const ib_weighted *const *XuE0, int64_t nE, int64_t *rows,
                  int64_t *cols, double *vals)
{
    static thread_local std::vector<int64_t> r, c;
    static thread_local std::vector<double> v;
    IB_TRY
    if (nsheet <= 0 || !XuE1 || !XuE0) fail(IB_ERR_ARG, "ib_e1ve0c: bad argument");
    if (!rows) {
        std::vector<Weighted *> a((size_t)nsheet), b((size_t)nsheet);
        for (int i = 0; i < nsheet; ++i) {
            if (!XuE1[i] || !XuE0[i]) fail(IB_ERR_ARG, "ib_e1ve0c: null matrix");
            a[i] = XuE1[i]->w.get(); b[i] = XuE0[i]->w.get();
        }
        e1ve0c(nsheet, a.data(), b.data(), nE, r, c, v, default_stream());
        return (int64_t)r.size();
    }
    if (!cols || !vals) fail(IB_ERR_ARG, "ib_e1ve0c: bad argument");
    std::copy(r.begin(), r.end(), rows);
    std::copy(c.begin(), c.end(), cols);
    std::copy(v.begin(), v.end(), vals);
    return (int64_t)r.size();
    IB_CATCH
}

int ib_w_combine_chunks(int32_t nchunk, const ib_weighted *const *W, int scale, int64_t *M_row, int64_t *M_col, double *M_val,
                        int64_t *wM_index, double *wM_val, int64_t *Mw_index, double *Mw_val)
{
    IB_TRY
    if (nchunk <= 0 || !W || !M_row || !M_col || !M_val || !wM_index || !wM_val || !Mw_index || !Mw_val)
        fail(IB_ERR_ARG, "ib_w_combine_chunks: bad argument");
    std::vector<Weighted *> w((size_t)nchunk);
    for (int i = 0; i < nchunk; ++i) {
        if (!W[i]) fail(IB_ERR_ARG, "ib_w_combine_chunks: null chunk");
        w[i] = W[i]->w.get();
    }
    combine_chunks(nchunk, w.data(), scale != 0, M_row, M_col, M_val, wM_index, wM_val, Mw_index, Mw_val, default_stream());
    return IB_OK;
    IB_CATCH
}

int64_t ib_exgrid_rectilinear(int32_t npoly, const int32_t *pptr, const double *pxy, const int64_t *iA, int32_t nx, const double *xe,
                              int32_t ny, const double *ye, int iI_major_x, int32_t *ex_ij, double *ex_area)
{
    IB_TRY
    return exgrid_rectilinear(npoly, pptr, pxy, iA, nx, xe, ny, ye, iI_major_x, ex_ij, ex_area, default_stream());
    IB_CATCH
}

int ib_merge_pack_dev(const double *dY, int64_t ldy, int32_t nvar, const int32_t *d_loc, int64_t nshared, double *d_buf, void *stream)
{
    IB_TRY
    ensure_device();
    merge_pack(dY, ldy, nvar, d_loc, nshared, d_buf, stream ? (cudaStream_t)stream : default_stream());
    return IB_OK;
    IB_CATCH
}
int ib_merge_unpack_dev(double *dY, int64_t ldy, int32_t nvar, int64_t nrow, const int32_t *d_slot_of_row, const double *d_buf,
                        const double *d_wtot, const double *d_wlocal, void *stream)
{
    IB_TRY
    ensure_device();
    merge_unpack(dY, ldy, nvar, nrow, d_slot_of_row, d_buf, d_wtot, d_wlocal, stream ? (cudaStream_t)stream : default_stream());
    return IB_OK;
    IB_CATCH
}

// ---------------------------------------------------------------------------------------- Hntr
int ib_hntr_create(ib_hntr **out, int imB, int jmB, double offiB, double dlatB, int imA, int jmA, double offiA, double dlatA,
                   double datmis)
{
    IB_TRY
    std::unique_ptr<ib_hntr> H(new ib_hntr);
    H->h = hntr_create(imB, jmB, offiB, dlatB, imA, jmA, offiA, dlatA, datmis);
    *out = H.release();
    return IB_OK;
    IB_CATCH
}
void ib_hntr_free(ib_hntr *h)
{
    if (h) { hntr_free(h->h); delete h; }
}
int ib_hntr_get(const ib_hntr *h, int which, double *dout, int32_t *iout)
{
    IB_TRY
    hntr_get(*h->h, which, dout, iout);
    return IB_OK;
    IB_CATCH
}
int ib_hntr_regrid_dev(const ib_hntr *h, const double *d_wta, double wtm, double wtb, const void *d_src, int src_dtype, double *d_dst,
                       int mean_polar, void *stream)
{
    IB_TRY
    if (!h || !d_src || !d_dst) fail(IB_ERR_ARG, "ib_hntr_regrid_dev: bad argument");
    hntr_regrid_dev(*h->h, d_wta, wtm, wtb, d_src, src_dtype, d_dst, mean_polar != 0, stream ? (cudaStream_t)stream : default_stream());
    return IB_OK;
    IB_CATCH
}
int ib_hntr_regrid(const ib_hntr *h, const double *wta, double wtm, double wtb, const void *src, int src_dtype, double *dst,
                   int mean_polar)
{
    IB_TRY
    if (!h || !src || !dst) fail(IB_ERR_ARG, "ib_hntr_regrid: bad argument");
    ensure_device();
    int d[4];
    hntr_dims(*h->h, d);
    const size_t nA = (size_t)d[0] * d[1], nB = (size_t)d[2] * d[3];
    const size_t esz = src_dtype == IB_SRC_I16 ? 2 : 8;
    cudaStream_t s = default_stream();
    DBuf<char> dsrc(nA * esz, s);
    DBuf<double> dw(wta ? nA : 0, s), dB(nB, s);
    IB_CUDA(cudaMemcpyAsync(dsrc.p, src, nA * esz, cudaMemcpyHostToDevice, s));
    if (wta) dw.upload(wta, nA);
    hntr_regrid_dev(*h->h, wta ? dw.p : nullptr, wtm, wtb, dsrc.p, src_dtype, dB.p, mean_polar != 0, s);
    dB.download(dst, nB);
    IB_CUDA(cudaStreamSynchronize(s));
    return IB_OK;
    IB_CATCH
}
int64_t ib_hntr_matrix(const ib_hntr *h, double R, int64_t *rows, int64_t *cols, double *vals)
{
    IB_TRY
    return hntr_matrix(*h->h, R, rows, cols, vals, default_stream());
    IB_CATCH
}

}   // extern "C"

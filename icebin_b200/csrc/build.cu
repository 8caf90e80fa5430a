// build.cu -- regrid-matrix construction on the device (compiled with -fmad=false: every value below is produced by
// the same IEEE -,/,* sequence as the reference/oracle, so triplet values are bit-reproducible).
//
// Replaces, for the L0 ice parameterisation (the only one implemented upstream, IceRegridder.cpp:121-132):
//   IceRegridder_L0::GvEp/GvI/GvAp            slib/icebin/IceRegridder_L0.cpp:100-214      (K1: per exchange cell)
//   IceRegridder::sApvA/sEpvE                 slib/icebin/IceRegridder.cpp:41-65           (ratioA diagonal)
//   MakeDenseEigenT / SparseSetT / TupleListT ibmisc, used RegridMatrices_Dynamic.cpp:75-300 (K2: ids, sort, reduce)
//   compute_AEvI / compute_IvAE / compute_EvA RegridMatrices_Dynamic.cpp:50-332            (weights + scaling)
//
// Because GvI('X') has exactly one entry per exchange-grid row, every sparse product in the three combiners
// collapses to "one or two terms per exchange cell -> relabel with dense ids -> stable sort by (row,col) ->
// sum runs in order -> diagonal scale" (SURVEY.md 8a "Algebraic reduction").  The per-term arithmetic keeps the
// reference's evaluation order, e.g. EvI term = ((s*w)*(1/s))*s  for  (EpvG * diag(sGvI)) * GvI.
#include <cmath>
#include <cstring>

#include "radix_sort.cuh"
#include "scan.cuh"
#include "segsum.cuh"
#include "types.cuh"
#include "util.cuh"

namespace ib {

// ------------------------------------------------------------------------------------------------
// Dim
__global__ void fill_i32_kernel(int32_t *p, int64_t n, int32_t v)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
__global__ void dim_scatter_kernel(const int64_t *sparse_of, int64_t n, int32_t *dense_of)
{
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d < n) dense_of[sparse_of[d]] = (int32_t)d;
}

void Dim::reserve(int64_t n, cudaStream_t s)
{
    if ((int64_t)sparse_of.n >= n) return;
    int64_t cap = std::max<int64_t>(n, 16);
    DBuf<int64_t> nb((size_t)cap, s);
    if (dense_extent > 0)
        IB_CUDA(cudaMemcpyAsync(nb.p, sparse_of.p, dense_extent * sizeof(int64_t), cudaMemcpyDeviceToDevice, s));
    sparse_of = std::move(nb);
}

void Dim::set_sparse_extent(int64_t n, cudaStream_t s)
{
    if (identity) {
        if (sparse_extent != n)
            fail(IB_ERR_ARG, "identity dimension of extent %ld used for a grid of extent %ld", (long)sparse_extent, (long)n);
        return;
    }
    if (sparse_extent >= 0 && sparse_extent != n && dense_of.n)
        fail(IB_ERR_ARG, "dimension with sparse_extent %ld used for a grid of extent %ld", (long)sparse_extent, (long)n);
    sparse_extent = n;
    if (dense_of.n == (size_t)n && n > 0) return;
    if (n >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "sparse extent too large");
    dense_of.alloc((size_t)n, s);
    if (n) {
        fill_i32_kernel<<<(unsigned)cdiv(n, 256), 256, 0, s>>>(dense_of.p, n, -1);
        IB_LAUNCHED();
    }
    if (!pending.empty()) {
        for (int64_t v : pending)
            if (v < 0 || v >= n) fail(IB_ERR_ARG, "dimension id %ld outside sparse extent %ld", (long)v, (long)n);
        dense_extent = 0;
        reserve((int64_t)pending.size(), s);
        dense_extent = (int64_t)pending.size();
        IB_CUDA(cudaMemcpyAsync(sparse_of.p, pending.data(), pending.size() * sizeof(int64_t), cudaMemcpyHostToDevice, s));
        IB_CUDA(cudaStreamSynchronize(s));
        pending.clear();
        dim_scatter_kernel<<<(unsigned)cdiv(dense_extent, 256), 256, 0, s>>>(sparse_of.p, dense_extent, dense_of.p);
        IB_LAUNCHED();
    }
}

std::vector<int64_t> Dim::ids(cudaStream_t s) const
{
    std::vector<int64_t> v((size_t)dense_extent);
    if (identity) {
        for (int64_t i = 0; i < dense_extent; ++i) v[i] = i;
    } else if (!pending.empty()) {
        v = pending;
    } else if (dense_extent) {
        IB_CUDA(cudaMemcpyAsync(v.data(), sparse_of.p, dense_extent * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
    }
    return v;
}

// ------------------------------------------------------------------------------------------------
// K1: one exchange cell -> its terms
enum Form : int { F_SIS = 0, F_SWIS = 1, F_SISW = 2, F_S = 3, F_SW = 4 };

struct Ctx {
    const int2 *ex_ij;
    const double *ex_area;
    const double *em;
    const double *hc;
    int nhc;
    int64_t nA;
    int kindB, kindA;
    int form;
    int interp;
    int useE;       // 1: two term slots per cell (position p = 2x+t), 0: one (p = x)
    int needPos;    // terms need area > 0 (a GvAp factor is involved, IceRegridder_L0.cpp:208)
    // dims (B = rows / dims[0], A = cols / dims[1])
    int32_t *dense_ofB, *dense_ofA;
    uint32_t *firstB, *firstA;
    int64_t *sparse_ofB, *sparse_ofA;
    int32_t extB, extA;
    int identB, identA;
    int cbits;
    int *err;       // [0] elevation out of bounds count, [1..2] offending value bits
};

struct Term {
    int64_t sB, sA;
    double val;
    bool dimB, dimA, term;
};

// std::lower_bound on hcdefs + the index/weight rules of linterp_1d_b / nearest_1d (IceRegridder_L0.cpp:43-93)
__device__ __forceinline__ int lower_bound_hc(const double *hc, int n, double x)
{
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (__ldg(hc + mid) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ Term eval_term(Ctx const &c, int64_t p)
{
    Term t;
    t.dimB = t.dimA = t.term = false;
    t.sB = t.sA = 0;
    t.val = 0;
    const int64_t x = c.useE ? (p >> 1) : p;
    const int slot = c.useE ? (int)(p & 1) : 0;
    const int2 ij = c.ex_ij[x];
    const double em = c.em[ij.y];
    if (isnan(em)) return t;
    const double s = c.ex_area[x];
    const bool pos = s > 0;

    int k = 0;
    double w = 1.0, w_other = 0.0;
    bool wvalid = true;
    if (c.useE) {
        const double elev = fmax(em, 0.0);
        int hi = lower_bound_hc(c.hc, c.nhc, elev);
        if (c.interp == IB_Z_INTERP) {
            if (hi <= 0) hi = 1;
            if (hi >= c.nhc) {
                if (atomicAdd(c.err, 1) == 0) {
                    c.err[1] = __double2loint(elev);
                    c.err[2] = __double2hiint(elev);
                }
                return t;
            }
            const int lo = hi - 1;
            const double h0 = __ldg(c.hc + lo), h1 = __ldg(c.hc + hi);
            const double ratio = (elev - h0) / (h1 - h0);
            const double w0 = 1.0 - ratio, w1 = ratio;
            k = slot ? hi : lo;
            w = slot ? w1 : w0;
            w_other = slot ? w0 : w1;
            wvalid = (w != 0);
        } else {   // ELEV_CLASS_INTERP: single term in slot 0
            if (slot) return t;
            if (hi <= 0) k = 0;
            else if (hi >= c.nhc) k = c.nhc - 1;
            else {
                const double d0 = fabs(elev - __ldg(c.hc + hi - 1));
                const double d1 = fabs(__ldg(c.hc + hi) - elev);
                k = (d0 <= d1) ? hi - 1 : hi;
            }
        }
    }
    auto sparse_id = [&](int kind) -> int64_t {
        switch (kind) {
            case KA: return ij.x;
            case KE: return (int64_t)k * c.nA + ij.x;
            case KI: return ij.y;
            default: return x;
        }
    };
    // does this position feed the first-appearance stream of a dim of `kind`?  
    auto in_stream = [&](int kind, int other) -> bool {
        switch (kind) {
            case KA: return slot == 0 && pos;                 // GvAp stream
            case KE: return wvalid;                           // GvEp stream, interleaved iE0,iE1
            case KI: return slot == 0;                        // GvI stream
            default:                                          // X comes from the Ur generator of the other side
                return other == KE ? wvalid : (slot == 0 && pos);
        }
    };
    t.sB = sparse_id(c.kindB);
    t.sA = sparse_id(c.kindA);
    t.dimB = in_stream(c.kindB, c.kindA);
    t.dimA = in_stream(c.kindA, c.kindB);
    t.term = wvalid && (c.needPos ? pos : true);
    switch (c.form) {
        case F_SIS: t.val = (s * (1. / s)) * s; break;
        case F_SWIS: t.val = ((s * w) * (1. / s)) * s; break;
        case F_SISW: {
            const double a = s * w;
            double rs = a;
            if (c.interp == IB_Z_INTERP && w_other != 0) rs = slot ? (s * w_other + a) : (a + s * w_other);
            t.val = (s * (1. / rs)) * a;
        } break;
        case F_S: t.val = s; break;
        default: t.val = s * w; break;
    }
    return t;
}

__global__ void k1_first_kernel(Ctx c, int64_t npos)
{
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npos) return;
    Term t = eval_term(c, p);
    if (t.dimB && !c.identB && c.dense_ofB[t.sB] < 0 && c.firstB[t.sB] > (uint32_t)p) atomicMin(&c.firstB[t.sB], (uint32_t)p);
    if (t.dimA && !c.identA && c.dense_ofA[t.sA] < 0 && c.firstA[t.sA] > (uint32_t)p) atomicMin(&c.firstA[t.sA], (uint32_t)p);
}

// scan element: x = new id in dims[0], y = new id in dims[1], z = triplet emitted
struct K1Load {
    Ctx c;
    __device__ uint4 operator()(int64_t p) const
    {
        Term t = eval_term(c, p);
        uint4 e = make_uint4(0, 0, 0, 0);
        if (t.dimB && !c.identB && c.dense_ofB[t.sB] < 0 && c.firstB[t.sB] == (uint32_t)p) e.x = 1;
        if (t.dimA && !c.identA && c.dense_ofA[t.sA] < 0 && c.firstA[t.sA] == (uint32_t)p) e.y = 1;
        if (t.term) e.z = 1;
        return e;
    }
};
// NOTE: K1Assign mutates dense_of while K1Load of other positions may read it; the only positions whose flags
// depend on dense_of[s] are those with first[s]==p, i.e. the single writer itself, so phase-3 flags equal phase-1's.
struct K1Assign {
    Ctx c;
    __device__ void operator()(int64_t p, uint4 prefix, uint4 e) const
    {
        if (!(e.x | e.y)) return;
        Term t = eval_term(c, p);
        if (e.x) { int32_t id = c.extB + (int32_t)prefix.x; c.dense_ofB[t.sB] = id; c.sparse_ofB[id] = t.sB; }
        if (e.y) { int32_t id = c.extA + (int32_t)prefix.y; c.dense_ofA[t.sA] = id; c.sparse_ofA[id] = t.sA; }
    }
};
// After K1Assign the x/y flags all read 0; only the z component of the scan is used here.
struct K1Emit {
    Ctx c;
    uint64_t *keys;
    double *vals;
    __device__ void operator()(int64_t p, uint4 prefix, uint4 e) const
    {
        if (!e.z) return;
        Term t = eval_term(c, p);
        uint32_t r = c.identB ? (uint32_t)t.sB : (uint32_t)c.dense_ofB[t.sB];
        uint32_t q = c.identA ? (uint32_t)t.sA : (uint32_t)c.dense_ofA[t.sA];
        keys[prefix.z] = ((uint64_t)r << c.cbits) | q;
        vals[prefix.z] = t.val;
    }
};

// Is the emitted stream already ordered by its row part / its column part?  (Dense ids of a fresh ice-grid dim are ranks
// of first appearance in this very stream, so with an ice-cell-major exchange grid the I side comes out sorted.)
// flags: bit 0 column parts non-decreasing, bit 1 row parts non-decreasing, bit 2 no run of more than LOCAL_SORT_MAX equal rows
constexpr int LOCAL_SORT_MAX = 32;
__global__ void order_check_kernel(const uint64_t *keys, int64_t n, int cbits, int *flags)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    const uint64_t cm = (uint64_t(1) << cbits) - 1;
    const uint64_t k = keys[i], p = keys[i - 1];
    int clear = 0;
    if ((k & cm) < (p & cm)) clear |= 1;
    if ((k >> cbits) < (p >> cbits)) clear |= 2;
    if (i >= LOCAL_SORT_MAX && (keys[i - LOCAL_SORT_MAX] >> cbits) == (k >> cbits)) clear |= 4;
    if (clear) atomicAnd(flags, ~clear);
}
// Rows already contiguous and in order: each row's few entries are put in column order by one thread (stable: equal
// keys keep their emission order, in which they are summed later)
__global__ void local_sort_kernel(const uint64_t *kin, const double *vin, int64_t n, int cbits, uint64_t *kout, double *vout)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint64_t r = kin[i] >> cbits;
    if (i > 0 && (kin[i - 1] >> cbits) == r) return;        // not the head of its row
    uint64_t k[LOCAL_SORT_MAX];
    double v[LOCAL_SORT_MAX];
    int m = 0;
    for (int64_t j = i; j < n && m < LOCAL_SORT_MAX && (kin[j] >> cbits) == r; ++j, ++m) {
        const uint64_t kk = kin[j];
        const double vv = vin[j];
        int q = m;
        while (q > 0 && k[q - 1] > kk) { k[q] = k[q - 1]; v[q] = v[q - 1]; --q; }
        k[q] = kk; v[q] = vv;
    }
    for (int q = 0; q < m; ++q) { kout[i + q] = k[q]; vout[i + q] = v[q]; }
}

__global__ void gather_i32_kernel(const uint32_t *perm, const int32_t *src, int64_t n, int32_t *dst)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

// diag[d] = ratioA[sparse id mod nA] for A/E dims (sApvA / sEpvE restricted to the dim), else unused
__global__ void ratio_kernel(const int64_t *sparse_of, int identity, int64_t n, int64_t nA, const double *ratioA, double *out)
{
    int64_t d = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= n) return;
    int64_t s = identity ? d : sparse_of[d];
    out[d] = ratioA[s % nA];
}

// Weight vectors and row factors (RegridMatrices_Dynamic.cpp:100-146, 201-233, 290-329)
//   rowmode 0: wM = rowsum,           f = 1/wM
//   rowmode 1: wM = ratio*rowsum,     f = (1/ratio)*(1/rowsum)      (compute_AEvI, correctA)
//   rowmode 2: wM = ratio*rowsum,     f = 1/wM                      (compute_EvA, correctA)
__global__ void row_weights_kernel(const double *rowsum, const double *ratio, int rowmode, int64_t n, double *wM, double *f)
{
    int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n) return;
    double rs = rowsum[r];
    if (rowmode == 0) { wM[r] = rs; f[r] = 1. / rs; }
    else {
        double w = ratio[r];
        double m = w * rs;
        wM[r] = m;
        f[r] = (rowmode == 1) ? (1. / w) * (1. / rs) : 1. / m;
    }
}
__global__ void col_weights_kernel(const double *colsum, const double *ratio, int use_ratio, int64_t n, double *Mw)
{
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    Mw[c] = use_ratio ? ratio[c] * colsum[c] : colsum[c];
}
__global__ void scale_vals_kernel(const int32_t *row, const int32_t *col, const double *f, const double *cratio, int scale,
                                  int use_cratio, int64_t nnz, double *val)
{
    int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nnz) return;
    double v = val[j];
    if (scale) v = f[row[j]] * v;
    if (use_cratio) v = v * cratio[col[j]];
    val[j] = v;
}

// ---- filter_cellsA: domain restriction of the exchange grids (GCMRegridder.cpp:154-164) ---------------------------
struct KeepLoad {
    const int2 *ex_ij;
    const uint8_t *keepA;
    __device__ uint32_t operator()(int64_t x) const { return keepA[ex_ij[x].x] ? 1u : 0u; }
};
struct KeepScatter {
    const int2 *ex_ij;
    const double *ex_area;
    int2 *out_ij;
    double *out_area;
    __device__ void operator()(int64_t x, uint32_t prefix, uint32_t keep) const
    {
        if (keep) { out_ij[prefix] = ex_ij[x]; out_area[prefix] = ex_area[x]; }
    }
};

// IceRegridder::filter_cellsA (IceRegridder.cpp:150-178) -> ExchangeGrid::filter_cellsB (AbbrGrid.cpp:33-49): keep the
// exchange cells whose GCM cell is kept, in their original order; the survivors are renumbered (the comment at
// IceRegridder.cpp:170-176 explains why that is harmless).  Upstream applies the predicate to the exchange cell's own id
// (IceRegridder.cpp:163) and reads the ids from the still-empty output vector (AbbrGrid.cpp:39); what is implemented here is
// the documented intent (GCMRegridder.hpp:380-389: "remove cells that do not interact with the cells we've kept").
void filter_cellsA(Gcm &g, const uint8_t *keepA_h, cudaStream_t s)
{
    ensure_device();
    DBuf<uint8_t> keep((size_t)g.nA, s);
    keep.upload(keepA_h, (size_t)g.nA);
    for (auto &shp : g.sheets) {
        Sheet &sh = *shp;
        int64_t nkeep = 0;
        if (sh.nX > 0) {
            ScanState<uint32_t> st;
            KeepLoad L{sh.ex_ij.p, keep.p};
            transform_scan<uint32_t>(st, sh.nX, L, AddU32{}, 0u, s);
            uint32_t tot = 0;
            IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
            IB_CUDA(cudaStreamSynchronize(s));
            nkeep = tot;
            DBuf<int2> nij((size_t)nkeep, s);
            DBuf<double> narea((size_t)nkeep, s);
            scan_apply<uint32_t>(st, L, AddU32{}, 0u, KeepScatter{sh.ex_ij.p, sh.ex_area.p, nij.p, narea.p}, s);
            IB_CUDA(cudaStreamSynchronize(s));
            sh.ex_ij = std::move(nij);
            sh.ex_area = std::move(narea);
        }
        sh.nX = nkeep;
        // agridA->filter_cells(keepA): the sheet's per-realised-A-cell data follow the GCM grid's realised set
        std::vector<double> proj;
        for (size_t d = 0; d < g.dimA_h.size(); ++d)
            if (keepA_h[g.dimA_h[d]]) proj.push_back(sh.gridA_proj_area_h[d]);
        sh.gridA_proj_area_h.swap(proj);
        std::vector<double> ratio((size_t)g.nA, 0.0);
        size_t k = 0;
        for (size_t d = 0; d < g.dimA_h.size(); ++d)
            if (keepA_h[g.dimA_h[d]]) { ratio[(size_t)g.dimA_h[d]] = g.native_area_h[d] / sh.gridA_proj_area_h[k]; ++k; }
        sh.ratioA.upload(ratio.data(), (size_t)g.nA);
        IB_CUDA(cudaStreamSynchronize(s));
    }
    std::vector<int64_t> dimA;
    std::vector<double> area;
    for (size_t d = 0; d < g.dimA_h.size(); ++d)
        if (keepA_h[g.dimA_h[d]]) { dimA.push_back(g.dimA_h[d]); area.push_back(g.native_area_h[d]); }
    g.dimA_h.swap(dimA);
    g.native_area_h.swap(area);
}

struct SpecPlan {
    int kindB, kindA, form, useE, needPos;
    int combiner;   // 0 AEvI, 1 IvAE, 2 EvA
};

static SpecPlan parse_spec(std::string const &spec)
{
    auto kind = [](char ch) -> int {
        switch (ch) { case 'A': return KA; case 'E': return KE; case 'I': return KI; case 'X': return KX; }
        return -1;
    };
    SpecPlan p{};
    int b = spec.size() == 3 && spec[1] == 'v' ? kind(spec[0]) : -1;
    int a = spec.size() == 3 && spec[1] == 'v' ? kind(spec[2]) : -1;
    bool bAE = (b == KA || b == KE), aAE = (a == KA || a == KE);
    bool bIX = (b == KI || b == KX), aIX = (a == KI || a == KX);
    if (b < 0 || a < 0 || !((bAE && aIX) || (bIX && aAE) || (bAE && aAE && a != b)))
        fail(IB_ERR_ARG, "map::at: no regrid matrix named '%s'", spec.c_str());   // RegridMatrices_Dynamic.cpp:419
    p.kindB = b; p.kindA = a;
    p.useE = (b == KE || a == KE);
    if (bAE && aIX) {
        p.combiner = 0;
        p.needPos = (b == KA);
        p.form = (a == KI) ? (b == KA ? F_SIS : F_SWIS) : (b == KA ? F_S : F_SW);
    } else if (bIX && aAE) {
        p.combiner = 1;
        p.needPos = (a == KA);
        p.form = (b == KI) ? (a == KA ? F_SIS : F_SISW) : (a == KA ? F_S : F_SW);
    } else {
        p.combiner = 2;
        p.needPos = 1;
        p.form = (b == KE) ? F_SWIS : F_SISW;   // EvA : AvE
    }
    return p;
}

std::unique_ptr<Weighted> build_matrix(RegridMatrices &rm, std::string const &spec, Dim *d0, Dim *d1, Params const &prm,
                                       bool own_dims)
{
    std::unique_ptr<Weighted> W(new Weighted);
    if (own_dims) { W->owned[0].reset(d0); W->owned[1].reset(d1); }   // before anything can throw
    W->dims[0] = d0; W->dims[1] = d1;
    ensure_device();
    SpecPlan sp = parse_spec(spec);
    Gcm &g = *rm.gcm;
    Sheet &sh = *g.sheets.at(rm.sheet);
    cudaStream_t s = rm.stream;
    W->stream = s;
    if (g.hcdefs_h.empty() && sp.useE) fail(IB_ERR_ARG, "IceRegridder_L0::GvEp(): hcdefs is zero-length!");
    if (d0 == d1) fail(IB_ERR_ARG, "the two dimensions of a matrix must be distinct objects");

    auto extent = [&](int kind) -> int64_t {
        switch (kind) { case KA: return g.nA; case KE: return g.nE(); case KI: return sh.nI; default: return sh.nX; }
    };
    d0->set_sparse_extent(extent(sp.kindB), s);
    d1->set_sparse_extent(extent(sp.kindA), s);

    const int64_t npos = sp.useE ? 2 * sh.nX : sh.nX;
    if (npos >= (int64_t(1) << 31)) fail(IB_ERR_ARG, "exchange grid too large");

    DBuf<uint32_t> firstB(d0->identity ? 0 : (size_t)d0->sparse_extent, s), firstA(d1->identity ? 0 : (size_t)d1->sparse_extent, s);
    firstB.fill_byte(0xff);
    firstA.fill_byte(0xff);
    DBuf<int> err(4, s);
    err.zero();

    Ctx c{};
    c.ex_ij = sh.ex_ij.p; c.ex_area = sh.ex_area.p; c.em = rm.elevmaskI.p; c.hc = g.hcdefs.p;
    c.nhc = (int)g.hcdefs_h.size(); c.nA = g.nA;
    c.kindB = sp.kindB; c.kindA = sp.kindA; c.form = sp.form; c.interp = sh.interp_style; c.useE = sp.useE; c.needPos = sp.needPos;
    c.dense_ofB = d0->dense_of.p; c.dense_ofA = d1->dense_of.p;
    c.firstB = firstB.p; c.firstA = firstA.p;
    c.extB = (int32_t)d0->dense_extent; c.extA = (int32_t)d1->dense_extent;
    c.identB = d0->identity; c.identA = d1->identity;
    c.err = err.p;

    uint4 tot = make_uint4(0, 0, 0, 0);
    ScanState<uint4> st;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    if (npos > 0) {
        k1_first_kernel<<<(unsigned)cdiv(npos, 256), 256, 0, s>>>(c, npos);
        IB_LAUNCHED();
        transform_scan<uint4>(st, npos, K1Load{c}, AddU32x4{}, zero4, s);
        int herr[4];
        IB_CUDA(cudaMemcpyAsync(&tot, st.total.p, sizeof tot, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaMemcpyAsync(herr, err.p, sizeof herr, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        if (herr[0]) {
            uint64_t bits = ((uint64_t)(uint32_t)herr[2] << 32) | (uint32_t)herr[1];
            double ev;
            memcpy(&ev, &bits, 8);
            fail(IB_ERR_ELEV, "Elevation %g out of bounds (%g, %g)", ev, g.hcdefs_h.front(), g.hcdefs_h.back());
        }
    }
    const int64_t ntrip = tot.z;
    const int64_t nrow = d0->dense_extent + tot.x, ncol = d1->dense_extent + tot.y;
    if (!d0->identity) d0->reserve(nrow, s);
    if (!d1->identity) d1->reserve(ncol, s);
    c.sparse_ofB = d0->sparse_of.p; c.sparse_ofA = d1->sparse_of.p;
    c.cbits = bits_for(ncol);
    const int rbits = bits_for(nrow);
    if (c.cbits + rbits > 64) fail(IB_ERR_ARG, "matrix too large for 64-bit keys");

    DBuf<uint64_t> k0((size_t)ntrip, s), k1((size_t)ntrip, s);
    DBuf<double> v0((size_t)ntrip, s), v1((size_t)ntrip, s);
    if (npos > 0) {
        scan_apply<uint4>(st, K1Load{c}, AddU32x4{}, zero4, K1Assign{c}, s);
        scan_apply<uint4>(st, K1Load{c}, AddU32x4{}, zero4, K1Emit{c, k0.p, v0.p}, s);
    }
    d0->dense_extent = nrow;
    d1->dense_extent = ncol;

    // ---- K2: sort + reduce-by-key -> COO sorted by (row, col).  The stream often arrives half sorted: ordered by ice
    // cell (the exchange grid is ice-cell-major and the ice dim is fresh) -- then a few radix passes or none are needed
    int order = 0;
    if (ntrip > 1) {
        DBuf<int> oflags(1, s);
        const int all = 7;
        IB_CUDA(cudaMemcpyAsync(oflags.p, &all, sizeof all, cudaMemcpyHostToDevice, s));
        order_check_kernel<<<(unsigned)cdiv(ntrip, 256), 256, 0, s>>>(k0.p, ntrip, c.cbits, oflags.p);
        IB_LAUNCHED();
        oflags.download(&order, 1);
        IB_CUDA(cudaStreamSynchronize(s));
    }
    int which;
    if ((order & 6) == 6) {                 // rows in order, short rows: one local sort per row, no radix pass
        local_sort_kernel<<<(unsigned)cdiv(ntrip, 256), 256, 0, s>>>(k0.p, v0.p, ntrip, c.cbits, k1.p, v1.p);
        IB_LAUNCHED();
        which = 1;
    } else if (order & 1) {                 // columns in order: a stable sort on the row bits alone
        which = radix_sort_pairs<uint64_t, double>(k0.p, v0.p, k1.p, v1.p, ntrip, c.cbits + rbits, s, c.cbits);
    } else {
        which = radix_sort_pairs<uint64_t, double>(k0.p, v0.p, k1.p, v1.p, ntrip, c.cbits + rbits, s);
    }
    const uint64_t *sk = which ? k1.p : k0.p;
    const double *sv = which ? v1.p : v0.p;

    int64_t nnz = 0;
    ScanState<uint32_t> hs;
    if (ntrip > 0) {
        transform_scan<uint32_t>(hs, ntrip, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, s);
        uint32_t h;
        IB_CUDA(cudaMemcpyAsync(&h, hs.total.p, sizeof h, cudaMemcpyDeviceToHost, s));
        IB_CUDA(cudaStreamSynchronize(s));
        nnz = h;
    }
    W->nrow = nrow; W->ncol = ncol; W->nnz = nnz;
    W->rowptr.alloc((size_t)nrow + 1, s);
    W->rowptr.zero();
    W->col.alloc((size_t)nnz, s);
    W->val.alloc((size_t)nnz, s);
    W->wM.alloc((size_t)nrow, s);
    W->Mw.alloc((size_t)ncol, s);
    DBuf<int32_t> row((size_t)nnz, s);
    DBuf<double> rowsum((size_t)nrow, s), colsum((size_t)ncol, s), rowf((size_t)nrow, s);
    if (nnz > 0) {
        DBuf<int32_t> segstart((size_t)nnz + 1, s);
        scan_apply<uint32_t>(hs, HeadLoad<uint64_t>{sk}, AddU32{}, 0u, ReduceApply{sk, c.cbits, row.p, W->col.p, segstart.p}, s);
        fill_i32_kernel<<<1, 32, 0, s>>>(segstart.p + nnz, 1, (int32_t)ntrip);      // sentinel (no host copy, no sync)
        IB_LAUNCHED();
        seg_sum(segstart.p, nullptr, sv, nnz, W->val.p, s);      // duplicate (row, col) triplets summed in emission order
        seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(row.p, nnz, nrow, W->rowptr.p);
        IB_LAUNCHED();
    }
    k0.release(); k1.release(); v0.release(); v1.release();

    // ---- weights: row sums in CSR order, column sums in CSC order (the oracle's traversal orders)
    seg_sum(W->rowptr.p, nullptr, W->val.p, nrow, rowsum.p, s);
    {
        DBuf<uint32_t> ck0((size_t)nnz, s), ck1((size_t)nnz, s), pm0((size_t)nnz, s), pm1((size_t)nnz, s);
        DBuf<int32_t> colptr((size_t)ncol + 1, s), csorted((size_t)nnz, s);
        colptr.zero();
        const uint32_t *perm = pm0.p;
        if (nnz > 0) {
            IB_CUDA(cudaMemcpyAsync(ck0.p, W->col.p, nnz * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
            iota_u32_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(pm0.p, nnz);
            IB_LAUNCHED();
            int w2 = radix_sort_pairs<uint32_t, uint32_t>(ck0.p, pm0.p, ck1.p, pm1.p, nnz, c.cbits, s);
            perm = w2 ? pm1.p : pm0.p;
            gather_i32_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(perm, W->col.p, nnz, csorted.p);
            IB_LAUNCHED();
            seg_ptr_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(csorted.p, nnz, ncol, colptr.p);
            IB_LAUNCHED();
        }
        seg_sum(colptr.p, perm, W->val.p, ncol, colsum.p, s);
    }

    // ---- diagonal ratios + scaling
    const bool rowAE = (sp.kindB == KA || sp.kindB == KE), colAE = (sp.kindA == KA || sp.kindA == KE);
    DBuf<double> rratio, cratio;
    int rowmode = 0, use_cratio = 0;
    if (prm.correctA) {
        if (rowAE && nrow) {
            rratio.alloc((size_t)nrow, s);
            ratio_kernel<<<(unsigned)cdiv(nrow, 256), 256, 0, s>>>(d0->sparse_of.p, d0->identity, nrow, g.nA, sh.ratioA.p, rratio.p);
            IB_LAUNCHED();
            rowmode = (sp.combiner == 0) ? 1 : 2;
        }
        if (colAE && ncol && sp.combiner != 0) {
            cratio.alloc((size_t)ncol, s);
            ratio_kernel<<<(unsigned)cdiv(ncol, 256), 256, 0, s>>>(d1->sparse_of.p, d1->identity, ncol, g.nA, sh.ratioA.p, cratio.p);
            IB_LAUNCHED();
            use_cratio = 1;
        }
    }
    if (nrow) { row_weights_kernel<<<(unsigned)cdiv(nrow, 256), 256, 0, s>>>(rowsum.p, rratio.p, rowmode, nrow, W->wM.p, rowf.p); IB_LAUNCHED(); }
    if (ncol) { col_weights_kernel<<<(unsigned)cdiv(ncol, 256), 256, 0, s>>>(colsum.p, cratio.p, use_cratio, ncol, W->Mw.p); IB_LAUNCHED(); }
    if (nnz && (prm.scale || use_cratio)) {
        scale_vals_kernel<<<(unsigned)cdiv(nnz, 256), 256, 0, s>>>(row.p, W->col.p, rowf.p, cratio.p, prm.scale ? 1 : 0, use_cratio, nnz, W->val.p);
        IB_LAUNCHED();
    }
    W->scaled = prm.scale;
    W->conservative = true;
    if (sp.combiner == 1 && prm.smooth()) {
        // the smoothing operator lives on the ice grid (smoother.cpp:43-99 walks agridI); upstream applies it to XvA / XvE
        // too, where rows are exchange cells and the product has no meaning: refused here instead of indexing the ice
        // grid's tables with exchange-cell ids
        if (sp.kindB != KI) fail(IB_ERR_ARG, "matrix %s: sigma smoothing applies to matrices whose rows are ice cells (IvA, IvE)", spec.c_str());
        W->conservative = false;              // RegridMatrices_Dynamic.cpp:167
        smooth_rows(rm, *W, prm);
    }
    IB_CUDA(cudaStreamSynchronize(s));
    return W;
}

}   // namespace ib

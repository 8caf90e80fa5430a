// types.cuh -- host-side objects behind the C ABI handles (device-resident state).
#pragma once
#include <array>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

namespace ib {

enum Kind : int { KA = 0, KE = 1, KI = 2, KX = 3 };

// SparseSetT (spsparse::SparseSet<long,int>) with both directions resident in HBM:
//   dense_of[sparse_extent] (int32, -1 = absent)  and  sparse_of[dense_extent] (int64).
// `identity` sets (id_sparse_set, IceCoupler.cpp:366) carry no tables.
struct Dim {
    int64_t sparse_extent = -1;
    int64_t dense_extent = 0;
    bool identity = false;
    DBuf<int32_t> dense_of;
    DBuf<int64_t> sparse_of;        // capacity may exceed dense_extent
    std::vector<int64_t> pending;   // ids given before the sparse extent was known (host)

    void set_sparse_extent(int64_t n, cudaStream_t s);   // allocates/validates the table
    void reserve(int64_t n, cudaStream_t s);             // sparse_of capacity
    std::vector<int64_t> ids(cudaStream_t s) const;
};

struct Sheet {
    std::string name;
    int64_t nI = 0, nX = 0;
    int interp_style = 0;
    // agridI (AbbrGrid.hpp:93-109)
    std::vector<int64_t> dimI_h;
    DBuf<int64_t> dimI;            // [n_real_I] dense->sparse
    DBuf<double> native_areaI;     // [n_real_I]
    DBuf<double> centroid_xy;      // [n_real_I][2] (may be empty)
    bool dimI_identity = false;
    double bbox[4] = {0, 0, 0, 0};   // x0, x1, y0, y1 of the centroids (smoothing bins)
    // aexgrid (AbbrGrid.hpp:40-78)
    DBuf<int2> ex_ij;              // [nX] (iA, iI)
    DBuf<double> ex_area;          // [nX]
    // native_area/gridA_proj_area per SPARSE A id (0 where not realised): the sApvA/sEpvE diagonal
    DBuf<double> ratioA;           // [nA]
    std::vector<double> gridA_proj_area_h;   // [n_real_A]
};

struct Gcm {
    int64_t nA = 0;
    int correctA = 0;
    std::vector<int64_t> dimA_h;
    std::vector<double> native_area_h;
    std::vector<double> hcdefs_h;
    DBuf<double> hcdefs;
    std::vector<std::unique_ptr<Sheet>> sheets;
    int64_t nE() const { return nA * (int64_t)hcdefs_h.size(); }
};

struct Params {
    bool scale = true;
    bool correctA = false;
    std::array<double, 3> sigma{{0, 0, 0}};
    bool smooth() const { return sigma[0] != 0; }   // RegridMatrices.hpp:31
};

struct RegridMatrices {
    Gcm *gcm = nullptr;
    int sheet = 0;
    Params params;
    DBuf<double> elevmaskI;   // private copy (RegridMatrices_Dynamic.cpp:352)
    cudaStream_t stream = nullptr;
};

// SpMM work decomposition (apply.cu): groups of short rows + chunks of long rows.  One apply at a time per matrix.
struct ApplyPlan {
    bool ready = false;
    int64_t nitems = 0;
    DBuf<int4> items;        // (first_row, nnz_begin, nnz_end, partial slot >= 0 | -(rows in group))
    int64_t nsplit_rows = 0;
    DBuf<int4> split_rows;   // (row, first_partial, nchunks, 0)
    int64_t npartials = 0;
    DBuf<int32_t> slot_srow; // partial slot -> split-row index
    int max_row_len = 0;     // longest row (X-stationary kernel eligibility)
    // distinct-column groups (matrices whose rows all have <= GC_S nonzeros, e.g. IvE, IvA): see gc_plan.cuh
    bool gc = false;
    int64_t gc_nitems = 0;
    DBuf<int4> gc_items;     // (first_row, record_begin, record_end, -(rows in group))
    DBuf<int32_t> gc_ucol;   // [gc_nitems][GC_S] distinct columns of the group (-1 = unused)
    DBuf<int4> gc_recs;      // GcRec (value, slot, flags) per nonzero in row order, one dummy per empty row
    DBuf<unsigned long long> gc_counter;   // dynamic unit counter of the kernel (one apply at a time per matrix)
};

// One-pass plan for wide matrices, alone or stacked with a partner that shares the input space (apply_stream.cu)
struct StreamPlan {
    bool tried = false, ok = false;
    uint64_t partner_uid = 0;   // uid of the stacked partner (0: none)
    int R = 0, R_asked = 0;     // runs (one per CTA and field tile); what the caller asked for (R <= R_asked)
    int64_t nrec = 0, nev = 0, nslots = 0, nfix = 0, nrow1 = 0, nrowS = 0;
    DBuf<int4> recs;            // StRec, four int4 each
    DBuf<int32_t> cols;         // [nrec] column of every record
    DBuf<int2> rowinfo;         // [nrowS] (first partial slot, slots) of rows with several sums; (-1, 0) = no entry at all
    DBuf<int32_t> fixrows;      // [nfix] the rows the fix-up kernel visits
    DBuf<double> partials;      // [nslots][ldp]
};

// Register-blocked plan for tall matrices whose neighbouring rows share columns (apply_group.cu)
struct GroupPlan {
    bool tried = false, ok = false;
    int64_t nitems = 0, nchunks = 0;
    DBuf<int4> items;               // (first row, first chunk, end chunk, rows)
    DBuf<unsigned char> chunks;     // chunk records: columns, presence masks, dense coefficient table
    DBuf<uint32_t> wprefix;         // [nitems + 1] running sum of the items' estimated cost
    DBuf<int> split;                // [split_grid + 1] first item of every CTA for the row block [split_lo, split_hi)
    int64_t split_lo = -1, split_hi = -1;
    int split_grid = 0;
};

uint64_t next_uid();

// linear::Weighted_Eigen: CSR matrix (rows = output space) + weights + dims.
struct Weighted {
    int64_t nrow = 0, ncol = 0, nnz = 0;
    bool conservative = true, scaled = true;
    DBuf<int32_t> rowptr;   // [nrow+1]
    DBuf<int32_t> col;      // [nnz]
    DBuf<double> val;       // [nnz]
    DBuf<double> wM, Mw;    // [nrow], [ncol]
    DBuf<double> cons_v;    // [ncol]  M^T wM (only for non-conservative matrices; conservation correction)
    Dim *dims[2] = {nullptr, nullptr};
    std::unique_ptr<Dim> owned[2];
    ApplyPlan plan;
    StreamPlan stream_plan;
    GroupPlan group_plan;
    uint64_t uid = next_uid();   // identifies the matrix in plans that involve a partner
    cudaStream_t stream = nullptr;
};

std::unique_ptr<Weighted> build_matrix(RegridMatrices &rm, std::string const &spec, Dim *d0, Dim *d1,
                                       Params const &p, bool own_dims);
// build.cu: GCMRegridder_Standard::filter_cellsA (GCMRegridder.cpp:154-164)
void filter_cellsA(Gcm &g, const uint8_t *keepA, cudaStream_t s);
// smooth.cu: M <- smoothI * M  (RegridMatrices_Dynamic.cpp:236-248)
void smooth_rows(RegridMatrices &rm, Weighted &W, Params const &p);
// apply.cu
void apply_dev(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, bool force_conservation,
               cudaStream_t s, int64_t row_begin = 0, int64_t row_end = -1);
// apply_group.cu: tall matrices with ~16 nonzeros per row (sigma-smoothed IvE / IvA); false = does not fit
bool apply_group(Weighted &W, const double *dX, int64_t ldx, int nvar, double *dY, int64_t ldy, int row_lo, int row_hi, cudaStream_t s);
// apply_stream.cu: [Y1; Y2] = [M1; M2] X in one pass over X (M2 may be null); false = the pair does not qualify
bool apply_stream(Weighted &M1, Weighted *M2, const double *dX, int64_t ldx, int nvar, double *dY1, int64_t ldy1, double *dY2,
                  int64_t ldy2, cudaStream_t s);

}   // namespace ib

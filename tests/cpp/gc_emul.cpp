// Host emulation of the distinct-column group plan and of spmm_gc_kernel's flat record walk (TEST ONLY).
// Runs the very routine the device runs (icebin_b200/csrc/gc_plan.cuh), checks the plan's invariants and replays the
// kernel's loop (slot table per item, records in order, mul-then-add) for every field.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../icebin_b200/csrc/gc_plan.cuh"

using namespace ib;

extern "C" int gc_emul(const int32_t *rowptr, const int32_t *col, const double *val, int64_t nrow, const double *X,
                       int nvar, double *Y, int64_t *stats /* items, distinct loads, records */)
{
    const int64_t nblk = (nrow + GC_BLOCK - 1) / GC_BLOCK;
    std::vector<uint32_t> ibase(nblk + 1, 0), rbase(nblk + 1, 0);
    for (int64_t b = 0; b < nblk; ++b) {
        const GcCount c = gc_block<false>(rowptr, col, val, nrow, b, 0u, 0u, nullptr, nullptr, nullptr);
        ibase[b + 1] = ibase[b] + c.items;
        rbase[b + 1] = rbase[b] + c.recs;
    }
    const uint32_t nitems = ibase[nblk], nrecs = rbase[nblk];
    std::vector<GcItem> items(nitems);
    std::vector<int32_t> ucol((size_t)nitems * GC_S, -7);
    std::vector<GcRec> recs(nrecs ? nrecs : 1);
    for (auto &r : recs) { r.v = -1e300; r.slot = 999; r.flags = 999; }
    for (int64_t b = 0; b < nblk; ++b) {
        const GcCount c = gc_block<true>(rowptr, col, val, nrow, b, ibase[b], rbase[b], items.data(), ucol.data(), recs.data());
        if (c.items != ibase[b + 1] - ibase[b] || c.recs != rbase[b + 1] - rbase[b]) return -1;
    }
    // structural invariants
    int64_t next_row = 0, next_rec = 0, ndist = 0;
    for (uint32_t k = 0; k < nitems; ++k) {
        const GcItem &it = items[k];
        const int nrows = -it.negrows;
        if (it.row != next_row || it.rbeg != next_rec) return -2;          // items tile the rows and the records
        if (nrows < 1 || nrows > GC_BLOCK || it.rend - it.rbeg > GC_MAXREC || it.rend - it.rbeg < nrows) return -3;
        if (it.row / GC_BLOCK != (it.row + nrows - 1) / GC_BLOCK) return -4;   // never crosses an aligned block
        int nu = 0;
        while (nu < GC_S && ucol[(size_t)k * GC_S + nu] >= 0) ++nu;
        for (int q = nu; q < GC_S; ++q) if (ucol[(size_t)k * GC_S + q] != -1) return -6;
        for (int q = 0; q < nu; ++q) for (int p = 0; p < q; ++p) if (ucol[(size_t)k * GC_S + q] == ucol[(size_t)k * GC_S + p]) return -7;
        // records against the CSR rows
        int rec = it.rbeg;
        for (int r = it.row; r < it.row + nrows; ++r) {
            if (rowptr[r + 1] == rowptr[r]) {
                if (recs[rec].flags != (GC_LAST | GC_ZERO)) return -8;
                ++rec;
            } else
                for (int j = rowptr[r]; j < rowptr[r + 1]; ++j, ++rec) {
                    const GcRec &rc = recs[rec];
                    if (rc.v != val[j] || (int)(rc.slot & 0xff) >= nu || ucol[(size_t)k * GC_S + (rc.slot & 0xff)] != col[j]) return -9;
                    if ((rc.slot >> 16) != 0) return -9;
                    if (rc.flags != ((j + 1 == rowptr[r + 1]) ? GC_LAST : 0u)) return -10;
                }
        }
        if (rec != it.rend) return -11;
        for (int q = it.rbeg; q < it.rend; ++q) {       // next-slot links inside the item
            const uint32_t want = q + 1 < it.rend ? (recs[q + 1].slot & 0xff) : 0u;
            if (((recs[q].slot >> 8) & 0xff) != want) return -14;
        }
        ndist += nu;
        next_row += nrows; next_rec = it.rend;
    }
    if (next_row != nrow || next_rec != (int64_t)nrecs) return -12;
    stats[0] = nitems; stats[1] = ndist; stats[2] = nrecs;
    // replay of the kernel's loop
    for (uint32_t k = 0; k < nitems; ++k) {
        const GcItem &it = items[k];
        for (int f = 0; f < nvar; ++f) {
            double xs[GC_S];
            for (int i = 0; i < GC_S; ++i) {
                const int c = ucol[(size_t)k * GC_S + i];
                xs[i] = c >= 0 ? X[(int64_t)c * nvar + f] : -7e77;
            }
            int64_t row = it.row;
            double acc = 0.0;
            uint32_t sl = recs[it.rbeg].slot & 0xff;          // the kernel learns a record's slot from its predecessor
            for (int q = it.rbeg; q < it.rend; ++q) {
                const GcRec &rc = recs[q];
                if (!(rc.flags & GC_ZERO)) {
                    const double t = rc.v * xs[sl & (GC_S - 1)];
                    acc = acc + t;
                }
                sl = (rc.slot >> 8) & 0xff;
                if (rc.flags & GC_LAST) { Y[row * nvar + f] = acc; acc = 0.0; ++row; }
            }
            if (row != it.row - it.negrows) return -13;
        }
    }
    return 0;
}

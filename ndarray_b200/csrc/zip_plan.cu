// The Zip planner: host-side replacement for ndarray's per-Zip layout logic
// (array_layout + Layout::intersect + layout_tendency + for_each_core's C/F/strided dispatch,
// reference src/zip/mod.rs:54-77,193-225,309-424,704-758 and src/layout/mod.rs).
//
// ndarray decides between one flat loop and an odometer over "all axes but one".  On the GPU the
// useful question is different: which axis do warps walk so that every producer is read in
// full sectors?  The planner therefore
//   1. drops length-1 axes,
//   2. reverses axes on which the OUTPUT runs backwards (legal for a map: order is irrelevant),
//   3. orders axes by the output's |stride| (slowest first),
//   4. merges neighbouring axes whenever EVERY producer steps through them as one run
//      (the joint form of merge_axes, src/dimension/mod.rs:739-763),
//   5. picks a kernel class (see ndb_kernel_class).
// Pure host code: callable with no GPU present.
#include <algorithm>
#include <cstdlib>

#include "common.h"
#include "zip_plan.h"

namespace ndb {

// reference: src/dimension/mod.rs:58-85 (dim_stride_overlap): a written view may not alias itself.
static bool self_overlaps(const ndb_view *v)
{
    int order[NDB_MAX_NDIM];
    int n = 0;
    for (int i = 0; i < v->ndim; ++i) {
        if (v->dim[i] == 0) return false;  // empty: nothing is written
        if (v->dim[i] > 1) order[n++] = i;
    }
    std::sort(order, order + n, [&](int a, int b) { return llabs(v->stride[a]) < llabs(v->stride[b]); });
    int64_t reach = 0;  // largest offset reachable with the axes seen so far
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        const int64_t s = llabs(v->stride[i]);
        if (s <= reach) return true;
        reach += s * (v->dim[i] - 1);
    }
    return false;
}

int make_zip_plan(int op, int n_parts, const ndb_view *parts, ZipPlan *zp)
{
    clear_error();
    if (n_parts < 1 || n_parts > NDB_MAX_PARTS || !parts || !zp) {
        set_error("Zip: between 1 and %d producers are supported, got %d", NDB_MAX_PARTS, n_parts);
        return NDB_ERR_UNSUPPORTED;
    }
    const int arity = op_arity(op);
    if (arity == -2) {
        set_error("Zip: unknown op %d", op);
        return NDB_ERR_UNSUPPORTED;
    }
    const int n_in = n_parts - 1;
    if ((arity >= 0 && n_in != arity) || (arity == -1 && (n_in < 2 || n_in > 8))) {
        set_error("Zip: op %d takes %d input producer(s), got %d", op, arity, n_in);
        return NDB_ERR_INVALID_ARG;
    }
    for (int p = 0; p < n_parts; ++p) NDB_TRY(validate_view(&parts[p], "Zip producer"));
    const ndb_view &o = parts[0];
    for (int p = 1; p < n_parts; ++p) {
        if (parts[p].dtype != o.dtype) {
            set_error("Zip: producer %d has dtype %d, expected %d", p, parts[p].dtype, o.dtype);
            return NDB_ERR_UNSUPPORTED;
        }
        bool same = parts[p].ndim == o.ndim;
        for (int i = 0; same && i < o.ndim; ++i) same = parts[p].dim[i] == o.dim[i];
        if (!same) {
            // ndarray panics here (src/zip/mod.rs:247-258); message kept verbatim for the host wrapper.
            std::string exp = "[", got = "[";
            for (int i = 0; i < o.ndim; ++i) exp += (i ? ", " : "") + std::to_string(o.dim[i]);
            for (int i = 0; i < parts[p].ndim; ++i) got += (i ? ", " : "") + std::to_string(parts[p].dim[i]);
            set_error("Zip: Producer dimension mismatch, expected: %s], got: %s]", exp.c_str(), got.c_str());
            return NDB_ERR_INCOMPATIBLE_SHAPE;
        }
    }
    if ((op == NDB_OP_SQRT || op == NDB_OP_MATH) && !dtype_is_float(o.dtype)) {
        set_error("Zip: sqrt and the float maths maps need a float element type");
        return NDB_ERR_UNSUPPORTED;
    }
    if ((op == NDB_OP_TF32_HI || op == NDB_OP_TF32_LO) && o.dtype != NDB_F32) {
        set_error("Zip: the tf32 split ops are f32-only");
        return NDB_ERR_UNSUPPORTED;
    }
    if (self_overlaps(&o)) {
        set_error("Zip: the output producer overlaps itself (zero or aliasing strides)");
        return NDB_ERR_ALIASING;
    }

    ZipPlan &z = *zp;
    memset(&z, 0, sizeof(z));
    z.n_parts = n_parts;
    z.dtype = o.dtype;
    z.op = op;
    z.tile_axis = -1;
    for (int p = 0; p < n_parts; ++p) z.ptr[p] = (char *)parts[p].ptr;

    // 1. empty / drop unit axes
    int axes[NDB_MAX_NDIM];
    int n = 0;
    for (int i = 0; i < o.ndim; ++i) {
        if (o.dim[i] == 0) {
            z.kernel = NDB_KERNEL_NOOP;
            z.ndim = 0;
            return NDB_OK;
        }
        if (o.dim[i] != 1) axes[n++] = i;
    }
    const int64_t esz = (int64_t)dtype_size(o.dtype);
    int64_t str[NDB_MAX_PARTS][NDB_MAX_NDIM];
    int64_t dim[NDB_MAX_NDIM];
    for (int k = 0; k < n; ++k) {
        dim[k] = o.dim[axes[k]];
        for (int p = 0; p < n_parts; ++p) str[p][k] = parts[p].stride[axes[k]];
    }
    // 2. run forwards along every axis of the output
    int64_t off[NDB_MAX_PARTS] = {0};
    for (int k = 0; k < n; ++k) {
        if (str[0][k] < 0) {
            for (int p = 0; p < n_parts; ++p) {
                off[p] += (dim[k] - 1) * str[p][k];
                str[p][k] = -str[p][k];
            }
        }
    }
    // 3. slowest output axis first (stable: equal strides keep logical order)
    int order[NDB_MAX_NDIM];
    for (int k = 0; k < n; ++k) order[k] = k;
    std::stable_sort(order, order + n, [&](int a, int b) { return str[0][a] > str[0][b]; });
    int64_t d2[NDB_MAX_NDIM], s2[NDB_MAX_PARTS][NDB_MAX_NDIM];
    for (int k = 0; k < n; ++k) {
        d2[k] = dim[order[k]];
        for (int p = 0; p < n_parts; ++p) s2[p][k] = str[p][order[k]];
    }
    // 4. joint merge, innermost outwards
    int m = 0;  // collapsed rank, built back to front in rd/rs
    int64_t rd[NDB_MAX_NDIM], rs[NDB_MAX_PARTS][NDB_MAX_NDIM];
    for (int k = n - 1; k >= 0; --k) {
        bool merged = false;
        if (m > 0) {
            bool ok = true;
            for (int p = 0; p < n_parts && ok; ++p) ok = s2[p][k] == rd[m - 1] * rs[p][m - 1];
            if (ok) {
                rd[m - 1] *= d2[k];
                merged = true;
            }
        }
        if (!merged) {
            rd[m] = d2[k];
            for (int p = 0; p < n_parts; ++p) rs[p][m] = s2[p][k];
            ++m;
        }
    }
    if (m == 0) {  // every axis had length 1 (or 0-d): one element
        rd[0] = 1;
        for (int p = 0; p < n_parts; ++p) rs[p][0] = 1;
        m = 1;
    }
    z.ndim = m;
    for (int k = 0; k < m; ++k) {  // reverse into slowest-first order
        z.dim[k] = rd[m - 1 - k];
        for (int p = 0; p < n_parts; ++p) z.stride[p][k] = rs[p][m - 1 - k];
    }
    for (int p = 0; p < n_parts; ++p) {
        z.offset[p] = off[p];
        z.ptr[p] += off[p] * esz;
    }
    z.total = 1;
    for (int k = 0; k < m; ++k) z.total *= z.dim[k];

    // 5. classify
    const int inner = m - 1;
    const int64_t vec = 16 / esz;
    bool all_unit = true, all_vec_ok = true;
    for (int p = 0; p < n_parts; ++p) {
        const int64_t s = z.stride[p][inner];
        all_unit &= (s == 1);
        // output: forwards only; inputs: forwards, constant along the row, or backwards
        all_vec_ok &= (s == 1 || (p > 0 && (s == 0 || s == -1)));
    }
    auto aligned = [&](int p) { return ((uintptr_t)z.ptr[p] % 16) == 0; };
    // a backwards row starts at the LAST element of a 16-byte vector
    auto aligned_rev = [&](int p) { return ((uintptr_t)(z.ptr[p] + esz) % 16) == 0; };
    if (m == 1 && all_unit) {
        bool al = true;
        for (int p = 0; p < n_parts; ++p) al &= aligned(p);
        z.kernel = al ? NDB_KERNEL_FLAT_VEC : NDB_KERNEL_FLAT;
        return NDB_OK;
    }
    if (all_vec_ok) {
        bool ok = z.dim[inner] % vec == 0;
        for (int p = 0; p < n_parts && ok; ++p) {
            const int64_t s = z.stride[p][inner];
            if (s == 0) continue;  // scalar-per-row producer: loaded once per thread
            ok &= (s == 1) ? aligned(p) : aligned_rev(p);
            for (int k = 0; k < inner && ok; ++k) ok &= (z.stride[p][k] % vec == 0);
        }
        z.kernel = ok ? NDB_KERNEL_ROWS_VEC : NDB_KERNEL_ROWS;
        return NDB_OK;
    }
    // A producer that is strided along the output's fastest axis but dense along another axis is
    // what ndarray's strided loop handles worst (src/zip/mod.rs:376-424: one useful element per cache
    // line).  Stage it through shared memory instead.
    // "Dense along k" = a stride of at most 4 elements (a stepped slice of a dense axis still packs >= 1/4 of
    // every sector) that is smaller than the producer's stride along the output's fastest axis.
    auto dense_along = [&](int p, int k) {
        const int64_t sk = llabs(z.stride[p][k]), si = llabs(z.stride[p][inner]);
        return sk >= 1 && sk <= 4 && si > sk && si > 4;
    };
    if (z.stride[0][inner] == 1 && m >= 2) {
        int best_axis = -1;
        int64_t best_score = 0;
        for (int k = 0; k < inner; ++k) {
            int64_t score = 0;  // prefer the axis that rescues the most producers, then the longest one
            for (int p = 1; p < n_parts; ++p)
                if (dense_along(p, k)) score += ((int64_t)1 << 40) / llabs(z.stride[p][k]);
            if (score > 0) score += z.dim[k];
            if (score > best_score) {
                best_score = score;
                best_axis = k;
            }
        }
        if (best_axis >= 0 && z.dim[best_axis] >= 16 && z.dim[inner] >= 16) {
            uint32_t mask = 0;
            int staged = 0;
            for (int p = 1; p < n_parts; ++p) {
                if (dense_along(p, best_axis)) {
                    mask |= 1u << p;
                    ++staged;
                }
            }
            if (staged <= ZIP_MAX_STAGED) {
                z.kernel = NDB_KERNEL_TILE;
                z.tile_axis = best_axis;
                z.staged_mask = mask;
                return NDB_OK;
            }
        }
    }
    z.kernel = NDB_KERNEL_ROWS;
    return NDB_OK;
}

}  // namespace ndb

using namespace ndb;

extern "C" {

int ndb_zip_plan(int op, int n_parts, const ndb_view *parts, ndb_zip_plan_t *plan)
{
    if (!plan) return NDB_ERR_INVALID_ARG;
    ZipPlan z;
    NDB_TRY(make_zip_plan(op, n_parts, parts, &z));
    memset(plan, 0, sizeof(*plan));
    plan->kernel = z.kernel;
    plan->ndim = z.ndim;
    for (int k = 0; k < z.ndim; ++k) plan->dim[k] = z.dim[k];
    for (int p = 0; p < z.n_parts; ++p) {
        plan->offset[p] = z.offset[p];
        for (int k = 0; k < z.ndim; ++k) plan->stride[p][k] = z.stride[p][k];
    }
    plan->tile_axis = z.tile_axis;
    plan->staged_mask = z.staged_mask;
    return NDB_OK;
}

// reference: src/zip/mod.rs:278-281 (prefer_f) over the layouts accumulated by build_and (:747-758)
int ndb_zip_prefer_f(int n_inputs, const ndb_view *inputs)
{
    if (n_inputs < 1 || !inputs) return 0;
    uint32_t layout = 0xF;
    int tendency = 0;
    for (int i = 0; i < n_inputs; ++i) {
        if (validate_view(&inputs[i], "prefer_f") != NDB_OK) return 0;
        const uint32_t l = array_layout(&inputs[i]);
        layout &= l;
        tendency += layout_tendency(l);
    }
    const bool c = layout & 1, f = layout & 2;
    return (!c && (f || tendency < 0)) ? 1 : 0;
}

}  // extern "C"

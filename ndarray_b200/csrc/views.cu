// Host-side view arithmetic: the (ptr, dim, stride) rewrites ndarray performs for slicing,
// broadcasting, axis permutation/inversion and splitting.  No device code here; these entry
// points work without a GPU so the Rust/Python host can share one implementation.
#include <algorithm>
#include <cstdlib>

#include "common.h"

using namespace ndb;

namespace {

inline void move_ptr(ndb_view *v, int64_t elems)
{
    v->ptr = (void *)((char *)v->ptr + elems * (int64_t)dtype_size(v->dtype));
}

// Python-style index normalisation of ndarray's abs_index (dimension/mod.rs:389-397).
inline int64_t abs_index(int64_t len, int64_t index) { return index < 0 ? len + index : index; }

}  // namespace

extern "C" {

// reference: src/dimension/mod.rs:399-495 (to_abs_slice + do_slice)
int ndb_view_slice_axis(ndb_view *v, int axis, int64_t start, int64_t end, int has_end, int64_t step)
{
    NDB_TRY(validate_view(v, "slice_axis"));
    if (axis < 0 || axis >= v->ndim) {
        set_error("slice_axis: axis %d out of bounds for %d-d view", axis, v->ndim);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    const int64_t len = v->dim[axis];
    int64_t s = abs_index(len, start);
    int64_t e = abs_index(len, has_end ? end : len);
    if (e < s) e = s;
    if (s < 0 || s > len) {
        set_error("Slice begin %lld is past end of axis of length %lld", (long long)s, (long long)len);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (e < 0 || e > len) {
        set_error("Slice end %lld is past end of axis of length %lld", (long long)e, (long long)len);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (step == 0) {
        set_error("Slice stride must not be zero");
        return NDB_ERR_INVALID_ARG;
    }
    const int64_t m = e - s;
    const int64_t st = v->stride[axis];
    int64_t offset = 0;
    if (m != 0) offset = (step < 0 ? (e - 1) : s) * st;
    const int64_t abs_step = step < 0 ? -step : step;
    const int64_t nd = (m + abs_step - 1) / abs_step;
    v->dim[axis] = nd;
    v->stride[axis] = nd <= 1 ? 0 : st * step;
    move_ptr(v, offset);
    return NDB_OK;
}

// reference: src/impl_methods.rs:1009-1081 (index_axis / collapse_axis + remove_axis)
int ndb_view_index_axis(ndb_view *v, int axis, int64_t index)
{
    NDB_TRY(validate_view(v, "index_axis"));
    if (axis < 0 || axis >= v->ndim) {
        set_error("index_axis: axis %d out of bounds for %d-d view", axis, v->ndim);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (index < 0 || index >= v->dim[axis]) {
        set_error("index_axis: index %lld out of bounds for axis of length %lld", (long long)index,
                  (long long)v->dim[axis]);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    move_ptr(v, index * v->stride[axis]);
    for (int i = axis; i + 1 < v->ndim; ++i) {
        v->dim[i] = v->dim[i + 1];
        v->stride[i] = v->stride[i + 1];
    }
    v->ndim -= 1;
    v->dim[v->ndim] = 0;
    v->stride[v->ndim] = 0;
    return NDB_OK;
}

// reference: src/impl_methods.rs:2371-2430 (broadcast / upcast): right-aligned, length-1 axes and
// missing leading axes get stride 0.
int ndb_view_broadcast(const ndb_view *v, int ndim, const int64_t *dim, ndb_view *out)
{
    NDB_TRY(validate_view(v, "broadcast"));
    if (ndim < 0 || ndim > NDB_MAX_NDIM || !out || (ndim > 0 && !dim)) {
        set_error("broadcast: bad target rank %d", ndim);
        return NDB_ERR_UNSUPPORTED;
    }
    if (ndim < v->ndim) {
        set_error("broadcast: cannot broadcast a %d-d view to %d-d", v->ndim, ndim);
        return NDB_ERR_INCOMPATIBLE_SHAPE;
    }
    ndb_view r;
    memset(&r, 0, sizeof(r));
    r.ptr = v->ptr;
    r.dtype = v->dtype;
    r.ndim = ndim;
    // size_of_shape_checked: product of non-zero axis lengths must fit isize
    {
        unsigned __int128 prod = 1;
        for (int i = 0; i < ndim; ++i) {
            if (dim[i] < 0) return NDB_ERR_INVALID_ARG;
            if (dim[i] != 0) prod *= (unsigned __int128)dim[i];
            if (prod > (unsigned __int128)INT64_MAX) {
                set_error("broadcast: shape overflows isize");
                return NDB_ERR_OVERFLOW;
            }
        }
    }
    for (int i = 0; i < ndim; ++i) r.dim[i] = dim[i];
    const int shift = ndim - v->ndim;
    for (int i = ndim - 1; i >= 0; --i) {
        const int j = i - shift;
        if (j < 0) {
            r.stride[i] = 0;
        } else if (dim[i] == v->dim[j]) {
            r.stride[i] = v->stride[j];
        } else if (v->dim[j] == 1) {
            r.stride[i] = 0;
        } else {
            set_error("broadcast: axis %d of length %lld cannot be broadcast to %lld", j, (long long)v->dim[j],
                      (long long)dim[i]);
            return NDB_ERR_INCOMPATIBLE_SHAPE;
        }
    }
    *out = r;
    return NDB_OK;
}

// reference: src/impl_methods.rs:2522-2560 (permuted_axes); t()/reversed_axes is perm = (n-1 .. 0)
int ndb_view_permute(const ndb_view *v, const int *perm, ndb_view *out)
{
    NDB_TRY(validate_view(v, "permute"));
    if (!perm || !out) return NDB_ERR_INVALID_ARG;
    bool seen[NDB_MAX_NDIM] = {false};
    for (int i = 0; i < v->ndim; ++i) {
        if (perm[i] < 0 || perm[i] >= v->ndim || seen[perm[i]]) {
            set_error("permuted_axes: not a permutation");
            return NDB_ERR_INVALID_ARG;
        }
        seen[perm[i]] = true;
    }
    ndb_view r;
    memset(&r, 0, sizeof(r));
    r.ptr = v->ptr;
    r.dtype = v->dtype;
    r.ndim = v->ndim;
    for (int i = 0; i < v->ndim; ++i) {
        r.dim[i] = v->dim[perm[i]];
        r.stride[i] = v->stride[perm[i]];
    }
    *out = r;
    return NDB_OK;
}

// reference: src/impl_methods.rs:2676-2690 (invert_axis): negate the stride, start from the last element.
int ndb_view_invert_axis(ndb_view *v, int axis)
{
    NDB_TRY(validate_view(v, "invert_axis"));
    if (axis < 0 || axis >= v->ndim) {
        set_error("invert_axis: axis %d out of bounds", axis);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    const int64_t m = v->dim[axis];
    if (m != 0) {
        move_ptr(v, (m - 1) * v->stride[axis]);
        v->stride[axis] = -v->stride[axis];
    }
    return NDB_OK;
}

// reference: src/dimension/mod.rs:739-763 (merge_axes)
int ndb_view_merge_axes(ndb_view *v, int take, int into)
{
    if (validate_view(v, "merge_axes") != NDB_OK) return -1;
    if (take < 0 || take >= v->ndim || into < 0 || into >= v->ndim) {
        set_error("merge_axes: axis out of bounds");
        return -NDB_ERR_OUT_OF_BOUNDS;
    }
    const int64_t into_len = v->dim[into], into_stride = v->stride[into];
    const int64_t take_len = v->dim[take], take_stride = v->stride[take];
    const int64_t merged = into_len * take_len;
    if (take_len <= 1) {
        v->dim[into] = merged;
        v->dim[take] = merged == 0 ? 0 : 1;
        return 1;
    }
    if (into_len <= 1) {
        v->stride[into] = take_stride;
        v->dim[into] = merged;
        v->dim[take] = merged == 0 ? 0 : 1;
        return 1;
    }
    if (take_stride == into_len * into_stride) {
        v->dim[into] = merged;
        v->dim[take] = 1;
        return 1;
    }
    return 0;
}

// reference: src/impl_methods.rs:1075-1081 (collapse_axis) -> dimension::do_collapse_axis (mod.rs:372-386)
int ndb_view_collapse_axis(ndb_view *v, int axis, int64_t index)
{
    NDB_TRY(validate_view(v, "collapse_axis"));
    if (axis < 0 || axis >= v->ndim) {
        set_error("collapse_axis: axis %d out of bounds", axis);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (index < 0 || index >= v->dim[axis]) {
        set_error("collapse_axis: Index %lld must be less than axis length %lld for array with shape", (long long)index,
                  (long long)v->dim[axis]);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    move_ptr(v, index * v->stride[axis]);
    v->dim[axis] = 1;
    return NDB_OK;
}

// reference: src/impl_raw_views.rs:113-135 (split_at): the right part starts `index` steps along `axis`,
// unless that would point past the allocation (index == len), in which case the pointer is left alone.
int ndb_view_split_at(const ndb_view *v, int axis, int64_t index, ndb_view *left, ndb_view *right)
{
    NDB_TRY(validate_view(v, "split_at"));
    if (!left || !right) return NDB_ERR_INVALID_ARG;
    if (axis < 0 || axis >= v->ndim) {
        set_error("split_at: axis %d out of bounds", axis);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    if (index < 0 || index > v->dim[axis]) {
        set_error("split_at: index %lld out of bounds for axis of length %lld", (long long)index,
                  (long long)v->dim[axis]);
        return NDB_ERR_OUT_OF_BOUNDS;
    }
    ndb_view l = *v, r = *v;
    l.dim[axis] = index;
    r.dim[axis] = v->dim[axis] - index;
    if (index != v->dim[axis]) move_ptr(&r, index * v->stride[axis]);
    *left = l;
    *right = r;
    return NDB_OK;
}

}  // extern "C"

namespace ndb {

// reference: src/dimension/mod.rs:687-711
bool is_layout_c(const ndb_view *v)
{
    if (v->ndim == 1) return v->stride[0] == 1 || v->dim[0] <= 1;
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] == 0) return true;
    int64_t contig = 1;
    for (int i = v->ndim - 1; i >= 0; --i) {
        if (v->dim[i] != 1) {
            if (v->stride[i] != contig) return false;
            contig *= v->dim[i];
        }
    }
    return true;
}

// reference: src/dimension/mod.rs:713-737
bool is_layout_f(const ndb_view *v)
{
    if (v->ndim == 1) return v->stride[0] == 1 || v->dim[0] <= 1;
    for (int i = 0; i < v->ndim; ++i)
        if (v->dim[i] == 0) return true;
    int64_t contig = 1;
    for (int i = 0; i < v->ndim; ++i) {
        if (v->dim[i] != 1) {
            if (v->stride[i] != contig) return false;
            contig *= v->dim[i];
        }
    }
    return true;
}

enum { L_CORDER = 1, L_FORDER = 2, L_CPREFER = 4, L_FPREFER = 8 };

// reference: src/zip/mod.rs:54-77 (array_layout)
uint32_t array_layout(const ndb_view *v)
{
    const int n = v->ndim;
    if (is_layout_c(v)) {
        int big = 0;
        for (int i = 0; i < n; ++i) big += v->dim[i] > 1;
        if (n <= 1 || big <= 1) return L_CORDER | L_CPREFER | L_FORDER | L_FPREFER;
        return L_CORDER | L_CPREFER;
    }
    if (n > 1 && is_layout_f(v)) return L_FORDER | L_FPREFER;
    if (n > 1) {
        if (v->dim[0] > 1 && v->stride[0] == 1) return L_FPREFER;
        if (v->dim[n - 1] > 1 && v->stride[n - 1] == 1) return L_CPREFER;
    }
    return 0;
}

int layout_tendency(uint32_t l)
{
    return (int)((l & L_CORDER) != 0) - (int)((l & L_FORDER) != 0) + (int)((l & L_CPREFER) != 0) -
           (int)((l & L_FPREFER) != 0);
}

}  // namespace ndb

extern "C" {

uint32_t ndb_view_layout(const ndb_view *v)
{
    if (validate_view(v, "layout") != NDB_OK) return 0;
    return array_layout(v);
}

// reference: src/dimension/dimension_trait.rs:294-323 (is_contiguous)
int ndb_view_is_contiguous(const ndb_view *v)
{
    if (validate_view(v, "is_contiguous") != NDB_OK) return 0;
    const int n = v->ndim;
    // default (C) strides, with the all-zero convention for empty shapes (dimension_trait.rs:128-150)
    {
        bool empty = false;
        for (int i = 0; i < n; ++i) empty |= v->dim[i] == 0;
        int64_t expect[NDB_MAX_NDIM];
        int64_t c = 1;
        for (int i = n - 1; i >= 0; --i) {
            expect[i] = empty ? 0 : c;
            c *= v->dim[i];
        }
        bool same = true;
        for (int i = 0; i < n; ++i) same &= expect[i] == v->stride[i];
        if (same) return 1;
    }
    if (n == 1) return v->dim[0] <= 1 || v->stride[0] == -1;
    int order[NDB_MAX_NDIM];
    for (int i = 0; i < n; ++i) order[i] = i;
    std::stable_sort(order, order + n, [&](int a, int b) { return llabs(v->stride[a]) < llabs(v->stride[b]); });
    int64_t c = 1;
    for (int k = 0; k < n; ++k) {
        const int i = order[k];
        if (v->dim[i] != 1 && llabs(v->stride[i]) != c) return 0;
        c *= v->dim[i];
    }
    return 1;
}

// reference: src/dimension/dimension_trait.rs:343-357 (generic: LAST axis wins ties, len-1 axes are NOT
// skipped) and :715-724 (Ix2: Axis(0) only if strictly smaller).
int ndb_view_min_stride_axis(const ndb_view *v)
{
    if (validate_view(v, "min_stride_axis") != NDB_OK || v->ndim == 0) return -1;
    const int n = v->ndim;
    if (n == 1) return 0;
    int best = n - 1;
    for (int i = n - 1; i >= 0; --i)
        if (llabs(v->stride[i]) < llabs(v->stride[best])) best = i;
    return best;
}

}  // extern "C"

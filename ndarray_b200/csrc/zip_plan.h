#pragma once
#include "common.h"

namespace ndb {

constexpr int ZIP_MAX_STAGED = 2;  // producers transposed through shared memory per tile kernel

struct ZipPlan {
    int kernel;  // ndb_kernel_class
    int ndim;    // collapsed rank (>= 1 unless NOOP)
    int n_parts;
    int dtype;
    int op;
    int tile_axis;
    uint32_t staged_mask;
    int64_t total;  // number of elements
    int64_t dim[NDB_MAX_NDIM];
    int64_t stride[NDB_MAX_PARTS][NDB_MAX_NDIM];
    int64_t offset[NDB_MAX_PARTS];
    char *ptr[NDB_MAX_PARTS];  // base pointers with `offset` applied
};

int make_zip_plan(int op, int n_parts, const ndb_view *parts, ZipPlan *zp);
// fused Zip + reduction over two dense aligned producers (reduce.cu); NDB_ERR_UNSUPPORTED if no kernel for (op, redop)
int zip2_reduce_dense(ndb_ctx *ctx, int op, int redop, int dtype, const void *x, const void *y, int64_t n, void *out_host);

int zip_fold_dense(ndb_ctx *ctx, const ZipPlan &z, int redop, const void *scalars, void *out_dev, bool peer);

bool is_layout_c(const ndb_view *v);
bool is_layout_f(const ndb_view *v);
uint32_t array_layout(const ndb_view *v);
int layout_tendency(uint32_t l);

}  // namespace ndb

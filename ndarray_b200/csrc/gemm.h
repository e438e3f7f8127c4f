#pragma once
#include "common.h"

namespace ndb {

// C[m,n] <- alpha * A[m,k] * B[k,n] + beta * C, signed element strides (row stride, column stride),
// exactly the parameter set of matrixmultiply::{s,d}gemm (reference src/linalg/impl_linalg.rs:512-546).
struct GemmArgs {
    const char *a;
    const char *b;
    char *c;
    int64_t m, n, k;
    int64_t rsa, csa, rsb, csb, rsc, csc;
};

bool gemm_f64_eligible(const GemmArgs &g);
int gemm_f64_launch(ndb_ctx *ctx, const GemmArgs &g, double alpha, double beta);
// TMA-fed f64 kernel (gemm_f64_tma.cu); NDB_ERR_UNSUPPORTED = shape/driver outside its rules, use the cp.async kernel
int gemm_f64_tma_launch(ndb_ctx *ctx, const double *a, int64_t lda, bool a_kc, const double *b, int64_t ldb, bool b_kc, double *c,
                        int64_t ldc, int64_t m, int64_t n, int64_t k, double alpha, double beta);
bool gemm_f32_eligible(const GemmArgs &g);
int gemm_f32_launch(ndb_ctx *ctx, const GemmArgs &g, float alpha, float beta);

}  // namespace ndb

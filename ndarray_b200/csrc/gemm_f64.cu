// f64 general_mat_mul on the FP64 tensor cores (DMMA).  Placeholder until the pipeline lands: nothing is
// eligible, so gemm.cu's tiled kernel serves every f64 call.
#include "gemm.h"

namespace ndb {
bool gemm_f64_eligible(const GemmArgs &) { return false; }
int gemm_f64_launch(ndb_ctx *, const GemmArgs &, double, double)
{
    set_error("gemm_f64: not built");
    return NDB_ERR_UNSUPPORTED;
}
}  // namespace ndb

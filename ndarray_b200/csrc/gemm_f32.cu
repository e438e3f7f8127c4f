// f32 general_mat_mul as 3xTF32 on tcgen05/TMEM.  Placeholder until the kernel lands: nothing is
// eligible, so gemm.cu's tiled kernel serves every f32 call.
#include "gemm.h"

namespace ndb {
bool gemm_f32_eligible(const GemmArgs &) { return false; }
int gemm_f32_launch(ndb_ctx *, const GemmArgs &, float, float)
{
    set_error("gemm_f32: not built");
    return NDB_ERR_UNSUPPORTED;
}
}  // namespace ndb

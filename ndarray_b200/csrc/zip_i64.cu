// Elementwise Zip kernels instantiated for int64_t (one translation unit per element type so the
// build parallelises).
#include "zip_kernels.cuh"

namespace ndb {
int zip_dispatch_i64(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    return zip_dispatch_typed<int64_t>(ctx, z, scalars);
}
}  // namespace ndb

// Elementwise Zip kernels instantiated for int32_t (one translation unit per element type so the
// build parallelises).
#include "zip_kernels.cuh"

namespace ndb {
int zip_dispatch_i32(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    return zip_dispatch_typed<int32_t>(ctx, z, scalars);
}
}  // namespace ndb

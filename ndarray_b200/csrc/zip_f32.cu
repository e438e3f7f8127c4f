// Elementwise Zip kernels instantiated for float (one translation unit per element type so the
// build parallelises).
#include "zip_kernels.cuh"

namespace ndb {
int zip_dispatch_f32(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    return zip_dispatch_typed<float>(ctx, z, scalars);
}
}  // namespace ndb

// Elementwise Zip kernels instantiated for double (one translation unit per element type so the
// build parallelises).
#include "zip_kernels.cuh"

namespace ndb {
int zip_dispatch_f64(ndb_ctx *ctx, const ZipPlan &z, const void *scalars)
{
    return zip_dispatch_typed<double>(ctx, z, scalars);
}
}  // namespace ndb

/* registration.cu -- placeholder until the registration kernels land (same session). */
#include "context.h"
extern "C" {
int m3d_ransac_registration(m3d_ctx *ctx, const double *, size_t, const double *, size_t, const size_t *,
                            const size_t *, size_t, double, int, double, double, uint32_t, double *, m3d_reg_stats *) {
    return ctx ? ctx->fail(M3D_ERR_INTERNAL, "m3d_ransac_registration: not built yet") : M3D_ERR_INVALID_ARG;
}
int m3d_least_squares_transform(m3d_ctx *ctx, const double *, const double *, size_t, int, double *) {
    return ctx ? ctx->fail(M3D_ERR_INTERNAL, "m3d_least_squares_transform: not built yet") : M3D_ERR_INVALID_ARG;
}
}

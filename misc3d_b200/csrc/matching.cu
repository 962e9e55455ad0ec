/* matching.cu -- placeholder until the brute-force L2 kernel lands (same session). */
#include "context.h"
extern "C" {
int m3d_match_correspondence(m3d_ctx *ctx, const double *, size_t, const double *, size_t, int, int, int, size_t *,
                             size_t *, size_t *, float *) {
    return ctx ? ctx->fail(M3D_ERR_INTERNAL, "m3d_match_correspondence: not built yet") : M3D_ERR_INVALID_ARG;
}
int m3d_nearest(m3d_ctx *ctx, const double *, size_t, const double *, size_t, int, size_t *, float *) {
    return ctx ? ctx->fail(M3D_ERR_INTERNAL, "m3d_nearest: not built yet") : M3D_ERR_INVALID_ARG;
}
}

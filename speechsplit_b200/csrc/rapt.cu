// rapt.cu - placeholder until the RAPT kernels land (stage a6).
#include "common.cuh"
namespace ssfe {
struct RaptTables { int dummy; };
int init_rapt(ssfe_ctx *) { return SSFE_OK; }
void free_rapt(ssfe_ctx *) {}
int rapt_run(ssfe_ctx *ctx, const float *, const int64_t *, const int64_t *, const int64_t *, int,
             const float *, const float *, float *)
{
    return set_error(ctx, SSFE_ERR_INVALID, "RAPT kernels not built yet");
}
}  // namespace ssfe

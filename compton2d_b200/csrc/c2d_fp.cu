// c2d_fp.cu -- Fokker-Planck update (placeholder until the batched Chang-Cooper kernel lands)
#include "c2d_ctx.h"
namespace c2d {
void Ctx::fp_free() {}
int Ctx::sigma_setup() { return fail(C2D_EINVAL, "sigma_mode 1 not built yet"); }
int Ctx::sigma_build() { return fail(C2D_EINVAL, "sigma_mode 1 not built yet"); }
}
extern "C" {
int c2d_ic_loss(c2d_ctx *, double *) { return c2d::fail(C2D_EINVAL, "FP not built yet"); }
int c2d_update_fp(c2d_ctx *, const c2d_fp_in *, c2d_fp_out *) { return c2d::fail(C2D_EINVAL, "FP not built yet"); }
int c2d_tally_bench(c2d_ctx *, int64_t, int, float *, int64_t *) { return c2d::fail(C2D_EINVAL, "not built yet"); }
}

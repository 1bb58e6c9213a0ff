// comm.cu -- multi-GPU domain decomposition (one process per GPU).  Round-1 state: single rank only;
// the multi-rank halo exchange over NVLink peer memory is being built on top of these entry points.
#include "lqcd_internal.cuh"

struct CommState { int dummy; };

int comm_destroy(lqcd_ctx *ctx) { delete ctx->comm; ctx->comm = nullptr; return LQCD_OK; }

int comm_allreduce_sum(lqcd_ctx *ctx, double *, int) {
    if (ctx->nranks == 1) return LQCD_OK;
    return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank reductions are not connected");
}

int comm_dslash(lqcd_ctx *ctx, const lqcd_op *, cplx *, const cplx *, int, const DslashFuse *) {
    return lqcd_fail(ctx, LQCD_ERR_COMM, "multi-rank dslash is not connected (call lqcd_comm_connect)");
}

extern "C" int lqcd_comm_export(lqcd_ctx *ctx, void *) { return lqcd_fail(ctx, LQCD_ERR_COMM, "not implemented"); }
extern "C" int lqcd_comm_connect(lqcd_ctx *ctx, const void *) { return lqcd_fail(ctx, LQCD_ERR_COMM, "not implemented"); }

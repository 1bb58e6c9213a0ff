// force.cu -- fermion force assembly (calc_UdSfdU!, src/md/AbstractMD.jl:129).  Placeholder until the
// device outer-product kernel lands; the solve part already runs on the device through lqcd_solve.
#include "lqcd_internal.cuh"
extern "C" int lqcd_fermion_force(lqcd_ctx *ctx, const lqcd_op *, const lqcd_fermion *, lqcd_fermion *, double, int,
                                  double *const[4], int *, double *) {
    return lqcd_fail(ctx, LQCD_ERR_ARG, "lqcd_fermion_force: not implemented yet");
}

/* examples/propagator.c -- the C ABI from plain C (the same calls the Julia ccall shim makes).
 *
 * Loads a gauge configuration in one of the reference's formats (or generates a warm synthetic one), builds the Wilson operator
 * with the parameters of test/test_wilson.toml, solves D x = b for the 12 spin-colour point sources at the origin in ONE batched
 * call -- what calc_quark_propagators_point_source (src/measurements/unusedfiles/measure_Pion_correlator.jl:333-409) does with
 * twelve solve_DinvX! calls -- and prints the pion correlator C(t) = sum_{x, sources, components} |S(x, t)|^2.
 *
 *   gcc -std=c99 -Iinclude examples/propagator.c -Llatticeqcd.jl_b200 -l:liblqcd_b200.so -Wl,-rpath,$PWD/latticeqcd.jl_b200 -lm -o propagator
 *   ./propagator 8 8 8 16 [conf.ildg]
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "lqcd_b200.h"

#define CHECK(call)                                                                            \
    do {                                                                                       \
        int st_ = (call);                                                                      \
        if (st_ != LQCD_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, st_, lqcd_last_error(ctx)); return 1; } \
    } while (0)

int main(int argc, char **argv) {
    int dims[4] = {8, 8, 8, 16}, pg[4] = {1, 1, 1, 1};
    lqcd_ctx *ctx = NULL;
    for (int i = 0; i < 4 && i + 1 < argc; i++) dims[i] = atoi(argv[i + 1]);
    CHECK(lqcd_ctx_create(dims, pg, 0, 0, &ctx));
    if (argc > 5) CHECK(lqcd_gauge_load(ctx, argv[5], LQCD_IO_ILDG));
    else          CHECK(lqcd_gauge_random(ctx, 111, 0.3));
    double plaq = 0.0;
    CHECK(lqcd_gauge_plaquette(ctx, &plaq));
    printf("plaquette %.12f\n", plaq);

    lqcd_op op;
    memset(&op, 0, sizeof op);
    op.kind = LQCD_WILSON; op.kappa = 0.141139; op.r = 1.0;            /* test/test_wilson.toml, parameter_structs.jl:126-133 */
    op.bc[0] = op.bc[1] = op.bc[2] = 1.0; op.bc[3] = -1.0;

    enum { NSRC = 12 };
    lqcd_fermion *b[NSRC], *x[NSRC];
    const size_t V = (size_t)dims[0] * dims[1] * dims[2] * dims[3];
    double *host = calloc(V * 12 * 2, sizeof(double));                  /* psi[c, x, y, z, t, alpha], complex */
    if (!host) return 1;
    for (int i = 0; i < NSRC; i++) {
        const int is = i % 4, ic = i / 4;                               /* measure_Pion_correlator.jl:360-361 */
        CHECK(lqcd_fermion_alloc(ctx, LQCD_WILSON, &b[i]));
        CHECK(lqcd_fermion_alloc(ctx, LQCD_WILSON, &x[i]));
        memset(host, 0, V * 12 * 2 * sizeof(double));
        host[2 * (ic + 3 * (V * (size_t)is))] = 1.0;                    /* value 1 at the origin */
        CHECK(lqcd_fermion_upload(ctx, b[i], host, 0));
        CHECK(lqcd_fermion_zero(ctx, x[i]));                            /* clear_fermion!(p): zero initial guess */
    }
    int iters[NSRC];
    double resid[NSRC];
    CHECK(lqcd_solve_multi(ctx, &op, x, (const lqcd_fermion *const *)b, NSRC, LQCD_SOLVER_CGNR, LQCD_OP_D, 1e-19, 3000, iters, resid));
    double *corr = calloc(dims[3], sizeof(double));
    for (int i = 0; i < NSRC; i++) {
        printf("source %2d: %d iterations, |r|^2 = %.3e\n", i, iters[i], resid[i]);
        CHECK(lqcd_fermion_download(ctx, x[i], host, 0));
        const size_t Vs = (size_t)dims[0] * dims[1] * dims[2];
        for (int al = 0; al < 4; al++)
            for (int t = 0; t < dims[3]; t++)
                for (size_t s = 0; s < Vs * 3 * 2; s++) {
                    const double v = host[(V * (size_t)al + Vs * (size_t)t) * 3 * 2 + s];
                    corr[t] += v * v;
                }
    }
    for (int t = 0; t < dims[3]; t++) printf("C(%d) = %.10e\n", t, corr[t]);
    for (int i = 0; i < NSRC; i++) { lqcd_fermion_free(ctx, b[i]); lqcd_fermion_free(ctx, x[i]); }
    free(host); free(corr);
    lqcd_ctx_destroy(ctx);
    return 0;
}

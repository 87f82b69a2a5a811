// Accessors shared by the translation units of libsimpimc_b200.so (not part of the public ABI;
// hidden from include/simpimc_b200.h).  capi.cu owns the context; comm.cc builds the multi-GPU
// data plane on the public entry points plus these.
#ifndef SIMPIMC_B200_INTERNAL_H_
#define SIMPIMC_B200_INTERNAL_H_

#include "../../include/simpimc_b200.h"

#ifdef __cplusplus
extern "C" {
#endif
int pimc_internal_fail(int code, const char *msg);  /* records the pimc_last_error() text, returns code */
int pimc_internal_device(const pimc_ctx *ctx);
int pimc_internal_n_clones(const pimc_ctx *ctx);
int pimc_internal_n_species(const pimc_ctx *ctx);
int pimc_internal_n_part(const pimc_ctx *ctx, int species); /* -1 when out of range */
/* which = 0 / 1 / 2: whole-path action / dU/dbeta / potential of n actions of one context into d_out[action][clone]
 * (device memory).  The first action runs on the context's stream, the others on side streams forked from it and joined
 * back (capturable): each whole-path kernel is one persistent CTA per SM, so the next action's CTAs start on the SMs the
 * previous one has already left instead of waiting for its slowest.  Sequential while per-kernel timing is on. */
int pimc_internal_evaluate_many(pimc_ctx *ctx, int which, pimc_action *const *actions, int n, double *d_out);
#ifdef __cplusplus
}
#endif
#endif

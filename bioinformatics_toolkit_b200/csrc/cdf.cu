// cdf.cu — scoreCDF / pValueToScore on the device (Motif.hs:213-214,282-326).  (stub: round-1 WIP)
#include "internal.h"
using namespace pwmscan;

extern "C" int pwmscan_score_cdf(pwmscan_ctx* ctx, const pwmscan_motifs*, int32_t, double**, double**, int64_t*) {
    return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_score_cdf: not implemented yet");
}
extern "C" int pwmscan_pvalue_to_score(pwmscan_ctx* ctx, const pwmscan_motifs*, double, double*) {
    return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_pvalue_to_score: not implemented yet");
}

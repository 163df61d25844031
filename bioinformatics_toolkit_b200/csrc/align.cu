// align.cu — all-pairs PWM alignment (Motif/Alignment.hs:83-132).  (stub: round-1 WIP)
#include "internal.h"
using namespace pwmscan;

extern "C" int pwmscan_align_all_pairs(pwmscan_ctx* ctx, const pwmscan_motifs*, int, double, int, double*, uint8_t*, int32_t*) {
    return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align_all_pairs: not implemented yet");
}
extern "C" int pwmscan_align_pairs(pwmscan_ctx* ctx, const pwmscan_motifs*, int64_t, const int32_t*, const int32_t*, int, double, int,
                                   double*, uint8_t*, int32_t*) {
    return set_err(ctx, PWMSCAN_EINVAL, "pwmscan_align_pairs: not implemented yet");
}

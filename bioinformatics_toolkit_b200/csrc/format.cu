// format.cu — host-side BED6 text of scanMotif hits: `toLine` of Data/Bed/Types.hs:139-149 applied to
// mkBed of Data/Bed/Utils.hs:109-111 (chrom, start+i, start+i+n, motif name, affinity score, +/-).
// Pure host code (the Haskell drop-in keeps this step in Haskell); multi-threaded over hit ranges.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "internal.h"

using namespace pwmscan;

namespace {
inline char* put_int(char* p, int64_t v) {
    char tmp[24];
    int n = 0;
    bool neg = v < 0;
    uint64_t u = neg ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v;
    do { tmp[n++] = (char)('0' + u % 10); u /= 10; } while (u);
    if (neg) *p++ = '-';
    while (n) *p++ = tmp[--n];
    return p;
}
}  // namespace

// chrom_names / motif_names: concatenated bytes with (count+1) offsets.  region_chrom[r] indexes chrom_names.
// Output: *text (malloc'ed, release with pwmscan_free) and *text_len; one line per hit, '\n' terminated.
extern "C" int pwmscan_format_bed6(int64_t nhits, const int32_t* segment, const int32_t* motif, const uint8_t* strand,
                                   const int64_t* pos, const int64_t* score, const int32_t* region_chrom,
                                   const int64_t* region_start, const char* chrom_names, const int64_t* chrom_off,
                                   const char* motif_names, const int64_t* motif_off, const int32_t* motif_len,
                                   char** text, int64_t* text_len) {
    if (nhits < 0 || !text || !text_len) return PWMSCAN_EINVAL;
    *text = nullptr; *text_len = 0;
    if (nhits == 0) return PWMSCAN_OK;
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), nhits / 65536 + 1));
    auto ndigits = [](int64_t v) { int n = v < 0 ? 2 : 1; uint64_t u = v < 0 ? (uint64_t)(-(v + 1)) + 1 : (uint64_t)v; while (u >= 10) { u /= 10; ++n; } return n; };
    auto line_len = [&](int64_t k) -> int64_t {           // exact
        const int32_t r = segment[k], m = motif[k];
        const int32_t c = region_chrom[r];
        const int64_t st = region_start[r] + pos[k];
        return (chrom_off[c + 1] - chrom_off[c]) + (motif_off[m + 1] - motif_off[m]) + ndigits(st) + ndigits(st + motif_len[m]) + ndigits(score[k]) + 7;
    };
    auto range = [&](int t, int64_t* a, int64_t* b) { *a = nhits * t / nthreads; *b = nhits * (t + 1) / nthreads; };
    std::vector<int64_t> start((size_t)nthreads + 1, 0);
    {   // pass 1: exact bytes per thread range, so that every thread can write straight into the one output buffer
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] { int64_t a, b, s = 0; range(t, &a, &b); for (int64_t k = a; k < b; ++k) s += line_len(k); start[t + 1] = s; });
        for (auto& x : th) x.join();
    }
    for (int t = 0; t < nthreads; ++t) start[t + 1] += start[t];
    const int64_t tot = start[nthreads];
    char* out = (char*)std::malloc((size_t)tot + 1);
    if (!out) return PWMSCAN_EINVAL;
    bool ok = true;
    {
        std::vector<std::thread> th;
        std::vector<char> good((size_t)nthreads, 1);
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] {
            int64_t a, b; range(t, &a, &b);
            char* p = out + start[t];
            for (int64_t k = a; k < b; ++k) {
                const int32_t r = segment[k], m = motif[k], c = region_chrom[r];
                const int64_t st = region_start[r] + pos[k];
                const int64_t cl = chrom_off[c + 1] - chrom_off[c], ml = motif_off[m + 1] - motif_off[m];
                std::memcpy(p, chrom_names + chrom_off[c], (size_t)cl); p += cl; *p++ = '\t';
                p = put_int(p, st); *p++ = '\t';
                p = put_int(p, st + motif_len[m]); *p++ = '\t';
                std::memcpy(p, motif_names + motif_off[m], (size_t)ml); p += ml; *p++ = '\t';
                p = put_int(p, score[k]); *p++ = '\t';
                *p++ = strand[k] == 0 ? '+' : '-';
                *p++ = '\n';
            }
            good[t] = (p == out + start[t + 1]);
        });
        for (auto& x : th) x.join();
        for (char g : good) ok = ok && g;
    }
    if (!ok) { std::free(out); return PWMSCAN_EINVAL; }
    out[tot] = 0;
    *text = out; *text_len = tot;
    return PWMSCAN_OK;
}

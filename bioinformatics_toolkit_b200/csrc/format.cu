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
    std::vector<int64_t> lens((size_t)nthreads + 1, 0);
    auto line_len = [&](int64_t k) -> int64_t {
        const int32_t r = segment[k], m = motif[k];
        const int32_t c = region_chrom[r];
        return (chrom_off[c + 1] - chrom_off[c]) + (motif_off[m + 1] - motif_off[m]) + 3 * 20 + 8;   // upper bound
    };
    auto range = [&](int t, int64_t* a, int64_t* b) { *a = nhits * t / nthreads; *b = nhits * (t + 1) / nthreads; };
    {   // upper bounds per thread
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] { int64_t a, b, s = 0; range(t, &a, &b); for (int64_t k = a; k < b; ++k) s += line_len(k); lens[t + 1] = s; });
        for (auto& x : th) x.join();
    }
    std::vector<char*> bufs(nthreads, nullptr);
    std::vector<int64_t> used(nthreads, 0);
    for (int t = 0; t < nthreads; ++t) { bufs[t] = (char*)std::malloc((size_t)lens[t + 1] + 1); if (!bufs[t]) { for (auto b : bufs) std::free(b); return PWMSCAN_EINVAL; } }
    {
        std::vector<std::thread> th;
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] {
            int64_t a, b; range(t, &a, &b);
            char* p = bufs[t];
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
            used[t] = p - bufs[t];
        });
        for (auto& x : th) x.join();
    }
    int64_t tot = 0;
    for (int t = 0; t < nthreads; ++t) tot += used[t];
    char* out = (char*)std::malloc((size_t)tot + 1);
    if (!out) { for (auto b : bufs) std::free(b); return PWMSCAN_EINVAL; }
    int64_t o = 0;
    for (int t = 0; t < nthreads; ++t) { std::memcpy(out + o, bufs[t], (size_t)used[t]); o += used[t]; std::free(bufs[t]); }
    *text = out; *text_len = tot;
    return PWMSCAN_OK;
}

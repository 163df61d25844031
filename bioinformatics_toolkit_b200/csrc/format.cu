// format.cu — host-side BED6 text of scanMotif hits: `toLine` of Data/Bed/Types.hs:139-149 applied to
// mkBed of Data/Bed/Utils.hs:109-111 (chrom, start+i, start+i+n, motif name, affinity score, +/-).
// Pure host code (the Haskell drop-in keeps this step in Haskell); multi-threaded over hit ranges.
#include <charconv>
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

// ---- `merge_motifs dist` text (app/MergeMotifs.hs:158-169): "name_a<TAB>name_b<TAB>toShortest d" per pair -------------
namespace {

// Data.Double.Conversion toShortest = the ECMAScript Number-to-string rules on the shortest round-trip digits:
// plain decimals for 1e-6 <= |x| < 1e21, otherwise d[.ddd]e[+-]n.
inline char* put_shortest(char* p, double x) {
    if (x != x) { std::memcpy(p, "NaN", 3); return p + 3; }
    if (x < 0) { *p++ = '-'; x = -x; }
    if (x == 0) { *p++ = '0'; return p; }
    if (x > 1.7976931348623157e308) { std::memcpy(p, "Infinity", 8); return p + 8; }
    char buf[40];
    const auto r = std::to_chars(buf, buf + sizeof(buf), x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest digits
    char digits[24];
    int k = 0;
    const char* q = buf;
    for (; q < r.ptr && *q != 'e'; ++q) if (*q != '.') digits[k++] = *q;
    int e10 = 0;
    {
        ++q;                                        // skip 'e'
        const bool neg = *q == '-';
        ++q;
        for (; q < r.ptr; ++q) e10 = e10 * 10 + (*q - '0');
        if (neg) e10 = -e10;
    }
    while (k > 1 && digits[k - 1] == '0') --k;
    const int n = e10 + 1;                          // value = 0.DIGITS x 10^n
    if (k <= n && n <= 21) {
        std::memcpy(p, digits, (size_t)k); p += k;
        for (int i = 0; i < n - k; ++i) *p++ = '0';
    } else if (0 < n && n <= 21) {
        std::memcpy(p, digits, (size_t)n); p += n; *p++ = '.';
        std::memcpy(p, digits + n, (size_t)(k - n)); p += k - n;
    } else if (-6 < n && n <= 0) {
        *p++ = '0'; *p++ = '.';
        for (int i = 0; i < -n; ++i) *p++ = '0';
        std::memcpy(p, digits, (size_t)k); p += k;
    } else {
        *p++ = digits[0];
        if (k > 1) { *p++ = '.'; std::memcpy(p, digits + 1, (size_t)(k - 1)); p += k - 1; }
        *p++ = 'e';
        const int e = n - 1;
        *p++ = e >= 0 ? '+' : '-';
        p = put_int(p, e >= 0 ? e : -e);
    }
    return p;
}

}  // namespace

// a == b == NULL: all i < j of the M names, packed upper triangle in row-major order (pwmscan_align_all_pairs' order);
// otherwise npairs explicit (a[k], b[k]).  names: concatenated bytes with M+1 offsets.
extern "C" int pwmscan_format_dist(int64_t npairs, int32_t M, const int32_t* a, const int32_t* b, const double* dist, const char* names,
                                   const int64_t* name_off, char** text, int64_t* text_len) {
    if (!text || !text_len || M < 0 || !name_off || (npairs > 0 && !dist) || ((a == nullptr) != (b == nullptr))) return PWMSCAN_EINVAL;
    *text = nullptr; *text_len = 0;
    const bool all = a == nullptr;
    if (all) npairs = (int64_t)M * (M - 1) / 2;
    if (npairs <= 0) return PWMSCAN_OK;
    const int nthreads = (int)std::max<int64_t>(1, std::min<int64_t>(std::thread::hardware_concurrency(), npairs / 65536 + 1));
    auto kstart = [&](int64_t i) { return i * M - i * (i + 1) / 2; };       // first pair of row i
    // work split: contiguous pair ranges; in the all-pairs mode they start at row boundaries
    std::vector<int64_t> cut((size_t)nthreads + 1, 0);
    std::vector<int64_t> row((size_t)nthreads + 1, 0);
    for (int t = 1; t < nthreads; ++t) {
        int64_t k = npairs * t / nthreads;
        if (all) {
            int64_t lo = 0, hi = M - 1;                                         // last row whose first pair is <= k
            while (hi - lo > 1) { const int64_t mid = (lo + hi) / 2; if (kstart(mid) <= k) lo = mid; else hi = mid; }
            row[t] = lo; k = kstart(lo);
        }
        cut[t] = k;
    }
    cut[nthreads] = npairs; row[nthreads] = all ? M - 1 : 0;
    auto nlen = [&](int32_t m) { return name_off[m + 1] - name_off[m]; };
    std::vector<char*> bufs((size_t)nthreads, nullptr);
    std::vector<int64_t> used((size_t)nthreads, 0);
    bool ok = true;
    {
        std::vector<std::thread> th;
        std::vector<char> good((size_t)nthreads, 1);
        for (int t = 0; t < nthreads; ++t) th.emplace_back([&, t] {
            int64_t cap = 0;                                                    // upper bound of this range's text
            if (all) { for (int64_t i = row[t]; i < (t + 1 < nthreads ? row[t + 1] : M - 1); ++i) { for (int64_t j = i + 1; j < M; ++j) cap += nlen((int32_t)i) + nlen((int32_t)j) + 30; } }
            else for (int64_t k = cut[t]; k < cut[t + 1]; ++k) cap += nlen(a[k]) + nlen(b[k]) + 30;
            char* buf = (char*)std::malloc((size_t)cap + 1);
            if (!buf) { good[t] = 0; return; }
            char* p = buf;
            auto line = [&](int32_t x, int32_t y, double d) {
                std::memcpy(p, names + name_off[x], (size_t)nlen(x)); p += nlen(x); *p++ = '\t';
                std::memcpy(p, names + name_off[y], (size_t)nlen(y)); p += nlen(y); *p++ = '\t';
                p = put_shortest(p, d); *p++ = '\n';
            };
            if (all) {
                int64_t k = cut[t];
                for (int64_t i = row[t]; i < (t + 1 < nthreads ? row[t + 1] : M - 1); ++i)
                    for (int64_t j = i + 1; j < M; ++j) line((int32_t)i, (int32_t)j, dist[k++]);
            } else {
                for (int64_t k = cut[t]; k < cut[t + 1]; ++k) line(a[k], b[k], dist[k]);
            }
            bufs[t] = buf; used[t] = p - buf;
        });
        for (auto& x : th) x.join();
        for (char g : good) ok = ok && g;
    }
    int64_t tot = 0;
    for (int t = 0; t < nthreads; ++t) tot += used[t];
    char* out = ok ? (char*)std::malloc((size_t)tot + 1) : nullptr;
    if (!out) { for (auto bfr : bufs) std::free(bfr); return PWMSCAN_EINVAL; }
    int64_t o = 0;
    for (int t = 0; t < nthreads; ++t) { std::memcpy(out + o, bufs[t], (size_t)used[t]); o += used[t]; std::free(bufs[t]); }
    out[tot] = 0;
    *text = out; *text_len = tot;
    return PWMSCAN_OK;
}

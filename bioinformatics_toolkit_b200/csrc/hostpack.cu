// hostpack.cu — the packed genome store: ASCII -> the 2-bit + invalid-base planes the device scans, on the HOST.
//
// A genome that is scanned more than once (every motif set, every p-value) does not have to cross PCIe as one byte per base
// every time: `mkIndex` (Seq/IO.hs:87-106) can keep, next to the sequence bytes, the two planes pack_kernel would make of
// them — 2 bits per base (16 bases per uint32, base t at bits 2t..2t+1; A C G T = 0 1 2 3, any other letter a position
// hash) and 1 bit per base "not A/C/G/T" — 0.375 bytes per base.  pwmscan_scan_packed_many / pwmscan_seq_upload_packed take
// slices of these planes (any start that is a multiple of 32 bases) straight into HBM; the device only fills in the padding
// around them.  Layout unit: 32 bases = two words of the 2-bit plane + one word of the mask plane (what one pack_kernel
// thread writes), so a sequence of `len` bases has units = ceil(len / 32), 2 * units + units words.
// The codes are exactly pack_kernel's (same table, same hash of the stored index), so a sequence uploaded packed from
// position 0 is bit-identical in HBM to the same sequence uploaded as ASCII.
#include <algorithm>
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "internal.h"
#include "scan_core.cuh"

using namespace pwmscan;

extern "C" int64_t pwmscan_packed_units(int64_t len) { return len < 0 ? -1 : (len + 31) / 32; }

// ascii[0, len) -> seq2_out[2 * units], mask_out[units].  *first_non_acgt / *first_non_iupac: fromBS's verdict (Seq.hs:86-95)
// for the `Basic` / `IUPAC` alphabets: offset of the first byte outside it, -1 if none (after upper-casing if
// PWMSCAN_SEQ_UPPERCASE).  nthreads <= 0: all host cores.
extern "C" int pwmscan_pack_host(const uint8_t* ascii, int64_t len, int flags, uint32_t* seq2_out, uint32_t* mask_out,
                                 int64_t* first_non_acgt, int64_t* first_non_iupac, int nthreads) {
    if (len < 0 || (len > 0 && (!ascii || !seq2_out || !mask_out))) return PWMSCAN_EINVAL;
    unsigned char lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = (unsigned char)iupac_code((unsigned char)i);
    const bool upper = (flags & PWMSCAN_SEQ_UPPERCASE) != 0;
    if (upper) for (int c = 'a'; c <= 'z'; ++c) lut[c] = lut[c - 32];
    const int64_t units = (len + 31) / 32;
    int nt = nthreads > 0 ? nthreads : (int)std::max(1u, std::thread::hardware_concurrency());
    nt = (int)std::max<int64_t>(1, std::min<int64_t>(nt, units / 32768 + 1));
    std::vector<int64_t> bad_a((size_t)nt, -1), bad_i((size_t)nt, -1);
    auto work = [&](int w) {
        const int64_t t0 = units * w / nt, t1 = units * (w + 1) / nt;
        int64_t ba = -1, bi = -1;
        const uint32_t fold = upper ? 0xDFDFDFDFu : 0xFFFFFFFFu;
        // four ASCII bytes -> their 2-bit codes in 8 bits + "not one of ACGT" flags, without a table (pack_kernel's pack4)
        auto pack4 = [fold](uint32_t w, uint32_t* bad) -> uint32_t {
            const uint32_t v = w & fold;
            const uint32_t x = (v >> 1) & 0x03030303u;
            const uint32_t m = (x >> 1) & ~x & 0x01010101u;
            const uint32_t expect = 0x41414141u + (x << 1) + (m << 4) - m;
            const uint32_t d = v ^ expect;
            *bad |= (((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) & 0x80808080u;
            const uint32_t code = x ^ ((x >> 1) & 0x01010101u);
            return (code * 0x01041040u) >> 24;
        };
        for (int64_t t = t0; t < t1; ++t) {
            const int64_t p0 = t * 32;
            if (p0 + 32 <= len) {                                  // fast path: 32 bytes that are all A/C/G/T
                uint32_t w[8], bad = 0, c[8];
                std::memcpy(w, ascii + p0, 32);
                for (int q = 0; q < 8; ++q) c[q] = pack4(w[q], &bad);
                if (bad == 0u) {
                    seq2_out[2 * t] = c[0] | (c[1] << 8) | (c[2] << 16) | (c[3] << 24);
                    seq2_out[2 * t + 1] = c[4] | (c[5] << 8) | (c[6] << 16) | (c[7] << 24);
                    mask_out[t] = 0u;
                    continue;
                }
            }
            uint32_t w0 = 0, w1 = 0, mk = 0;
            const int lim = (int)std::min<int64_t>(32, len - p0);
            for (int k = 0; k < 32; ++k) {
                const unsigned code = k < lim ? lut[ascii[p0 + k]] : 15u;
                uint32_t c2;
                if (code < 4) c2 = code;
                else {
                    c2 = pad_hash(p0 + k + kPadFront);           // the stored index pack_kernel hashes
                    mk |= 1u << k;
                    if (k < lim) {
                        if (ba < 0) ba = p0 + k;
                        if (code == 15 && bi < 0) bi = p0 + k;
                    }
                }
                if (k < 16) w0 |= c2 << (2 * k); else w1 |= c2 << (2 * (k - 16));
            }
            seq2_out[2 * t] = w0; seq2_out[2 * t + 1] = w1; mask_out[t] = mk;
        }
        bad_a[(size_t)w] = ba; bad_i[(size_t)w] = bi;
    };
    if (nt == 1) work(0);
    else {
        std::vector<std::thread> th;
        for (int w = 0; w < nt; ++w) th.emplace_back(work, w);
        for (auto& t : th) t.join();
    }
    int64_t ba = -1, bi = -1;
    for (int w = 0; w < nt; ++w) { if (ba < 0) ba = bad_a[(size_t)w]; if (bi < 0) bi = bad_i[(size_t)w]; }
    if (first_non_acgt) *first_non_acgt = ba;
    if (first_non_iupac) *first_non_iupac = bi;
    return PWMSCAN_OK;
}

// inflate_fast.hpp -- raw DEFLATE (RFC 1951) decoder for BGZF blocks.
//
// BGZF members are at most 64 KiB of output each and arrive whole, with the uncompressed size known from the
// gzip trailer, so the decoder can be a single pass over two flat buffers: a 64-bit bit buffer refilled with one
// unaligned 8-byte load, an 11-bit first-level literal/length table (8-bit for distances) with second-level
// tables for longer codes, table entries that already carry base value and extra-bit count, and word-wise match
// copies into an output buffer that has slack behind it.  About 2x the throughput of zlib's inflate() on BAM data,
// which is what bounds the staging host once every block is inflated only once.
//
// vtx_inflate_raw returns true iff `in` is a complete, valid DEFLATE stream that produces exactly `out_len` bytes.
// Requirements: `in` must be readable for kInflateInPad bytes beyond in_len (any content) and `out` writable for
// kInflateOutPad bytes beyond out_len (they may be clobbered).
#pragma once
#include <cstdint>
#include <cstring>

namespace vtxhost {

constexpr size_t kInflateInPad = 16, kInflateOutPad = 64;

namespace inflate_detail {

constexpr int kLitBits = 11, kDistBits = 8;
constexpr uint32_t kTypeLiteral = 0u << 30, kTypeBase = 1u << 30, kTypeEnd = 2u << 30, kTypeSub = 3u << 30, kTypeMask = 3u << 30;
// entry: [31:30] type  [29:13] payload (literal / base value / subtable offset)  [12:8] extra bits (or subtable bits)  [7:0] code length
inline uint32_t make_entry(uint32_t type, uint32_t payload, uint32_t extra, uint32_t len) { return type | (payload << 13) | (extra << 8) | len; }

struct Tables {
    uint32_t lit[(1 << kLitBits) + 1024];      // worst case of second-level entries for 288 symbols of <= 15 bits
    uint32_t dist[(1 << kDistBits) + 512];
};

inline uint32_t reverse_bits(uint32_t v, int n)
{
    uint32_t r = 0;
    for (int i = 0; i < n; ++i) { r = (r << 1) | (v & 1); v >>= 1; }
    return r;
}

// Canonical Huffman code lengths -> two-level lookup table indexed by the next bits of the stream (LSB first).
// sym_entry(s) gives the entry of symbol s without its length field.  Returns false for over-subscribed codes and
// for incomplete ones (except the single-code case zlib also accepts).
template <class EntryOf>
inline bool build_table(const uint8_t* lens, int n_sym, int main_bits, uint32_t* tab, size_t tab_cap, EntryOf sym_entry)
{
    int count[16] = { 0 };
    for (int s = 0; s < n_sym; ++s) count[lens[s]]++;
    count[0] = 0;
    int max_len = 15;
    while (max_len > 0 && count[max_len] == 0) --max_len;
    const uint32_t main_size = 1u << main_bits;
    if (max_len == 0) {                          // no codes at all: every lookup is an error
        for (uint32_t i = 0; i < main_size; ++i) tab[i] = 0;
        return true;
    }
    uint32_t code = 0, next_code[16];
    int64_t left = 1;
    for (int l = 1; l <= 15; ++l) {
        left <<= 1; left -= count[l];
        if (left < 0) return false;              // over-subscribed
        code = (code + uint32_t(count[l - 1])) << 1;
        next_code[l] = code;
    }
    int n_codes = 0;
    for (int l = 1; l <= 15; ++l) n_codes += count[l];
    if (left > 0 && n_codes != 1) return false;  // incomplete
    for (uint32_t i = 0; i < main_size; ++i) tab[i] = 0;
    // second-level tables: one per main-table prefix that has longer codes, sized by the longest code under it
    uint8_t sub_bits[1 << kLitBits];
    if (max_len > main_bits) {
        memset(sub_bits, 0, main_size);
        uint32_t nc[16];
        memcpy(nc, next_code, sizeof(nc));
        for (int s = 0; s < n_sym; ++s) {
            const int l = lens[s];
            if (l == 0) continue;
            const uint32_t c = nc[l]++;
            if (l > main_bits) {
                const uint32_t prefix = reverse_bits(c >> (l - main_bits), main_bits);
                if (uint8_t(l - main_bits) > sub_bits[prefix]) sub_bits[prefix] = uint8_t(l - main_bits);
            }
        }
        size_t off = main_size;
        for (uint32_t p = 0; p < main_size; ++p)
            if (sub_bits[p]) {
                if (off + (size_t(1) << sub_bits[p]) > tab_cap) return false;
                tab[p] = make_entry(kTypeSub, uint32_t(off), sub_bits[p], uint32_t(main_bits));
                for (size_t i = 0; i < (size_t(1) << sub_bits[p]); ++i) tab[off + i] = 0;
                off += size_t(1) << sub_bits[p];
            }
    }
    for (int s = 0; s < n_sym; ++s) {
        const int l = lens[s];
        if (l == 0) continue;
        const uint32_t c = next_code[l]++;
        const uint32_t rev = reverse_bits(c, l);
        const uint32_t e = sym_entry(s) | uint32_t(l);
        if (l <= main_bits) {
            for (uint32_t i = rev; i < main_size; i += 1u << l) tab[i] = e;
        } else {
            const uint32_t prefix = rev & (main_size - 1);
            const uint32_t sub = tab[prefix];
            const uint32_t off = (sub >> 13) & 0x1FFFF, sb = (sub >> 8) & 31;
            for (uint32_t i = rev >> main_bits; i < (1u << sb); i += 1u << (l - main_bits)) tab[off + i] = e;
        }
    }
    return true;
}

inline uint32_t litlen_entry(int s)
{
    static const uint16_t base[29] = { 3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258 };
    static const uint8_t extra[29] = { 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0 };
    if (s < 256) return make_entry(kTypeLiteral, uint32_t(s), 0, 0);
    if (s == 256) return make_entry(kTypeEnd, 0, 0, 0);
    if (s > 285) return make_entry(kTypeEnd, 1, 0, 0);        // 286, 287: invalid in data (payload 1 marks the error)
    return make_entry(kTypeBase, base[s - 257], extra[s - 257], 0);
}
inline uint32_t dist_entry(int s)
{
    static const uint16_t base[30] = { 1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073,
                                       4097, 6145, 8193, 12289, 16385, 24577 };
    static const uint8_t extra[30] = { 0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13 };
    if (s > 29) return make_entry(kTypeEnd, 1, 0, 0);          // 30, 31: invalid
    return make_entry(kTypeBase, base[s], extra[s], 0);
}

}  // namespace inflate_detail

inline bool vtx_inflate_raw(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len)
{
    using namespace inflate_detail;
    static thread_local Tables T;
    static thread_local Tables fixedT;
    static thread_local bool fixed_ready = false;

    uint64_t bitbuf = 0;
    int bitcnt = 0;
    size_t ip = 0, op = 0;
    // refill to >= 56 valid bits with one unaligned load; reading past in_len is allowed (padding), *using* those bits is
    // detected at the end through ip
    bool bad = false;
    auto refill = [&]() {
        if (ip + 8 > in_len + kInflateInPad) { bad = true; return; }      // would read beyond the padding: corrupt input
        uint64_t w;
        memcpy(&w, in + ip, 8);
        bitbuf |= w << bitcnt;
        const int take = (63 - bitcnt) >> 3;
        ip += size_t(take);
        bitcnt += take << 3;
    };
    auto overrun = [&]() { return ip > in_len + 8 || (ip > in_len && size_t(bitcnt) < (ip - in_len) * 8); };

    for (;;) {
        refill();
        if (bad) return false;
        const uint32_t final_block = uint32_t(bitbuf & 1), type = uint32_t((bitbuf >> 1) & 3);
        bitbuf >>= 3; bitcnt -= 3;
        if (type == 0) {                                         // stored
            const int drop = bitcnt & 7;
            bitbuf >>= drop; bitcnt -= drop;
            // give whole bytes of the bit buffer back to the input
            ip -= size_t(bitcnt >> 3); bitbuf = 0; bitcnt = 0;
            if (ip + 4 > in_len) return false;
            const uint32_t len = uint32_t(in[ip]) | (uint32_t(in[ip + 1]) << 8), nlen = uint32_t(in[ip + 2]) | (uint32_t(in[ip + 3]) << 8);
            ip += 4;
            if ((len ^ 0xFFFF) != nlen || ip + len > in_len || op + len > out_len) return false;
            memcpy(out + op, in + ip, len);
            ip += len; op += len;
        } else if (type == 1 || type == 2) {
            const Tables* tb;
            if (type == 1) {
                if (!fixed_ready) {
                    uint8_t ll[288], dl[32];
                    for (int i = 0; i < 144; ++i) ll[i] = 8;
                    for (int i = 144; i < 256; ++i) ll[i] = 9;
                    for (int i = 256; i < 280; ++i) ll[i] = 7;
                    for (int i = 280; i < 288; ++i) ll[i] = 8;
                    for (int i = 0; i < 32; ++i) dl[i] = 5;
                    build_table(ll, 288, kLitBits, fixedT.lit, sizeof(fixedT.lit) / 4, litlen_entry);
                    build_table(dl, 32, kDistBits, fixedT.dist, sizeof(fixedT.dist) / 4, dist_entry);
                    fixed_ready = true;
                }
                tb = &fixedT;
            } else {
                const uint32_t hlit = uint32_t(bitbuf & 31) + 257, hdist = uint32_t((bitbuf >> 5) & 31) + 1, hclen = uint32_t((bitbuf >> 10) & 15) + 4;
                bitbuf >>= 14; bitcnt -= 14;
                if (hlit > 286 || hdist > 30) return false;
                static const uint8_t order[19] = { 16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15 };
                uint8_t pl[19] = { 0 };
                for (uint32_t i = 0; i < hclen; ++i) {
                    if (bitcnt < 3) { refill(); if (bad) return false; }
                    pl[order[i]] = uint8_t(bitbuf & 7);
                    bitbuf >>= 3; bitcnt -= 3;
                }
                uint32_t pre[1 << 7];
                if (!build_table(pl, 19, 7, pre, 1 << 7, [](int s) { return make_entry(kTypeLiteral, uint32_t(s), 0, 0); })) return false;
                uint8_t lens[288 + 32 + 140];
                uint32_t n = 0;
                while (n < hlit + hdist) {
                    refill();
                    if (bad) return false;
                    const uint32_t e = pre[bitbuf & 127];
                    const uint32_t l = e & 0xFF;
                    if (l == 0) return false;
                    bitbuf >>= l; bitcnt -= int(l);
                    const uint32_t sym = (e >> 13) & 0x1FFFF;
                    if (sym < 16) { lens[n++] = uint8_t(sym); continue; }
                    uint32_t rep, val = 0;
                    if (sym == 16) { if (n == 0) return false; val = lens[n - 1]; rep = 3 + uint32_t(bitbuf & 3); bitbuf >>= 2; bitcnt -= 2; }
                    else if (sym == 17) { rep = 3 + uint32_t(bitbuf & 7); bitbuf >>= 3; bitcnt -= 3; }
                    else { rep = 11 + uint32_t(bitbuf & 127); bitbuf >>= 7; bitcnt -= 7; }
                    if (n + rep > hlit + hdist) return false;
                    memset(lens + n, int(val), rep);
                    n += rep;
                }
                if (lens[256] == 0) return false;                // no end-of-block code
                uint8_t ll[288], dl[32];
                memset(ll, 0, sizeof(ll)); memset(dl, 0, sizeof(dl));
                memcpy(ll, lens, hlit); memcpy(dl, lens + hlit, hdist);
                if (!build_table(ll, 288, kLitBits, T.lit, sizeof(T.lit) / 4, litlen_entry)) return false;
                if (!build_table(dl, 32, kDistBits, T.dist, sizeof(T.dist) / 4, dist_entry)) return false;
                tb = &T;
            }
            const uint32_t* lit = tb->lit;
            const uint32_t* dst = tb->dist;
            for (;;) {
                refill();                                        // >= 56 bits: one length/distance pair (<= 48) or three literals
                if (bad) return false;
                uint32_t e = lit[bitbuf & ((1u << kLitBits) - 1)];
                if ((e & kTypeMask) == kTypeSub) e = lit[((e >> 13) & 0x1FFFF) + ((bitbuf >> kLitBits) & ((1u << ((e >> 8) & 31)) - 1))];
                uint32_t l = e & 0xFF;
                if (l == 0) return false;
                bitbuf >>= l; bitcnt -= int(l);
                if ((e & kTypeMask) == kTypeLiteral) {
                    if (op >= out_len) return false;
                    out[op++] = uint8_t(e >> 13);
                    // up to two more literals from the same refill (3 x 15 bits <= 56)
                    for (int k = 0; k < 2; ++k) {
                        uint32_t e2 = lit[bitbuf & ((1u << kLitBits) - 1)];
                        if ((e2 & kTypeMask) != kTypeLiteral) break;
                        const uint32_t l2 = e2 & 0xFF;
                        if (l2 == 0 || op >= out_len) break;
                        bitbuf >>= l2; bitcnt -= int(l2);
                        out[op++] = uint8_t(e2 >> 13);
                    }
                    continue;
                }
                if ((e & kTypeMask) == kTypeEnd) {
                    if ((e >> 13) & 0x1FFFF) return false;       // symbols 286 / 287
                    break;
                }
                const uint32_t xb = (e >> 8) & 31;
                const uint32_t length = ((e >> 13) & 0x1FFFF) + uint32_t(bitbuf & ((1u << xb) - 1));
                bitbuf >>= xb; bitcnt -= int(xb);
                uint32_t d = dst[bitbuf & ((1u << kDistBits) - 1)];
                if ((d & kTypeMask) == kTypeSub) d = dst[((d >> 13) & 0x1FFFF) + ((bitbuf >> kDistBits) & ((1u << ((d >> 8) & 31)) - 1))];
                l = d & 0xFF;
                if (l == 0 || (d & kTypeMask) != kTypeBase) return false;
                bitbuf >>= l; bitcnt -= int(l);
                const uint32_t db = (d >> 8) & 31;
                const uint32_t distance = ((d >> 13) & 0x1FFFF) + uint32_t(bitbuf & ((1u << db) - 1));
                bitbuf >>= db; bitcnt -= int(db);
                if (distance > op || op + length > out_len) return false;
                uint8_t* o = out + op;
                const uint8_t* s = o - distance;
                if (distance >= 8) {                             // word-wise; may write up to 7 bytes past the match (slack)
                    for (uint32_t i = 0; i < length; i += 8) { uint64_t w; memcpy(&w, s + i, 8); memcpy(o + i, &w, 8); }
                } else if (distance == 1) {
                    memset(o, s[0], length);
                } else {
                    for (uint32_t i = 0; i < length; ++i) o[i] = s[i];
                }
                op += length;
            }
        } else {
            return false;
        }
        if (overrun()) return false;
        if (final_block) break;
    }
    // bits consumed must lie inside the input
    return op == out_len && !overrun();
}

}  // namespace vtxhost

// Memory-safety check of the device staging kernels' bodies (vartrix_b200/csrc/vtx_stage.cuh) and of the device DEFLATE
// decoder's bit-stream half (vtx_inflate.cuh) on DAMAGED input, built with -fsanitize=address,undefined and run by
// tests/test_host_staging_cpu.py: a shard dumped by `vartrix_b200 --gpu-stage --dump-staged` is inflated, then bytes of the
// inflated stream (record headers, CIGARs, aux fields) and of the compressed members are flipped at random, and the same call
// sequence as vtx_submit_bam runs over it.  Whatever the damage, nothing may be read or written outside the buffers the
// engine would have allocated (the stream plus its padding, arrays sized by the counts of the first pass).
//   usage: stage_dev_fuzz DUMP N_ROUNDS SEED
#include <zlib.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>

#include "../vartrix_b200/csrc/vtx_inflate.cuh"
#include "../vartrix_b200/csrc/vtx_stage.cuh"

using namespace vtx::stage;

struct Member { uint64_t in_off; uint32_t in_len, out_len; uint64_t out_off; uint32_t crc, pad; };
struct Shard {
    int64_t tid = 0;
    std::vector<uint32_t> row; std::vector<int64_t> start, end; std::vector<Member> members; std::vector<uint8_t> comp; std::vector<uint64_t> entry;
};

static bool read_dump(const char* path, std::vector<Shard>* out)
{
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    std::vector<uint8_t> d; uint8_t buf[1 << 16]; size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) d.insert(d.end(), buf, buf + n);
    fclose(f);
    size_t p = 16;
    auto take = [&](std::vector<uint8_t>* v) { uint64_t nb; memcpy(&nb, d.data() + p, 8); p += 8; v->assign(d.begin() + long(p), d.begin() + long(p + nb)); p += nb; };
    while (p + 4 <= d.size()) {
        if (memcmp(d.data() + p, "VTXD", 4) != 0) return false;
        p += 4;
        Shard s; std::vector<uint8_t> b;
        take(&b); memcpy(&s.tid, b.data(), 8);
        take(&b); s.row.resize(b.size() / 4); memcpy(s.row.data(), b.data(), b.size());
        take(&b); s.start.resize(b.size() / 8); memcpy(s.start.data(), b.data(), b.size());
        take(&b); s.end.resize(b.size() / 8); memcpy(s.end.data(), b.data(), b.size());
        take(&b); s.members.resize(b.size() / sizeof(Member)); memcpy(s.members.data(), b.data(), b.size());
        take(&s.comp);
        take(&b); s.entry.resize(b.size() / 8); memcpy(s.entry.data(), b.data(), b.size());
        for (int k = 0; k < 5; ++k) take(&b);                       // windows: not needed here
        out->push_back(std::move(s));
    }
    return true;
}

// the call sequence of vtx_submit_bam over a (possibly damaged) stream; arrays sized like the engine sizes them
static int run_stage(const std::vector<uint8_t>& stream_padded, uint64_t s_len, const Shard& sh, int want_umi)
{
    Params P{};
    P.s = stream_padded.data(); P.s_len = s_len; P.tid = int32_t(sh.tid); P.mapq_min = 0; P.primary_only = 0; P.no_duplicates = 0; P.want_umi = want_umi;
    P.tag0 = 'C'; P.tag1 = 'B';
    const uint32_t n_seg = sh.entry.size() >= 2 ? uint32_t(sh.entry.size() - 1) : 0, nl = uint32_t(sh.row.size());
    uint32_t err = 0, max_span = 0, max_read = 0;
    std::vector<uint32_t> seg_count(n_seg + 1, 0), seg_first(n_seg + 2, 0);
    DirectFetch F{ P.s };
    for (uint32_t k = 0; k < n_seg; ++k) walk_segment(P, k, sh.entry.data(), 0, seg_count.data(), nullptr, nullptr, &err, F);
    if (err & (kErrWalk | kErrRecord)) return 1;                    // the engine stops here
    for (uint32_t k = 0; k < n_seg; ++k) seg_first[k + 1] = seg_first[k] + seg_count[k];
    const uint32_t n_rec = seg_first[n_seg];
    std::vector<uint64_t> rec_off(n_rec + 1);
    std::vector<int32_t> rec_tid(n_rec + 1), rec_pos(n_rec + 1), rec_end(n_rec + 1);
    std::vector<uint32_t> rec_fm(n_rec + 1), used(n_rec + 1, 0);
    for (uint32_t k = 0; k < n_seg; ++k) walk_segment(P, k, sh.entry.data(), 1, nullptr, seg_first.data(), rec_off.data(), &err, F);
    for (uint32_t i = 0; i < n_rec; ++i) parse_record(P, i, rec_off.data(), rec_tid.data(), rec_pos.data(), rec_end.data(), rec_fm.data(), &max_span);
    std::vector<uint32_t> cand_count(nl + 1, 0), cand_first(nl + 2, 0);
    LocusMetrics met{};
    for (uint32_t l = 0; l < nl; ++l)
        locus_cands(P, l, sh.start.data(), sh.end.data(), n_rec, rec_off.data(), rec_tid.data(), rec_pos.data(), rec_end.data(), rec_fm.data(), &max_span, &max_read,
                    0, cand_count.data(), nullptr, nullptr, nullptr, &met);
    for (uint32_t l = 0; l < nl; ++l) cand_first[l + 1] = cand_first[l] + cand_count[l];
    const size_t n_cand = cand_first[nl];
    std::vector<uint32_t> cand_rec(n_cand + 1);
    for (uint32_t l = 0; l < nl; ++l)
        locus_cands(P, l, sh.start.data(), sh.end.data(), n_rec, rec_off.data(), rec_tid.data(), rec_pos.data(), rec_end.data(), rec_fm.data(), &max_span, &max_read,
                    1, nullptr, cand_first.data(), cand_rec.data(), used.data(), &met);
    std::vector<uint64_t> read_off(n_rec + 1), read_umi(n_rec + 1);
    std::vector<uint32_t> read_len(n_rec + 1), read_cb_off(n_rec + 1);
    std::vector<uint16_t> read_cb_len(n_rec + 1);
    for (uint32_t i = 0; i < n_rec; ++i)
        read_emit(P, i, rec_off.data(), used.data(), read_off.data(), read_len.data(), read_cb_off.data(), read_cb_len.data(), read_umi.data(), &err);
    // what the pipeline would dereference next: bases and tag bytes of the used records must lie inside the stream
    for (uint32_t i = 0; i < n_rec; ++i) {
        if (!used[i]) continue;
        if (read_off[i] + (read_len[i] + 1) / 2 > s_len) return 100;
        if (read_cb_off[i] != kNoCb && uint64_t(read_cb_off[i]) + read_cb_len[i] > s_len) return 101;
    }
    return 0;
}

int main(int argc, char** argv)
{
    if (argc < 4) { fprintf(stderr, "usage: %s DUMP ROUNDS SEED\n", argv[0]); return 2; }
    std::vector<Shard> shards;
    if (!read_dump(argv[1], &shards) || shards.empty()) { fprintf(stderr, "cannot read %s\n", argv[1]); return 2; }
    const int rounds = atoi(argv[2]);
    std::mt19937_64 rng(strtoull(argv[3], nullptr, 10));
    long refused = 0, taken = 0, inflate_refused = 0, inflate_ok = 0;
    for (const Shard& sh : shards) {
        uint64_t s_len = 0;
        for (const Member& m : sh.members) s_len += m.out_len;
        std::vector<uint8_t> stream(s_len + 4096 + 64, 0);                      // the engine's padding behind the stream
        for (const Member& m : sh.members) {
            uLongf dl = m.out_len;
            z_stream zs{}; inflateInit2(&zs, -15);
            zs.next_in = const_cast<Bytef*>(sh.comp.data() + m.in_off); zs.avail_in = m.in_len; zs.next_out = stream.data() + m.out_off; zs.avail_out = uInt(dl);
            if (inflate(&zs, Z_FINISH) != Z_STREAM_END) { fprintf(stderr, "dump does not inflate\n"); return 2; }
            inflateEnd(&zs);
        }
        if (run_stage(stream, s_len, sh, 1) != 0) { fprintf(stderr, "the intact shard is refused\n"); return 3; }
        for (int r = 0; r < rounds && s_len; ++r) {
            std::vector<uint8_t> bad = stream;
            const int flips = 1 + int(rng() % 6);
            for (int k = 0; k < flips; ++k) {
                // aim at record heads half of the time: block_size, l_read_name, n_cigar, l_seq are what the walk trusts least
                uint64_t at = rng() % s_len;
                if ((rng() & 1) && !sh.entry.empty()) at = std::min<uint64_t>(s_len - 1, sh.entry[rng() % sh.entry.size()] + rng() % 40);
                bad[at] = uint8_t(rng());
            }
            const int rc = run_stage(bad, s_len, sh, int(rng() & 1));
            if (rc >= 100) { fprintf(stderr, "out-of-stream reference after damage (code %d)\n", rc); return 4; }
            (rc ? refused : taken) += 1;
        }
        // the device decoder's bit-stream logic on damaged members (what lane 0 runs)
        for (int r = 0; r < rounds && !sh.members.empty(); ++r) {
            const Member& m = sh.members[rng() % sh.members.size()];
            std::vector<uint32_t> in((m.in_len + 3) / 4 + 4, 0);
            memcpy(in.data(), sh.comp.data() + m.in_off, m.in_len);
            uint8_t* ib = reinterpret_cast<uint8_t*>(in.data());
            const int flips = 1 + int(rng() % 4);
            for (int k = 0; k < flips && m.in_len; ++k) ib[rng() % m.in_len] ^= uint8_t(1u << (rng() % 8));
            using namespace vtx::inflate;
            State st; state_init(st, ib, m.in_len, m.out_len);
            static Tables T; uint8_t lens[512]; Sym batch[32];
            std::vector<uint8_t> out(m.out_len + 1);
            size_t op = 0; bool ok = true; long guard = 0;
            while (st.status == kOk && st.phase != 3 && ok) {
                const int n = decode_batch(st, T, lens, batch, 31);
                for (int k = 0; k < n && ok; ++k) {
                    const Sym& sy = batch[k];
                    if (op + sy.len > m.out_len) { ok = false; break; }                    // decode_batch promises this never happens
                    if (sy.kind == 0) out[op] = uint8_t(sy.arg);
                    else if (sy.kind == 1) { if (sy.arg > op) { ok = false; break; } for (uint32_t i = 0; i < sy.len; ++i) out[op + i] = out[op - sy.arg + (sy.arg >= sy.len ? i : i % sy.arg)]; }
                    else { if (uint64_t(sy.arg) + sy.len > m.in_len) { ok = false; break; } memcpy(out.data() + op, ib + sy.arg, sy.len); }
                    op += sy.len;
                }
                if (++guard > 200000) { fprintf(stderr, "decoder does not terminate\n"); return 5; }
            }
            if (!ok) { fprintf(stderr, "decoder emitted a symbol outside its buffers\n"); return 6; }
            (st.status == kOk ? inflate_ok : inflate_refused) += 1;
        }
    }
    printf("shards %zu: damaged streams refused %ld, taken %ld; damaged members refused %ld, decoded %ld\n", shards.size(), refused, taken, inflate_refused, inflate_ok);
    return 0;
}

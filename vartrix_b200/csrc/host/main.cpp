// vartrix_b200 CLI -- keeps the `vartrix` flag surface (/root/reference/src/main.rs:40-135) and the
// `_main` flow (main.rs:163-418): inputs are decoded and filtered on host threads (stager.hpp), every shard
// of loci is handed to the GPU engine through the C ABI (include/vartrix_b200.h), outputs are the same
// Matrix-Market / label files.  `--dump-staged` stops after staging (no GPU needed; used by the tests).
#include <atomic>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <malloc.h>
#include <unistd.h>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <sys/stat.h>
#include <chrono>
#include <future>
#include <thread>

#include "stager.hpp"

using namespace vtxhost;

namespace {

int g_log = 0;      // 0 error, 1 info, 2 debug   (--log-level, main.rs:102-106)
void logf(int level, const char* tag, const char* fmt, ...)
{
    if (level > g_log) return;
    va_list ap; va_start(ap, fmt);
    fprintf(stderr, "[%s] ", tag); vfprintf(stderr, fmt, ap); fputc('\n', stderr);
    va_end(ap);
}
#define LOG_ERR(...) logf(0, "ERROR", __VA_ARGS__)
#define LOG_INFO(...) logf(1, "INFO", __VA_ARGS__)

double now_s()
{
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}

bool exists(const std::string& p) { struct stat st; return stat(p.c_str(), &st) == 0; }

struct Opts {
    std::string vcf, bam, fasta, barcodes, out_matrix = "out_matrix.mtx", ref_matrix = "ref_matrix.mtx", out_variants, out_barcodes;
    std::string scoring = "consensus", bam_tag = "CB", valid_chars = "ATGCatgc", dump_staged;
    long padding = 100, threads = 1, mapq = 0, device = 0, shard_loci = 0;      // 0: chosen from the number of loci and threads
    bool primary = false, no_dups = false, umi = false, ref_matrix_given = false;
};

void usage()
{
    puts("vartrix_b200 -- Variant assignment for single cell genomics (B200-native engine)\n"
         "USAGE: vartrix_b200 --vcf FILE --bam FILE --fasta FILE --cell-barcodes FILE [OPTIONS]\n"
         "  -v, --vcf FILE              Called variant file (VCF)\n"
         "  -b, --bam FILE              Cellranger BAM file\n"
         "  -f, --fasta FILE            Genome fasta file\n"
         "  -c, --cell-barcodes FILE    File with cell barcodes to be evaluated\n"
         "  -o, --out-matrix FILE       Output Matrix Market file [out_matrix.mtx]\n"
         "      --out-variants FILE     Output variant file\n"
         "      --out-barcodes FILE     Output cell barcode file\n"
         "  -p, --padding INT           Padding on both sides of the variant [100]\n"
         "  -s, --scoring-method M      consensus | coverage | alt_frac [consensus]\n"
         "      --ref-matrix FILE       Reference matrix (coverage mode) [ref_matrix.mtx]\n"
         "      --log-level L           info | debug | error [error]\n"
         "      --threads INT           Host staging threads [1]\n"
         "      --mapq INT              Minimum mapping quality [0]\n"
         "      --primary-alignments    Use primary alignments only\n"
         "      --no-duplicates         Do not consider duplicate alignments\n"
         "      --umi                   Consider UMI information\n"
         "      --bam-tag TAG           BAM tag marking cells [CB]\n"
         "      --valid-chars CHARS     Valid characters in an alternative haplotype [ATGCatgc]\n"
         "      --device INT            CUDA device ordinal [0]\n"
         "      --shard-loci INT        VCF records per staged shard [up to 2048, fewer for short VCFs]\n"
         "      --dump-staged FILE      Stage only, write the shards to FILE (no GPU)");
}

bool parse(int argc, char** argv, Opts* o)
{
    auto need = [&](int& i) -> const char* { if (i + 1 >= argc) { fprintf(stderr, "error: %s needs a value\n", argv[i]); exit(1); } return argv[++i]; };
    for (int i = 1; i < argc; ++i) {
        std::string a = argv[i];
        std::string val; bool has_eq = false;
        if (a.rfind("--", 0) == 0) { size_t e = a.find('='); if (e != std::string::npos) { val = a.substr(e + 1); a = a.substr(0, e); has_eq = true; } }
        auto v = [&]() -> std::string { return has_eq ? val : std::string(need(i)); };
        if (a == "-v" || a == "--vcf") o->vcf = v();
        else if (a == "-b" || a == "--bam") o->bam = v();
        else if (a == "-f" || a == "--fasta") o->fasta = v();
        else if (a == "-c" || a == "--cell-barcodes") o->barcodes = v();
        else if (a == "-o" || a == "--out-matrix") o->out_matrix = v();
        else if (a == "--out-variants") o->out_variants = v();
        else if (a == "--out-barcodes") o->out_barcodes = v();
        else if (a == "-p" || a == "--padding") o->padding = atol(v().c_str());
        else if (a == "-s" || a == "--scoring-method") o->scoring = v();
        else if (a == "--ref-matrix") { o->ref_matrix = v(); o->ref_matrix_given = true; }
        else if (a == "--log-level") { std::string l = v(); if (l == "info") g_log = 1; else if (l == "debug") g_log = 2; else if (l == "error") g_log = 0; else { puts("Log level not valid"); exit(1); } }
        else if (a == "--threads") o->threads = atol(v().c_str());
        else if (a == "--mapq") o->mapq = atol(v().c_str());
        else if (a == "--primary-alignments") o->primary = true;
        else if (a == "--no-duplicates") o->no_dups = true;
        else if (a == "--umi") o->umi = true;
        else if (a == "--bam-tag") o->bam_tag = v();
        else if (a == "--valid-chars") o->valid_chars = v();
        else if (a == "--device") o->device = atol(v().c_str());
        else if (a == "--shard-loci") o->shard_loci = atol(v().c_str());
        else if (a == "--dump-staged") o->dump_staged = v();
        else if (a == "-h" || a == "--help") { usage(); exit(0); }
        else if (a == "-V" || a == "--version") { puts("vartrix_b200 0.1 (vartrix 1.1.22 surface)"); exit(0); }
        else { fprintf(stderr, "error: unknown argument %s\n", argv[i]); return false; }
    }
    if (o->vcf.empty() || o->bam.empty() || o->fasta.empty() || o->barcodes.empty()) { fprintf(stderr, "error: --vcf, --bam, --fasta and --cell-barcodes are required\n"); return false; }
    if (o->scoring != "consensus" && o->scoring != "coverage" && o->scoring != "alt_frac") { fprintf(stderr, "error: invalid --scoring-method\n"); return false; }
    if (o->bam_tag.size() != 2) { fprintf(stderr, "error: --bam-tag must have two characters\n"); return false; }
    if (o->threads < 1) o->threads = 1;
    if (o->shard_loci < 0) o->shard_loci = 0;
    return true;
}

// validate_output_path (main.rs:475-491): refuse to overwrite, parent directory must exist
void validate_output_path(const std::string& p)
{
    if (exists(p)) { LOG_ERR("Output path already exists"); exit(1); }
    size_t s = p.find_last_of('/');
    if (s != std::string::npos && s > 0 && !exists(p.substr(0, s))) { LOG_ERR("Output directory \"%s\" does not exist", p.substr(0, s).c_str()); exit(1); }
}

// check_inputs_exist (main.rs:493-542)
void check_inputs_exist(const Opts& o)
{
    for (const std::string* p : { &o.fasta, &o.vcf, &o.bam, &o.barcodes })
        if (!exists(*p)) { LOG_ERR("Input file %s does not exist", p->c_str()); exit(1); }
    if (o.dump_staged.empty()) { validate_output_path(o.out_matrix); validate_output_path(o.ref_matrix); }
    if (!exists(o.fasta + ".fai")) { LOG_ERR("File %s.fai does not exist", o.fasta.c_str()); exit(1); }
    const size_t dot = o.bam.find_last_of('.');
    const std::string ext = dot == std::string::npos ? "" : o.bam.substr(dot + 1);
    if (ext == "bam") {
        const bool bai = exists(o.bam + ".bai") || exists(o.bam.substr(0, dot) + ".bai");
        if (!bai && exists(o.bam + ".csi")) { LOG_ERR("%s.csi: CSI indices are not supported by this build (no htslib); create a BAI index (samtools index -b)", o.bam.c_str()); exit(1); }
        if (!bai) { LOG_ERR("BAM index does not exist. Expecting %s.bai", o.bam.c_str()); exit(1); }
    } else if (ext == "cram") {
        LOG_ERR("CRAM input is not supported by this build (no htslib); convert to BAM"); exit(1);
    } else { LOG_ERR("BAM file did not end in .bam or .cram. Unable to validate"); exit(1); }
}

void dump_shard(FILE* f, const StagedShard& s)
{
    auto put = [&](const void* p, size_t bytes) { uint64_t n = bytes; fwrite(&n, 8, 1, f); if (bytes) fwrite(p, 1, bytes, f); };
#define PUTV(v) put((v).data(), (v).size() * sizeof((v)[0]))
    fwrite("VTXS", 1, 4, f);
    PUTV(s.locus_row); PUTV(s.hap_bytes); PUTV(s.ref_off); PUTV(s.ref_len); PUTV(s.alt_off); PUTV(s.alt_len); PUTV(s.cand_start);
    PUTV(s.read_nib); PUTV(s.read_off); PUTV(s.read_len); PUTV(s.cb_bytes); PUTV(s.read_cb_off); PUTV(s.read_cb_len);
    PUTV(s.read_umi_key); PUTV(s.cand_read);
#undef PUTV
    uint64_t m[7] = { s.met.num_reads, s.met.num_low_mapq, s.met.num_non_primary, s.met.num_duplicates, s.met.num_not_useful,
                      s.met.num_invalid_recs, s.met.num_multiallelic_recs };
    fwrite(m, 8, 7, f);
}

// pinned arena holding one shard for the asynchronous copy
struct Arena {
    uint8_t* base = nullptr; size_t cap = 0;
    bool ensure(size_t bytes)
    {
        if (bytes <= cap) return true;
        if (base) vtx_host_free(base);
        void* p = nullptr;
        if (vtx_host_alloc(&p, bytes + bytes / 4) != VTX_OK) { base = nullptr; cap = 0; return false; }
        base = static_cast<uint8_t*>(p); cap = bytes + bytes / 4;
        return true;
    }
    ~Arena() { if (base) vtx_host_free(base); }
};

// Copies the shard's arrays into the pinned arena.  This runs on the submitting thread while up to `--threads` workers
// stage; with many workers a single memcpy stream (~10 GB/s) would cap the whole pipeline near 70 M reads/s, so shards
// above a few megabytes are copied in 4 MB pieces by the caller plus up to three helper threads.
void stage_into_arena(const StagedShard& s, Arena& a, vtx_batch* b)
{
    struct Job { uint8_t* dst; const uint8_t* src; size_t bytes; };
    std::vector<Job> jobs;
    size_t off = 0, total = 0;
    auto put = [&](const void* p, size_t bytes) -> const void* {
        off = (off + 15) & ~size_t(15);
        uint8_t* d = a.base + off;
        constexpr size_t kPiece = size_t(4) << 20;
        for (size_t o = 0; o < bytes; o += kPiece) jobs.push_back({ d + o, static_cast<const uint8_t*>(p) + o, std::min(kPiece, bytes - o) });
        off += bytes; total += bytes;
        return d;
    };
    s.fill(b);
#define MV(field, vec) b->field = static_cast<decltype(b->field)>(put((vec).data(), (vec).size() * sizeof((vec)[0])))
    MV(locus_row, s.locus_row); MV(hap_bytes, s.hap_bytes); MV(ref_off, s.ref_off); MV(ref_len, s.ref_len); MV(alt_off, s.alt_off);
    MV(alt_len, s.alt_len); MV(cand_start, s.cand_start); MV(read_nib, s.read_nib); MV(read_off, s.read_off); MV(read_len, s.read_len);
    MV(cb_bytes, s.cb_bytes); MV(read_cb_off, s.read_cb_off); MV(read_cb_len, s.read_cb_len); MV(read_umi_key, s.read_umi_key);
    MV(cand_read, s.cand_read);
#undef MV
    std::atomic<size_t> next{ 0 };
    auto run = [&]() {
        for (size_t j; (j = next.fetch_add(1)) < jobs.size();) memcpy(jobs[j].dst, jobs[j].src, jobs[j].bytes);
    };
    const size_t helpers = total >= (size_t(8) << 20) ? std::min<size_t>(3, jobs.size() - 1) : 0;
    std::vector<std::thread> pool;
    for (size_t t = 0; t < helpers; ++t) pool.emplace_back(run);
    run();
    for (auto& t : pool) t.join();
}

}  // namespace

int main(int argc, char** argv)
{
    // staging grows multi-megabyte vectors on many threads: keep them on the heap arenas instead of
    // mmap/munmap per reallocation (which serialises the threads on the process address-space lock)
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    Opts o;
    if (!parse(argc, argv, &o)) { usage(); return 1; }
    check_inputs_exist(o);
    std::string err;

    BarcodeList bcs;
    if (!load_barcodes(o.barcodes, &bcs, &err)) { LOG_ERR("%s", err.c_str()); return 1; }
    LOG_INFO("Loaded %zu barcodes", bcs.keys.size());

    // CUDA context creation takes ~1 s: start it now, in the background, while the VCF is parsed and the first
    // shards are staged
    vtx_ctx* ctx = nullptr;
    std::string engine_err;
    std::future<int> engine_ready;
    if (o.dump_staged.empty()) {
        engine_ready = std::async(std::launch::async, [&]() -> int {
            vtx_config cfg{};
            cfg.device = int(o.device);
            cfg.mode = o.scoring == "consensus" ? VTX_MODE_CONSENSUS : o.scoring == "coverage" ? VTX_MODE_COVERAGE : VTX_MODE_ALT_FRAC;
            cfg.flags = VTX_F_VALUES_ONLY;       // the writers need row, col and the matrix values only
            cfg.use_umi = o.umi; cfg.match = 1; cfg.mismatch = -5; cfg.gap_open = -5; cfg.gap_extend = -1; cfg.min_score = 25;
            if (vtx_create(&cfg, &ctx) != VTX_OK) { engine_err = vtx_last_error(nullptr); return 1; }
            if (vtx_set_barcodes(ctx, bcs.bytes.data(), bcs.off.data(), uint32_t(bcs.keys.size())) != VTX_OK) { engine_err = vtx_last_error(ctx); return 1; }
            return 0;
        });
    }

    std::vector<VcfRecord> recs;
    if (!read_vcf(o.vcf, &recs, &err)) { printf("Vartrix error.\nError: %s\n", err.c_str()); return 1; }
    if (recs.empty()) LOG_ERR("Warning! Zero variants found in input VCF. Output matrices will be by definition empty but will still be generated.");
    LOG_INFO("Initialized a %zu variants x %zu cell barcodes matrix", recs.size(), bcs.keys.size());
    LOG_INFO("[%.3f s] inputs parsed", now_s());

    // validate_inputs (main.rs:545-594): contigs present in FASTA and BAM, record end inside the contig
    Fasta fa0;
    if (!fa0.open(o.fasta, &err)) { LOG_ERR("%s", err.c_str()); return 1; }
    {
        BamFile b0;
        if (!b0.open(o.bam, &err)) { printf("Vartrix error.\nError: error opening bam file: %s (%s)\n", o.bam.c_str(), err.c_str()); return 1; }
        for (const VcfRecord& r : recs) {
            if (!fa0.has(r.chrom)) { LOG_ERR("Sequence %s not seen in FASTA", r.chrom.c_str()); return 1; }
            if (b0.tid_of(r.chrom) < 0) { LOG_ERR("Sequence %s not seen in BAM", r.chrom.c_str()); return 1; }
            const int64_t end = r.pos0 + int64_t(r.alleles[0].size());
            if (end > fa0.length(r.chrom)) {
                LOG_ERR("Record %s:%lld has end position %lld, which is larger than the chromosome length (%lld). Does your FASTA match your VCF?",
                        r.chrom.c_str(), (long long)r.pos0, (long long)end, (long long)fa0.length(r.chrom));
                return 1;
            }
        }
    }

    StageArgs sa;
    sa.padding = o.padding; sa.mapq = uint32_t(o.mapq); sa.primary_only = o.primary; sa.no_duplicates = o.no_dups;
    sa.bam_tag[0] = o.bam_tag[0]; sa.bam_tag[1] = o.bam_tag[1];
    for (unsigned char c : o.valid_chars) sa.valid[c] = true;

    // ---- staging: worker threads produce shards of `shard_loci` records; the main thread consumes them in order ----
    // default shard size: 2048 loci (~100 k candidates at 50x, enough to fill the GPU), smaller when the VCF is short so that
    // every staging thread still gets ~10 shards (load balance; the GPU is idle most of the time anyway)
    if (o.shard_loci == 0) o.shard_loci = long(std::min<size_t>(2048, std::max<size_t>(128, recs.size() / (size_t(o.threads) * 10 + 1))));
    const size_t n_shards = (recs.size() + size_t(o.shard_loci) - 1) / size_t(o.shard_loci);
    std::vector<std::unique_ptr<StagedShard>> ready(n_shards);
    std::mutex mu; std::condition_variable cv;
    std::atomic<size_t> next{ 0 };
    size_t consumed = 0;                       // guarded by mu
    const size_t window = size_t(o.threads) * 2 + 2;
    bool failed = false; std::string fail_msg;
    UmiInterner umis;
    std::vector<std::unique_ptr<StagedShard>> recycled;      // guarded by mu
    auto worker = [&]() {
        Fasta fa; BamFile bam; std::string e;
        if (!fa.open(o.fasta, &e) || !bam.open(o.bam, &e)) { std::lock_guard<std::mutex> g(mu); failed = true; fail_msg = e; cv.notify_all(); return; }
        for (;;) {
            const size_t k = next.fetch_add(1);
            if (k >= n_shards) break;
            { std::unique_lock<std::mutex> g(mu); cv.wait(g, [&] { return failed || k < consumed + window; }); if (failed) return; }
            std::unique_ptr<StagedShard> sh;
            { std::lock_guard<std::mutex> g(mu); if (!recycled.empty()) { sh = std::move(recycled.back()); recycled.pop_back(); } }
            if (!sh) sh = std::make_unique<StagedShard>();
            const size_t lo = k * size_t(o.shard_loci), hi = std::min(recs.size(), lo + size_t(o.shard_loci));
            if (!stage_loci(recs, lo, hi, fa, bam, sa, umis, sh.get(), &e)) { std::lock_guard<std::mutex> g(mu); failed = true; fail_msg = e; cv.notify_all(); return; }
            { std::lock_guard<std::mutex> g(mu); ready[k] = std::move(sh); }
            cv.notify_all();
        }
    };
    HostMetrics hm;
    FILE* dump = nullptr;
    Arena arenas[3];
    if (!o.dump_staged.empty()) {            // before the pool starts: an early return must not leave joinable threads behind
        dump = fopen(o.dump_staged.c_str(), "wb");
        if (!dump) { LOG_ERR("cannot write %s", o.dump_staged.c_str()); return 1; }
        uint64_t hdr[2] = { recs.size(), bcs.keys.size() };
        fwrite(hdr, 8, 2, dump);
    }
    std::vector<std::thread> pool;
    for (long t = 0; t < o.threads; ++t) pool.emplace_back(worker);

    if (!dump && engine_ready.get() != 0) {
        printf("Vartrix error.\nError: %s\n", engine_err.c_str());
        { std::lock_guard<std::mutex> g(mu); failed = true; }
        cv.notify_all();
        for (auto& t : pool) t.join();
        return 1;
    }
    LOG_INFO("[%.3f s] engine ready, staging on %ld thread(s)", now_s(), o.threads);
    int rc = 0;
    for (size_t k = 0; k < n_shards && rc == 0; ++k) {
        std::unique_ptr<StagedShard> sh;
        {
            std::unique_lock<std::mutex> g(mu);
            cv.wait(g, [&] { return failed || ready[k]; });
            if (failed) { rc = 1; break; }
            sh = std::move(ready[k]);
            consumed = k + 1;
        }
        cv.notify_all();
        hm.add(sh->met);
        auto recycle = [&]() { sh->clear(); std::lock_guard<std::mutex> g(mu); recycled.push_back(std::move(sh)); };
        if (dump) { dump_shard(dump, *sh); recycle(); continue; }
        Arena& ar = arenas[k % 3];
        if (k >= 3 && vtx_wait_copies(ctx) != VTX_OK) { rc = 1; break; }     // the arena's previous copy must have landed
        if (!ar.ensure(sh->bytes())) { LOG_ERR("pinned allocation failed"); rc = 1; break; }
        vtx_batch b;
        stage_into_arena(*sh, ar, &b);
        if (vtx_submit(ctx, &b) != VTX_OK) { printf("Vartrix error.\nError: %s\n", vtx_last_error(ctx)); rc = 1; }
        recycle();
    }
    if (rc) { std::lock_guard<std::mutex> g(mu); failed = true; }
    cv.notify_all();
    for (auto& t : pool) t.join();
    if (failed && !fail_msg.empty()) { printf("Vartrix error.\nError: %s\n", fail_msg.c_str()); rc = 1; }
    if (dump) { fclose(dump); return rc; }
    if (rc) { vtx_destroy(ctx); return rc; }

    LOG_INFO("[%.3f s] all shards staged and submitted", now_s());
    vtx_result res{};
    if (vtx_finish(ctx, &res) != VTX_OK) { printf("Vartrix error.\nError: %s\n", vtx_last_error(ctx)); vtx_destroy(ctx); return 1; }

    LOG_INFO("[%.3f s] triplets on the host", now_s());
    // metrics (main.rs:350-379)
    LOG_INFO("Number of alignments evaluated: %llu", (unsigned long long)hm.num_reads);
    LOG_INFO("Number of alignments skipped due to low mapping quality: %llu", (unsigned long long)hm.num_low_mapq);
    LOG_INFO("Number of alignments skipped due to not being primary: %llu", (unsigned long long)hm.num_non_primary);
    LOG_INFO("Number of alignments skipped due to being duplicates: %llu", (unsigned long long)hm.num_duplicates);
    LOG_INFO("Number of alignments skipped due to not being associated with a cell barcode: %llu", (unsigned long long)res.metrics.num_not_cell_bc);
    LOG_INFO("Number of alignments skipped due to not intersecting variant: %llu", (unsigned long long)hm.num_not_useful);
    LOG_INFO("Number of alignments skipped due to not having a UMI: %llu", (unsigned long long)res.metrics.num_non_umi);
    LOG_INFO("Number of VCF records skipped due to having invalid characters in the alternative haplotype: %llu", (unsigned long long)hm.num_invalid_recs);
    LOG_INFO("Number of VCF records skipped due to being multi-allelic: %llu", (unsigned long long)hm.num_multiallelic_recs);
    LOG_INFO("Number of (read, locus) pairs scored on the GPU: %llu", (unsigned long long)res.metrics.num_scored);

    if (!write_mtx(o.out_matrix, recs.size(), bcs.keys.size(), res.n, res.row, res.col, res.val, &err, unsigned(o.threads))) { printf("Vartrix error.\nError: Error writing out-matrix\n"); rc = 1; }
    if (o.scoring == "coverage")        // clap-2 default_value counts as present (main.rs:100, 385)
        if (!write_mtx(o.ref_matrix, recs.size(), bcs.keys.size(), res.n, res.row, res.col, res.val2, &err, unsigned(o.threads))) { printf("Vartrix error.\nError: Error writing ref-matrix\n"); rc = 1; }

    if (!o.out_variants.empty()) {      // write_variants (main.rs:1166-1179): chrom_pos0
        validate_output_path(o.out_variants);
        FILE* f = fopen(o.out_variants.c_str(), "wb");
        if (!f) { LOG_ERR("error writing variants file"); rc = 1; }
        else { for (const VcfRecord& r : recs) fprintf(f, "%s_%lld\n", r.chrom.c_str(), (long long)r.pos0); fclose(f); }
    }
    if (!o.out_barcodes.empty()) {      // write_barcodes (main.rs:1181-1195): index order
        validate_output_path(o.out_barcodes);
        FILE* f = fopen(o.out_barcodes.c_str(), "wb");
        if (!f) { LOG_ERR("error writing barcodes file"); rc = 1; }
        else { for (const std::string& k : bcs.keys) { fwrite(k.data(), 1, k.size(), f); fputc('\n', f); } fclose(f); }
    }
    LOG_INFO("[%.3f s] outputs written", now_s());
    double sum = 0;
    for (uint64_t k = 0; k < res.n; ++k) sum += res.val[k];
    if (sum == 0.0) LOG_ERR("The resulting matrix has a sum of 0. Did you use the --umi flag on data without UMIs?");   // main.rs:410-415
    // every output is closed; tearing the CUDA context and the pinned arenas down costs 0.5-0.9 s that a one-shot CLI
    // does not need to spend (the driver reclaims everything at process exit)
    fflush(nullptr);
    _exit(rc);
}

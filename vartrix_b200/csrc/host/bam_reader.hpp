// bam_reader.hpp -- BGZF + BAM + BAI region reader for the staging host (zlib only; htslib is not in
// this image).  Replaces what rust-htslib 0.36 / C htslib give the reference at
// /root/reference/src/main.rs:262, 470 (IndexedReader::from_path), 822-829 (fetch + records),
// 742/753 (Record::aux), 793-796 (cigar), 896 (seq).  Region semantics = the htslib iterator: every
// record of the contig with pos < end and bam_endpos > beg, in file order.
#pragma once
#include <zlib.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <string>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace vtxhost {

inline uint16_t rd16(const uint8_t* p) { return uint16_t(p[0] | (p[1] << 8)); }
inline uint32_t rd32(const uint8_t* p) { return uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24); }
inline uint64_t rd64(const uint8_t* p) { return uint64_t(rd32(p)) | (uint64_t(rd32(p + 4)) << 32); }

// ---------------------------------------------------------------------------------------------
// BGZF: concatenated gzip members (<= 64 KiB each) addressed by virtual offsets (coffset << 16 | uoffset)
// ---------------------------------------------------------------------------------------------
class Bgzf {
public:
    ~Bgzf() { close(); }
    bool open(const std::string& path, std::string* err)
    {
        close();
        fd_ = ::open(path.c_str(), O_RDONLY);
        if (fd_ < 0) { *err = "cannot open " + path; return false; }
        struct stat st; fstat(fd_, &st); size_ = uint64_t(st.st_size);
        memset(&zs_, 0, sizeof(zs_));
        if (inflateInit2(&zs_, -15) != Z_OK) { *err = "inflateInit2 failed"; return false; }
        zinit_ = true;
        cbuf_.resize(1 << 16); ubuf_.resize(1 << 16);
        block_coff_ = 0; block_len_ = 0; block_pos_ = 0; next_coff_ = 0; have_block_ = false;
        return true;
    }
    void close()
    {
        if (zinit_) { inflateEnd(&zs_); zinit_ = false; }
        if (fd_ >= 0) { ::close(fd_); fd_ = -1; }
    }
    bool seek(uint64_t voff)
    {
        const uint64_t coff = voff >> 16; const uint32_t uoff = uint32_t(voff & 0xFFFF);
        if (!have_block_ || coff != block_coff_) { if (!load(coff)) return false; }
        if (uoff > block_len_) return false;
        block_pos_ = uoff;
        return true;
    }
    uint64_t tell()
    {   // htslib convention: at the end of a block the position is the start of the next one
        if (have_block_ && block_pos_ == block_len_ && block_len_ > 0) return next_coff_ << 16;
        return (block_coff_ << 16) | block_pos_;
    }
    // read exactly n bytes; false at EOF / error
    bool read(void* dst, size_t n)
    {
        uint8_t* d = static_cast<uint8_t*>(dst);
        while (n) {
            if (!have_block_ || block_pos_ == block_len_) {
                if (!load(have_block_ ? next_coff_ : 0)) return false;
                if (block_len_ == 0) { if (eof_) return false; continue; }   // empty (EOF marker) block
            }
            const size_t k = std::min<size_t>(n, block_len_ - block_pos_);
            memcpy(d, ubuf_.data() + block_pos_, k);
            d += k; n -= k; block_pos_ += uint32_t(k);
        }
        return true;
    }
    bool at_eof() const { return eof_; }

private:
    // neighbouring loci fetch overlapping file ranges: keep the last few inflated blocks
    struct Cached { uint64_t coff = ~0ull, next = 0; std::vector<uint8_t> data; };
    static constexpr int kCache = 6;
    Cached cache_[kCache];
    int cache_rr_ = 0;

    bool load(uint64_t coff)
    {
        eof_ = false;
        for (Cached& c : cache_)
            if (c.coff == coff) {
                if (ubuf_.size() < c.data.size()) ubuf_.resize(c.data.size());
                memcpy(ubuf_.data(), c.data.data(), c.data.size());
                block_coff_ = coff; block_len_ = uint32_t(c.data.size()); block_pos_ = 0; next_coff_ = c.next; have_block_ = true;
                return true;
            }
        if (coff >= size_) { eof_ = true; have_block_ = true; block_coff_ = coff; block_len_ = 0; block_pos_ = 0; next_coff_ = coff; return false; }
        uint8_t hdr[18];
        if (pread(fd_, hdr, 18, off_t(coff)) != 18) { eof_ = true; return false; }
        if (hdr[0] != 31 || hdr[1] != 139) return false;
        const uint32_t xlen = rd16(hdr + 10);
        // locate the BC subfield (normally the first and only one)
        uint32_t bsize = 0;
        if (xlen == 6 && hdr[12] == 66 && hdr[13] == 67) bsize = rd16(hdr + 16);
        else {
            std::vector<uint8_t> x(xlen);
            if (pread(fd_, x.data(), xlen, off_t(coff + 12)) != ssize_t(xlen)) return false;
            for (uint32_t q = 0; q + 4 <= xlen;) {
                const uint32_t slen = rd16(x.data() + q + 2);
                if (x[q] == 66 && x[q + 1] == 67) bsize = rd16(x.data() + q + 4);
                q += 4 + slen;
            }
        }
        const uint32_t total = bsize + 1;
        if (total < 12 + xlen + 8) return false;
        const uint32_t clen = total - 12 - xlen - 8;
        if (cbuf_.size() < total) cbuf_.resize(total);
        if (pread(fd_, cbuf_.data(), clen + 8, off_t(coff + 12 + xlen)) != ssize_t(clen + 8)) return false;
        const uint32_t isize = rd32(cbuf_.data() + clen + 4);
        if (isize > ubuf_.size()) ubuf_.resize(isize);
        if (isize) {
            inflateReset(&zs_);
            zs_.next_in = cbuf_.data(); zs_.avail_in = clen;
            zs_.next_out = ubuf_.data(); zs_.avail_out = isize;
            if (inflate(&zs_, Z_FINISH) != Z_STREAM_END) return false;
        }
        block_coff_ = coff; block_len_ = isize; block_pos_ = 0; next_coff_ = coff + total; have_block_ = true;
        Cached& c = cache_[cache_rr_]; cache_rr_ = (cache_rr_ + 1) % kCache;
        c.coff = coff; c.next = next_coff_; c.data.assign(ubuf_.begin(), ubuf_.begin() + isize);
        return true;
    }
    int fd_ = -1;
    uint64_t size_ = 0;
    z_stream zs_;
    bool zinit_ = false, have_block_ = false, eof_ = false;
    std::vector<uint8_t> cbuf_, ubuf_;
    uint64_t block_coff_ = 0, next_coff_ = 0;
    uint32_t block_len_ = 0, block_pos_ = 0;
};

// ---------------------------------------------------------------------------------------------
// BAM record view + BAI index
// ---------------------------------------------------------------------------------------------
struct BamRecord {
    std::vector<uint8_t> data;    // the record without its block_size prefix
    uint64_t voff = 0;            // virtual offset of the record (identity of the record inside the file)
    int32_t refid() const { return int32_t(rd32(data.data())); }
    int32_t pos() const { return int32_t(rd32(data.data() + 4)); }
    uint32_t l_read_name() const { return data[8]; }
    uint32_t mapq() const { return data[9]; }
    uint32_t n_cigar() const { return rd16(data.data() + 12); }
    uint32_t flag() const { return rd16(data.data() + 14); }
    int32_t l_seq() const { return int32_t(rd32(data.data() + 16)); }
    const uint8_t* cigar() const { return data.data() + 32 + l_read_name(); }
    const uint8_t* seq() const { return cigar() + 4 * n_cigar(); }
    const uint8_t* aux() const { return seq() + (l_seq() + 1) / 2 + l_seq(); }
    const uint8_t* end() const { return data.data() + data.size(); }
    // htslib bam_endpos: pos + reference length of the CIGAR; pos + 1 when unmapped / no CIGAR / zero length
    int64_t endpos() const
    {
        int64_t rlen = 0;
        if (!(flag() & 4)) {
            const uint8_t* c = cigar();
            for (uint32_t i = 0; i < n_cigar(); ++i) {
                const uint32_t v = rd32(c + 4 * i), op = v & 0xF;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += v >> 4;
            }
        }
        return int64_t(pos()) + (rlen > 0 ? rlen : 1);
    }
    // first aux field named `tag`: value bytes when its type is 'Z' (Record::aux -> Aux::String), else nullptr
    const uint8_t* aux_z(const char tag[2], uint32_t* len) const
    {
        const uint8_t* p = aux(); const uint8_t* e = end();
        while (p + 3 <= e) {
            const bool hit = p[0] == uint8_t(tag[0]) && p[1] == uint8_t(tag[1]);
            const uint8_t ty = p[2];
            p += 3;
            if (ty == 'Z' || ty == 'H') {
                const uint8_t* q = static_cast<const uint8_t*>(memchr(p, 0, size_t(e - p)));
                if (!q) return nullptr;
                if (hit) { if (ty == 'Z') { *len = uint32_t(q - p); return p; } return nullptr; }
                p = q + 1;
            } else {
                size_t sz;
                switch (ty) {
                case 'A': case 'c': case 'C': sz = 1; break;
                case 's': case 'S': sz = 2; break;
                case 'i': case 'I': case 'f': sz = 4; break;
                case 'B': {
                    if (p + 5 > e) return nullptr;
                    const uint8_t sub = p[0]; const uint32_t cnt = rd32(p + 1);
                    const size_t es = (sub == 'c' || sub == 'C') ? 1 : (sub == 's' || sub == 'S') ? 2 : 4;
                    sz = 5 + size_t(cnt) * es; break;
                }
                default: return nullptr;
                }
                if (hit) return nullptr;
                p += sz;
            }
        }
        return nullptr;
    }
};

struct BaiChunk { uint64_t beg, end; };
struct BaiRef {
    std::vector<uint32_t> bin_ids;
    std::vector<std::vector<BaiChunk>> bin_chunks;
    std::vector<uint64_t> linear;
};

class BamFile {
public:
    bool open(const std::string& path, std::string* err)
    {
        path_ = path;
        if (!bg_.open(path, err)) return false;
        uint8_t magic[4];
        if (!bg_.read(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) { *err = path + " is not a BAM file"; return false; }
        uint8_t b4[4];
        if (!bg_.read(b4, 4)) { *err = "truncated BAM header"; return false; }
        std::vector<uint8_t> text(rd32(b4));
        if (!text.empty() && !bg_.read(text.data(), text.size())) { *err = "truncated BAM header"; return false; }
        if (!bg_.read(b4, 4)) { *err = "truncated BAM header"; return false; }
        const uint32_t n_ref = rd32(b4);
        for (uint32_t i = 0; i < n_ref; ++i) {
            if (!bg_.read(b4, 4)) { *err = "truncated BAM header"; return false; }
            std::vector<char> name(rd32(b4));
            if (!bg_.read(name.data(), name.size()) || !bg_.read(b4, 4)) { *err = "truncated BAM header"; return false; }
            ref_names.emplace_back(name.data());
            ref_lens.push_back(rd32(b4));
        }
        return load_index(err);
    }
    int tid_of(const std::string& chrom) const
    {
        for (size_t i = 0; i < ref_names.size(); ++i) if (ref_names[i] == chrom) return int(i);
        return -1;
    }
    std::vector<std::string> ref_names;
    std::vector<uint32_t> ref_lens;

    // htslib-style region iterator: call fetch(), then next() until it returns false
    bool fetch(int tid, int64_t beg, int64_t end)
    {
        chunks_.clear(); cur_ = 0; positioned_ = false; tid_ = tid; beg_ = beg; end_ = end; done_ = false;
        if (tid < 0 || size_t(tid) >= refs_.size()) { done_ = true; return false; }
        if (beg < 0) beg = 0;
        if (end <= beg) { done_ = true; return true; }
        const BaiRef& r = refs_[tid];
        uint64_t min_off = 0;
        if (!r.linear.empty()) {
            size_t w = size_t(beg >> 14);
            if (w >= r.linear.size()) w = r.linear.size() - 1;
            min_off = r.linear[w];
            // htslib walks back over empty windows; offsets are monotone so the entry itself is a safe lower bound
        }
        const int64_t e1 = end - 1;
        auto add_bin = [&](uint32_t bin) {
            auto it = std::lower_bound(r.bin_ids.begin(), r.bin_ids.end(), bin);
            if (it == r.bin_ids.end() || *it != bin) return;
            for (const BaiChunk& c : r.bin_chunks[size_t(it - r.bin_ids.begin())]) if (c.end > min_off) chunks_.push_back(c);
        };
        add_bin(0);
        for (int64_t k = 1 + (beg >> 26); k <= 1 + (e1 >> 26); ++k) add_bin(uint32_t(k));
        for (int64_t k = 9 + (beg >> 23); k <= 9 + (e1 >> 23); ++k) add_bin(uint32_t(k));
        for (int64_t k = 73 + (beg >> 20); k <= 73 + (e1 >> 20); ++k) add_bin(uint32_t(k));
        for (int64_t k = 585 + (beg >> 17); k <= 585 + (e1 >> 17); ++k) add_bin(uint32_t(k));
        for (int64_t k = 4681 + (beg >> 14); k <= 4681 + (e1 >> 14); ++k) add_bin(uint32_t(k));
        std::sort(chunks_.begin(), chunks_.end(), [](const BaiChunk& a, const BaiChunk& b) { return a.beg < b.beg; });
        size_t w = 0;                         // merge overlapping / adjacent chunks so no record is visited twice
        for (size_t i = 0; i < chunks_.size(); ++i) {
            if (w && chunks_[i].beg <= chunks_[w - 1].end) chunks_[w - 1].end = std::max(chunks_[w - 1].end, chunks_[i].end);
            else chunks_[w++] = chunks_[i];
        }
        chunks_.resize(w);
        if (chunks_.empty()) done_ = true;
        return true;
    }
    bool next(BamRecord* rec)
    {
        while (!done_) {
            if (!positioned_) {
                if (cur_ >= chunks_.size()) { done_ = true; break; }
                if (!bg_.seek(chunks_[cur_].beg)) { done_ = true; break; }
                positioned_ = true;
            }
            if (bg_.tell() >= chunks_[cur_].end) { ++cur_; positioned_ = false; continue; }
            const uint64_t voff = bg_.tell();
            uint8_t b4[4];
            if (!bg_.read(b4, 4)) { done_ = true; break; }
            const uint32_t bs = rd32(b4);
            rec->data.resize(bs);
            if (bs < 32 || !bg_.read(rec->data.data(), bs)) { done_ = true; break; }
            rec->voff = voff;
            if (rec->refid() != tid_ || int64_t(rec->pos()) >= end_) { done_ = true; break; }   // sorted file: nothing further can overlap
            if (rec->endpos() > beg_) return true;
        }
        return false;
    }

private:
    bool load_index(std::string* err)
    {
        std::string p = path_ + ".bai";
        FILE* f = fopen(p.c_str(), "rb");
        if (!f) {
            std::string q = path_;
            if (q.size() > 4 && q.substr(q.size() - 4) == ".bam") { q = q.substr(0, q.size() - 4) + ".bai"; f = fopen(q.c_str(), "rb"); }
        }
        if (!f) { *err = "BAM index does not exist. Expecting " + path_ + ".bai (CSI indices are not supported by this reader)"; return false; }
        std::vector<uint8_t> buf;
        fseek(f, 0, SEEK_END); const long sz = ftell(f); fseek(f, 0, SEEK_SET);
        buf.resize(size_t(sz));
        if (sz > 0 && fread(buf.data(), 1, size_t(sz), f) != size_t(sz)) { fclose(f); *err = "cannot read BAM index"; return false; }
        fclose(f);
        if (buf.size() < 8 || memcmp(buf.data(), "BAI\1", 4) != 0) { *err = "bad BAI magic"; return false; }
        size_t o = 4;
        const uint32_t n_ref = rd32(buf.data() + o); o += 4;
        refs_.resize(n_ref);
        for (uint32_t r = 0; r < n_ref; ++r) {
            if (o + 4 > buf.size()) { *err = "truncated BAI"; return false; }
            const uint32_t n_bin = rd32(buf.data() + o); o += 4;
            std::vector<std::pair<uint32_t, std::vector<BaiChunk>>> bins;
            for (uint32_t b = 0; b < n_bin; ++b) {
                if (o + 8 > buf.size()) { *err = "truncated BAI"; return false; }
                const uint32_t bin = rd32(buf.data() + o); const uint32_t n_chunk = rd32(buf.data() + o + 4); o += 8;
                if (o + size_t(n_chunk) * 16 > buf.size()) { *err = "truncated BAI"; return false; }
                std::vector<BaiChunk> cs(n_chunk);
                for (uint32_t c = 0; c < n_chunk; ++c) { cs[c].beg = rd64(buf.data() + o); cs[c].end = rd64(buf.data() + o + 8); o += 16; }
                if (bin != 37450) bins.emplace_back(bin, std::move(cs));      // 37450 = metadata pseudo-bin
            }
            std::sort(bins.begin(), bins.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
            for (auto& b : bins) { refs_[r].bin_ids.push_back(b.first); refs_[r].bin_chunks.push_back(std::move(b.second)); }
            if (o + 4 > buf.size()) { *err = "truncated BAI"; return false; }
            const uint32_t n_intv = rd32(buf.data() + o); o += 4;
            if (o + size_t(n_intv) * 8 > buf.size()) { *err = "truncated BAI"; return false; }
            refs_[r].linear.resize(n_intv);
            for (uint32_t i = 0; i < n_intv; ++i) { refs_[r].linear[i] = rd64(buf.data() + o); o += 8; }
        }
        return true;
    }
    std::string path_;
    Bgzf bg_;
    std::vector<BaiRef> refs_;
    std::vector<BaiChunk> chunks_;
    size_t cur_ = 0;
    bool positioned_ = false, done_ = true;
    int tid_ = -1;
    int64_t beg_ = 0, end_ = 0;
};

}  // namespace vtxhost

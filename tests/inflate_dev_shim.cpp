// test shim: the bit-stream half of the DEVICE DEFLATE decoder (vartrix_b200/csrc/vtx_inflate.cuh: table builders, block
// headers, symbol batches -- all __host__ __device__) driven on the CPU with a serial stand-in for the warp's apply step.
#include <cstring>
#include <vector>
#include "../vartrix_b200/csrc/vtx_inflate.cuh"
extern "C" int vtx_test_inflate_dev(const unsigned char* in, unsigned long in_len, unsigned char* out, unsigned long out_len)
{
    using namespace vtx::inflate;
    std::vector<uint32_t> buf((in_len + 3) / 4 + 4, 0xA5A5A5A5u);          // 4-byte aligned copy with >= 8 bytes of padding behind it
    memcpy(buf.data(), in, in_len);
    State st;
    state_init(st, reinterpret_cast<const uint8_t*>(buf.data()), uint32_t(in_len), uint32_t(out_len));
    static thread_local Tables T;
    uint8_t lens[512];
    Sym batch[32];
    size_t op = 0;
    while (st.status == kOk && st.phase != 3) {
        const uint32_t word0 = st.word; const int phase0 = st.phase, cnt0 = st.bitcnt; const uint32_t left0 = st.stored_left;
        const int n = decode_batch(st, T, lens, batch, 32);
        // what the kernel relies on: a call reads less than kMaxInputPerCall bytes beyond where it started
        if (st.phase != 2 && st.word > word0 + kMaxInputPerCall / 4 + 2) return -2;
        for (int k = 0; k < n; ++k) {
            const Sym& sy = batch[k];
            if (op + sy.len > out_len) return kBadSize;
            if (sy.kind == 0) out[op] = (unsigned char)sy.arg;
            else if (sy.kind == 1) { for (uint32_t i = 0; i < sy.len; ++i) out[op + i] = out[op - sy.arg + (sy.arg >= sy.len ? i : i % sy.arg)]; }
            else memcpy(out + op, reinterpret_cast<const uint8_t*>(buf.data()) + sy.arg, sy.len);
            op += sy.len;
        }
        if (n == 0 && st.status == kOk && st.phase != 3 && st.word == word0 && st.phase == phase0 && st.bitcnt == cnt0 && st.stored_left == left0)
            return -1;                                                   // no progress: would loop forever
    }
    return st.status;
}

// test shim: exposes the host's DEFLATE decoder (vartrix_b200/csrc/host/inflate_fast.hpp) to ctypes
#include "../vartrix_b200/csrc/host/inflate_fast.hpp"
extern "C" int vtx_test_inflate(const unsigned char* in, unsigned long in_len, unsigned char* out, unsigned long out_len)
{
    return vtxhost::vtx_inflate_raw(in, in_len, out, out_len) ? 1 : 0;
}
extern "C" unsigned long vtx_test_inflate_in_pad() { return vtxhost::kInflateInPad; }
extern "C" unsigned long vtx_test_inflate_out_pad() { return vtxhost::kInflateOutPad; }

// test shim: exposes pieces of the staging host (vartrix_b200/csrc/host/inputs.hpp) to ctypes
#include "../vartrix_b200/csrc/host/crc32_fast.hpp"
#include "../vartrix_b200/csrc/host/inputs.hpp"
extern "C" int vtx_test_write_mtx(const char* path, unsigned long n_rows, unsigned long n_cols, unsigned long n, const uint32_t* row,
                                  const uint32_t* col, const double* val, unsigned threads)
{
    std::string err;
    return vtxhost::write_mtx(path, n_rows, n_cols, n, row, col, val, &err, threads) ? 1 : 0;
}

// the carry-less-multiplication CRC-32 of the staging host (csrc/host/crc32_fast.hpp)
extern "C" uint32_t vtx_test_crc32(const unsigned char* p, unsigned long n) { return vtx_crc::crc32_of(p, n); }
extern "C" int vtx_test_crc32_uses_clmul() {
#if VTX_CRC_CLMUL
    return vtx_crc::have_clmul() ? 1 : 0;
#else
    return 0;
#endif
}

// test shim: exposes pieces of the staging host (vartrix_b200/csrc/host/inputs.hpp) to ctypes
#include "../vartrix_b200/csrc/host/inputs.hpp"
extern "C" int vtx_test_write_mtx(const char* path, unsigned long n_rows, unsigned long n_cols, unsigned long n, const uint32_t* row,
                                  const uint32_t* col, const double* val, unsigned threads)
{
    std::string err;
    return vtxhost::write_mtx(path, n_rows, n_cols, n, row, col, val, &err, threads) ? 1 : 0;
}

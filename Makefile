# vartrix_b200 -- build the sm_100a engine library (and the oracle used by the tests).
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CSRC      := vartrix_b200/csrc
LIBDIR    := vartrix_b200/lib
LIB       := $(LIBDIR)/libvartrix_b200.so

all: $(LIB) oracle

$(LIB): $(CSRC)/vtx_api.cu $(CSRC)/vtx_sw.cuh $(CSRC)/vtx_pipeline.cuh include/vartrix_b200.h
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CSRC)/vtx_api.cu -ldl 2> $(LIBDIR)/ptxas.log || (cat $(LIBDIR)/ptxas.log; exit 1)
	@grep -E "error|warning" $(LIBDIR)/ptxas.log | grep -v "ptxas info" || true

oracle:
	$(MAKE) -s -C oracle

clean:
	rm -f $(LIB) $(LIBDIR)/ptxas.log
	$(MAKE) -s -C oracle clean

.PHONY: all oracle clean

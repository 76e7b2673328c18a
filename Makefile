# Build of the B200 batched-inversion engine.
#   make lib     -> cuda_matrix_inversion_b200/lib/libinvgpu.so   (C ABI: include/*.h)
#   make cli     -> bin/inverse_bench bin/gauss_bench             (reference-compatible CLIs)
#   make oracle  -> oracle/liboracle.so (+ oracle/_ref when /root/reference is mounted) -- checker only
# sm_100a only; -lineinfo so ncu source pages map to the .cu/.cuh files.

NVCC     ?= /usr/local/cuda/bin/nvcc
CC        = gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v $(EXTRA_NVFLAGS)
CSRC     := cuda_matrix_inversion_b200/csrc
LIBDIR   := cuda_matrix_inversion_b200/lib
LIB      := $(LIBDIR)/libinvgpu.so
HDRS     := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)

all: lib cli

lib: $(LIB)

CUOBJS   := $(LIBDIR)/capi.o $(patsubst $(CSRC)/%.cu,$(LIBDIR)/%.o,$(wildcard $(CSRC)/inst_*.cu))

$(LIBDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/ptxas_$*.log || (cat $(LIBDIR)/ptxas_$*.log; false)

$(LIBDIR)/mats_io.o: $(CSRC)/mats_io.c include/helper_cpu.h include/types.h
	@mkdir -p $(LIBDIR)
	$(CC) -O2 -fPIC -std=gnu11 -c $< -o $@

$(LIB): $(CUOBJS) $(LIBDIR)/mats_io.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

cli: bin/inverse_bench bin/gauss_bench

bin/%: $(CSRC)/%.c $(CSRC)/bench_common.h $(LIB) $(HDRS)
	@mkdir -p bin
	$(CC) -O2 -std=gnu11 -fopenmp -o $@ $< -L$(LIBDIR) -linvgpu -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -ldl -lm

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR) bin

.PHONY: all lib cli oracle clean

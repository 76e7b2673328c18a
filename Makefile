# Build of the B200 batched-inversion engine.
#   make lib     -> cuda_matrix_inversion_b200/lib/libinvgpu.so   (C ABI: include/*.h)
#   make cli     -> bin/inverse_bench bin/gauss_bench             (reference-compatible CLIs)
#   make oracle  -> oracle/liboracle.so (+ oracle/_ref when /root/reference is mounted) -- checker only
# sm_100a only; -lineinfo so ncu source pages map to the .cu/.cuh files.

NVCC     ?= /usr/local/cuda/bin/nvcc
CC        = gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v $(EXTRA_NVFLAGS)
# `make log=1`: the reference's detailed-logging build (Makefile:115-117, -DDETAILED_LOGGING): every *_gpu wrapper
# prints its phase timers "<name>_mem_htod|_ker|_mem_dtoh,batch,n,ms,ns" (same effect at run time: INVGPU_DETAILED_LOGGING=1)
ifeq ($(log),1)
NVFLAGS  += -DDETAILED_LOGGING
endif
# `make lab=1`: also build the measured-but-not-default kernel generations (csrc/tile_configs.h, INVGPU_LAB)
ifeq ($(lab),1)
NVFLAGS  += -DINVGPU_LAB=1
endif
CSRC     := cuda_matrix_inversion_b200/csrc
LIBDIR   := cuda_matrix_inversion_b200/lib
LIB      := $(LIBDIR)/libinvgpu.so
HDRS     := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)

all: lib cli tools

lib: $(LIB)

CUOBJS   := $(LIBDIR)/capi.o $(patsubst $(CSRC)/%.cu,$(LIBDIR)/%.o,$(wildcard $(CSRC)/inst_*.cu))

$(LIBDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(LIBDIR)/ptxas_$*.log || (cat $(LIBDIR)/ptxas_$*.log; false)
	@sed -i '/Compile time/d' $(LIBDIR)/ptxas_$*.log

$(LIBDIR)/mats_io.o: $(CSRC)/mats_io.c include/helper_cpu.h include/types.h
	@mkdir -p $(LIBDIR)
	$(CC) -O2 -fPIC -std=gnu11 -c $< -o $@

$(LIB): $(CUOBJS) $(LIBDIR)/mats_io.o
	$(NVCC) $(ARCH) -shared -o $@ $^ -cudart static

cli: bin/inverse_bench bin/gauss_bench

bin/%: $(CSRC)/%.c $(CSRC)/bench_common.h $(LIB) $(HDRS)
	@mkdir -p bin
	$(CC) -O2 -std=gnu11 -fopenmp -o $@ $< -L$(LIBDIR) -linvgpu -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)' -ldl -lm

tools: tools/bin/xfer_bench tools/bin/microbench

# transfer micro-benchmark (reference src/bench.cu counterpart) and the pipe-rate micro-benchmark
tools/bin/xfer_bench: tools/xfer_bench.cu $(CSRC)/host_numa.h include/invgpu.h $(LIB)
	@mkdir -p tools/bin
	$(NVCC) -O2 -std=c++17 $(ARCH) -o $@ $< -L$(LIBDIR) -linvgpu -Xlinker -rpath -Xlinker '$$ORIGIN/../../$(LIBDIR)' -lpthread

tools/bin/microbench: tools/microbench.cu
	@mkdir -p tools/bin
	$(NVCC) -O3 -std=c++17 $(ARCH) -o $@ $<

# Sweeps of the reference's run-inverse-bench / run-gauss-bench targets (Makefile:202-220): every fixture size x
# MATRIX_DUPLICATES in {1,2,4,8,16}, CSV lines concatenated into results/{inverse,gauss}-bench.txt (what
# results/generate_plots.m:1,22,47 reads).  Sizes without a fixture directory are skipped.
BENCH_NUM_THREADS ?= $(shell nproc)
BENCH_REPS        ?= 10
BENCH_TESTS       ?= tests/golden/reference
BENCH_SIZES       ?= 8 16 32 64 128
BENCH_DUPS        ?= 1 2 4 8 16

run-inverse-bench: bin/inverse_bench
	@mkdir -p results; : > results/inverse-bench.txt
	@for n in $(BENCH_SIZES); do for d in $(BENCH_DUPS); do \
	    dir=$(BENCH_TESTS)/inverse_100_$${n}x$${n}; [ -d $$dir ] || continue; \
	    echo "OMP_NUM_THREADS=$(BENCH_NUM_THREADS) bin/inverse_bench $$dir $(BENCH_REPS) $$d -csv"; \
	    OMP_NUM_THREADS=$(BENCH_NUM_THREADS) bin/inverse_bench $$dir $(BENCH_REPS) $$d -csv >> results/inverse-bench.txt || exit 1; \
	done; done

run-gauss-bench: bin/gauss_bench
	@mkdir -p results; : > results/gauss-bench.txt
	@for n in $(BENCH_SIZES); do for d in $(BENCH_DUPS); do \
	    dir=$(BENCH_TESTS)/gaussian_100_$${n}x$${n}; [ -d $$dir ] || continue; \
	    echo "OMP_NUM_THREADS=$(BENCH_NUM_THREADS) bin/gauss_bench $$dir $(BENCH_REPS) $$d -csv"; \
	    OMP_NUM_THREADS=$(BENCH_NUM_THREADS) bin/gauss_bench $$dir $(BENCH_REPS) $$d -csv >> results/gauss-bench.txt || exit 1; \
	done; done

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(LIBDIR) bin

.PHONY: all lib cli tools oracle clean run-inverse-bench run-gauss-bench

# Builds the product library (CUDA, sm_100a only), the host tools and the test-only oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
# EXTRA: e.g. -DLLC_ROLE_TIMING (per-role cycle counters of the fused coder, see scripts/role_stats.py)
EXTRA ?=
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v $(EXTRA)
SRC := $(wildcard llcomp_b200/csrc/*.cu)
HDR := $(wildcard llcomp_b200/csrc/*.cuh) include/llcomp_b200.h
LIB := llcomp_b200/lib/libllcomp_b200.so

all: $(LIB) tools oracle

$(LIB): $(SRC) $(HDR)
	@mkdir -p llcomp_b200/lib
	$(NVCC) $(NVFLAGS) -shared -cudart static -o $@ $(SRC) 2> llcomp_b200/lib/ptxas.log || (cat llcomp_b200/lib/ptxas.log; false)
	@grep -E "error|warning" llcomp_b200/lib/ptxas.log || true

tools: $(LIB)
	$(MAKE) -C llcomp_b200/host

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf llcomp_b200/lib llcomp_b200/host/llcompc llcomp_b200/host/llcompd
	$(MAKE) -C oracle clean

.PHONY: all tools oracle clean

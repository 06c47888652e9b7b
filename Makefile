# Builds the product library (CUDA kernels + C ABI + C++ host layer) in-tree for sm_100a, and the CPU
# oracle (test infrastructure).  No CPU fallback exists: the product library needs a GPU at run time.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX := $(shell test -x /usr/bin/g++ && echo /usr/bin/g++ || echo g++)

CSRC := bevyray_b200/csrc
CU_SRCS := $(wildcard $(CSRC)/*.cu)
HOST_SRCS := $(wildcard $(CSRC)/host/*.cpp)
HDRS := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/host/*.hpp) $(wildcard include/*.h)
LIB ?= bevyray_b200/libbevyray_b200.so
# compile-time variants for A/B runs: make LIB=/path/other.so DEFS="-DBVR_FAR_GENERIC=0"; BEVYRAY_B200_LIB picks it at run time
DEFS ?=

NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
	--fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
	-ccbin $(HOSTCXX) -Xcompiler -fPIC,-fopenmp,-ffp-contract=off,-fno-fast-math,-Wall \
	-Xptxas -v $(DEFS)

all: $(LIB) oracle

$(LIB): $(CU_SRCS) $(HOST_SRCS) $(HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CU_SRCS) $(HOST_SRCS) -lgomp 2> build.log || (cat build.log; exit 1)
	@grep -E "error|warning" build.log | grep -v "^ptxas info" || true

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) build.log
	$(MAKE) -C oracle clean

.PHONY: all oracle clean

# Builds catears_b200/libce_gpu.so -- the C-ABI shared library (include/ce_gpu.h) -- for sm_100a,
# and the test-only oracle libraries.  Everything is built in-tree so that the .so files
# travel to the GPU box with the repository snapshot.

NVCC      ?= nvcc
CXX       ?= g++
CUDA_HOME ?= /usr/local/cuda
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-Wall,-Wno-unused-function -Iinclude \
             -Icatears_b200/csrc --expt-relaxed-constexpr
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -Iinclude -Icatears_b200/csrc -I$(CUDA_HOME)/include
BUILD     := build
LIB       := catears_b200/libce_gpu.so

CU_SRCS   := $(wildcard catears_b200/csrc/*.cu)
CC_SRCS   := $(wildcard catears_b200/csrc/*.cc)
OBJS      := $(patsubst catears_b200/csrc/%.cu,$(BUILD)/%.cu.o,$(CU_SRCS)) \
             $(patsubst catears_b200/csrc/%.cc,$(BUILD)/%.cc.o,$(CC_SRCS))
HDRS      := $(wildcard catears_b200/csrc/*.h catears_b200/csrc/*.cuh) include/ce_gpu.h

.PHONY: all lib oracle clean
all: lib oracle
lib: $(LIB)

$(BUILD)/%.cu.o: catears_b200/csrc/%.cu $(HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> $(BUILD)/$*.ptxas.log || (cat $(BUILD)/$*.ptxas.log; false)

$(BUILD)/%.cc.o: catears_b200/csrc/%.cc $(HDRS)
	@mkdir -p $(BUILD)
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart_static -ldl -lpthread -lrt

oracle:
	$(MAKE) -C oracle port
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref ref-sse4 && $(MAKE) lib && $(MAKE) -C oracle dropin; else echo "reference absent: keeping prebuilt oracle/_ref"; fi

clean:
	rm -rf $(BUILD) $(LIB)

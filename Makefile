# Builds every native artefact in-tree (so the .so files travel with gpurun):
#   bam_parallel_b200/lib/libbamgpu.so   product: CUDA kernels + C-ABI host reader (sm_100a)
#   bam_parallel_b200/lib/libbamharness.so  per-record consumer loop for bench.py (e2e_next_rec)
#   bam_parallel_b200/lib/libbamgen.so   synthetic BAM generator (test/bench input only)
#   bam_parallel_b200/lib/bamgen         generator CLI
#   oracle/_build/liboracle.so           CPU oracle (test infrastructure only)
# oracle/_ref is NOT built: the reference is Rust and there is no Rust toolchain
# in this image (see DESIGN.md "Oracle").

NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
LIBDIR    := bam_parallel_b200/lib
CSRC      := bam_parallel_b200/csrc
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             -Xcompiler -fPIC,-Wall,-Wno-unused-function,-Wno-unknown-pragmas -Wno-deprecated-gpu-targets -Iinclude -I$(CSRC) $(EXTRA)
CXXFLAGS  := -O2 -std=c++17 -fPIC -Wall -pthread

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CPP_SRCS  := $(wildcard $(CSRC)/*.cpp)
HDRS      := $(wildcard $(CSRC)/*.h $(CSRC)/*.cuh include/*.h)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,build/%.cu.o,$(CU_SRCS))
CPP_OBJS  := $(patsubst $(CSRC)/%.cpp,build/%.cpp.o,$(CPP_SRCS))

all: $(LIBDIR)/libbamgpu.so $(LIBDIR)/libbamharness.so $(LIBDIR)/libbamgen.so $(LIBDIR)/bamgen oracle

build/%.cu.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

build/%.cpp.o: $(CSRC)/%.cpp $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -x cu -c $< -o $@

$(LIBDIR)/libbamgpu.so: $(CU_OBJS) $(CPP_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) -Wno-deprecated-gpu-targets -shared -o $@ $^ -lpthread

# bench.py's per-record consumer loop (what the Rust shim's next_rec() does), linked against the product library
$(LIBDIR)/libbamharness.so: tools/next_rec_harness.c $(LIBDIR)/libbamgpu.so include/bamgpu.h
	$(CC) -O2 -fPIC -Wall -shared -Iinclude -o $@ $< -L$(LIBDIR) -lbamgpu -Wl,-rpath,'$$ORIGIN'

$(LIBDIR)/libbamgen.so: tools/bamgen.cpp
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -shared -o $@ $< -lz

$(LIBDIR)/bamgen: tools/bamgen.cpp
	@mkdir -p $(LIBDIR)
	$(CXX) $(CXXFLAGS) -DBAMGEN_MAIN -o $@ $< -lz

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIBDIR) oracle/_build

.PHONY: all oracle clean

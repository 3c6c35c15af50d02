# Build of the B200 demultiplex path.  Everything lands in-tree (git-ignored,
# but shipped to the GPU box):
#   mgikit_b200/lib/libmgikit_gpu.so   CUDA kernels + C ABI (include/mgikit_gpu.h)
#   mgikit_b200/bin/mgikit             the product CLI (host pipeline + CUDA path)
#   oracle/_ref/liboracle.so           CPU oracle (tests only)
#   oracle/_ref/mgikit_oracle          host pipeline bound to the oracle (tests only)
#   oracle/_ref/mgikit                 the reference's own prebuilt binary, when /root/reference exists
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
CC        ?= gcc
CUDA_HOME ?= /usr/local/cuda
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wextra -Xptxas -v
CXXFLAGS  := -O2 -std=c++17 -Wall -Wextra -pthread -fPIC
H         := mgikit_b200/csrc/host
G         := mgikit_b200/csrc/gpu
HOST_SRC  := $(H)/cli.cpp $(H)/pipeline.cpp $(H)/report.cpp $(H)/run_info.cpp $(H)/sample_sheet.cpp $(H)/log.cpp $(H)/template_cmd.cpp $(H)/inflate.cpp $(H)/gz_reader.cpp
HOST_HDR  := $(wildcard $(H)/*.hpp) include/mgikit_gpu.h
GPU_SRC   := $(wildcard $(G)/*.cu)
GPU_HDR   := $(wildcard $(G)/*.cuh) include/mgikit_gpu.h
REF_BIN   := /root/reference/bins/mgikit

all: gpu host oracle

gpu: mgikit_b200/lib/libmgikit_gpu.so
host: mgikit_b200/bin/mgikit
oracle: oracle/_ref/liboracle.so oracle/_ref/mgikit_oracle oracle/_ref/mgikit oracle/_ref/libemu.so oracle/_ref/mgikit_emu oracle/_ref/libhostcodec.so

mgikit_b200/lib/libmgikit_gpu.so: $(GPU_SRC) $(GPU_HDR)
	@mkdir -p mgikit_b200/lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(GPU_SRC) -lcudart -ldl 2> mgikit_b200/lib/ptxas.log || (cat mgikit_b200/lib/ptxas.log; false)

mgikit_b200/bin/mgikit: $(HOST_SRC) $(H)/main_gpu.cpp $(HOST_HDR) mgikit_b200/lib/libmgikit_gpu.so
	@mkdir -p mgikit_b200/bin
	$(CXX) $(CXXFLAGS) -o $@ $(H)/main_gpu.cpp $(HOST_SRC) -Lmgikit_b200/lib -lmgikit_gpu -lz \
	    -Wl,-rpath,'$$ORIGIN/../lib' -Wl,-rpath,$(CUDA_HOME)/lib64

oracle/_ref/liboracle.so: oracle/mgikit_oracle.c oracle/mgikit_oracle.h include/mgikit_gpu.h
	@mkdir -p oracle/_ref
	$(CC) -O2 -std=c11 -Wall -Wextra -fPIC -shared -o $@ oracle/mgikit_oracle.c

oracle/_ref/mgikit_oracle.o: oracle/mgikit_oracle.c oracle/mgikit_oracle.h include/mgikit_gpu.h
	@mkdir -p oracle/_ref
	$(CC) -O2 -std=c11 -Wall -Wextra -fPIC -c -o $@ oracle/mgikit_oracle.c

oracle/_ref/mgikit_oracle: oracle/oracle_cli.cpp oracle/_ref/mgikit_oracle.o $(HOST_SRC) $(HOST_HDR)
	$(CXX) $(CXXFLAGS) -o $@ oracle/oracle_cli.cpp $(HOST_SRC) oracle/_ref/mgikit_oracle.o -lz

# CPU emulation of the device logic (tests only): device_logic.cuh compiled by g++
EMU_DEP := tests/emu/emu.cpp mgikit_b200/csrc/gpu/device_logic.cuh mgikit_b200/csrc/gpu/emit_logic.cuh mgikit_b200/csrc/gpu/params_build.hpp include/mgikit_gpu.h
oracle/_ref/libemu.so: $(EMU_DEP)
	@mkdir -p oracle/_ref
	$(CXX) $(CXXFLAGS) -shared -o $@ tests/emu/emu.cpp

oracle/_ref/mgikit_emu: tests/emu/emu_cli.cpp $(EMU_DEP) $(HOST_SRC) $(HOST_HDR)
	$(CXX) $(CXXFLAGS) -o $@ tests/emu/emu_cli.cpp tests/emu/emu.cpp $(HOST_SRC) -lz

# C hooks over the product's host codec (inflate, member-parallel reader) for the CPU tests
oracle/_ref/libhostcodec.so: tests/emu/hostcodec.cpp $(H)/inflate.cpp $(H)/inflate.hpp $(H)/gz_reader.cpp $(H)/gz_reader.hpp
	@mkdir -p oracle/_ref
	$(CXX) $(CXXFLAGS) -shared -o $@ tests/emu/hostcodec.cpp $(H)/inflate.cpp $(H)/gz_reader.cpp

# The reference cannot be compiled here (no cargo/rustc, no vendored crates);
# its own prebuilt, version-matched binary is the strongest oracle we can get.
oracle/_ref/mgikit:
	@mkdir -p oracle/_ref
	@if [ -x $(REF_BIN) ]; then cp $(REF_BIN) $@; echo "copied reference binary"; \
	 else echo "reference binary not available here (GPU box): skipping"; fi

clean:
	rm -rf mgikit_b200/lib mgikit_b200/bin oracle/_ref/liboracle.so oracle/_ref/libemu.so oracle/_ref/mgikit_oracle oracle/_ref/mgikit_emu oracle/_ref/libhostcodec.so oracle/_ref/*.o

.PHONY: all gpu host oracle clean

# Builds the sm_100a shared library (C ABI) and the host CLI, in-tree.
#   make            -> miekki_b200/libmiekki_b200.so , miekki_b200/cli/miekki
#   make oracle     -> oracle/_build/libmiekki_oracle.so (+ oracle/_ref/Miekki when the reference is present)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX := $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -ccbin $(CXX) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wextra,-Wno-unused-parameter -Xptxas -v
SRC := miekki_b200/csrc
OBJ := build/obj
LIB := miekki_b200/libmiekki_b200.so
CLI := miekki_b200/cli/miekki
CU := $(SRC)/api.cu $(SRC)/sketch.cu $(SRC)/scan.cu $(SRC)/scan_tiled.cu $(SRC)/topk.cu $(SRC)/exact.cu
OBJS := $(patsubst $(SRC)/%.cu,$(OBJ)/%.o,$(CU)) $(OBJ)/pack.o
HDRS := $(SRC)/common.cuh $(SRC)/scan_common.cuh $(SRC)/kernels.h $(SRC)/pack.h include/miekki_b200.h

TOOLS := benchmarks/make_dump

all: $(LIB) $(CLI) $(TOOLS)

# bench tooling (see the header of the source): index of a synthetic configuration as a dump file
benchmarks/make_dump: benchmarks/make_dump.cpp include/miekki_b200.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -Wextra -Iinclude -o $@ benchmarks/make_dump.cpp \
	    -Lmiekki_b200 -lmiekki_b200 -Wl,-rpath,'$$ORIGIN/../miekki_b200'

$(OBJ)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; false)

# host-side sequence packer: plain C++ (its AVX2 kernel is selected at run time)
$(OBJ)/pack.o: $(SRC)/pack.cpp $(SRC)/pack.h
	@mkdir -p $(OBJ)
	$(CXX) -O3 -std=c++17 -fPIC -Wall -Wextra -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart static -ccbin $(CXX) -o $@ $(OBJS) -lpthread

$(CLI): miekki_b200/cli/miekki_cli.cpp miekki_b200/cli/fasta.hpp include/miekki_b200.h $(LIB)
	$(CXX) -O2 -std=c++17 -Wall -Wextra -fopenmp -Iinclude -o $@ miekki_b200/cli/miekki_cli.cpp \
	    -Lmiekki_b200 -lmiekki_b200 -lz -Wl,-rpath,'$$ORIGIN/..'

oracle:
	$(MAKE) -C oracle
	@if [ -d /root/reference ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -rf build $(LIB) $(CLI) $(TOOLS)

.PHONY: all oracle clean

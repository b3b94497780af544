# Top-level build: everything is compiled in-tree for sm_100a only.
#   make lib      -> kmer_id_b200/libkmerid_b200.so   (CUDA kernels + C-ABI, include/kmer_id.h)
#   make host     -> kmer_id_b200/bin/nk10            (C++ drop-in for the reference's ./nk10 <dir/>)
#   make tools    -> tools/libkidsynth.so, tools/kid_synth (synthetic DB / read generators)
#   make oracle   -> oracle/liboracle.so (+ oracle/_ref/* when /root/reference is present)
NVCC      ?= nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall $(if $(TUNE),-DKID_TUNE_VARIANTS) $(if $(THREADS),-DKID_CLASSIFY_THREADS=$(THREADS)) $(if $(MINBLOCKS),-DKID_CLASSIFY_MINBLOCKS=$(MINBLOCKS))
CXXFLAGS  := -O3 -std=c++17 -Wall -Wextra -fPIC -pthread
CUDA_HOME ?= /usr/local/cuda

CSRC := kmer_id_b200/csrc
LIB  := kmer_id_b200/libkmerid_b200.so
LIB_OBJS := $(CSRC)/kid_api.o $(CSRC)/kid_classify.o $(CSRC)/kid_classify3.o $(CSRC)/kid_pack.o $(CSRC)/kid_pack_host.o $(CSRC)/kid_build_sorted.o $(CSRC)/kid_sample.o $(CSRC)/kid_ingest.o

all: lib tools host oracle

lib: $(LIB)

$(CSRC)/%.o: $(CSRC)/%.cu $(CSRC)/kid_common.cuh $(CSRC)/kid_table2.cuh $(CSRC)/kid_kernels.cuh $(CSRC)/kid_readprep.cuh $(CSRC)/kid_internal.cuh $(CSRC)/kid_inflate.cuh $(CSRC)/kid_inflate_chain.hpp include/kmer_id.h
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(CSRC)/%.o: $(CSRC)/%.cpp include/kmer_id.h
	$(CXX) $(CXXFLAGS) -c $< -o $@

$(LIB): $(LIB_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(LIB_OBJS)

host: lib
	@if [ -f kmer_id_b200/host/Makefile ]; then $(MAKE) --no-print-directory -C kmer_id_b200/host; fi

tools: lib
	@if [ -f tools/Makefile ]; then $(MAKE) --no-print-directory -C tools; fi

oracle:
	$(MAKE) --no-print-directory -C oracle all

clean:
	rm -f $(LIB_OBJS) $(LIB)
	-$(MAKE) -C oracle clean
.PHONY: all lib host tools oracle clean

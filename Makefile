# Builds libgminer_b200.so (CUDA kernels + C ABI) for sm_100a, the CLI drop-ins, and the CPU oracle.
NVCC    ?= /usr/local/cuda/bin/nvcc
CCBIN   ?= /usr/bin/g++
ARCH     = -gencode arch=compute_100a,code=sm_100a
NVFLAGS  = $(ARCH) -O3 -std=c++17 -lineinfo -ccbin $(CCBIN) -Xcompiler -fPIC,-fopenmp,-Wall,-Wno-unknown-pragmas \
           --expt-relaxed-constexpr --expt-extended-lambda -Iinclude $(EXTRA_NVFLAGS)
CSRC     = graphminer_b200/csrc
LIB      = graphminer_b200/libgminer_b200.so
CU_SRCS  = $(CSRC)/graph.cu $(CSRC)/rank.cu $(CSRC)/tc.cu $(CSRC)/batch.cu $(CSRC)/patterns.cu $(CSRC)/clique_bitmap.cu $(CSRC)/support.cu $(CSRC)/cycle4.cu $(CSRC)/solvers.cu $(CSRC)/gen.cu
CC_SRCS  = $(CSRC)/host_graph.cc
OBJS     = $(CU_SRCS:.cu=.o) $(CC_SRCS:.cc=.o)
HDRS     = $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h include/gm/*.cuh)
APPS     = tc_gpu_base tc_multigpu clique_gpu_base kcl_gpu_base clique_multigpu sgl_gpu_base sgl_multigpu \
           motif_gpu_base motif_gpu_formula motif_multigpu

all: $(LIB) apps oracle

$(CSRC)/%.o: $(CSRC)/%.cu $(HDRS)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(CSRC)/%.o: $(CSRC)/%.cc $(HDRS)
	$(CCBIN) -O3 -std=c++17 -fPIC -fopenmp -Wall -Iinclude -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -ccbin $(CCBIN) -o $@ $(OBJS) -lgomp -ldl

apps: $(addprefix bin/,$(APPS)) bin/gm_ops_selftest
bin/%: $(CSRC)/apps/%.cc $(CSRC)/apps/app_common.h $(CSRC)/solvers.h $(LIB)
	@mkdir -p bin
	$(CCBIN) -O2 -std=c++17 -fopenmp -Iinclude -I$(CSRC) $< -o $@ -Lgraphminer_b200 -lgminer_b200 -Wl,-rpath,'$$ORIGIN/../graphminer_b200'
bin/kcl_gpu_base: bin/clique_gpu_base
	cp $< $@
# the header-only device operator API used from a user's own kernels (no library needed)
bin/gm_ops_selftest: $(CSRC)/apps/gm_ops_selftest.cu $(HDRS)
	@mkdir -p bin
	$(NVCC) $(ARCH) -O2 -std=c++17 -lineinfo -ccbin $(CCBIN) -Iinclude $< -o $@

oracle:
	$(MAKE) -C oracle all

clean:
	rm -f $(OBJS) $(LIB); rm -rf bin
.PHONY: all apps oracle clean

# Builds the C-ABI CUDA library (sm_100a only) and, with `make host`, the C++
# host mirror of folve's SoundProcessor / zita-config / ProcessorPool on top of it.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  ?= g++
ARCH  = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-Wall -Xptxas -v $(EXTRA_NVFLAGS)
CSRC = folve_b200/csrc
LIB  = folve_b200/libfolve_b200.so

all: $(LIB)

# one object per kernel family (they compile in parallel: `make -j`), ptxas -v output kept per object
CU_SRCS = fcv_engine fcv_nonuniform fcv_k_fft fcv_k_fft13 fcv_k_mac fcv_k_mac_tma fcv_k_fused13
CU_OBJS = $(CU_SRCS:%=$(CSRC)/%.o)
CU_HDRS = $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/folve_b200.h

$(CSRC)/%.o: $(CSRC)/%.cu $(CU_HDRS)
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> $(CSRC)/$*.ptxas.log || (cat $(CSRC)/$*.ptxas.log; exit 1)
	@grep -E "error|warning" $(CSRC)/$*.ptxas.log || true

$(LIB): $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(CU_OBJS)
	@cat $(CU_SRCS:%=$(CSRC)/%.ptxas.log) > $(CSRC)/ptxas.log

clean:
	rm -f $(LIB) $(CU_OBJS) $(CSRC)/*ptxas.log

.PHONY: all clean

# ---- C++ host layer: SoundProcessor / filter-config / ProcessorPool + test harness
HOST = folve_b200/host
HOSTLIB = folve_b200/libfolve_host.so
HOST_SRCS = $(HOST)/sound-processor.cc $(HOST)/filter-config.cc $(HOST)/processor-pool.cc \
            $(HOST)/batch-convolver.cc \
            $(HOST)/harness.cc $(HOST)/sndfile_shim/sndfile_shim.cc
HOST_HDRS = $(HOST)/sound-processor.h $(HOST)/filter-config.h $(HOST)/processor-pool.h \
            $(HOST)/batch-convolver.h \
            $(HOST)/sndfile_shim/sndfile.h include/folve_b200.h
# same optimisation flags as oracle/Makefile so that float expressions in the
# config loader (hilbert taps, gains) round identically on both sides
HOST_FLAGS = -O3 -march=x86-64-v3 -fno-math-errno -fno-trapping-math -std=c++17 -fPIC -Wall -Wextra \
             -I$(HOST)/sndfile_shim -I$(HOST) -Iinclude

host: $(HOSTLIB)

$(HOSTLIB): $(HOST_SRCS) $(HOST_HDRS) $(LIB)
	$(CXX) $(HOST_FLAGS) -shared -Wl,-Bsymbolic -o $@ $(HOST_SRCS) -L folve_b200 -lfolve_b200 -Wl,-rpath,'$$ORIGIN' -lpthread

all: host

# Builds the C-ABI CUDA library (sm_100a only) and, with `make host`, the C++
# host mirror of folve's SoundProcessor / zita-config / ProcessorPool on top of it.
NVCC ?= /usr/local/cuda/bin/nvcc
CXX  ?= g++
ARCH  = -gencode arch=compute_100a,code=sm_100a
NVFLAGS = $(ARCH) -O3 -lineinfo -std=c++17 --expt-relaxed-constexpr -Xcompiler -fPIC,-Wall -Xptxas -v
CSRC = folve_b200/csrc
LIB  = folve_b200/libfolve_b200.so

all: $(LIB)

$(LIB): $(CSRC)/fcv_engine.cu $(CSRC)/fcv_fft.cuh $(CSRC)/fcv_mac.cuh include/folve_b200.h
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CSRC)/fcv_engine.cu 2> $(CSRC)/ptxas.log || (cat $(CSRC)/ptxas.log; exit 1)
	@grep -E "error|warning" $(CSRC)/ptxas.log || true

clean:
	rm -f $(LIB) $(CSRC)/ptxas.log

.PHONY: all clean

# Builds libsolarosc_b200.so, the C-ABI library a host (Julia via ccall, Python via ctypes, C++) binds.
# In a SolarOscFitter checkout this target sits next to the existing libnufast_earth.so rule
# (reference Makefile:18,38-54) and writes src/oscillations/libsolarosc_b200.so; see INTEGRATION.md.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS ?= -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-O3 -Xptxas -v
PKG       := dune_solaroscfitter_b200
SRC       := $(PKG)/csrc/solarosc_b200.cu
HDR       := $(PKG)/csrc/solarosc_device.cuh $(PKG)/csrc/solarosc_kernels.cuh $(PKG)/csrc/solarosc_pipe.cuh $(PKG)/csrc/solarosc_grad.cuh $(PKG)/csrc/solarosc_adjoint.cuh $(PKG)/csrc/solarosc_adjoint_math.cuh $(PKG)/csrc/solarosc_general.cuh $(PKG)/csrc/solarosc_gengrad.cuh $(PKG)/csrc/solarosc_setup.cuh $(PKG)/csrc/solarosc_batcher.hpp $(PKG)/csrc/solarosc_comm.hpp include/solarosc_b200.h
LIB       := $(PKG)/libsolarosc_b200.so

all: $(LIB) oracle

$(LIB): $(SRC) $(HDR)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(SRC) -lcudart -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean

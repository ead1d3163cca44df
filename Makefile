# Builds the B200 C-ABI library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC      ?= /usr/local/cuda/bin/nvcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr $(EXTRA_NVFLAGS)
CSRC      := aether_primitives_b200/csrc
LIBDIR    := aether_primitives_b200/lib
OBJDIR    := build/obj
SRCS      := api elementwise fft fir chain chain_x2 spectral
OBJS      := $(addprefix $(OBJDIR)/,$(addsuffix .o,$(SRCS)))
LIB       := $(LIBDIR)/libaether_b200.so
ORACLE    := oracle/liboracle.so

all: $(LIB) $(ORACLE)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/aether_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(OBJS)

# -ffp-contract=off: like rustc, never fuse a*b+c.  x86-64-v3 so the .so also runs on the GPU box's host.
$(ORACLE): oracle/aether_oracle.cpp
	$(CXX) -O3 -march=x86-64-v3 -ffp-contract=off -fno-fast-math -std=c++17 -fPIC -shared -pthread -o $@ $<

oracle: $(ORACLE)
lib: $(LIB)

clean:
	rm -rf build $(LIB) $(ORACLE)

.PHONY: all clean oracle lib

# C++ host-mirror test (links the C-ABI library only; run it on a GPU box)
CPPTEST := build/host_mirror_test
$(CPPTEST): tests/cpp/host_mirror_test.cpp include/aether_b200.hpp include/aether_b200.h $(LIB)
	@mkdir -p build
	$(CXX) -std=c++17 -O2 -Iinclude $< -o $@ -L$(LIBDIR) -laether_b200 -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'
cpptest: $(CPPTEST)
.PHONY: cpptest

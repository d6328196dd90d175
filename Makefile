# Builds the product library (sm_100a only) and the CPU oracle (test infrastructure).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v --expt-relaxed-constexpr
CSRC      := vnext_b200/csrc
LIBDIR    := vnext_b200/lib
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,build/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) include/vnext_b200.h

all: $(LIBDIR)/libvnext_b200.so oracle

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIBDIR)/libvnext_b200.so: $(OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

oracle: oracle/_build/libmsda_oracle.so

oracle/_build/libmsda_oracle.so: oracle/msda_oracle.c
	@mkdir -p oracle/_build
	gcc -O2 -ffp-contract=off -fPIC -shared -o $@ $< -lm

clean:
	rm -rf build $(LIBDIR)/*.so oracle/_build

.PHONY: all oracle clean

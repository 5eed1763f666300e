# Builds the product library haslr_b200/libhaslr_b200.so (CUDA, sm_100a only) in-tree.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC     := haslr_b200/csrc
OBJDIR   := build/obj
LIB      := haslr_b200/libhaslr_b200.so
TUS      := api poa k12
HDRS     := $(wildcard $(CSRC)/*.cuh) include/haslr_b200.h

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(foreach t,$(TUS),$(OBJDIR)/$(t).o)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart

clean:
	rm -rf build $(LIB)

.PHONY: all clean

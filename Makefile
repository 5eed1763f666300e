# Builds the product library haslr_b200/libhaslr_b200.so (CUDA, sm_100a only) in-tree.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC     := haslr_b200/csrc
OBJDIR   := build/obj
LIB      := haslr_b200/libhaslr_b200.so
TUS      := api poa k12 coords paf
HDRS     := $(wildcard $(CSRC)/*.cuh) include/haslr_b200.h

HOSTSRC  := $(wildcard haslr_b200/host/*.cpp)
BIN      := bin/haslr_assemble
PATHLIB  := haslr_b200/libhaslr_path.so
GEN      := bin/gen_synth

all: $(LIB) $(BIN) $(PATHLIB) $(GEN)

# the whole path behind one C call (include/haslr_path.h): the binary's host code without main()
$(PATHLIB): $(filter-out haslr_b200/host/main.cpp,$(HOSTSRC)) haslr_b200/host/haslr.hpp include/haslr_path.h $(LIB)
	g++ -std=c++17 -O2 -Wall -fPIC -shared -o $@ $(filter-out haslr_b200/host/main.cpp,$(HOSTSRC)) -Lhaslr_b200 -lhaslr_b200 -Wl,-rpath,'$$ORIGIN' -lz -lpthread

# seeded synthetic datasets (SURVEY 8d generator: genome, SRC contigs, long reads, exact PAF)
$(GEN): tools/gen_synth.cpp
	@mkdir -p bin
	g++ -std=c++17 -O2 -Wall -o $@ $<

# the drop-in binary: C++ host code above the C ABI
$(BIN): $(HOSTSRC) haslr_b200/host/haslr.hpp include/haslr_path.h $(LIB)
	@mkdir -p bin
	g++ -std=c++17 -O2 -Wall -o $@ $(HOSTSRC) -Lhaslr_b200 -lhaslr_b200 -Wl,-rpath,'$$ORIGIN/../haslr_b200' -lz -lpthread

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(foreach t,$(TUS),$(OBJDIR)/$(t).o)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcudart

clean:
	rm -rf build $(LIB) $(BIN) $(PATHLIB) $(GEN)

.PHONY: all clean

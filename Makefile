# Top-level build.  Knob names follow the reference Makefile (/root/reference/Makefile:24-41):
#   STATES=DNA  NUM_ACCELERATORS=9  INPUT_SRC=mem|gen  PLIO_LAYOUT=Comb|Sep  WINDOW_SIZE=8192
#   NO_PRERUN_CHECK / NO_CORRECTNESS_CHECK / NO_INTERMEDIATE_RESULTS  (Makefile:145-161)
# and its run-time parameters DEVICE, ALIGNMENTS, PLF_CALLS, INSTANCES_USED (Makefile:14-18).
#
#   make lib      libb200plf.so            (CUDA kernels + C ABI, sm_100a only)
#   make host     host_mem.exe host_gen.exe (C++ host drop-ins on top of the C ABI)
#   make oracle   oracle/liboracle.so (+ oracle/_ref when /root/reference is present)
#   make run      ./host_<INPUT_SRC>.exe <config> <DEVICE> <ALIGNMENTS> <PLF_CALLS> <INSTANCES_USED>

PKG      := amd-versal-phylogenetic-likelihood-function_b200
NVCC     ?= nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC
CUDA_HOME ?= /usr/local/cuda

STATES ?= DNA
AIE_TYPE ?= window
WINDOW_SIZE ?= 8192
PLIO_LAYOUT ?= Comb
NUM_ACCELERATORS ?= 9
INPUT_SRC ?= mem

DEVICE ?= 0
ALIGNMENTS ?= 100
PLF_CALLS ?= 1
INSTANCES_USED ?= 1

NO_PRERUN_CHECK ?= 1
NO_CORRECTNESS_CHECK ?= 0
NO_INTERMEDIATE_RESULTS ?= 0

ifeq ($(AIE_TYPE), stream)
AIE_COMS_METHOD := $(AIE_TYPE)
else
AIE_COMS_METHOD := $(AIE_TYPE)$(WINDOW_SIZE)
endif
# Same artefact naming as the reference (Makefile:37-39); the host parses the knobs back out of it.
CONFIG := plf_128x$(NUM_ACCELERATORS)$(STATES)$(AIE_COMS_METHOD)$(PLIO_LAYOUT)_$(INPUT_SRC)$(STATES)$(AIE_TYPE)$(PLIO_LAYOUT)

LIB  := $(PKG)/libb200plf.so
HOSTDEFS := -DNO_PRERUN_CHECK=$(NO_PRERUN_CHECK) -DNO_CORRECTNESS_CHECK=$(NO_CORRECTNESS_CHECK) \
            -DNO_INTERMEDIATE_RESULTS=$(NO_INTERMEDIATE_RESULTS)
HOSTFLAGS := -O2 -g -Wall -std=c++17 -ffp-contract=off -pthread -Iinclude $(HOSTDEFS)

all: lib host oracle

lib:
	@$(MAKE) --no-print-directory -j8 $(LIB)

CSRC := $(PKG)/csrc
OBJDIR := $(PKG)/build
KHDRS := $(CSRC)/plf_kernels.cuh $(CSRC)/plf_registry.h include/b200plf.h
OBJS := $(OBJDIR)/plf_capi.o $(OBJDIR)/plf_tree.o $(OBJDIR)/plf_evaluate.o $(OBJDIR)/plf_protein.o $(OBJDIR)/plf_protein_tc.o $(OBJDIR)/plf_multi.o $(OBJDIR)/sel_ldg_strict.o $(OBJDIR)/sel_ldg_fma.o \
        $(OBJDIR)/sel_tma_strict.o $(OBJDIR)/sel_tma_fma.o $(OBJDIR)/sel_dyn_strict.o $(OBJDIR)/sel_dyn_fma.o

$(OBJDIR)/plf_capi.o: $(CSRC)/plf_capi.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(OBJDIR)/plf_tree.o: $(CSRC)/plf_tree.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(OBJDIR)/plf_evaluate.o: $(CSRC)/plf_evaluate.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

$(OBJDIR)/plf_protein.o: $(CSRC)/plf_protein.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
$(OBJDIR)/plf_protein_tc.o: $(CSRC)/plf_protein_tc.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<
# NCCL: header only at build time (types and prototypes); the symbols are resolved with dlopen at run time
$(OBJDIR)/plf_multi.o: $(CSRC)/plf_multi.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c -o $@ $<

# the kernel instantiations: one translation unit per (kernel family, arithmetic mode)
$(OBJDIR)/sel_ldg_strict.o: $(CSRC)/plf_sel_ldg.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_MATH=MathStrict -DPLF_SEL_NAME=select_ldg_strict -c -o $@ $<
$(OBJDIR)/sel_ldg_fma.o: $(CSRC)/plf_sel_ldg.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_MATH=MathFma -DPLF_SEL_NAME=select_ldg_fma -c -o $@ $<
$(OBJDIR)/sel_tma_strict.o: $(CSRC)/plf_sel_tma.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_MATH=MathStrict -DPLF_SEL_NAME=select_tma_strict -c -o $@ $<
$(OBJDIR)/sel_tma_fma.o: $(CSRC)/plf_sel_tma.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_MATH=MathFma -DPLF_SEL_NAME=select_tma_fma -c -o $@ $<

$(OBJDIR)/sel_dyn_strict.o: $(CSRC)/plf_sel_tma.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_DYNAMIC -DPLF_SEL_MATH=MathStrict -DPLF_SEL_NAME=select_tma_dyn_strict -c -o $@ $<
$(OBJDIR)/sel_dyn_fma.o: $(CSRC)/plf_sel_tma.cu $(KHDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -DPLF_SEL_DYNAMIC -DPLF_SEL_MATH=MathFma -DPLF_SEL_NAME=select_tma_dyn_fma -c -o $@ $<

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -ldl

host: $(PKG)/host_mem.exe $(PKG)/host_gen.exe $(PKG)/host_stream.exe $(PKG)/host_states.exe

$(PKG)/host_mem.exe: $(PKG)/host/host_mem.cpp $(PKG)/host/golden_plf.cpp $(wildcard $(PKG)/host/*.h) $(LIB)
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/host_mem.cpp $(PKG)/host/golden_plf.cpp \
	    -L$(PKG) -lb200plf -Wl,-rpath,'$$ORIGIN'

$(PKG)/host_gen.exe: $(PKG)/host/host_gen.cpp $(wildcard $(PKG)/host/*.h) $(LIB)
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/host_gen.cpp -L$(PKG) -lb200plf -Wl,-rpath,'$$ORIGIN'

$(PKG)/host_stream.exe: $(PKG)/host/host_stream.cpp $(PKG)/host/golden_plf.cpp $(wildcard $(PKG)/host/*.h) $(LIB)
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/host_stream.cpp $(PKG)/host/golden_plf.cpp \
	    -L$(PKG) -lb200plf -Wl,-rpath,'$$ORIGIN'

$(PKG)/host_states.exe: $(PKG)/host/host_states.cpp $(PKG)/host/golden_plf.cpp $(wildcard $(PKG)/host/*.h) $(LIB)
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/host_states.cpp $(PKG)/host/golden_plf.cpp \
	    -L$(PKG) -lb200plf -Wl,-rpath,'$$ORIGIN'

oracle:
	$(MAKE) -C oracle all

run: host
	./$(PKG)/host_$(INPUT_SRC).exe $(CONFIG) $(DEVICE) $(ALIGNMENTS) $(PLF_CALLS) $(INSTANCES_USED)

clean:
	rm -rf $(LIB) $(PKG)/build $(PKG)/host_mem.exe $(PKG)/host_gen.exe $(PKG)/host_stream.exe $(PKG)/host_states.exe
	$(MAKE) -C oracle clean

.PHONY: all lib host oracle run clean

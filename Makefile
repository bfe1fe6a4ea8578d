# Builds the CUDA library for sm_100a in-tree (the .so travels to the GPU box with the snapshot).
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -Wno-deprecated-gpu-targets
CSRC := ectrans_b200/csrc
OBJ := $(CSRC)/_build
LIB := ectrans_b200/lib/libectrans_b200.so
SRCS := api.cu legendre.cu legendre_tc.cu fourier.cu fft_plan.cu host_plan.cu gp_partition.cu transi.cu
OBJS := $(patsubst %.cu,$(OBJ)/%.o,$(SRCS))
HDRS := $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)

all: $(LIB)

$(OBJ)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJ)
	$(NVCC) $(NVFLAGS) $(EXTRA_$*) -Xptxas -v -c $< -o $@ 2> $(OBJ)/$*.ptxas.log || (cat $(OBJ)/$*.ptxas.log; exit 1)

# the table kernel is kept free of FMA contraction so that it reproduces the reference recurrence bit for bit
EXTRA_legendre :=

$(LIB): $(OBJS)
	@mkdir -p ectrans_b200/lib
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS) -ldl

emu:
	@mkdir -p tests/hostemu/_build
	$(NVCC) -std=c++17 -O2 -shared -Xcompiler -fPIC -Wno-deprecated-gpu-targets -o tests/hostemu/_build/libemu.so tests/hostemu/emu.cu $(CSRC)/fft_plan.cu

# compile check of the round-2 tcgen05 probe (tools/probes; not part of the library)
probe:
	@mkdir -p tools/probes/_build
	$(NVCC) -O2 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC -shared -Wno-deprecated-gpu-targets -o tools/probes/_build/libtcprobe.so tools/probes/tcgen05_tf32_probe.cu

# The reference's own transi test program (tests/transi/transi_test_program.c), compiled UNCHANGED from where it lies
# against include/ectrans/transi.h and linked with the CUDA library.  Only possible where /root/reference exists (the
# build container); the binary lands in oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).
REF ?= /root/reference
transi_ref:
	@mkdir -p oracle/_ref
	@if [ -f $(REF)/tests/transi/transi_test_program.c ]; then \
	  gcc -O1 -std=gnu99 -Iinclude -I$(REF)/tests/transi -o oracle/_ref/transi_test_program \
	    $(REF)/tests/transi/transi_test_program.c $(REF)/tests/transi/transi_test.c \
	    -Lectrans_b200/lib -lectrans_b200 -lm -Wl,-rpath,'$$ORIGIN/../../ectrans_b200/lib' && echo "built oracle/_ref/transi_test_program"; \
	else echo "reference tree not present: oracle/_ref/transi_test_program not rebuilt"; fi

clean:
	rm -rf $(OBJ) $(LIB) tests/hostemu/_build

# Builds libsonic_b200.so (CUDA, sm_100a only) in-tree, plus the CPU oracle's C restatement.
NVCC ?= nvcc
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v $(EXTRA)
CSRC := sonic_b200/csrc
OBJS := $(CSRC)/build/capi.o $(CSRC)/build/msm.o $(CSRC)/build/msm_acc_compact.o $(CSRC)/build/msm_acc_regs.o $(CSRC)/build/msm_acc_affine.o $(CSRC)/build/msm_reduce.o $(CSRC)/build/srs.o $(CSRC)/build/poly.o $(CSRC)/build/prove.o $(CSRC)/build/selftest.o $(CSRC)/build/g2srs.o
HDRS := $(wildcard $(CSRC)/*.cuh) $(CSRC)/internal.h include/sonic_b200.h

all: sonic_b200/libsonic_b200.so oracle

sonic_b200/libsonic_b200.so: $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart -lnccl -ldl -lpthread

$(CSRC)/build/g2srs.o: $(CSRC)/g2srs.cu $(HDRS)
	@mkdir -p $(CSRC)/build
	$(NVCC) $(NVCCFLAGS) -Xptxas -O1 -c -o $@ $< 2> $(CSRC)/build/g2srs.ptxas.log || (cat $(CSRC)/build/g2srs.ptxas.log; false)

$(CSRC)/build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(CSRC)/build
	$(NVCC) $(NVCCFLAGS) -c -o $@ $< 2> $(CSRC)/build/$*.ptxas.log || (cat $(CSRC)/build/$*.ptxas.log; false)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(CSRC)/build sonic_b200/libsonic_b200.so

.PHONY: all oracle clean

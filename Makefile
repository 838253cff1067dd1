# Builds libbellman_b200.so (sm_100a) in-tree.  `make -j8` compiles the translation units in
# parallel; the G2 unit dominates (~3.5 min).
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xptxas -v
SRC       := bellman_mpc_b200/csrc
OBJDIR    := build/obj
UNITS     := api multi ntt msm_sort group_g1 group_g2 prove r1cs
OBJS      := $(UNITS:%=$(OBJDIR)/%.o)
LIB       := bellman_mpc_b200/libbellman_b200.so
HDRS      := $(wildcard $(SRC)/*.cuh) $(SRC)/internal.h include/bellman_b200.h

all: $(LIB)

$(OBJDIR)/%.o: $(SRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR) build/ptxas
	$(NVCC) $(NVFLAGS) -c $< -o $@ 2> build/ptxas/$*.log || (cat build/ptxas/$*.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf build/obj build/ptxas $(LIB)

.PHONY: all clean

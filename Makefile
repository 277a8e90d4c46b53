# Builds the product artefacts without Python (the same two commands megalania_b200/build.py runs):
#   make            ->  megalania_b200/_build/libmegalania_cuda.so  (CUDA kernels + C ABI, sm_100a only)
#                       megalania_b200/_build/megalania              (drop-in C CLI)
#   make oracle     ->  the CPU checker (test infrastructure only; needs /root/reference for oracle/_ref)
NVCC ?= nvcc
CC   ?= gcc
OUT  := megalania_b200/_build
LIB  := $(OUT)/libmegalania_cuda.so
CLI  := $(OUT)/megalania
CSRC := megalania_b200/csrc
HOST := megalania_b200/host

all: $(LIB) $(CLI)

$(LIB): $(CSRC)/mg_api.cu $(CSRC)/mg_device.cuh $(CSRC)/mg_finder.cuh $(CSRC)/mg_kernels.cuh $(CSRC)/mg_comm.inc include/megalania_cuda.h
	mkdir -p $(OUT)
	$(NVCC) -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -ldl -o $@ $<

$(CLI): $(HOST)/main.c $(HOST)/host_io.c $(LIB)
	$(CC) -O2 -std=gnu11 -Wall -Wextra -Iinclude -o $@ $(HOST)/main.c $(HOST)/host_io.c -L$(OUT) -lmegalania_cuda -Wl,-rpath,'$$ORIGIN' -lm -lpthread

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf $(OUT)

.PHONY: all oracle clean

# Builds the product library and the test infrastructure.
#
#   make lib      chemtensor_b200/libchemtensor_b200.so : host C (chemtensor_b200/host) + sm_100a CUDA layer (chemtensor_b200/csrc)
#   make emu      tests/emu/libctb_hostlogic_emu.so     : the same host C linked against the CPU TEST DOUBLE of the CUDA
#                                                         layer (tests/emu/ctbd_emu.c) -- test infrastructure only
#   make oracle   oracle/_ref/libchemtensor_ref.so      : the unmodified reference, only when /root/reference exists
#
# Built artefacts are git-ignored but travel to the GPU box with the working tree.

NVCC     ?= /usr/local/cuda/bin/nvcc
CC       ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fopenmp -Iinclude -Ichemtensor_b200/csrc --expt-relaxed-constexpr
CFLAGS   := -std=gnu11 -O2 -fPIC -Wall -Wno-unused-function -Iinclude -Ichemtensor_b200/host
BUILD    := build_tmp

HOST_SRC := $(wildcard chemtensor_b200/host/*.c)
CUDA_SRC := $(wildcard chemtensor_b200/csrc/*.cu)
HOST_OBJ := $(patsubst chemtensor_b200/host/%.c,$(BUILD)/host/%.o,$(HOST_SRC))
CUDA_OBJ := $(patsubst chemtensor_b200/csrc/%.cu,$(BUILD)/csrc/%.o,$(CUDA_SRC))

LIB := chemtensor_b200/libchemtensor_b200.so
EMU := tests/emu/libctb_hostlogic_emu.so

all: lib emu
lib: $(LIB)
emu: $(EMU)

$(BUILD)/host/%.o: chemtensor_b200/host/%.c chemtensor_b200/host/ctb_internal.h $(wildcard include/*.h)
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -c $< -o $@

$(BUILD)/csrc/%.o: chemtensor_b200/csrc/%.cu chemtensor_b200/csrc/ctbd_common.cuh include/ctb_device.h
	@mkdir -p $(dir $@)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(HOST_OBJ) $(CUDA_OBJ)
	$(NVCC) $(ARCH) -shared -Xlinker -Bsymbolic -Xlinker -soname=libchemtensor_b200.so -Xlinker --no-undefined -o $@ $^ -lm -lgomp -ldl

$(BUILD)/emu/ctbd_emu.o: tests/emu/ctbd_emu.c include/ctb_device.h
	@mkdir -p $(dir $@)
	$(CC) $(CFLAGS) -c $< -o $@

$(EMU): $(HOST_OBJ) $(BUILD)/emu/ctbd_emu.o
	$(CC) -shared -Wl,-Bsymbolic -Wl,-soname,libctb_hostlogic_emu.so -Wl,--no-undefined -o $@ $^ -lm -lrt

oracle:
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle; else echo "reference sources absent: using prebuilt oracle/_ref"; fi

# the reference relinked against the engine; needs lib, emu and the oracle objects
dropin: lib emu oracle
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle dropin; else echo "reference sources absent: using prebuilt oracle/_ref"; fi

clean:
	rm -rf $(BUILD) $(LIB) $(EMU)

.PHONY: all lib emu oracle dropin clean

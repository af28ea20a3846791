# Builds libmarlin_b200.so (CUDA kernels for sm_100a + the C ABI) in-tree.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall
SRC       := marlin_b200/csrc
BUILD     := marlin_b200/_build
LIB       := marlin_b200/libmarlin_b200.so
CU        := $(wildcard $(SRC)/*.cu)
OBJ       := $(patsubst $(SRC)/%.cu,$(BUILD)/%.o,$(CU))
HDR       := $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h) include/marlin_b200.h

all: $(LIB)

$(BUILD)/%.o: $(SRC)/%.cu $(HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

emu: tests/emu/_build/emu_fft_test
tests/emu/_build/emu_fft_test: tests/emu/emu_fft_test.cpp tests/emu/cuda_emu.h $(wildcard $(SRC)/*.cuh)
	@mkdir -p tests/emu/_build
	g++ -std=c++17 -O1 -Itests/emu -o $@ $<

clean:
	rm -rf $(BUILD) $(LIB) tests/emu/_build
.PHONY: all clean emu

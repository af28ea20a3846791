# Builds libmarlin_b200.so (CUDA kernels for sm_100a + the C ABI) in-tree.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
SRC       := marlin_b200/csrc
BUILD     := marlin_b200/_build
NVFLAGS   := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -I$(BUILD)
LIB       := marlin_b200/libmarlin_b200.so
CU        := $(wildcard $(SRC)/*.cu)
CPP       := $(wildcard $(SRC)/*.cpp)
OBJ       := $(patsubst $(SRC)/%.cu,$(BUILD)/%.o,$(CU)) $(patsubst $(SRC)/%.cpp,$(BUILD)/%.o,$(CPP))
HDR       := $(wildcard $(SRC)/*.cuh) $(wildcard $(SRC)/*.h) include/marlin_b200.h

EMBED     := $(BUILD)/mrl_embedded_headers.inc
EMBED_SRC := $(SRC)/mrl_fft.cuh $(SRC)/mrl_passes.cuh $(SRC)/mrl_tma.cuh $(SRC)/mrl_passes_tma.cuh

# host side: MOOSE-style objects (TensorProblem / TensorOperator / TensorSolver ...) + the driver executable
HOST_SRC  := $(wildcard host/src/*.C) $(wildcard host/shim/*.C)
HOST_OBJ  := $(patsubst host/%.C,host/_build/%.o,$(HOST_SRC))
HOST_HDR  := $(wildcard host/include/*.h) $(wildcard host/shim/*.h) include/marlin_b200.h
HOST_FLAGS:= -std=c++17 -O2 -g -fPIC -Wall -Wno-unused-function -ffp-contract=off -Ihost/include -Ihost/shim -Iinclude
APP       := marlin_b200/marlin_b200-opt

all: $(LIB) $(APP)

host/_build/%.o: host/%.C $(HOST_HDR)
	@mkdir -p $(dir $@)
	g++ $(HOST_FLAGS) -c $< -o $@

$(APP): $(HOST_OBJ) $(LIB)
	g++ -o $@ $(HOST_OBJ) -Lmarlin_b200 -lmarlin_b200 -lz -pthread -Wl,-rpath,'$$ORIGIN' -Wl,-rpath-link,/usr/local/cuda/lib64

$(EMBED): $(EMBED_SRC) tools/embed_headers.py
	@mkdir -p $(BUILD)
	python3 tools/embed_headers.py $@ $(EMBED_SRC)

$(BUILD)/mrl_expr_zfwd.o: $(EMBED)

$(BUILD)/%.o: $(SRC)/%.cu $(HDR)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(BUILD)/%.o: $(SRC)/%.cpp $(HDR)
	@mkdir -p $(BUILD)
	g++ -std=c++17 -O2 -fPIC -Wall -I/usr/local/cuda/include -c $< -o $@

$(LIB): $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl

emu: tests/emu/_build/emu_fft_test
tests/emu/_build/emu_fft_test: tests/emu/emu_fft_test.cpp tests/emu/cuda_emu.h $(wildcard $(SRC)/*.cuh)
	@mkdir -p tests/emu/_build
	g++ -std=c++17 -O1 -Itests/emu -o $@ $<

clean:
	rm -rf $(BUILD) $(LIB) $(APP) host/_build tests/emu/_build
.PHONY: all clean emu

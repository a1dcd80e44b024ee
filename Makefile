# Builds the C-ABI shared library (sm_100a) and the oracle's native pieces.
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xfatbin=-compress-all -Xcompiler -fPIC -Iinclude -Ifancy_gym_b200/csrc --expt-relaxed-constexpr
SRC_DIR   := fancy_gym_b200/csrc
BUILD_DIR := build/obj
LIB       := fancy_gym_b200/lib/libfancygym_b200.so
SRCS      := $(wildcard $(SRC_DIR)/*.cu)
OBJS      := $(patsubst $(SRC_DIR)/%.cu,$(BUILD_DIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(SRC_DIR)/*.cuh) $(wildcard $(SRC_DIR)/*.h) include/fancy_gym_b200.h

all: $(LIB)

$(BUILD_DIR)/%.o: $(SRC_DIR)/%.cu $(HDRS)
	@mkdir -p $(BUILD_DIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $(LIB))
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

clean:
	rm -rf build $(LIB)

.PHONY: all clean

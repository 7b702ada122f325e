# Build of libcpm_b200.so (sm_100a only), the CPU oracle and the reference-derived RNG checker.
PKG := correlated-photon-mapping-for-interactive-global-illumination-of-time-varying-volumetric-data_b200
NVCC ?= /usr/local/cuda/bin/nvcc
CUDA_HOME ?= /usr/local/cuda
NVFLAGS := -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false \
           -Xcompiler -fPIC,-fvisibility=hidden,-Wall -Iinclude -I$(PKG)/csrc
SRCS := $(wildcard $(PKG)/csrc/*.cu)
OBJS := $(patsubst $(PKG)/csrc/%.cu,build/%.o,$(SRCS))
HDRS := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard include/*.h)
NCCL_INC ?= $(shell python -c "import nvidia.nccl, os; print(os.path.join(os.path.dirname(nvidia.nccl.__file__), 'include'))" 2>/dev/null)

LIB := $(PKG)/libcpm_b200.so
HOSTLIB := $(PKG)/libcpm_host.so
HOST_SRCS := $(wildcard $(PKG)/host/*.cpp)
HOST_HDRS := $(wildcard $(PKG)/host/*.h)
HOST_CXX ?= $(shell [ -x /usr/bin/g++ ] && echo /usr/bin/g++ || echo g++)

all: $(LIB) $(HOSTLIB) oracle

# Host mirror of the reference's Inviwo processors: plain C++17 on top of the C ABI only.
$(HOSTLIB): $(HOST_SRCS) $(HOST_HDRS) $(LIB) include/cpm_b200.h
	$(HOST_CXX) -std=c++17 -O2 -fPIC -fvisibility=hidden -Wall -shared -Iinclude -I$(PKG)/host \
	    -o $@ $(HOST_SRCS) -L$(PKG) -lcpm_b200 -Wl,-rpath,'$$ORIGIN'

build/%.o: $(PKG)/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) $(if $(NCCL_INC),-I$(NCCL_INC)) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) -gencode arch=compute_100a,code=sm_100a -cudart static -ldl

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB) $(HOSTLIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean

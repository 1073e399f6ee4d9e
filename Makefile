# Builds the three native pieces in-tree (built .so files are git-ignored but travel to the GPU box):
#   advanced-cpu-raytracing_b200/libdorktracer.so   CUDA hot path + C ABI (include/dorktracer.h), sm_100a only
#   advanced-cpu-raytracing_b200/libdthost.so       host mirror of the reference's scene layer (include/dorktracer_host.h)
#   oracle/libdtoracle.so                            CPU restatement of the reference algorithm (TEST INFRASTRUCTURE ONLY)
PKG      := advanced-cpu-raytracing_b200
NVCC     ?= /usr/local/cuda/bin/nvcc
CXX      := /usr/bin/g++
CC       := /usr/bin/gcc

# -fmad=false: the reference build has no FMA instructions; bit-exact hit parity needs unfused float math
# (SURVEY.md 8a).  Device code that wants FMAs (conservative box tests) uses explicit __fmaf_* intrinsics.
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false -std=c++17 \
             -Xcompiler -fPIC,-O2,-ffp-contract=off -Iinclude --expt-relaxed-constexpr
HOSTFLAGS := -O2 -fPIC -std=c++14 -ffp-contract=off -Iinclude -Wall -Wno-unused-function
ORACLEFLAGS := -O2 -fPIC -std=c99 -ffp-contract=off -Iinclude -Wall -Wno-unused-function

CUDA_SRCS := $(wildcard $(PKG)/csrc/*.cu)
CUDA_HDRS := $(wildcard $(PKG)/csrc/*.cuh) $(wildcard $(PKG)/csrc/*.h) include/dorktracer.h
HOST_SRCS := $(wildcard $(PKG)/host/*.cpp)
HOST_HDRS := $(wildcard $(PKG)/host/*.h) include/dorktracer.h include/dorktracer_host.h

all: host oracle cuda cli

host: $(PKG)/libdthost.so
oracle: oracle/libdtoracle.so
cuda: $(PKG)/libdorktracer.so
cli: $(PKG)/raytracer_gpu

$(PKG)/libdthost.so: $(filter-out $(PKG)/host/dth_main.cpp,$(HOST_SRCS)) $(HOST_HDRS)
	$(CXX) $(HOSTFLAGS) -shared -o $@ $(filter-out $(PKG)/host/dth_main.cpp,$(HOST_SRCS)) -lz -lpthread

oracle/libdtoracle.so: oracle/dt_oracle.c include/dorktracer.h
	$(CC) $(ORACLEFLAGS) -shared -o $@ oracle/dt_oracle.c -lm -lpthread

$(PKG)/libdorktracer.so: $(CUDA_SRCS) $(CUDA_HDRS)
	$(NVCC) $(NVCCFLAGS) -shared -o $@ $(CUDA_SRCS) -lcudart

$(PKG)/raytracer_gpu: $(PKG)/host/dth_main.cpp $(PKG)/libdthost.so $(PKG)/libdorktracer.so
	$(CXX) $(HOSTFLAGS) -o $@ $(PKG)/host/dth_main.cpp -L$(PKG) -ldthost -ldorktracer -lpthread -Wl,-rpath,'$$ORIGIN'

clean:
	rm -f $(PKG)/*.so oracle/*.so $(PKG)/raytracer_gpu

.PHONY: all host oracle cuda cli clean

# A/B variants of the CUDA library for on-GPU measurements (tests/_perf_ab.py): make variant NAME=x DEFS="-DFOO"
variant:
	$(NVCC) $(NVCCFLAGS) $(DEFS) -shared -o $(PKG)/libdorktracer_$(NAME).so $(CUDA_SRCS) -lcudart

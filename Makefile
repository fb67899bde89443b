# Builds libvofod_cuda.so (the product), the CPU oracle (test infrastructure) and the synthetic scan generator.
NVCC     ?= nvcc
CXX      ?= g++
ARCH     := -gencode arch=compute_100a,code=sm_100a
# -fmad=false: every fp32/fp64 mul+add stays separately rounded, like the reference's x86-64 build without -march
NVFLAGS  := -std=c++17 -O3 $(ARCH) -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off,-Wall
CSRC     := vofod_b200/csrc
OBJ      := build/ctx.o build/raycast.o build/voxelgrid.o build/cluster.o build/pipeline.o build/classify.o build/sepclusters.o build/slab.o build/apriori.o build/sensor.o
HDR      := $(CSRC)/common.cuh $(CSRC)/prims.cuh include/vofod_cuda.h

all: vofod_b200/libvofod_cuda.so oracle synth

build/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c $< -o $@

vofod_b200/libvofod_cuda.so: $(OBJ)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart

oracle:
	$(MAKE) -C oracle

synth: vofod_b200/synth/libvofod_synth.so
vofod_b200/synth/libvofod_synth.so: vofod_b200/synth/scene.cpp include/vofod_cuda.h
	$(CXX) -std=c++17 -O2 -fPIC -shared -pthread -o $@ $<

clean:
	rm -rf build vofod_b200/libvofod_cuda.so vofod_b200/synth/libvofod_synth.so; $(MAKE) -C oracle clean
.PHONY: all oracle synth clean

// hap_b200/csrc/simt.h
//
// The one header every kernel file includes.  Under nvcc it is just the CUDA runtime plus a launch
// macro.  Under HAPB200_EMU (tests/emu/, g++ only, never part of libhap_b200.so) it maps the CUDA
// execution model onto cooperative fibers so kernel logic can be unit-tested in a container without
// a GPU; that build is a development aid for tests only and is not a CPU fallback of the product.
#pragma once

#ifdef HAPB200_EMU
#include "simt_emu.h"
#else
#include <cuda_runtime.h>
#include <stdint.h>

#define HAP_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define HAP_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#endif

#define HAP_FULL_MASK 0xffffffffu

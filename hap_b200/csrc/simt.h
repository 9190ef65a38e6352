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

// Shared-memory accesses through a 32-bit shared-window address held in a register (for serial pointer-chasing
// loops: with generic addressing ptxas re-derives the window base -- an S2R -- in every iteration).
typedef uint32_t hap_saddr_t;
__device__ __forceinline__ hap_saddr_t hap_smem_addr(const void *p) { return (hap_saddr_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t hap_lds_u8(hap_saddr_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a)); return v; }
__device__ __forceinline__ void hap_sts_u16(hap_saddr_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "h"((unsigned short)v)); }
#endif

#define HAP_FULL_MASK 0xffffffffu

// ob2_platform.h -- what the kernel sources need from the compiler.
//
// Product build: nvcc, sm_100a, everything below maps 1:1 onto CUDA.
// Test-only build (-DOB2_EMU, tests/emu/): the same kernel sources are compiled
// by g++ with every CUDA thread of one block running as a host thread
// (tests/emu/cuda_emu.h), so kernel LOGIC (indexing, pivoting, table build,
// barrier placement under ThreadSanitizer) can be checked without a GPU.  The
// emulation is never linked into liboblas_b200.so.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef OB2_EMU
#include "cuda_emu.h"
#else
#include <cuda_runtime.h>
#define OB2_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define OB2_LAUNCH(kernel, grid, block, smem, stream, ...)                      \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// barrier over the first `nthreads` threads of the block (multiple of 32), id 1..15
#ifndef OB2_EMU
#define OB2_NAMED_BAR(id, nthreads) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory")
#else
#define OB2_NAMED_BAR(id, nthreads) emu::named_bar((id), (nthreads))
#endif

#define OB2_HD __host__ __device__ __forceinline__
#define OB2_D __device__ __forceinline__

// field ids = gfmat_type of the reference (gfmat.h:6)
enum { OB2_GF2 = 0, OB2_GF4 = 1, OB2_GF16 = 2, OB2_GF256 = 3 };

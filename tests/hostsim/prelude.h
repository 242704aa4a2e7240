// prelude.h -- TEST INFRASTRUCTURE: what a host (g++) build of the product's device headers needs: CUDA vector types and host
// stand-ins for the few device intrinsics the headers use. A "block" is one thread: barriers are no-ops, votes see one lane.
#pragma once
#include <cstring>
#include <cstdint>
#include <cmath>
#include <vector>
#include <thread>
#include <atomic>
#include <cuda_runtime.h>      // vector types + make_float3 (host-usable)

// host stand-ins for the few device intrinsics the headers use
template<typename T> static inline T __ldg(const T* p) { return *p; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
template<typename T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
// a "block" of one thread: barriers are no-ops, block votes / ballots see only this lane (the split-phase functions of zr_rpt.cuh
// are included for their per-thread arithmetic; wave-scope results are not compared on the host)
static inline void __syncthreads() {}
static inline int __syncthreads_or(int p) { return p; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
template<typename T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
static inline int __ffs(int x) { return __builtin_ffs(x); }
// kernels-only builtins that some device helpers touch (per-tile cost accounting): thread (0,0,0) of block (0,0,0), a clock that stands still
static const uint3 hostsim_threadIdx = { 0, 0, 0 }, hostsim_blockIdx = { 0, 0, 0 };
#define threadIdx hostsim_threadIdx
#define blockIdx hostsim_blockIdx
static inline long long clock64() { return 0; }
template<typename T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }


// common.cuh -- shared device helpers for the bzip2 kernels (sm_100a).
#pragma once
#ifdef BZ_SIM
#include "cusim.h"
#else
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#define KLAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char _dyn_smem_raw[]; \
  type *name = reinterpret_cast<type *>(_dyn_smem_raw)
#endif

typedef long long i64;
typedef unsigned long long u64;
typedef unsigned int u32;
typedef unsigned short u16;
typedef unsigned char u8;

#define FULL_MASK 0xffffffffu
#define BZ_GROUP 50          // symbols per selector group (BJ:1348)
#define BZ_MAX_GROUPS 6      // BJ:1347
#define BZ_MAX_SYMS 258      // BJ:1343
#define BZ_MAX_CODE 20       // BJ:1342
#define BZ_MAGIC_BLOCK 0x314159265359ULL  // BJ:1350
#define BZ_MAGIC_END 0x177245385090ULL    // BJ:1351
#define BZ_CRC_POLY 0x04C11DB7u

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

template <typename T>
__device__ __forceinline__ T warp_incl_sum(T v) {
  int lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(FULL_MASK, v, d);
    if (lane >= d) v += t;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_incl_max(T v) {
  int lane = lane_id();
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    T t = __shfl_up_sync(FULL_MASK, v, d);
    if (lane >= d && t > v) v = t;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL_MASK, v, d);
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_min(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T t = __shfl_xor_sync(FULL_MASK, v, d);
    if (t < v) v = t;
  }
  return v;
}
template <typename T>
__device__ __forceinline__ T warp_max(T v) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    T t = __shfl_xor_sync(FULL_MASK, v, d);
    if (t > v) v = t;
  }
  return v;
}

// Block-wide exclusive sum.  ws: >= 33 entries of shared memory.  blockDim.x % 32 == 0.
// Every thread must call it; ends with a barrier so ws can be reused at once.
template <typename T>
__device__ __forceinline__ T block_excl_sum(T v, T &total, T *ws) {
  int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  T inc = warp_incl_sum(v);
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    T x = lane < nw ? ws[lane] : (T)0;
    T xi = warp_incl_sum(x);
    ws[lane] = xi - x;
    if (lane == 31) ws[32] = xi;
  }
  __syncthreads();
  T res = ws[w] + inc - v;
  total = ws[32];
  __syncthreads();
  return res;
}
// Block-wide exclusive max (identity `lo`), same contract.
template <typename T>
__device__ __forceinline__ T block_excl_max(T v, T lo, T &total, T *ws) {
  int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  T inc = warp_incl_max(v);
  T prev = __shfl_up_sync(FULL_MASK, inc, 1);
  if (lane == 0) prev = lo;
  if (lane == 31) ws[w] = inc;
  __syncthreads();
  if (w == 0) {
    T x = lane < nw ? ws[lane] : lo;
    T xi = warp_incl_max(x);
    T xe = __shfl_up_sync(FULL_MASK, xi, 1);
    if (lane == 0) xe = lo;
    ws[lane] = xe;
    if (lane == 31) ws[32] = xi;
  }
  __syncthreads();
  T base = ws[w];
  T res = prev > base ? prev : base;
  total = ws[32];
  __syncthreads();
  return res;
}
template <typename T>
__device__ __forceinline__ T block_min(T v, T *ws) {
  int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  v = warp_min(v);
  if (lane == 0) ws[w] = v;
  __syncthreads();
  T x = ws[0];
  for (int i = 1; i < nw; i++) if (ws[i] < x) x = ws[i];
  __syncthreads();
  return x;
}
template <typename T>
__device__ __forceinline__ T block_max(T v, T *ws) {
  int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  v = warp_max(v);
  if (lane == 0) ws[w] = v;
  __syncthreads();
  T x = ws[0];
  for (int i = 1; i < nw; i++) if (ws[i] > x) x = ws[i];
  __syncthreads();
  return x;
}
template <typename T>
__device__ __forceinline__ T block_sum(T v, T *ws) {
  int lane = lane_id(), w = warp_id(), nw = blockDim.x >> 5;
  v = warp_sum(v);
  if (lane == 0) ws[w] = v;
  __syncthreads();
  T x = 0;
  for (int i = 0; i < nw; i++) x += ws[i];
  __syncthreads();
  return x;
}

// ---- ordered single-pass compaction: decoupled look-back over tiles -----------------------------
// status[t]: bits 63..62 = state (1: the tile's own count, 2: inclusive prefix), low 32 bits = value.
// Tiles take their index from an atomic ticket, so every tile a CTA waits for is already running.
#define LB_AGG (1ull << 62)
#define LB_INC (2ull << 62)
// All 32 lanes of ONE warp of the CTA call this; returns the number of items in the tiles before `tile`.
__device__ __forceinline__ u32 lookback_warp(volatile u64 *status, u32 tile, u32 aggregate) {
  const int lane = lane_id();
  if (tile == 0) {
    if (lane == 0) { __threadfence(); status[0] = LB_INC | aggregate; }
    return 0;
  }
  if (lane == 0) { __threadfence(); status[tile] = LB_AGG | aggregate; }
  u32 excl = 0;
  i64 t = (i64)tile - 1;
  while (true) {
    i64 my = t - lane;
    u64 v = LB_INC;  // before tile 0: inclusive prefix 0
    if (my >= 0) {
      do { v = status[my]; } while ((v >> 62) == 0);
    }
    u32 inc_mask = __ballot_sync(FULL_MASK, (v >> 62) == 2);
    int first_inc = inc_mask ? __ffs((int)inc_mask) - 1 : 32;
    u32 contrib = lane <= first_inc ? (u32)(v & 0xffffffffull) : 0u;
    excl += warp_sum<u32>(contrib);
    if (inc_mask) break;
    t -= 32;
  }
  if (lane == 0) { __threadfence(); status[tile] = LB_INC | (u64)(excl + aggregate); }
  return excl;
}

// ---- bzip2 CRC (MSB-first CRC-32, poly 0x04C11DB7; BJ:1013-1067) algebra ----
// A register value is a polynomial over GF(2) with the coefficient of x^k in bit k.
__host__ __device__ __forceinline__ u32 crc_mulmod(u32 a, u32 b) {
  u32 r = 0;
  for (int i = 0; i < 32; i++) {
    r = (r & 0x80000000u) ? (r << 1) ^ BZ_CRC_POLY : (r << 1);
    if (b & 0x80000000u) r ^= a;
    b <<= 1;
  }
  return r;
}
__host__ __device__ __forceinline__ u32 crc_xpow(u64 n) {  // x^n mod P
  u32 r = 1, base = 2;
  while (n) {
    if (n & 1) r = crc_mulmod(r, base);
    base = crc_mulmod(base, base);
    n >>= 1;
  }
  return r;
}
__host__ __device__ __forceinline__ u32 crc_table_entry(u32 b) {
  u32 c = b << 24;
  for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ BZ_CRC_POLY : (c << 1);
  return c;
}

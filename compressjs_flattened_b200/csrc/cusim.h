// cusim.h -- TEST-ONLY CPU stand-in for the CUDA execution model.
//
// There is no GPU in the build container, so the kernels in this directory are
// also compiled with g++ (-DBZ_SIM) against this header into
// tests/sim/libbz2b200_sim.so, where every CTA runs as a set of ucontext fibers
// (one per CUDA thread) with working __syncthreads / warp collectives.  That
// library exists only so `pytest -m "not gpu"` can exercise kernel LOGIC; it is
// never loaded by the product package and is not a fallback: libbz2b200.so is
// built by nvcc for sm_100a only and fails loudly without a GPU.
#pragma once
#ifndef BZ_SIM
#error "cusim.h is only for the -DBZ_SIM test build"
#endif
#include <algorithm>
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <vector>

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint2 { unsigned x, y; };
struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return uint4{a, b, c, d}; }
static inline uint2 make_uint2(unsigned a, unsigned b) { return uint2{a, b}; }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __shared__ static
#define __launch_bounds__(...)
#define __constant__ static
#define __align__(x) __attribute__((aligned(x)))

typedef int cudaError_t;
typedef void *cudaStream_t;
typedef void *cudaEvent_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };

// minimal x86-64 SysV fiber switch (callee-saved registers only; no signal mask syscalls)
extern "C" void cusim_switch(void **save_sp, void *load_sp);
asm(R"(
.text
.globl cusim_switch
.type cusim_switch,@function
cusim_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cusim_switch,.-cusim_switch
)");

namespace cusim {
struct Warp {
  uint64_t vals[2][32];
  unsigned arrived = 0, gen = 0, lanes = 32;
};
struct State {
  dim3 threadIdx, blockIdx, blockDim, gridDim;
  unsigned nthreads = 0, alive = 0, bar_arrived = 0, bar_gen = 0;
  unsigned cur = 0;
  std::vector<void *> ctx;
  std::vector<char *> stacks;
  std::vector<unsigned char> done;
  std::vector<Warp> warps;
  void *main_ctx = nullptr;
  const std::function<void()> *body = nullptr;
  unsigned char *dyn_smem = nullptr;
  size_t dyn_cap = 0;
};
inline State &S() { static State s; return s; }
static const size_t kStack = 64 * 1024;

inline void yield() {
  State &s = S();
  cusim_switch(&s.ctx[s.cur], s.main_ctx);
}
inline void bar_release_check() {
  State &s = S();
  if (s.alive && s.bar_arrived == s.alive) { s.bar_arrived = 0; s.bar_gen++; }
}
inline void entry() {
  State &s = S();
  (*s.body)();
  s.done[s.cur] = 1;
  s.alive--;
  bar_release_check();  // exited threads no longer hold a barrier back
  cusim_switch(&s.ctx[s.cur], s.main_ctx);
  abort();  // a finished fiber is never resumed
}
inline void set_tid(unsigned t) {
  State &s = S();
  s.cur = t;
  s.threadIdx.x = t % s.blockDim.x;
  s.threadIdx.y = (t / s.blockDim.x) % s.blockDim.y;
  s.threadIdx.z = t / (s.blockDim.x * s.blockDim.y);
}
inline std::mutex &launch_mutex() { static std::mutex m; return m; }
inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body) {
  std::lock_guard<std::mutex> one_kernel_at_a_time(launch_mutex());  // the shard scheduler launches from several host threads
  State &s = S();
  unsigned nt = block.x * block.y * block.z;
  assert(nt >= 1 && nt <= 1024);
  if (smem > s.dyn_cap) { free(s.dyn_smem); s.dyn_smem = (unsigned char *)aligned_alloc(128, (smem + 127) / 128 * 128); s.dyn_cap = smem; }
  while (s.stacks.size() < nt) s.stacks.push_back((char *)malloc(kStack));
  s.ctx.resize(nt);
  s.done.assign(nt, 0);
  s.gridDim = grid; s.blockDim = block; s.nthreads = nt; s.body = &body;
  for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
      for (unsigned bx = 0; bx < grid.x; bx++) {
        s.blockIdx = dim3(bx, by, bz);
        s.alive = nt; s.bar_arrived = 0; s.bar_gen = 0;
        s.warps.assign((nt + 31) / 32, Warp());
        for (unsigned w = 0; w < s.warps.size(); w++) s.warps[w].lanes = std::min(32u, nt - w * 32);
        std::fill(s.done.begin(), s.done.end(), 0);
        for (unsigned t = 0; t < nt; t++) {
          uintptr_t top = ((uintptr_t)s.stacks[t] + kStack) & ~(uintptr_t)15;
          void **sp = (void **)(top - 16);  // return-address slot, 16-byte aligned
          sp[1] = nullptr;
          sp[0] = (void *)entry;
          sp -= 6;                          // r15 r14 r13 r12 rbx rbp
          for (int q = 0; q < 6; q++) sp[q] = nullptr;
          s.ctx[t] = (void *)sp;
        }
        while (s.alive) {
          unsigned before = s.alive;
          bool progressed = false;
          for (unsigned t = 0; t < nt; t++) {
            if (s.done[t]) continue;
            set_tid(t);
            cusim_switch(&s.main_ctx, s.ctx[t]);
            progressed = true;
          }
          (void)before; (void)progressed;
        }
      }
  s.body = nullptr;
}
inline void syncthreads() {
  State &s = S();
  unsigned g = s.bar_gen;
  s.bar_arrived++;
  bar_release_check();
  while (s.bar_gen == g) yield();
}
// all lanes of the calling warp deposit v and receive everybody's value
inline void exchange(uint64_t v, uint64_t out[32]) {
  State &s = S();
  Warp &w = s.warps[s.cur / 32];
  unsigned lane = s.cur % 32, g = w.gen;
  w.vals[g & 1][lane] = v;
  if (++w.arrived == w.lanes) { w.arrived = 0; w.gen++; }
  else while (w.gen == g) yield();
  for (unsigned l = 0; l < 32; l++) out[l] = l < w.lanes ? w.vals[g & 1][l] : 0;
}
inline unsigned lanes_in_warp() { State &s = S(); return s.warps[s.cur / 32].lanes; }
}  // namespace cusim

#define threadIdx (cusim::S().threadIdx)
#define blockIdx (cusim::S().blockIdx)
#define blockDim (cusim::S().blockDim)
#define gridDim (cusim::S().gridDim)
#define warpSize 32

static inline void __syncthreads() { cusim::syncthreads(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { (void)mask; uint64_t o[32]; cusim::exchange(0, o); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}
static inline void __threadfence_system() {}

template <typename T> static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
  (void)mask; uint64_t o[32], x = 0; memcpy(&x, &v, sizeof(T)); cusim::exchange(x, o);
  unsigned lane = cusim::S().cur % 32; int base = (lane / width) * width;
  T r; uint64_t y = o[base + ((unsigned)src % (unsigned)width)]; memcpy(&r, &y, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_up_sync(unsigned mask, T v, unsigned d, int width = 32) {
  (void)mask; (void)width; uint64_t o[32], x = 0; memcpy(&x, &v, sizeof(T)); cusim::exchange(x, o);
  unsigned lane = cusim::S().cur % 32; uint64_t y = lane >= d ? o[lane - d] : x; T r; memcpy(&r, &y, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_down_sync(unsigned mask, T v, unsigned d, int width = 32) {
  (void)mask; (void)width; uint64_t o[32], x = 0; memcpy(&x, &v, sizeof(T)); cusim::exchange(x, o);
  unsigned lane = cusim::S().cur % 32; uint64_t y = lane + d < 32 ? o[lane + d] : x; T r; memcpy(&r, &y, sizeof(T)); return r;
}
template <typename T> static inline T __shfl_xor_sync(unsigned mask, T v, int m, int width = 32) {
  (void)mask; (void)width; uint64_t o[32], x = 0; memcpy(&x, &v, sizeof(T)); cusim::exchange(x, o);
  unsigned lane = cusim::S().cur % 32; uint64_t y = o[lane ^ (unsigned)m]; T r; memcpy(&r, &y, sizeof(T)); return r;
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
  (void)mask; uint64_t o[32]; cusim::exchange(pred ? 1 : 0, o); unsigned r = 0;
  for (unsigned l = 0; l < cusim::lanes_in_warp(); l++) if (o[l]) r |= 1u << l; return r;
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __all_sync(unsigned mask, int pred) {
  unsigned b = __ballot_sync(mask, pred), n = cusim::lanes_in_warp();
  return b == (n == 32 ? 0xffffffffu : ((1u << n) - 1));
}
template <typename T> static inline unsigned __match_any_sync(unsigned mask, T v) {
  (void)mask; uint64_t o[32], x = 0; memcpy(&x, &v, sizeof(T)); cusim::exchange(x, o); unsigned r = 0;
  for (unsigned l = 0; l < cusim::lanes_in_warp(); l++) if (o[l] == x) r |= 1u << l; return r;
}
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {
  (void)mask; uint64_t o[32]; cusim::exchange(v, o); unsigned r = 0;
  for (unsigned l = 0; l < cusim::lanes_in_warp(); l++) r |= (unsigned)o[l]; return r;
}
// position of the offset-th (1-based) set bit of mask at or above bit base; 0xffffffff if there is none (offset > 0 only)
static inline unsigned __fns(unsigned mask, unsigned base, int offset) {
  for (unsigned b = base; b < 32; b++) if ((mask >> b) & 1u) { if (--offset == 0) return b; }
  return 0xffffffffu;
}
static inline unsigned __activemask() { unsigned n = cusim::lanes_in_warp(); return n == 32 ? 0xffffffffu : ((1u << n) - 1); }

static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline unsigned __brev(unsigned v) { unsigned r = 0; for (int i = 0; i < 32; i++) if (v & (1u << i)) r |= 1u << (31 - i); return r; }
static inline unsigned __byte_perm(unsigned a, unsigned b, unsigned s) {
  uint64_t v = ((uint64_t)b << 32) | a; unsigned r = 0;
  for (int i = 0; i < 4; i++) { unsigned sel = (s >> (4 * i)) & 7; r |= (unsigned)((v >> (8 * sel)) & 0xff) << (8 * i); }
  return r;
}
static inline unsigned __funnelshift_l(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (hi << sh) | (lo >> (32 - sh)) : hi; }
static inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned sh) { sh &= 31; return sh ? (lo >> sh) | (hi << (32 - sh)) : lo; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline unsigned __vcmpne4(unsigned a, unsigned b) { unsigned r = 0; for (int i = 0; i < 4; i++) if (((a >> (8 * i)) & 0xff) != ((b >> (8 * i)) & 0xff)) r |= 0xffu << (8 * i); return r; }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) { return (unsigned long long)(((unsigned __int128)a * b) >> 64); }

template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
template <typename T> static inline T atomicSub(T *p, T v) { T o = *p; *p = o - v; return o; }
template <typename T> static inline T atomicMax(T *p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <typename T> static inline T atomicMin(T *p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <typename T> static inline T atomicOr(T *p, T v) { T o = *p; *p = o | v; return o; }
template <typename T> static inline T atomicAnd(T *p, T v) { T o = *p; *p = o & v; return o; }
template <typename T> static inline T atomicExch(T *p, T v) { T o = *p; *p = v; return o; }
template <typename T> static inline T atomicCAS(T *p, T c, T v) { T o = *p; if (o == c) *p = v; return o; }
#ifndef BZ_SIM_NO_MINMAX
using std::max;
using std::min;
#endif

// ---- runtime API subset (host memory stands in for device memory) ----
static inline cudaError_t cudaMalloc(void **p, size_t n) { *p = aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
static inline cudaError_t cudaFree(void *p) { free(p); return 0; }
static inline cudaError_t cudaMallocHost(void **p, size_t n) { return cudaMalloc(p, n); }
static inline cudaError_t cudaFreeHost(void *p) { free(p); return 0; }
static inline cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t = 0) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return 0; }
static inline cudaError_t cudaMemsetAsync(void *d, int v, size_t n, cudaStream_t = 0) { memset(d, v, n); return 0; }
static inline cudaError_t cudaMemset(void *d, int v, size_t n) { memset(d, v, n); return 0; }
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline cudaError_t cudaDeviceSynchronize() { return 0; }
static inline cudaError_t cudaStreamCreate(cudaStream_t *s) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = nullptr; return 0; }
static inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned = 0) { return 0; }
#define cudaStreamNonBlocking 1
static inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
static inline cudaError_t cudaGetLastError() { return 0; }
static inline cudaError_t cudaSetDevice(int) { return 0; }
static inline cudaError_t cudaGetDevice(int *d) { *d = 0; return 0; }
static inline cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return 0; }
static inline const char *cudaGetErrorString(cudaError_t) { return "cusim"; }
static inline cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = nullptr; return 0; }
static inline cudaError_t cudaEventDestroy(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t = 0) { return 0; }
static inline cudaError_t cudaEventSynchronize(cudaEvent_t) { return 0; }
static inline cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0; return 0; }
#define cudaFuncSetAttribute(...) 0
#define cudaDevAttrMultiProcessorCount 16
static inline int cudaDeviceGetAttribute(int *v, int, int) { *v = 148; return 0; }
#define cudaFuncAttributeMaxDynamicSharedMemorySize 0

// kernel launch: KLAUNCH(kernel, grid, block, smem_bytes, stream, args...)
#define KLAUNCH(kern, grid, block, smem, stream, ...) \
  do { std::function<void()> _b = [=]() { kern(__VA_ARGS__); }; cusim::launch(dim3(grid), dim3(block), (smem), _b); } while (0)
#define DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(cusim::S().dyn_smem)

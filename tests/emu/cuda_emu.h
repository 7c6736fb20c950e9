// TEST INFRASTRUCTURE (not product code): a minimal host emulation of the CUDA execution model, just enough to run the
// kernels of handwritten-ocr_b200/csrc/image_fast.cuh thread by thread on the CPU (g++ -std=c++20 -pthread).
// One std::thread per CUDA thread of a CTA, CTAs one after the other; __syncthreads is a std::barrier, __shared__ is a
// function-level static (CTAs run serially, so one copy is enough), warp shuffles go through a per-warp slot array.
// It exists so that the indexing / byte-permute / packed-lane logic of those kernels is checked against the oracle in
// the `-m "not gpu"` suite, before any GPU time is spent; the GPU tests remain the parity tests proper.
#pragma once
#define OCRB_EMU 1
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static
#define __constant__ static const
#define __align__(n) __attribute__((aligned(n)))

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct alignas(16) uint4 { uint32_t x, y, z, w; };
struct alignas(8) uint2 { uint32_t x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(8) float2 { float x, y; };
struct alignas(8) int2 { int x, y; };
static inline int2 make_int2(int x, int y) { return int2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace emu {
struct Cta {
  std::barrier<> *bar;
  std::vector<std::unique_ptr<std::barrier<>>> *warp_bar;
  uint64_t (*slots)[32];
};
inline thread_local Cta *cta = nullptr;
inline thread_local unsigned char *dyn_smem = nullptr;
}  // namespace emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

#define OCRB_DYN_SMEM(T, name) T *name = reinterpret_cast<T *>(emu::dyn_smem)

static inline void __syncthreads() { emu::cta->bar->arrive_and_wait(); }

template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int o) {
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned lin = threadIdx.x, warp = lin >> 5, lane = lin & 31;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu::cta->slots[warp][lane] = bits;
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  const uint64_t got = emu::cta->slots[warp][lane ^ (unsigned)o];
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &got, sizeof(T));
  return r;
}

static inline int atomicAdd(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
static inline uint32_t atomicAdd(uint32_t *p, uint32_t v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
// CUDA's overloaded integer min / max
static inline int min(int a, int b) { return a < b ? a : b; }
static inline int max(int a, int b) { return a > b ? a : b; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }
// lane + d / lane - d of the warp; lanes without a source keep their own value
template <class T>
static inline T __shfl_down_sync(unsigned m, T v, unsigned d) {
  const unsigned lane = threadIdx.x & 31;
  (void)m;
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  const unsigned warp = threadIdx.x >> 5;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu::cta->slots[warp][lane] = bits;
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  const uint64_t o = emu::cta->slots[warp][lane + d < 32 ? lane + d : lane];
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &o, sizeof(T));
  return r;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  emu::cta->slots[warp][lane] = bits;
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  const uint64_t o = emu::cta->slots[warp][lane >= d ? lane - d : lane];
  (*emu::cta->warp_bar)[warp]->arrive_and_wait();
  T r;
  std::memcpy(&r, &o, sizeof(T));
  return r;
}
// per-byte |a - b|
static inline uint32_t __vabsdiffu4(uint32_t a, uint32_t b) {
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const int x = (int)((a >> (8 * i)) & 0xffu), y = (int)((b >> (8 * i)) & 0xffu);
    r |= (uint32_t)(x > y ? x - y : y - x) << (8 * i);
  }
  return r;
}

static inline int atomicOr(int *p, int v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }

template <class T>
static inline T __ldg(const T *p) { return *p; }
template <class T>
static inline T __ldcg(const T *p) { return *p; }
static inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }

static inline uint32_t __byte_perm(uint32_t x, uint32_t y, uint32_t s) {
  const uint64_t v = ((uint64_t)y << 32) | x;
  uint32_t r = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t sel = (s >> (4 * i)) & 0x7u;
    r |= (uint32_t)((v >> (8 * sel)) & 0xffu) << (8 * i);
  }
  return r;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh) {
  const uint64_t v = ((uint64_t)hi << 32) | lo;
  return (uint32_t)(v >> (sh & 31u));
}
static inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
static inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
static inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
struct __nv_bfloat16 { uint16_t x; };
static inline __nv_bfloat16 __float2bfloat16_rn(float f) {          // round to nearest even (NaN not needed here)
  uint32_t u;
  std::memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return __nv_bfloat16{(uint16_t)(u >> 16)};
}
static inline float __bfloat162float(__nv_bfloat16 h) {
  const uint32_t u = (uint32_t)h.x << 16;
  float f;
  std::memcpy(&f, &u, 4);
  return f;
}
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }     // CUDA's rsqrtf is within 2 ulp of this; not bit-pinned
static inline float __fdiv_rn(float a, float b) { volatile float r = a / b; return r; }
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline double __dmul_rn(double a, double b) { volatile double r = a * b; return r; }
static inline double __dadd_rn(double a, double b) { volatile double r = a + b; return r; }
static inline double __dsub_rn(double a, double b) { volatile double r = a - b; return r; }
static inline int __float2int_rn(float f) { return (int)std::nearbyintf(f); }
static inline int __double2int_rn(double f) { return (int)std::nearbyint(f); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }

// unsigned dp4a: four byte products added to c
static inline uint32_t __dp4a(uint32_t a, uint32_t b, uint32_t c) {
  for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
}

// per 16-bit lane, signed: max(min(a + b, c), 0)   (VIADDMNMX.S16x2.RELU)
static inline uint32_t __viaddmin_s16x2_relu(uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r = 0;
  for (int h = 0; h < 2; ++h) {
    const int16_t s = (int16_t)(uint16_t)(((a >> (16 * h)) & 0xffffu) + ((b >> (16 * h)) & 0xffffu));
    const int16_t cc = (int16_t)(uint16_t)((c >> (16 * h)) & 0xffffu);
    int v = s < cc ? s : cc;
    if (v < 0) v = 0;
    r |= (uint32_t)(uint16_t)v << (16 * h);
  }
  return r;
}

namespace ocrb {
// dp2a.{lo,hi}.s32.u32: two signed 16-bit halves of w2 times bytes (0,1) / (2,3) of b4, added to acc
static inline int dp2a_lo_s16u8(uint32_t w2, uint32_t b4, int acc) {
  return acc + (int)(int16_t)(w2 & 0xffffu) * (int)(b4 & 0xffu) + (int)(int16_t)(w2 >> 16) * (int)((b4 >> 8) & 0xffu);
}
static inline int dp2a_hi_s16u8(uint32_t w2, uint32_t b4, int acc) {
  return acc + (int)(int16_t)(w2 & 0xffffu) * (int)((b4 >> 16) & 0xffu) + (int)(int16_t)(w2 >> 16) * (int)(b4 >> 24);
}
}  // namespace ocrb

namespace emu {
// run f() as every thread of every CTA of the grid (block.x threads; block.y == block.z == 1)
template <class F>
void launch(dim3 grid, dim3 block, size_t dyn_bytes, F f) {
  const unsigned nt = block.x, nwarp = (nt + 31) / 32;
  std::vector<unsigned char> dyn(dyn_bytes + 64);
  unsigned char *dyn_aligned = reinterpret_cast<unsigned char *>(((uintptr_t)dyn.data() + 63) & ~(uintptr_t)63);
  std::vector<uint64_t> slot_store(32 * nwarp);
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        std::barrier<> bar(nt);
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (unsigned w = 0; w < nwarp; ++w) wb.emplace_back(new std::barrier<>(std::min(32u, nt - 32 * w)));
        Cta c{&bar, &wb, reinterpret_cast<uint64_t(*)[32]>(slot_store.data())};
        std::vector<std::thread> th;
        th.reserve(nt);
        for (unsigned t = 0; t < nt; ++t)
          th.emplace_back([&, t] {
            threadIdx = dim3(t, 0, 0);
            blockIdx = dim3(bx, by, bz);
            blockDim = block;
            gridDim = grid;
            cta = &c;
            dyn_smem = dyn_aligned;
            f();
            wb[t >> 5]->arrive_and_drop();
            bar.arrive_and_drop();
          });
        for (auto &x : th) x.join();
      }
}
}  // namespace emu

// Minimal stand-in for <cuda_runtime.h> (tests only): lets g++ compile the .cuh sources of libpyflow_b200 so that
// kernel bodies can be executed on the host.
//
// Two execution styles (emu_kernels.cpp):
//  * kernels whose threads do not communicate run one emulated thread after the other on the calling thread;
//  * kernels with block reductions (and the cooperative plain CG kernel, on a grid of ONE block) run one OS thread per
//    CUDA thread of a block: __syncthreads is a barrier over the block, warp shuffles exchange through a per-warp
//    buffer between two warp barriers, the atomics are real atomics.  Blocks of a grid run one after the other, which
//    is a legal CUDA schedule for every non-cooperative kernel here ("the last block folds the partials").
#pragma once
#include <math.h>
#include <pthread.h>
#include <sched.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#define __device__
#define __host__
#define __global__ static inline
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __shared__ static

struct uint3 { unsigned x, y, z; };
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct double2 { double x, y; };
static inline double2 make_double2(double a, double b) { double2 v; v.x = a; v.y = b; return v; }
typedef void* cudaStream_t;

extern thread_local uint3 threadIdx, blockIdx;
extern thread_local dim3 blockDim, gridDim;

#define __align__(n)

namespace emu {
constexpr int MAX_THREADS = 1024;
struct Team {                       // one block executing with one OS thread per CUDA thread
  bool threaded = false;
  pthread_barrier_t block;
  pthread_barrier_t warp[MAX_THREADS / 32];
  double xchg[MAX_THREADS];
  double* dyn = nullptr;            // dynamic shared memory of the block (kernels launched with several blocks at once)
  double small[4][128];             // its small __shared__ arrays
  void* slots[48] = {nullptr};      // per-block storage of rewritten `__shared__` declarations (see shared_slot)
  pthread_mutex_t slot_mutex = PTHREAD_MUTEX_INITIALIZER;
};
extern Team team;                   // the block of a one-block-at-a-time launch
extern thread_local Team* mine;     // the block this OS thread belongs to
inline void block_sync() { if (mine->threaded) pthread_barrier_wait(&mine->block); }
extern pthread_barrier_t* grid_barrier;   // all threads of ALL blocks of a concurrently running grid, or null
// cg::grid_group::sync(): a real grid-wide barrier when the launcher set one up, else the grid is one block
inline void grid_sync() { if (grid_barrier) pthread_barrier_wait(grid_barrier); else block_sync(); }
inline double shuffle(double v, int src_lane) {
  Team& t_ = *mine;
  if (!t_.threaded) return v;
  const int t = (int)threadIdx.x, w = t >> 5;
  t_.xchg[t] = v;
  pthread_barrier_wait(&t_.warp[w]);
  const double r = t_.xchg[(w << 5) | (src_lane & 31)];
  pthread_barrier_wait(&t_.warp[w]);
  return r;
}
// `__shared__ T name[...]` of a generated kernel copy becomes a reference to the block's own zero-initialised storage,
// so that several blocks (or several ranks' grids) can run at once in one process.
inline void* shared_slot(int id, size_t bytes) {
  Team& t = *mine;
  pthread_mutex_lock(&t.slot_mutex);
  if (!t.slots[id]) t.slots[id] = calloc(1, bytes < 8 ? 8 : bytes);
  void* p = t.slots[id];
  pthread_mutex_unlock(&t.slot_mutex);
  return p;
}
inline double* dyn_shared() { return mine->dyn; }
inline double* blk_shared(int k) { return mine->small[k]; }
}  // namespace emu

static inline void __syncthreads() { emu::block_sync(); }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __shfl_xor_sync(unsigned, double v, int o) { return emu::shuffle(v, ((int)threadIdx.x & 31) ^ o); }
static inline double __shfl_up_sync(unsigned, double v, int d) {
  const int lane = (int)threadIdx.x & 31;
  const double r = emu::shuffle(v, lane - d < 0 ? lane : lane - d);
  return r;
}
static inline double __shfl_down_sync(unsigned, double v, int d) {
  const int lane = (int)threadIdx.x & 31;
  return emu::shuffle(v, lane + d > 31 ? lane : lane + d);
}
static inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
  if (emu::mine->threaded) pthread_barrier_wait(&emu::mine->warp[(int)threadIdx.x >> 5]);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src_lane) {      // 4- and 8-byte integer types
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  double carrier = 0.0;
  __builtin_memcpy(&carrier, &v, sizeof(T));
  carrier = emu::shuffle(carrier, src_lane);
  T out;
  __builtin_memcpy(&out, &carrier, sizeof(T));
  return out;
}
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST);
}
static inline unsigned long long atomicMin(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long o = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > o && !__atomic_compare_exchange_n(p, &o, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return o;
}
static inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) {
  return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST);
}
static inline void emu_yield() { sched_yield(); }     // pollers / half-written packets give way (pf_common.cuh LL words)
long long clock64();                                 // SM clock stand-in: nanoseconds x 2 (emu_kernels.cpp)
static inline double __longlong_as_double(long long v) { double d; __builtin_memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; __builtin_memcpy(&v, &d, 8); return v; }

// TEST INFRASTRUCTURE (see cuda_runtime.h in this directory).
//
// SIMT execution of a kernel on the host.  A launch runs block after block;
// inside a block the threads are either called one after the other (SIMPLE:
// kernels whose threads never cooperate) or run as fibers that yield to each
// other at every warp shuffle / vote / __syncthreads (COOP), which gives the
// CUDA semantics for convergent code.  A cooperative primitive reached from a
// SIMPLE launch aborts, so a wrong launch mode cannot pass silently.
#pragma once
#include <cstdint>
#include <cstring>
#include <functional>
#include <tuple>

namespace plb_emu {

extern uint3 g_threadIdx, g_blockIdx;
extern dim3 g_blockDim, g_gridDim;

enum LaunchMode { SIMPLE = 0, COOP = 1 };
// Enqueues the launch on `stream` (null: runs it now); `body` is called once
// per thread with threadIdx / blockIdx set.
void launch(int mode, dim3 grid, dim3 block, cudaStream_t stream,
            std::function<void()> body);
void sleep_hook();

void sync_block();
void sync_warp();
// exchange of one 64-bit payload inside the caller's warp: returns the value
// published by lane `src` (or the caller's own when src is out of range or
// that lane has exited)
uint64_t warp_exchange(uint64_t mine, int src);
unsigned warp_ballot(bool pred, unsigned *active = nullptr);
long long clock_ticks();

template <typename T>
inline uint64_t to_bits(T v)
{
    static_assert(sizeof(T) <= 8, "payload");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <typename T>
inline T from_bits(uint64_t b)
{
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}

}  // namespace plb_emu

#define threadIdx (plb_emu::g_threadIdx)
#define blockIdx (plb_emu::g_blockIdx)
#define blockDim (plb_emu::g_blockDim)
#define gridDim (plb_emu::g_gridDim)

// launch macro of the kernels' translation unit (PLB_LAUNCH in plb_kernels.cu)
// (arguments are evaluated and copied NOW, the kernel may run later)
#define PLB_EMU_LAUNCH(mode, kernel, grid, block, stream, ...)                 \
    plb_emu::launch(plb_emu::mode, dim3(grid), dim3(block), stream,            \
                    [plb_emu_args = std::make_tuple(__VA_ARGS__)]() {          \
                        std::apply(kernel, plb_emu_args);                      \
                    })

// ---- device intrinsics used by the kernels -----------------------------------
static inline void __syncthreads() { plb_emu::sync_block(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { plb_emu::sync_warp(); }
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned delta)
{
    const int lane = int(threadIdx.x & 31);
    return plb_emu::from_bits<T>(
        plb_emu::warp_exchange(plb_emu::to_bits(v), lane - int(delta)));
}
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta)
{
    const int lane = int(threadIdx.x & 31);
    return plb_emu::from_bits<T>(
        plb_emu::warp_exchange(plb_emu::to_bits(v), lane + int(delta)));
}
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src)
{
    return plb_emu::from_bits<T>(
        plb_emu::warp_exchange(plb_emu::to_bits(v), src & 31));
}
static inline unsigned __ballot_sync(unsigned, bool p) { return plb_emu::warp_ballot(p); }
static inline bool __all_sync(unsigned, bool p)
{
    unsigned active = 0;
    const unsigned votes = plb_emu::warp_ballot(p, &active);
    return votes == active;
}
static inline bool __any_sync(unsigned, bool p) { return plb_emu::warp_ballot(p) != 0; }
template <typename T>
static inline T __ldcg(const T *p) { return *p; }
template <typename T>
static inline T __ldcs(const T *p) { return *p; }
template <typename T>
static inline void __stcs(T *p, T v) { *p = v; }
static inline long long clock64() { return plb_emu::clock_ticks(); }
static inline void __nanosleep(unsigned) { plb_emu::sleep_hook(); }
static inline void __threadfence_system() {}
static inline void __threadfence() {}
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
[[noreturn]] static inline void __trap() { __builtin_trap(); }
static inline unsigned atomicAdd(unsigned *p, unsigned v)
{
    const unsigned old = *p;
    *p = old + v;
    return old;
}
static inline unsigned long long atomicExch(unsigned long long *p,
                                            unsigned long long v)
{
    const unsigned long long old = *p;
    *p = v;
    return old;
}

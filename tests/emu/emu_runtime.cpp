// TEST INFRASTRUCTURE (see include/cuda_runtime.h): the host-side stand-in for
// the CUDA runtime and the fiber scheduler behind simt.h.
#include <cstdio>
#include <cstdlib>
#include <ucontext.h>
#include <vector>

#include <cuda_runtime.h>

// ===========================================================================
// "runtime": device memory is host memory, a stream runs in issue order
// ===========================================================================
struct plb_emu_stream { int id; };
struct plb_emu_event { int id; };

extern "C" {

const char *cudaGetErrorString(cudaError_t e)
{
    return e == cudaSuccess ? "no error"
         : e == cudaErrorMemoryAllocation ? "out of memory"
         : e == cudaErrorNotSupported ? "not supported by the emulator"
         : "error";
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 1000000; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaDeviceGetPCIBusId(char *buf, int len, int)
{
    snprintf(buf, size_t(len), "0000:00:00.0");
    return cudaSuccess;
}
cudaError_t plb_emu_malloc(void **p, size_t n)
{
    // uninitialised on purpose (0xA5 pattern): reading memory nobody wrote
    // shows up as a wild value, like on the device
    *p = malloc(n ? n : 1);
    if (!*p) return cudaErrorMemoryAllocation;
    memset(*p, 0xA5, n);
    return cudaSuccess;
}
cudaError_t plb_emu_host_alloc(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t)
{
    memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new plb_emu_stream{0}; return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new plb_emu_stream{1}; return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new plb_emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { *e = new plb_emu_event{0}; return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *, void *) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void **, cudaIpcMemHandle_t, unsigned) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

}  // extern "C"

// ===========================================================================
// SIMT fibers
// ===========================================================================
namespace plb_emu {

uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);

namespace {

constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = size_t(256) << 10;

struct Barrier {
    int count = 0;
    unsigned gen = 0;
};

struct Fiber {
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = true;
};

bool g_coop = false;
int g_n = 0, g_cur = 0;
int g_alive_block = 0, g_alive_warp[MAX_THREADS / 32];
Barrier g_block_barrier, g_warp_barrier[MAX_THREADS / 32];
uint64_t g_slot[MAX_THREADS];
bool g_pred[MAX_THREADS];
Fiber g_fiber[MAX_THREADS];
ucontext_t g_sched;
const std::function<void()> *g_body = nullptr;
long long g_idle = 0, g_ticks = 0;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "plb_emu: %s (block %u, thread %u)\n", what, g_blockIdx.x,
            g_threadIdx.x);
    abort();
}

void yield()
{
    if (++g_idle > 50000000LL) die("dead lock: fibers wait at a barrier that cannot complete");
    swapcontext(&g_fiber[g_cur].ctx, &g_sched);
}

void wait(Barrier &b, const int &alive)
{
    const unsigned gen = b.gen;
    ++b.count;
    while (b.gen == gen) {
        if (b.count >= alive) {
            b.count = 0;
            ++b.gen;
            g_idle = 0;
            break;
        }
        yield();
    }
}

void trampoline()
{
    (*g_body)();
    const int me = g_cur;
    g_fiber[me].done = true;
    --g_alive_block;
    --g_alive_warp[me / 32];
    g_idle = 0;
    // returning switches to uc_link = the scheduler
}

void need_coop(const char *what)
{
    if (!g_coop) die(what);
}

}  // namespace

void sync_block()
{
    need_coop("__syncthreads in a kernel launched in SIMPLE mode");
    wait(g_block_barrier, g_alive_block);
}

void sync_warp()
{
    need_coop("__syncwarp in a kernel launched in SIMPLE mode");
    const int w = g_cur / 32;
    wait(g_warp_barrier[w], g_alive_warp[w]);
}

uint64_t warp_exchange(uint64_t mine, int src)
{
    need_coop("warp shuffle in a kernel launched in SIMPLE mode");
    const int w = g_cur / 32;
    g_slot[g_cur] = mine;
    wait(g_warp_barrier[w], g_alive_warp[w]);
    uint64_t r = mine;
    if (src >= 0 && src < 32) {
        const int t = w * 32 + src;
        if (t < g_n && !g_fiber[t].done) r = g_slot[t];
    }
    wait(g_warp_barrier[w], g_alive_warp[w]);
    return r;
}

// Votes and the mask of participating lanes are taken between the same two
// barriers: a lane that leaves the kernel right after the vote must not change
// what a slower lane sees.
unsigned warp_ballot(bool pred, unsigned *active)
{
    need_coop("warp vote in a kernel launched in SIMPLE mode");
    const int w = g_cur / 32;
    g_pred[g_cur] = pred;
    wait(g_warp_barrier[w], g_alive_warp[w]);
    unsigned m = 0, act = 0;
    for (int l = 0; l < 32; ++l) {
        const int t = w * 32 + l;
        if (t < g_n && !g_fiber[t].done) {
            act |= 1u << l;
            if (g_pred[t]) m |= 1u << l;
        }
    }
    wait(g_warp_barrier[w], g_alive_warp[w]);
    if (active) *active = act;
    return m;
}

long long clock_ticks() { return g_ticks += 1000; }

void launch(int mode, dim3 grid, dim3 block, const std::function<void()> &body)
{
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1)
        die("only one-dimensional launches are emulated");
    if (block.x < 1 || block.x > unsigned(MAX_THREADS)) die("bad block size");
    g_gridDim = grid;
    g_blockDim = block;
    for (unsigned b = 0; b < grid.x; ++b) {
        g_blockIdx = uint3{b, 0, 0};
        if (mode == SIMPLE) {
            g_coop = false;
            for (unsigned t = 0; t < block.x; ++t) {
                g_threadIdx = uint3{t, 0, 0};
                body();
            }
            continue;
        }
        g_coop = true;
        g_n = int(block.x);
        g_body = &body;
        g_alive_block = g_n;
        g_block_barrier = Barrier();
        for (int w = 0; w < (g_n + 31) / 32; ++w) {
            g_alive_warp[w] = (g_n - w * 32 < 32) ? g_n - w * 32 : 32;
            g_warp_barrier[w] = Barrier();
        }
        for (int t = 0; t < g_n; ++t) {
            Fiber &f = g_fiber[t];
            if (!f.stack) f.stack = static_cast<char *>(malloc(STACK_BYTES));
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = STACK_BYTES;
            f.ctx.uc_link = &g_sched;
            f.done = false;
            makecontext(&f.ctx, trampoline, 0);
        }
        g_idle = 0;
        while (g_alive_block > 0) {
            for (int t = 0; t < g_n; ++t) {
                if (g_fiber[t].done) continue;
                g_cur = t;
                g_threadIdx = uint3{unsigned(t), 0, 0};
                swapcontext(&g_sched, &g_fiber[t].ctx);
            }
        }
        g_coop = false;
    }
}

}  // namespace plb_emu

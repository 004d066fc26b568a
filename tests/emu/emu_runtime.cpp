// TEST INFRASTRUCTURE (see include/cuda_runtime.h): the host-side stand-in for
// the CUDA runtime, the fiber executor behind simt.h, and an in-process
// stand-in for the few NCCL calls libplb makes.
//
// Streams are queues of operations (kernel launches, copies, event records /
// waits, sends / receives).  Whoever touches the runtime "pumps": it executes
// the head of every queue until nothing can make progress.  An operation can
// be BLOCKED -- an event that has not fired, a receive whose message has not
// been sent, a kernel that spins on a flag another rank has not raised yet --
// and then simply stays at the head of its queue.  That is all it takes to run
// several ranks (one host thread each, like one process each on the GPU box)
// in one process: a rank that synchronises while its neighbour has not issued
// its step yet waits on a condition variable until the neighbour's thread
// enqueues the work and pumps.  A wait that nobody ends is reported as a dead
// lock instead of hanging the test.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <deque>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <sys/mman.h>
#include <ucontext.h>
#include <vector>

#include <cuda_runtime.h>
#include <nccl.h>

namespace plb_emu {

uint3 g_threadIdx = {0, 0, 0}, g_blockIdx = {0, 0, 0};
dim3 g_blockDim(1, 1, 1), g_gridDim(1, 1, 1);

namespace {

std::mutex g_mu;                       // one big lock: the "device"
std::condition_variable g_cv;          // "something made progress"
using Lock = std::unique_lock<std::mutex>;
constexpr int DEADLOCK_SECONDS = 90;

[[noreturn]] void die(const char *what)
{
    fprintf(stderr, "plb_emu: %s (block %u, thread %u)\n", what, g_blockIdx.x,
            g_threadIdx.x);
    abort();
}

// ---------------------------------------------------------------------------
// operations and streams
// ---------------------------------------------------------------------------
struct Op {
    virtual ~Op() {}
    virtual bool run() = 0;            // false: blocked, try again later
};

}  // namespace
}  // namespace plb_emu

struct plb_emu_stream {
    std::deque<std::unique_ptr<plb_emu::Op>> q;
};
struct plb_emu_event {
    unsigned long long enqueued = 0, done = 0;   // sequence numbers of records
};

namespace plb_emu {
namespace {

std::vector<plb_emu_stream *> g_streams;
unsigned long long g_seq = 0;
std::map<char *, std::pair<char *, size_t>> g_allocs;   // pointer -> (mapping, bytes)

// Executes queue heads until nothing moves.  Caller holds the lock.
bool pump()
{
    bool any = false, progress = true;
    while (progress) {
        progress = false;
        for (size_t i = 0; i < g_streams.size(); ++i) {
            plb_emu_stream *s = g_streams[i];
            while (!s->q.empty()) {
                if (!s->q.front()->run()) break;
                s->q.pop_front();
                progress = any = true;
            }
        }
    }
    if (any) g_cv.notify_all();
    return any;
}

// Blocks the calling host thread until pred() holds, pumping in between.
template <typename Pred>
void host_wait(Lock &lk, Pred pred, const char *what)
{
    auto last_progress = std::chrono::steady_clock::now();
    for (;;) {
        if (pump()) last_progress = std::chrono::steady_clock::now();
        if (pred()) return;
        if (g_cv.wait_for(lk, std::chrono::milliseconds(50)) == std::cv_status::no_timeout)
            last_progress = std::chrono::steady_clock::now();
        if (std::chrono::steady_clock::now() - last_progress >
            std::chrono::seconds(DEADLOCK_SECONDS)) {
            fprintf(stderr, "plb_emu: dead lock while waiting for %s\n", what);
            abort();
        }
    }
}

void enqueue(plb_emu_stream *s, Op *op)
{
    Lock lk(g_mu);
    if (!s) {                            // legacy default stream: right now
        std::unique_ptr<Op> own(op);
        host_wait(lk, [&] { return own->run(); }, "an operation on the default stream");
        return;
    }
    s->q.emplace_back(op);
    pump();
    g_cv.notify_all();
}

struct FnOp : Op {
    std::function<bool()> fn;
    explicit FnOp(std::function<bool()> f) : fn(std::move(f)) {}
    bool run() override { return fn(); }
};

// ---------------------------------------------------------------------------
// SIMT fibers
// ---------------------------------------------------------------------------
constexpr int MAX_THREADS = 1024;
constexpr size_t STACK_BYTES = size_t(256) << 10;

struct Barrier {
    int count = 0;
    unsigned gen = 0;
};

// Fiber switch.  swapcontext() makes two signal-mask system calls per switch,
// which dominates a run that switches at every warp shuffle; on x86-64 a
// fiber is therefore just a saved stack pointer (callee-saved registers,
// MXCSR and the x87 control word live on its stack).
#if defined(__x86_64__)
struct Context {
    void *sp = nullptr;
};
extern "C" void plb_emu_switch(Context *from, Context *to);
asm(R"(
    .text
    .globl plb_emu_switch
    .type plb_emu_switch,@function
plb_emu_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    subq $8, %rsp
    stmxcsr (%rsp)
    fnstcw 4(%rsp)
    movq %rsp, (%rdi)
    movq (%rsi), %rsp
    ldmxcsr (%rsp)
    fldcw 4(%rsp)
    addq $8, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size plb_emu_switch,.-plb_emu_switch
)");
void context_init(Context &c, Context &, char *stack, size_t bytes, void (*entry)())
{
    uintptr_t top = (reinterpret_cast<uintptr_t>(stack) + bytes) & ~uintptr_t(15);
    uint64_t *sp = reinterpret_cast<uint64_t *>(top);
    *--sp = 0;                                       // return address of entry(): never used
    *--sp = reinterpret_cast<uint64_t>(entry);       // "return" of the first switch
    for (int i = 0; i < 6; ++i) *--sp = 0;           // rbp rbx r12 r13 r14 r15
    *--sp = 0x0000037F00001F80ull;                   // MXCSR | x87 control word << 32
    c.sp = sp;
}
void context_switch(Context &from, Context &to) { plb_emu_switch(&from, &to); }
#else
struct Context {
    ucontext_t uc;
};
void context_init(Context &c, Context &back, char *stack, size_t bytes, void (*entry)())
{
    getcontext(&c.uc);
    c.uc.uc_stack.ss_sp = stack;
    c.uc.uc_stack.ss_size = bytes;
    c.uc.uc_link = &back.uc;
    makecontext(&c.uc, entry, 0);
}
void context_switch(Context &from, Context &to) { swapcontext(&from.uc, &to.uc); }
#endif

struct Fiber {
    Context ctx;
    char *stack = nullptr;
    bool done = true;
};

bool g_coop = false, g_blocked = false;
int g_n = 0, g_cur = 0;
int g_alive_block = 0, g_alive_warp[MAX_THREADS / 32];
Barrier g_block_barrier, g_warp_barrier[MAX_THREADS / 32];
uint64_t g_slot[MAX_THREADS];
bool g_pred[MAX_THREADS];
Fiber g_fiber[MAX_THREADS];
Context g_sched;
const std::function<void()> *g_body = nullptr;
long long g_idle = 0, g_ticks = 0;

void yield()
{
    if (++g_idle > 50000000LL) die("dead lock: fibers wait at a barrier that cannot complete");
    context_switch(g_fiber[g_cur].ctx, g_sched);
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
    context_switch(g_fiber[me].ctx, g_sched);        // for good
    die("a finished fiber was resumed");
}

void need_coop(const char *what)
{
    if (!g_coop) die(what);
}

// One block; false if a thread of it went to sleep on a condition that only
// other work can change (the block is abandoned and will be run again).
bool run_block(int mode, unsigned b, dim3 block, const std::function<void()> &body)
{
    g_blockIdx = uint3{b, 0, 0};
    g_blocked = false;
    if (mode == SIMPLE) {
        g_coop = false;
        for (unsigned t = 0; t < block.x; ++t) {
            g_threadIdx = uint3{t, 0, 0};
            body();
        }
        return true;
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
        context_init(f.ctx, g_sched, f.stack, STACK_BYTES, trampoline);
        f.done = false;
    }
    g_idle = 0;
    while (g_alive_block > 0 && !g_blocked) {
        for (int t = 0; t < g_n && !g_blocked; ++t) {
            if (g_fiber[t].done) continue;
            g_cur = t;
            g_threadIdx = uint3{unsigned(t), 0, 0};
            context_switch(g_sched, g_fiber[t].ctx);
        }
    }
    g_coop = false;
    return !g_blocked;
}

struct LaunchOp : Op {
    int mode;
    dim3 grid, block;
    std::function<void()> body;
    unsigned next_block = 0;
    bool run() override
    {
        g_gridDim = grid;
        g_blockDim = block;
        // PLB_EMU_BLOCK_ORDER=reverse runs the blocks last to first: a result
        // that depends on the order has two writers for one location
        const char *order = getenv("PLB_EMU_BLOCK_ORDER");
        const bool reverse = order && order[0] == 'r';
        for (; next_block < grid.x; ++next_block) {
            const unsigned b = reverse ? grid.x - 1 - next_block : next_block;
            if (!run_block(mode, b, block, body)) return false;
        }
        return true;
    }
};

}  // namespace

void launch(int mode, dim3 grid, dim3 block, cudaStream_t stream,
            std::function<void()> body)
{
    if (grid.y != 1 || grid.z != 1 || block.y != 1 || block.z != 1)
        die("only one-dimensional launches are emulated");
    if (block.x < 1 || block.x > unsigned(MAX_THREADS)) die("bad block size");
    LaunchOp *op = new LaunchOp();
    op->mode = mode;
    op->grid = grid;
    op->block = block;
    op->body = std::move(body);
    enqueue(stream, op);
}

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

// A kernel thread sleeps because it polls memory that only other work (another
// rank's kernels) can change: give the block up, the launch is retried later.
void sleep_hook()
{
    need_coop("__nanosleep (a polling kernel) in a kernel launched in SIMPLE mode");
    g_blocked = true;
    context_switch(g_fiber[g_cur].ctx, g_sched);
    die("an abandoned fiber was resumed");
}

}  // namespace plb_emu

using plb_emu::enqueue;
using plb_emu::FnOp;
using plb_emu::g_mu;
using plb_emu::Lock;

// ===========================================================================
// "runtime": device memory is host memory
// ===========================================================================
extern "C" {

const char *cudaGetErrorString(cudaError_t e)
{
    return e == cudaSuccess ? "no error"
         : e == cudaErrorMemoryAllocation ? "out of memory"
         : e == cudaErrorNotSupported ? "not supported by the emulator"
         : "error";
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaMemGetInfo(size_t *free_b, size_t *total_b)
{
    *free_b = *total_b = size_t(180) << 30;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetAttribute(int *v, cudaDeviceAttr, int) { *v = 1000000; return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -1; return cudaSuccess; }
cudaError_t cudaDeviceGetPCIBusId(char *buf, int len, int)
{
    snprintf(buf, size_t(len), "0000:00:00.0");
    return cudaSuccess;
}
// "Device" allocations sit between two inaccessible guard pages, flush against
// the upper one (PLB_EMU_GUARD=lo: against the lower one), so that a kernel
// that reads or writes past an end of a buffer dies on the spot instead of
// quietly using a neighbouring allocation.  Contents start as a 0xA5 pattern:
// reading memory nobody wrote shows up as a wild value, like on the device.
cudaError_t plb_emu_malloc(void **p, size_t n)
{
    const size_t page = 4096;
    const size_t body = (std::max<size_t>(n, 1) + page - 1) / page * page;
    const size_t total = body + 2 * page;
    char *base = static_cast<char *>(mmap(nullptr, total, PROT_READ | PROT_WRITE,
                                          MAP_PRIVATE | MAP_ANONYMOUS, -1, 0));
    if (base == MAP_FAILED) return cudaErrorMemoryAllocation;
    mprotect(base, page, PROT_NONE);
    mprotect(base + page + body, page, PROT_NONE);
    const char *mode = getenv("PLB_EMU_GUARD");
    char *ptr = base + page;
    if (!(mode && mode[0] == 'l'))
        ptr += body - (std::max<size_t>(n, 1) + 15) / 16 * 16;   // 16-byte aligned end
    memset(ptr, 0xA5, n);
    {
        Lock lk(g_mu);
        plb_emu::g_allocs[ptr] = {base, total};
    }
    *p = ptr;
    return cudaSuccess;
}
cudaError_t plb_emu_host_alloc(void **p, size_t n)
{
    *p = malloc(n ? n : 1);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p)
{
    if (!p) return cudaSuccess;
    Lock lk(g_mu);
    plb_emu::pump();
    auto it = plb_emu::g_allocs.find(static_cast<char *>(p));
    if (it == plb_emu::g_allocs.end()) plb_emu::die("cudaFree of a pointer cudaMalloc did not return");
    munmap(it->second.first, it->second.second);
    plb_emu::g_allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n)
{
    Lock lk(g_mu);
    plb_emu::pump();
    memset(p, v, n);
    return cudaSuccess;
}
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t s)
{
    enqueue(s, new FnOp([=] { memset(p, v, n); return true; }));
    return cudaSuccess;
}
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind)
{
    Lock lk(g_mu);
    plb_emu::pump();
    memmove(d, s, n);
    return cudaSuccess;
}
cudaError_t cudaMemcpyAsync(void *d, const void *s, size_t n, cudaMemcpyKind, cudaStream_t st)
{
    enqueue(st, new FnOp([=] { memmove(d, s, n); return true; }));
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned)
{
    Lock lk(g_mu);
    *s = new plb_emu_stream();
    plb_emu::g_streams.push_back(*s);
    return cudaSuccess;
}
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned f, int)
{
    return cudaStreamCreateWithFlags(s, f);
}
cudaError_t cudaStreamSynchronize(cudaStream_t s)
{
    if (!s) return cudaSuccess;
    Lock lk(g_mu);
    plb_emu::host_wait(lk, [&] { return s->q.empty(); }, "cudaStreamSynchronize");
    return cudaSuccess;
}
cudaError_t cudaStreamDestroy(cudaStream_t s)
{
    cudaStreamSynchronize(s);
    Lock lk(g_mu);
    auto &v = plb_emu::g_streams;
    for (size_t i = 0; i < v.size(); ++i)
        if (v[i] == s) { v.erase(v.begin() + long(i)); break; }
    delete s;
    return cudaSuccess;
}
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new plb_emu_event(); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
// events are never freed: an operation still queued may refer to one
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s)
{
    unsigned long long seq;
    {
        Lock lk(g_mu);
        seq = e->enqueued = ++plb_emu::g_seq;
    }
    enqueue(s, new FnOp([=] { e->done = seq; return true; }));
    return cudaSuccess;
}
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned)
{
    unsigned long long target;
    {
        Lock lk(g_mu);
        target = e->enqueued;
    }
    if (target) enqueue(s, new FnOp([=] { return e->done >= target; }));
    return cudaSuccess;
}
cudaError_t cudaEventSynchronize(cudaEvent_t e)
{
    Lock lk(g_mu);
    plb_emu::host_wait(lk, [&] { return e->done >= e->enqueued; }, "cudaEventSynchronize");
    return cudaSuccess;
}
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 1.0f; return cudaSuccess; }
// "IPC": every rank lives in this address space, the handle is the pointer
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p)
{
    if (getenv("PLB_EMU_NO_IPC")) return cudaErrorNotSupported;
    memset(h, 0, sizeof *h);
    memcpy(h->reserved, &p, sizeof p);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned)
{
    memcpy(p, h.reserved, sizeof *p);
    return *p ? cudaSuccess : cudaErrorInvalidValue;
}
cudaError_t cudaIpcCloseMemHandle(void *) { return cudaSuccess; }

}  // extern "C"

// ===========================================================================
// NCCL stand-in: ranks are host threads of this process
// ===========================================================================
namespace {

struct Message {
    std::vector<char> bytes;
};
struct Group {
    int n = 0, joined = 0;
    std::map<std::pair<int, int>, std::deque<Message>> mail;   // (src, dst) -> FIFO
    struct Reduce {
        int arrived = 0, left = 0;
        std::vector<std::vector<char>> parts;
    };
    std::map<unsigned long long, Reduce> reductions;
};
std::map<std::string, Group *> g_groups;

size_t type_size(ncclDataType_t t)
{
    return t == ncclChar ? 1 : t == ncclInt ? 4 : 8;
}

struct PendingP2p {
    bool send;
    const void *src;
    void *dst;
    size_t bytes;
    int peer;
    ncclComm_t comm;
    cudaStream_t stream;
};
thread_local int t_group_depth = 0;
thread_local std::vector<PendingP2p> t_pending;

}  // namespace

struct ncclComm {
    Group *group;
    int rank;
    unsigned long long reduce_seq = 0;
};

namespace {

void issue(const PendingP2p &p)
{
    Group *g = p.comm->group;
    const int me = p.comm->rank;
    if (p.send) {
        enqueue(p.stream, new FnOp([=] {
            Message m;
            m.bytes.assign(static_cast<const char *>(p.src),
                           static_cast<const char *>(p.src) + p.bytes);
            g->mail[{me, p.peer}].push_back(std::move(m));
            return true;
        }));
    } else {
        enqueue(p.stream, new FnOp([=] {
            auto &box = g->mail[{p.peer, me}];
            if (box.empty()) return false;               // not sent yet
            if (box.front().bytes.size() != p.bytes)
                plb_emu::die("ncclRecv size differs from the matching ncclSend");
            memcpy(p.dst, box.front().bytes.data(), p.bytes);
            box.pop_front();
            return true;
        }));
    }
}

}  // namespace

extern "C" {

ncclResult_t ncclGetUniqueId(ncclUniqueId *id)
{
    static unsigned long long counter = 0;
    Lock lk(g_mu);
    memset(id, 0, sizeof *id);
    snprintf(id->internal, sizeof id->internal, "plb-emu-%llu", ++counter);
    return ncclSuccess;
}
ncclResult_t ncclCommInitRank(ncclComm_t *comm, int n, ncclUniqueId id, int rank)
{
    Lock lk(g_mu);
    Group *&g = g_groups[std::string(id.internal)];
    if (!g) {
        g = new Group();
        g->n = n;
    }
    ++g->joined;
    plb_emu::g_cv.notify_all();
    Group *group = g;
    plb_emu::host_wait(lk, [&] { return group->joined >= group->n; }, "ncclCommInitRank");
    *comm = new ncclComm{group, rank};
    return ncclSuccess;
}
ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
ncclResult_t ncclGroupStart(void) { ++t_group_depth; return ncclSuccess; }
// inside a group all sends are issued before the receives, like NCCL, which
// progresses the operations of a group together
ncclResult_t ncclGroupEnd(void)
{
    if (--t_group_depth > 0) return ncclSuccess;
    std::vector<PendingP2p> ops;
    ops.swap(t_pending);
    for (const PendingP2p &p : ops) if (p.send) issue(p);
    for (const PendingP2p &p : ops) if (!p.send) issue(p);
    return ncclSuccess;
}
ncclResult_t ncclSend(const void *buf, size_t count, ncclDataType_t t, int peer,
                      ncclComm_t comm, cudaStream_t s)
{
    PendingP2p p{true, buf, nullptr, count * type_size(t), peer, comm, s};
    if (t_group_depth > 0) t_pending.push_back(p);
    else issue(p);
    return ncclSuccess;
}
ncclResult_t ncclRecv(void *buf, size_t count, ncclDataType_t t, int peer,
                      ncclComm_t comm, cudaStream_t s)
{
    PendingP2p p{false, nullptr, buf, count * type_size(t), peer, comm, s};
    if (t_group_depth > 0) t_pending.push_back(p);
    else issue(p);
    return ncclSuccess;
}
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t t,
                           ncclRedOp_t op, ncclComm_t comm, cudaStream_t s)
{
    if (t != ncclInt && t != ncclDouble) return ncclUnhandledCudaError;
    Group *g = comm->group;
    const unsigned long long seq = comm->reduce_seq++;
    const size_t bytes = count * type_size(t);
    auto deposited = std::make_shared<bool>(false);
    enqueue(s, new FnOp([=] {
        Group::Reduce &r = g->reductions[seq];
        if (!*deposited) {
            r.parts.emplace_back(static_cast<const char *>(send),
                                 static_cast<const char *>(send) + bytes);
            ++r.arrived;
            r.left = g->n;
            *deposited = true;
        }
        if (r.arrived < g->n) return false;
        for (size_t i = 0; i < count; ++i) {
            if (t == ncclInt) {
                int acc = 0;
                for (int k = 0; k < g->n; ++k) {
                    int v;
                    memcpy(&v, r.parts[size_t(k)].data() + 4 * i, 4);
                    acc = k == 0 ? v : (op == ncclMin ? (v < acc ? v : acc) : acc + v);
                }
                memcpy(static_cast<char *>(recv) + 4 * i, &acc, 4);
            } else {
                double acc = 0;
                for (int k = 0; k < g->n; ++k) {
                    double v;
                    memcpy(&v, r.parts[size_t(k)].data() + 8 * i, 8);
                    acc = k == 0 ? v : (op == ncclMin ? (v < acc ? v : acc) : acc + v);
                }
                memcpy(static_cast<char *>(recv) + 8 * i, &acc, 8);
            }
        }
        return true;
    }));
    return ncclSuccess;
}
const char *ncclGetErrorString(ncclResult_t r) { return r == ncclSuccess ? "no error" : "error"; }

}  // extern "C"

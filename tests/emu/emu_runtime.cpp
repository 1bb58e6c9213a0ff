// emu_runtime.cpp -- TEST INFRASTRUCTURE ONLY (see cuda_emu.h): fiber-based SIMT executor and fake CUDA runtime.
//
// Execution model: a launch runs its CTAs one after the other in blockIdx order; the threads of a CTA are ucontext
// fibers scheduled round-robin, switching only at __syncthreads / warp-synchronous intrinsics.  "Device" memory is host
// memory with red zones (checked on free and on every synchronize) and a 0xFF fill so that reads of uninitialised
// device memory surface as NaNs.  With LQCD_EMU_SHM=1 allocations live in POSIX shared memory and cudaIpc* handles
// carry the segment name, so multi-rank tests can run as separate processes exactly like the GPU ranks do.
#include "cuda_emu.h"
#include <cstdarg>
#include <fcntl.h>
#include <map>
#include <string>
#include <sys/mman.h>
#include <sys/stat.h>
#include <time.h>
#include <ucontext.h>
#include <unistd.h>
#include <vector>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

namespace {
const size_t STACK_BYTES = 256 << 10;
const int MAX_THREADS = 1024;
const size_t REDZONE = 256;

[[noreturn]] void die(const char *fmt, ...) __attribute__((format(printf, 1, 2)));
void die(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    fprintf(stderr, "[cuda_emu] FATAL: ");
    vfprintf(stderr, fmt, ap);
    fprintf(stderr, "\n");
    va_end(ap);
    fflush(stderr);
    abort();
}

// Context switch: a six-register x86-64 stack switch (swapcontext costs two sigprocmask system calls per switch, which
// dominated the run time); AddressSanitizer builds keep ucontext, which ASan intercepts and understands.
#if defined(__x86_64__) && !defined(__SANITIZE_ADDRESS__)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
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
.size emu_switch,.-emu_switch
)");
#else
#define EMU_FAST_SWITCH 0
#endif

enum Wait { RUN = 0, WAIT_CTA = 1, WAIT_WARP = 2, DONE = 3 };
struct Fiber {
#if EMU_FAST_SWITCH
    void *sp;
#else
    ucontext_t ctx;
#endif
    int state;
    unsigned gen;       // generation of the barrier it waits on
};
struct Cta {
    int nthreads, alive;
    Fiber fib[MAX_THREADS];
    unsigned cta_gen, cta_count;
    unsigned warp_gen[32], warp_count[32], warp_alive[32];
    uint64_t warp_buf[32][32];
    int cur;
#if EMU_FAST_SWITCH
    void *sched_sp;
#else
    ucontext_t sched;
#endif
    const std::function<void()> *body;
};
Cta *C = nullptr;
char *stacks = nullptr;
std::vector<unsigned char> dyn_smem_buf;
unsigned char *dyn_smem_ptr = nullptr;
cudaError_t last_error = cudaSuccess;

void set_tid(int t) {
    threadIdx.x = t % blockDim.x;
    threadIdx.y = (t / blockDim.x) % blockDim.y;
    threadIdx.z = t / (blockDim.x * blockDim.y);
}

void to_sched(Fiber &f) {
#if EMU_FAST_SWITCH
    emu_switch(&f.sp, C->sched_sp);
#else
    swapcontext(&f.ctx, &C->sched);
#endif
}
void to_fiber(Fiber &f) {
#if EMU_FAST_SWITCH
    emu_switch(&C->sched_sp, f.sp);
#else
    swapcontext(&C->sched, &f.ctx);
#endif
}

void trampoline() {
    (*C->body)();
    Fiber &f = C->fib[C->cur];
    f.state = DONE;
    to_sched(f);
    die("resumed a finished fiber");
}

void yield_fiber() {
    Fiber &f = C->fib[C->cur];
    const int me = C->cur;
    to_sched(f);
    C->cur = me;
    set_tid(me);
}

void release_barriers() {
    if (C->cta_count > 0 && (int)C->cta_count == C->alive) { C->cta_gen++; C->cta_count = 0; }
    const int nw = (C->nthreads + 31) / 32;
    for (int w = 0; w < nw; w++)
        if (C->warp_count[w] > 0 && C->warp_count[w] == C->warp_alive[w]) { C->warp_gen[w]++; C->warp_count[w] = 0; }
}

void run_cta() {
    const int n = C->nthreads;
    C->alive = n;
    C->cta_gen = 0; C->cta_count = 0;
    for (int w = 0; w < 32; w++) { C->warp_gen[w] = 0; C->warp_count[w] = 0; C->warp_alive[w] = 0; }
    for (int t = 0; t < n; t++) {
        Fiber &f = C->fib[t];
#if EMU_FAST_SWITCH
        {   // initial frame: six callee-saved registers, then the entry address that emu_switch's `ret` jumps to; the
            // stack pointer at entry is 8 mod 16 as after a call
            uintptr_t top = ((uintptr_t)(stacks + (size_t)(t + 1) * STACK_BYTES)) & ~(uintptr_t)15;
            void **sp = (void **)top;
            *--sp = nullptr;                       // fake return address of trampoline()
            *--sp = (void *)&trampoline;
            for (int k = 0; k < 6; k++) *--sp = nullptr;
            f.sp = sp;
        }
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = stacks + (size_t)t * STACK_BYTES;
        f.ctx.uc_stack.ss_size = STACK_BYTES;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, trampoline, 0);
#endif
        f.state = RUN; f.gen = 0;
        C->warp_alive[t >> 5]++;
    }
    // LQCD_EMU_THREAD_ORDER=reverse: the runnable threads of a CTA are resumed from the highest thread index down (a missing
    // __syncthreads / __syncwarp between a shared-memory write and a read by another thread shows up in one of the two orders)
    static int trev = -1;
    if (trev < 0) { const char *e = getenv("LQCD_EMU_THREAD_ORDER"); trev = (e && e[0] == 'r') ? 1 : 0; }
    while (C->alive > 0) {
        bool progress = false;
        for (int tt = 0; tt < n; tt++) {
            const int t = trev ? n - 1 - tt : tt;
            Fiber &f = C->fib[t];
            if (f.state == DONE) continue;
            if (f.state == WAIT_CTA) { if (f.gen == C->cta_gen) continue; f.state = RUN; }
            if (f.state == WAIT_WARP) { if (f.gen == C->warp_gen[t >> 5]) continue; f.state = RUN; }
            C->cur = t;
            set_tid(t);
            to_fiber(f);
            progress = true;
            if (f.state == DONE) { C->alive--; C->warp_alive[t >> 5]--; }
            release_barriers();
        }
        if (!progress)
            die("deadlock in CTA %u of %u: %d threads alive, %u at __syncthreads (divergent barrier?)", blockIdx.x, gridDim.x,
                C->alive, C->cta_count);
    }
}
}   // namespace

namespace emu {
// LQCD_EMU_TRACE=1: per-kernel launch histogram on stderr at exit (which kernels a test really exercised)
static std::map<std::string, long> launch_hist;
static void dump_hist() {
    for (auto &kv : launch_hist) fprintf(stderr, "[cuda_emu] launches %-40s %ld\n", kv.first.c_str(), kv.second);
}
void launch(dim3 grid, dim3 block, size_t dyn_smem, cudaStream_t, const std::function<void()> &body, const char *name) {
    static int trace = -1;
    if (trace < 0) { const char *e = getenv("LQCD_EMU_TRACE"); trace = e && atoi(e) != 0; if (trace) atexit(dump_hist); }
    if (trace) launch_hist[name]++;
    const size_t nthreads = (size_t)block.x * block.y * block.z;
    if (nthreads == 0 || nthreads > (size_t)MAX_THREADS || grid.x == 0 || dyn_smem > (227u << 10)) { last_error = cudaErrorInvalidValue; return; }
    if (C && C->body) die("nested kernel launch");
    if (!C) {
        C = new Cta();
        stacks = (char *)mmap(nullptr, STACK_BYTES * MAX_THREADS, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (stacks == MAP_FAILED) die("mmap of fiber stacks failed");
    }
    dyn_smem_buf.resize(dyn_smem + 256);
    dyn_smem_ptr = (unsigned char *)(((uintptr_t)dyn_smem_buf.data() + 127) & ~(uintptr_t)127);
    C->nthreads = (int)nthreads;
    C->body = &body;
    blockDim = block; gridDim = grid;
    // LQCD_EMU_CTA_ORDER=reverse: run the CTAs of every launch in DESCENDING blockIdx order.  A kernel whose CTAs exchange data
    // within one launch (a CTA reading what another one writes: in-place stencils, missing double buffering) gives different
    // results in the two orders; kernels that rely on dispatch order across ranks (halo flags) are single-rank-only in this mode.
    static int reverse = -1;
    if (reverse < 0) { const char *e = getenv("LQCD_EMU_CTA_ORDER"); reverse = (e && e[0] == 'r') ? 1 : 0; }
    const unsigned long long ncta = (unsigned long long)grid.x * grid.y * grid.z;
    for (unsigned long long i = 0; i < ncta; i++) {
        const unsigned long long c = reverse ? ncta - 1 - i : i;
        blockIdx.x = (unsigned)(c % grid.x); blockIdx.y = (unsigned)((c / grid.x) % grid.y); blockIdx.z = (unsigned)(c / ((unsigned long long)grid.x * grid.y));
        if (dyn_smem) memset(dyn_smem_ptr, 0xFF, dyn_smem);      // shared memory starts uninitialised
        run_cta();
    }
    C->body = nullptr;
}
void cta_barrier() {
    Fiber &f = C->fib[C->cur];
    f.state = WAIT_CTA; f.gen = C->cta_gen; C->cta_count++;
    yield_fiber();
}
void warp_barrier() {
    Fiber &f = C->fib[C->cur];
    const int w = C->cur >> 5;
    f.state = WAIT_WARP; f.gen = C->warp_gen[w]; C->warp_count[w]++;
    yield_fiber();
}
void yield_now() { yield_fiber(); }      // stay runnable, let the other fibers of the CTA run (spin loops on CTA-local state)
uint64_t *warp_slot(int lane) { return &C->warp_buf[C->cur >> 5][lane & 31]; }
unsigned char *dynamic_smem() { return dyn_smem_ptr; }
int lane_id() { return C->cur & 31; }

// mbarrier model: arrivals + transaction bytes; the phase flips when both reach zero.  LQCD_EMU_BULK=late defers every
// bulk copy until a thread actually waits on its barrier (the latest legal completion), =early (default) performs it at
// issue (the earliest): running a kernel under both catches missing waits and premature slot reuse.
struct PendingCopy { void *dst; const void *src; uint32_t bytes; uint64_t *bar; };
static std::vector<PendingCopy> pending;
static bool bulk_late() { static int v = -1; if (v < 0) { const char *e = getenv("LQCD_EMU_BULK"); v = e && !strcmp(e, "late"); } return v; }
}   // namespace emu

// The kernel reserves 8 bytes per mbarrier; the model needs 16, so the state lives in a side table keyed by address.
namespace {
std::map<uint64_t *, emu::MBar> mbars;
void mbar_try_flip(emu::MBar &m) {
    if (m.arrived >= m.expected && m.tx_pending == 0) { m.phase ^= 1u; m.arrived = 0; }
}
void complete_copy(const emu::PendingCopy &c) {
    memcpy(c.dst, c.src, c.bytes);
    emu::MBar &m = mbars[c.bar];
    if (m.tx_pending < c.bytes) die("mbarrier transaction underflow");
    m.tx_pending -= c.bytes;
    mbar_try_flip(m);
}
}   // namespace
namespace emu {
void mbar_init(uint64_t *bar, uint32_t count) { MBar m = {count, 0, 0, 0}; mbars[bar] = m; }
void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    auto it = mbars.find(bar);
    if (it == mbars.end()) die("mbarrier used before init");
    it->second.tx_pending += bytes;
    it->second.arrived++;
    // no flip here: the expected bytes are still outstanding
}
void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    if (bytes % 16 != 0 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) die("cp.async.bulk needs 16-byte aligned size and addresses");
    PendingCopy c = {dst, src, bytes, bar};
    if (bulk_late()) pending.push_back(c); else complete_copy(c);
}
void mbar_wait(uint64_t *bar, uint32_t parity) {
    auto it = mbars.find(bar);
    if (it == mbars.end()) die("mbarrier waited on before init");
    for (int spin = 0;; spin++) {
        if (it->second.phase != parity) return;               // phase `parity` has completed
        // complete deferred copies that target this barrier, then let other warps run
        bool any = false;
        for (size_t i = 0; i < pending.size();)
            if (pending[i].bar == bar) { PendingCopy c = pending[i]; pending.erase(pending.begin() + i); complete_copy(c); any = true; }
            else i++;
        if (it->second.phase != parity) return;
        if (!any) {
            if (spin > 100000) die("mbarrier wait never completes (missing arrive / copy?)");
            emu::yield_now();
        }
    }
}
}   // namespace emu

// ---- fake runtime -----------------------------------------------------------------------------------------------------
namespace {
struct Alloc { size_t bytes; char *raw; size_t raw_bytes; std::string shm; bool host; };
std::map<void *, Alloc> allocs;
std::map<void *, std::pair<size_t, size_t>> ipc_maps;      // mapped peer pointer -> (raw base offset, raw bytes)
int shm_counter = 0;
bool use_shm() { static int v = -1; if (v < 0) { const char *e = getenv("LQCD_EMU_SHM"); v = e && atoi(e) != 0; } return v; }

void check_redzones() {
    for (auto &kv : allocs) {
        const unsigned char *lo = (const unsigned char *)kv.second.raw, *hi = (const unsigned char *)kv.first + kv.second.bytes;
        for (size_t i = 0; i < REDZONE; i++)
            if (lo[i] != 0xA5 || hi[i] != 0xA5)
                die("out-of-bounds WRITE detected next to a %zu-byte device allocation (%s red zone, offset %zu)", kv.second.bytes,
                    lo[i] != 0xA5 ? "lower" : "upper", i);
    }
}
double now_ms() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6; }

void unlink_all_shm() {
    for (auto &kv : allocs)
        if (!kv.second.shm.empty()) shm_unlink(kv.second.shm.c_str());
}

cudaError_t alloc_common(void **p, size_t bytes, bool host) {
    if (!p) return cudaErrorInvalidValue;
    static bool registered = false;
    if (!registered) { registered = true; atexit(unlink_all_shm); }
    const size_t raw_bytes = ((bytes + 255) & ~(size_t)255) + 2 * REDZONE + 256;
    Alloc a; a.bytes = bytes; a.raw_bytes = raw_bytes; a.host = host;
    if (use_shm() && !host) {
        char name[64];
        snprintf(name, sizeof name, "/lqcd_emu_%d_%d", (int)getpid(), shm_counter++);
        int fd = shm_open(name, O_CREAT | O_RDWR | O_EXCL, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)raw_bytes) != 0) return cudaErrorMemoryAllocation;
        a.raw = (char *)mmap(nullptr, raw_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (a.raw == MAP_FAILED) return cudaErrorMemoryAllocation;
        a.shm = name;
    } else {
        void *r = nullptr;
        if (posix_memalign(&r, 256, raw_bytes) != 0) return cudaErrorMemoryAllocation;
        a.raw = (char *)r;
    }
    memset(a.raw, 0xA5, raw_bytes);
    char *user = a.raw + REDZONE;
    memset(user, 0xFF, bytes);
    allocs[user] = a;
    *p = user;
    return cudaSuccess;
}
}   // namespace

struct emu_stream { int id; };
struct emu_event { double t; };

const char *cudaGetErrorString(cudaError_t e) {
    switch (e) {
    case cudaSuccess: return "no error";
    case cudaErrorInvalidValue: return "invalid argument (emu)";
    case cudaErrorMemoryAllocation: return "out of memory (emu)";
    case cudaErrorNotSupported: return "operation not supported (emu)";
    default: return "unknown error (emu)";
    }
}
cudaError_t cudaGetLastError() { cudaError_t e = last_error; last_error = cudaSuccess; return e; }
cudaError_t cudaGetDeviceCount(int *n) { *n = 8; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDeviceProperties(cudaDeviceProp *p, int) {
    memset(p, 0, sizeof *p);
    p->major = 10; p->minor = 0; p->multiProcessorCount = 148; p->clockRate = 1965000; p->sharedMemPerBlockOptin = 227 << 10;
    // LQCD_EMU_SMS: pretend to have fewer SMs, so that persistent kernels (grid = a multiple of the SM count) loop over several
    // tasks per CTA on the small lattices the emulation can afford
    if (const char *e = getenv("LQCD_EMU_SMS")) { const int v = atoi(e); if (v >= 1 && v <= 148) p->multiProcessorCount = v; }
    snprintf(p->name, sizeof p->name, "emulated sm_100 (tests/emu)");
    return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { check_redzones(); return cudaSuccess; }
cudaError_t cudaDeviceGetStreamPriorityRange(int *lo, int *hi) { *lo = 0; *hi = -5; return cudaSuccess; }
cudaError_t cudaDeviceEnablePeerAccess(int, unsigned) { return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned) { *s = new emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned, int) { *s = new emu_stream(); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { delete s; return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { check_redzones(); return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = new emu_event(); (*e)->t = 0; return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned) { return cudaEventCreate(e); }
cudaError_t cudaEventDestroy(cudaEvent_t e) { delete e; return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) { e->t = now_ms(); return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t a, cudaEvent_t b) { *ms = (float)(b->t - a->t); return cudaSuccess; }
cudaError_t emu_malloc(void **p, size_t bytes) { return alloc_common(p, bytes, false); }
cudaError_t emu_malloc_host(void **p, size_t bytes) { return alloc_common(p, bytes, true); }
cudaError_t cudaFree(void *p) {
    if (!p) return cudaSuccess;
    auto it = allocs.find(p);
    if (it == allocs.end()) die("cudaFree of a pointer that was never allocated (or double free): %p", p);
    check_redzones();
    if (!it->second.shm.empty()) { munmap(it->second.raw, it->second.raw_bytes); shm_unlink(it->second.shm.c_str()); }
    else free(it->second.raw);
    allocs.erase(it);
    return cudaSuccess;
}
cudaError_t cudaFreeHost(void *p) { return cudaFree(p); }
cudaError_t cudaHostRegister(void *, size_t, unsigned) { return cudaSuccess; }
cudaError_t cudaHostUnregister(void *) { return cudaSuccess; }
cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }

// IPC: the handle carries the shared-memory segment name and the user offset (needs LQCD_EMU_SHM=1)
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t *h, void *p) {
    auto it = allocs.find(p);
    if (it == allocs.end() || it->second.shm.empty()) return cudaErrorNotSupported;
    memset(h, 0, sizeof *h);
    snprintf(h->reserved, sizeof h->reserved, "%s|%zu", it->second.shm.c_str(), it->second.raw_bytes);
    return cudaSuccess;
}
cudaError_t cudaIpcOpenMemHandle(void **p, cudaIpcMemHandle_t h, unsigned) {
    char name[64];
    size_t raw_bytes = 0;
    const char *bar = strchr(h.reserved, '|');
    if (!bar || (size_t)(bar - h.reserved) >= sizeof name) return cudaErrorInvalidValue;
    memcpy(name, h.reserved, bar - h.reserved);
    name[bar - h.reserved] = 0;
    raw_bytes = strtoull(bar + 1, nullptr, 10);
    int fd = shm_open(name, O_RDWR, 0600);
    if (fd < 0) return cudaErrorInvalidValue;
    char *raw = (char *)mmap(nullptr, raw_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (raw == MAP_FAILED) return cudaErrorMemoryAllocation;
    *p = raw + REDZONE;
    ipc_maps[*p] = std::make_pair(REDZONE, raw_bytes);
    return cudaSuccess;
}
cudaError_t cudaIpcCloseMemHandle(void *p) {
    auto it = ipc_maps.find(p);
    if (it == ipc_maps.end()) return cudaErrorInvalidValue;
    munmap((char *)p - it->second.first, it->second.second);
    ipc_maps.erase(it);
    return cudaSuccess;
}

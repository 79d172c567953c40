// xb200_chain.cu -- host side of the picture-level decision pass (xb200_analyze_picture and friends, include/xeve_b200.h).
// Own translation unit with its own (static) copies of the constant tables, like xb200_intra.cu.
//
// xb200_analyze_picture only QUEUES a picture.  One scheduler thread per device (shared by every context of the process) owns the
// chain server -- a long-lived grid of chain workers with a device-side task queue (xb200_chain.cuh: k_chain_server) -- and publishes
// a picture to it when its reference pictures are complete and the workers can take all its chains; when the last chain has finished
// (a word in host-mapped memory) it launches the optional copy of the unfiltered reconstruction, both loop-filter passes and the
// border expansion on one of the context's streams, whose completion callback makes the picture a usable reference.  The host
// never waits per CU or CTU; xb200_picture_fetch waits per picture.  Rules that keep a grid that outlives its work from dead-locking
// are collected in DESIGN.md section 2 and next to server_init / xb200_chain_free below.
#define XB200_CONST_LINKAGE static
#define XB200_CHAIN_TU
#define XB200_T64 128
#include "xb200_ctx.h"
#include "xb200_chain.cuh"
#include <math.h>
#include <deque>
#include <map>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

#ifdef XB200_CHAIN_DEBUG
static volatile int *h_dbg_words;   // host view of the kernel's progress words (host-mapped memory)
#endif

namespace {

constexpr int N_STREAMS = 64;

struct Job;
struct PicMaps {            // frame maps of a decided (or adopted) picture; heap-allocated, the address is stable for the context's life
    uint32_t *scu = nullptr;
    int8_t   *ipm = nullptr, *refi = nullptr;
    int16_t  *mv = nullptr;
    uint8_t  *flags = nullptr;
    uint64_t  issued = 0;                 // pictures enqueued into this handle so far (its "lives")
    std::atomic<uint64_t> ready{0};       // lives that are complete: decided, filtered, border-expanded -- a usable reference picture
};
struct JobBufs {            // outputs and working set of one picture in flight (recycled)
    xb200_scu_rec *scu = nullptr;
    int16_t       *coef = nullptr;
    ChState       *ctu_state = nullptr;
    double        *ctu_cost = nullptr;
    int           *done = nullptr;
    unsigned long long *counts = nullptr;
    ChainWs       *ws = nullptr;
    int            ws_chains = 0;
    xb200_cu_item    *cu_log = nullptr;
    xb200_intra_item *intra_log = nullptr;
    long long      cu_cap = 0, intra_cap = 0;
    cudaEvent_t    ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
};
struct Dep { PicMaps *m; uint64_t life; };   // satisfied when m->ready >= life
enum { JOB_PENDING = 0, JOB_RUNNING = 1, JOB_CHAIN_DONE = 2, JOB_READY = 3, JOB_FAILED = 4 };
struct Job {
    JobBufs     b;
    xb200_ctx  *c = nullptr;
    int         rec_pic = -1;
    int         n_chain = 0, bps = 1;   // its CTAs, and how many of them an SM can hold (registers / shared memory of its kernel variant)
    int         b_fast = 1, b_dense = 1;   // CTAs per SM of k_chain<2> / k_chain<3> with this picture's shared memory
    bool        dense = false;
    ChainPic    P;                      // kernel arguments, complete at enqueue
    size_t      smem = 0;
    Pic         rp, up;                 // copies of the context's picture descriptors (the table may grow while the job waits)
    bool        has_up = false, deblock = false;
    PicMaps    *m = nullptr;
    uint64_t    life = 0;
    Dep         deps[2 * XB200_MAX_REFP + 1];
    int         n_dep = 0;
    int         stream_no = -1;
    int         slot = -1;              // picture slot of the chain server while its chains run
    unsigned    seq = 0;
    unsigned long long t_chain[2] = {0, 0};   // %globaltimer: first chain started, last chain finished
    std::atomic<int> state{JOB_PENDING};
};

// ---- the device scheduler ------------------------------------------------------------------------------------------------------------
// One scheduler thread per device owns the chain server (xb200_chain.cuh) and serves every context of the process:
// xb200_analyze_picture only queues the picture.  A picture is PUBLISHED to the workers when its reference pictures are complete
// (host side; a chain that has to wait for a reference would hold a worker) and the workers can take all its chains at once; when
// its last chain has finished (a word in host-mapped memory, polled) the loop filter and the border expansion are launched on one of
// the context's streams, and their completion callback makes the picture a usable reference.
struct DeviceSched {
    std::mutex              mu;              // scheduler state below; held by the scheduler thread while it launches
    std::condition_variable cv_done;         // a picture became complete (waiters: xb200_picture_fetch & co.)
    // Completion callbacks run on a CUDA-internal thread and must never wait for a thread that is inside a CUDA call: they only append
    // to `events` under their own small mutex, which no one holds across a CUDA call.
    std::mutex              cb_mu;
    std::condition_variable cv_work;
    std::vector<std::pair<Job *, int>> events;   // {job, 0: decision kernel finished | 1: picture complete}
    std::deque<Job *>       pending;         // enqueue order, all contexts
    std::vector<Job *>      running;         // published to the chain server, chains not all finished
    int                     chains = 0;      // chains published and not finished
    int                     n_ctx = 0;       // contexts with a decision pass on this device
    int                     drain = 0;       // contexts being torn down: nothing is published, the grid ends when its chains have finished
    // the chain server (xb200_chain.cuh: k_chain_server)
    ChainQueue             *dq = nullptr;    // device queue
    volatile unsigned      *h_done = nullptr;// host-mapped completion words, one per picture slot
    unsigned               *d_done = nullptr;// ... their device address
    ChainPicTask           *h_slot = nullptr;// pinned staging of the picture records
    ChainTask              *h_tasks = nullptr;   // pinned mirror of the task ring
    unsigned               *h_tail = nullptr;    // pinned ring of published tail values (sources of the async copies)
    int                    *h_flag = nullptr;    // pinned {0, 1}
    unsigned long long     *h_times = nullptr;   // pinned [slot][2]: device timer when the first chain started / the last one finished
    int8_t                 *d_tm64 = nullptr;
    cudaStream_t            feed = nullptr, srv = nullptr;
    bool                    server_on = false;
    size_t                  server_smem = 0;
    int                     workers = 0;
    unsigned                tail = 0, seq = 0;
    std::vector<int>        free_slots;
    int                     sms = 0, device = 0;
    bool                    wake = false, stop = false, started = false;   // wake / stop: guarded by cb_mu
    std::thread             th;
};
// Never destroyed: the (detached) scheduler threads wait on these condition variables until the process exits, and destroying a
// condition variable that has a waiter blocks (glibc) -- a static array's destructor would hang every process at exit.
DeviceSched *const g_sched = new DeviceSched[64];

struct ChainCtx {
    std::mutex   mu;                    // host-side bookkeeping: one thread may enqueue pictures while another fetches results
    cudaStream_t streams[N_STREAMS] = {};
    int          stream_busy[N_STREAMS] = {};   // pictures launched on the stream and not yet complete (guarded by the scheduler mutex)
    cudaStream_t copy = nullptr;
    std::vector<PicMaps *> maps;
    std::map<int, Job *> jobs;          // by rec_pic
    std::vector<JobBufs> pool;
    int16_t     *zero_mv = nullptr;     // colocated map of a reference picture without one (all zero)
    long long    log_cu = 0, log_intra = 0;
    cudaEvent_t  ev_span0 = nullptr;    // device time span of a batch of pictures: first launch after a reset ...
    bool         span_on = false;
    float        span_ms = 0.f;         // ... to the latest completion among the pictures fetched since
    bool         ready = false;
    int          n_lcu = 0, w_lcu = 0, h_lcu = 0, w_scu = 0, h_scu = 0;
    size_t       f_scu = 0;
};

ChainCtx *cc_of(xb200_ctx *c) { return static_cast<ChainCtx *>(c->chain); }
void sched_thread(DeviceSched *D);
void sched_atexit();

int chain_init(xb200_ctx *c)
{
    if(c->chain && cc_of(c)->ready) return XB200_OK;
    ChainCtx *k = c->chain ? cc_of(c) : new ChainCtx();
    c->chain = k;
    k->w_lcu = (c->seq.w + 63) >> 6; k->h_lcu = (c->seq.h + 63) >> 6; k->n_lcu = k->w_lcu * k->h_lcu;
    k->w_scu = (c->seq.w + 3) >> 2; k->h_scu = (c->seq.h + 3) >> 2; k->f_scu = (size_t)k->w_scu * k->h_scu;
    for(int i = 0; i < N_STREAMS; i++) CK(cudaStreamCreateWithFlags(&k->streams[i], cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&k->copy, cudaStreamNonBlocking));
    CK(cudaEventCreate(&k->ev_span0));
    // Device-wide state of this translation unit -- constant tables, the worker kernel's shared-memory attribute -- is set ONCE per
    // device, by the first context: a later context may be created while the chain server is alive, and nothing it does may wait for
    // the device (cudaMemcpyToSymbol, cudaMemset and cudaDeviceSynchronize do, one way or another; the worker grid only ends when the
    // scheduler says so).
    static std::mutex once_mu;
    static bool       once_done[64] = {};
    {
    std::lock_guard<std::mutex> ol(once_mu);
    if(!once_done[c->device & 63]) {
    {   // constant tables of this translation unit
        static int8_t tm[64 * 64];
        xb200_gen_tm64(tm);
        const int16_t l[4][8] = XB200_MC_L_TAPS;
        const int16_t ch[8][4] = XB200_MC_C_TAPS;
        const int32_t qs[6] = XB200_QUANT_SCALE, dq[6] = XB200_DEQUANT_SCALE;
        static int64_t es[7][6][7];
        for(int b = 0; b < 7; b++)
            for(int q = 0; q < 6; q++)
                for(int l2 = 0; l2 < 7; l2++) es[b][q][l2] = xb200_err_scale(q, l2, b + 8);
        static const uint8_t mpm[6][6][5] = XB200_MPM_TABLE;
        static uint16_t scan[16 + 64 + 256 + 1024 + 4096];
        static int32_t  eb[1024];
        int             off = 0;
        for(int l2 = 2; l2 <= 6; l2++) { xb200_gen_scan(scan + off, l2, l2); off += 1 << (2 * l2); }
        for(int i = 0; i < 1024; i++) {     // xeve_init_bits_est (src_base/xeve_mode.c:304-313), host libm
            const double p = (512 * (i + 0.5)) / 1024;
            eb[i] = (int32_t)(-32768 * (log(p) / log(2.0) - 9));
        }
        CK(cudaMemcpyToSymbol(c_tm64, tm, sizeof(tm)));
        CK(cudaMemcpyToSymbol(c_mc_l, l, sizeof(l)));
        CK(cudaMemcpyToSymbol(c_mc_c, ch, sizeof(ch)));
        CK(cudaMemcpyToSymbol(c_quant_scale, qs, sizeof(qs)));
        CK(cudaMemcpyToSymbol(c_dequant_scale, dq, sizeof(dq)));
        CK(cudaMemcpyToSymbol(c_err_scale, es, sizeof(es)));
        CK(cudaMemcpyToSymbol(c_mpm_tbl, mpm, sizeof(mpm)));
        CK(cudaMemcpyToSymbol(g_scan, scan, sizeof(scan)));
        CK(cudaMemcpyToSymbol(g_entropy_bits, eb, sizeof(eb)));
    }
#ifdef XB200_CHAIN_DEBUG
    if(!h_dbg_words) {
        int *hp = nullptr, *dp = nullptr;
        CK(cudaHostAlloc(&hp, 64 * sizeof(int), cudaHostAllocMapped));
        memset(hp, 0, 64 * sizeof(int));
        CK(cudaHostGetDevicePointer(&dp, hp, 0));
        h_dbg_words = hp;
        volatile int *dv = dp;
        CK(cudaMemcpyToSymbol(g_dbg, &dv, sizeof(dv)));
    }
#endif
    {   // opt-in shared memory of the worker kernel: the maximum (the attribute belongs to the function, not to a launch)
        int optin = 0;
        cudaFuncAttributes fa;
        CK(cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
        CK(cudaFuncGetAttributes(&fa, k_chain_server<3>));
        CK(cudaFuncSetAttribute(k_chain_server<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, optin - (int)fa.sharedSizeBytes));
    }
    CK(cudaDeviceSynchronize());   // first context of the device: no worker grid can be alive yet
    once_done[c->device & 63] = true;
    }
    }
    CK(cudaMalloc(&k->zero_mv, k->f_scu * 8));
    CK(cudaMemsetAsync(k->zero_mv, 0, k->f_scu * 8, k->copy));
    if(!c->d_err) { CK(cudaMalloc(&c->d_err, sizeof(int))); CK(cudaMemsetAsync(c->d_err, 0, sizeof(int), k->copy)); }
    CK(cudaStreamSynchronize(k->copy));
    k->ready = true;
    {
        DeviceSched &D = g_sched[c->device & 63];
        std::lock_guard<std::mutex> dl(D.mu);
        D.n_ctx++;
        D.sms = c->sms; D.device = c->device;
        if(!D.started) {
            D.started = true;
            static std::once_flag once;
            std::call_once(once, [] { atexit(sched_atexit); });
            D.th = std::thread(sched_thread, &D);
            D.th.detach();
        }
    }
    return XB200_OK;
}

int maps_of(xb200_ctx *c, int pic, PicMaps **out)
{
    ChainCtx *k = cc_of(c);
    if((int)k->maps.size() <= pic) k->maps.resize(pic + 1, nullptr);
    if(!k->maps[pic]) k->maps[pic] = new PicMaps();
    PicMaps &m = *k->maps[pic];
    if(!m.scu) {
        const size_t f = k->f_scu;
        CK(cudaMalloc(&m.scu, f * 4)); CK(cudaMalloc(&m.ipm, f)); CK(cudaMalloc(&m.refi, f * 2)); CK(cudaMalloc(&m.mv, f * 8));
        CK(cudaMalloc(&m.flags, f));
    }
    *out = &m;
    return XB200_OK;
}

int bufs_get(xb200_ctx *c, int n_chain, JobBufs *out)
{
    ChainCtx *k = cc_of(c);
    JobBufs   b;
    if(!k->pool.empty()) { b = k->pool.back(); k->pool.pop_back(); }
    const size_t n = (size_t)k->n_lcu;
    if(!b.scu) {
        CK(cudaMalloc(&b.scu, n * 256 * sizeof(xb200_scu_rec)));
        CK(cudaMalloc(&b.coef, n * 6144 * sizeof(int16_t)));
        CK(cudaMalloc(&b.ctu_state, n * 2 * sizeof(ChState)));
        CK(cudaMalloc(&b.ctu_cost, n * sizeof(double)));
        CK(cudaMalloc(&b.done, n * sizeof(int)));
        CK(cudaMalloc(&b.counts, 2 * sizeof(unsigned long long)));
        CK(cudaEventCreate(&b.ev0)); CK(cudaEventCreate(&b.ev1)); CK(cudaEventCreate(&b.ev2));
    }
    if(b.ws_chains < n_chain) {
        if(b.ws) cudaFree(b.ws);
        CK(cudaMalloc(&b.ws, (size_t)n_chain * sizeof(ChainWs)));
        b.ws_chains = n_chain;
    }
    if(b.cu_cap < k->log_cu) {
        if(b.cu_log) cudaFree(b.cu_log);
        CK(cudaMalloc(&b.cu_log, (size_t)k->log_cu * sizeof(xb200_cu_item)));
        b.cu_cap = k->log_cu;
    }
    if(b.intra_cap < k->log_intra) {
        if(b.intra_log) cudaFree(b.intra_log);
        CK(cudaMalloc(&b.intra_log, (size_t)k->log_intra * sizeof(xb200_intra_item)));
        b.intra_cap = k->log_intra;
    }
    *out = b;
    return XB200_OK;
}
void bufs_free(JobBufs &b)
{
    for(void *p : {(void *)b.scu, (void *)b.coef, (void *)b.ctu_state, (void *)b.ctu_cost, (void *)b.done, (void *)b.counts, (void *)b.ws,
                   (void *)b.cu_log, (void *)b.intra_log})
        if(p) cudaFree(p);
    if(b.ev0) { cudaEventDestroy(b.ev0); cudaEventDestroy(b.ev1); cudaEventDestroy(b.ev2); }
    b = JobBufs();
}

// completion callback of a picture's loop filter + border expansion (runs on a CUDA-internal thread: no CUDA calls, no waiting for
// the scheduler)
void post_event(Job *j, int kind)
{
    DeviceSched &D = g_sched[j->c->device & 63];
    {
        std::lock_guard<std::mutex> cl(D.cb_mu);
        D.events.emplace_back(j, kind);
        D.wake = true;
    }
    D.cv_work.notify_one();
}
void CUDART_CB cb_ready(void *arg) { post_event(static_cast<Job *>(arg), 1); }
void sched_kick(DeviceSched &D)
{
    { std::lock_guard<std::mutex> cl(D.cb_mu); D.wake = true; }
    D.cv_work.notify_one();
}

// XB200_SCHED_DEBUG=1: the scheduler's steps on stderr (time in ms since its first line)
static bool sched_debug() { static const bool on = getenv("XB200_SCHED_DEBUG") && atoi(getenv("XB200_SCHED_DEBUG")); return on; }
static double sched_ms()
{
    static const auto t0 = std::chrono::steady_clock::now();
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
}
#define SDBG(...) do { if(sched_debug()) { fprintf(stderr, "[xb200 sched %9.3f] ", sched_ms()); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); fflush(stderr); } } while(0)
#define CKJ(call) do { cudaError_t e_ = (call); if(e_ != cudaSuccess) { fprintf(stderr, "xeve_b200: CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); return false; } } while(0)

// queue, staging and streams of a device's chain server (once)
bool server_init(DeviceSched &D, xb200_ctx *c)
{
    if(D.dq) return true;
    CKJ(cudaMalloc(&D.dq, sizeof(ChainQueue)));
    CKJ(cudaMemset(D.dq, 0, sizeof(ChainQueue)));
    unsigned *hd = nullptr;
    CKJ(cudaHostAlloc(&hd, CH_Q_SLOTS * sizeof(unsigned), cudaHostAllocMapped));
    memset(hd, 0, CH_Q_SLOTS * sizeof(unsigned));
    D.h_done = hd;
    CKJ(cudaHostGetDevicePointer(&D.d_done, hd, 0));
    CKJ(cudaHostAlloc(&D.h_slot, CH_Q_SLOTS * sizeof(ChainPicTask), cudaHostAllocDefault));
    CKJ(cudaHostAlloc(&D.h_tasks, CH_Q_TASKS * sizeof(ChainTask), cudaHostAllocDefault));
    CKJ(cudaHostAlloc(&D.h_tail, 4096 * sizeof(unsigned), cudaHostAllocDefault));
    CKJ(cudaHostAlloc(&D.h_flag, 2 * sizeof(int), cudaHostAllocDefault));
    CKJ(cudaHostAlloc(&D.h_times, CH_Q_SLOTS * 2 * sizeof(unsigned long long), cudaHostAllocDefault));
    D.h_flag[0] = 0; D.h_flag[1] = 1;
    CKJ(cudaMalloc(&D.d_tm64, 4096));
    CKJ(cudaMemcpy(D.d_tm64, c->d_tm64, 4096, cudaMemcpyDeviceToDevice));
    CKJ(cudaStreamCreateWithFlags(&D.feed, cudaStreamNonBlocking));
    CKJ(cudaStreamCreateWithFlags(&D.srv, cudaStreamNonBlocking));
    for(int i = CH_Q_SLOTS - 1; i >= 0; i--) D.free_slots.push_back(i);
    // Everything that will be launched while the worker grid is alive must be LOADED before it starts: with lazy module loading the
    // first launch of a kernel loads its code, which can synchronise the device -- and the grid only ends when the host says so.
    if(xb200_preload_api_kernels() || xb200_preload_frame_kernels()) return false;
    {   // the driver's own memset / copy kernels: run each kind once (1-, 2- and 4-byte patterns, 2-D copy, small and large)
        unsigned char *scratch = nullptr;
        CKJ(cudaMalloc(&scratch, 1 << 20));
        for(size_t n : {(size_t)4, (size_t)4096, (size_t)1 << 20, (size_t)(1 << 20) - 3}) CKJ(cudaMemsetAsync(scratch, 0, n, D.feed));
        CKJ(cudaMemset2DAsync(scratch, 2048, 0, 1000, 256, D.feed));
        CKJ(cudaMemcpy2DAsync(scratch, 2048, scratch + (1 << 19), 2048, 1000, 128, cudaMemcpyDeviceToDevice, D.feed));
        CKJ(cudaMemcpyAsync(scratch, scratch + (1 << 19), 1 << 18, cudaMemcpyDeviceToDevice, D.feed));
        CKJ(cudaMemcpyAsync(scratch, D.h_flag, 8, cudaMemcpyHostToDevice, D.feed));
        CKJ(cudaMemcpyAsync(D.h_times, scratch, 16, cudaMemcpyDeviceToHost, D.feed));
        CKJ(cudaStreamSynchronize(D.feed));
        CKJ(cudaFree(scratch));
    }
    return true;
}
// tell the workers to leave and wait until the grid is gone (only called when no chain is in flight)
bool server_stop(DeviceSched &D)
{
    if(!D.server_on) return true;
    CKJ(cudaMemcpyAsync(&D.dq->stop, &D.h_flag[1], sizeof(int), cudaMemcpyHostToDevice, D.feed));
    CKJ(cudaStreamSynchronize(D.feed));
    SDBG("server stop requested");
    CKJ(cudaStreamSynchronize(D.srv));
    D.server_on = false;
    SDBG("server stopped");
    return true;
}
// The worker grid: as many CTAs as fit with `smem` bytes each, minus a reserve -- the workers live as long as there is work, and
// the short kernels around them (upload conversion, loop filter, border expansion) need SM slots of their own to run at all.
bool server_start(DeviceSched &D, size_t smem)
{
    int bps = 0;
    CKJ(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_chain_server<3>, CH_T, smem));
    if(bps < 1) return false;
    int reserve = D.sms / 8;
    if(reserve < 8) reserve = 8;
    D.workers = bps * D.sms - reserve;
    if(const char *e = getenv("XB200_CHAIN_WORKERS")) { const int v = atoi(e); if(v > 0 && v < D.workers) D.workers = v; }
    if(D.workers < 1) D.workers = 1;
    // head = tail = stop = 0: every ticket of the previous grid is void
    CKJ(cudaMemsetAsync(D.dq, 0, offsetof(ChainQueue, tasks), D.feed));
    CKJ(cudaStreamSynchronize(D.feed));
    D.tail = 0;
    k_chain_server<3><<<D.workers, CH_T, smem, D.srv>>>(D.dq, D.d_tm64);
    CKJ(cudaGetLastError());
    D.server_on = true;
    D.server_smem = smem;
    SDBG("server started: %d workers x %d threads, %zu bytes of shared memory each (%d per SM)", D.workers, CH_T, smem, bps);
    return true;
}

// hand a picture to the chain server: its record, the cleared maps and n_chain tasks, then the new tail -- all on the feed stream, so
// the workers see the tail move after everything it announces.  Scheduler thread, scheduler mutex held.
bool publish_job(DeviceSched &D, Job *j)
{
    xb200_ctx *c = j->c;
    ChainCtx  *k = cc_of(c);
    PicMaps   *m = j->m;
    const size_t f = k->f_scu;
    const int  slot = D.free_slots.back();
    D.free_slots.pop_back();
    if(++D.seq == 0) ++D.seq;
    j->slot = slot; j->seq = D.seq;
    ChainPicTask &T = D.h_slot[slot];
    memset(&T, 0, sizeof(T));
    T.P = j->P; T.pics = c->d_pics; T.sq = c->sq; T.err_flag = c->d_err; T.seq = j->seq; T.finished = 0;
    T.h_done = D.d_done + slot; T.t_first = ~0ull; T.t_last = 0;
    D.h_done[slot] = 0;
    cudaStream_t s = D.feed;
    if(!k->span_on) { CKJ(cudaEventRecord(k->ev_span0, s)); k->span_on = true; k->span_ms = 0.f; }
    CKJ(cudaMemcpyAsync(&D.dq->slot[slot], &T, sizeof(T), cudaMemcpyHostToDevice, s));
    CKJ(cudaMemsetAsync(m->scu, 0, f * 4, s)); CKJ(cudaMemsetAsync(m->ipm, 0, f, s)); CKJ(cudaMemsetAsync(m->refi, 0, f * 2, s));
    CKJ(cudaMemsetAsync(m->mv, 0, f * 8, s)); CKJ(cudaMemsetAsync(m->flags, 0, f, s));
    CKJ(cudaMemsetAsync(j->b.done, 0, (size_t)k->n_lcu * sizeof(int), s));
    CKJ(cudaMemsetAsync(j->b.counts, 0, 2 * sizeof(unsigned long long), s));
    for(int q = 0; q < j->n_chain; q++) {
        const unsigned pos = (D.tail + (unsigned)q) % CH_Q_TASKS;
        D.h_tasks[pos].slot = slot; D.h_tasks[pos].chain = q;
        CKJ(cudaMemcpyAsync(&D.dq->tasks[pos], &D.h_tasks[pos], sizeof(ChainTask), cudaMemcpyHostToDevice, s));
    }
    D.tail += (unsigned)j->n_chain;
    unsigned *src = &D.h_tail[D.seq % 4096];
    *src = D.tail;
    CKJ(cudaMemcpyAsync(&D.dq->tail, src, sizeof(unsigned), cudaMemcpyHostToDevice, s));
    D.chains += j->n_chain;
    D.running.push_back(j);
    j->state.store(JOB_RUNNING);
    SDBG("published POC %d (%d chains) in slot %d, seq %u, tail %u; %d chains in flight", j->P.pp.poc, j->n_chain, slot, j->seq, D.tail, D.chains);
    return true;
}
// the chains of a picture have all finished: copy of the unfiltered picture (optional), both loop-filter passes, border expansion on
// one of the context's streams, then the completion callback
bool finish_job(DeviceSched &D, Job *j)
{
    xb200_ctx *c = j->c;
    ChainCtx  *k = cc_of(c);
    PicMaps   *m = j->m;
    int        sn = 0;
    for(int i = 1; i < N_STREAMS; i++) if(k->stream_busy[i] < k->stream_busy[sn]) sn = i;
    cudaStream_t s = k->streams[sn];
    j->stream_no = sn;
    k->stream_busy[sn]++;
    CKJ(cudaMemcpyAsync(&D.h_times[2 * j->slot], &D.dq->slot[j->slot].t_first, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CKJ(cudaEventRecord(j->b.ev1, s));
    if(j->has_up)
        for(int q = 0; q < 3; q++)
            CKJ(cudaMemcpy2DAsync(j->up.buf[q] + (size_t)j->up.pad[q] * j->up.s[q] + j->up.pad[q], (size_t)j->up.s[q] * 2,
                                  j->rp.buf[q] + (size_t)j->rp.pad[q] * j->rp.s[q] + j->rp.pad[q], (size_t)j->rp.s[q] * 2, (size_t)j->rp.w[q] * 2,
                                  j->rp.h[q], cudaMemcpyDeviceToDevice, s));
    if(j->deblock && xb200_deblock_dev(c, j->rp, &j->P.pp.df, m->scu, m->refi, m->mv, m->flags, s)) return false;
    if(xb200_pad_planes(c, j->rp, s)) return false;
    CKJ(cudaEventRecord(j->b.ev2, s));
    CKJ(cudaLaunchHostFunc(s, cb_ready, j));
    SDBG("POC %d: chains finished, loop filter launched", j->P.pp.poc);
    return true;
}

// process exit with a worker grid still alive (an application that never fetched its pictures, a failed run): tell the workers to
// leave, or the runtime's teardown waits for them forever
void sched_atexit()
{
    for(int d = 0; d < 64; d++) {
        DeviceSched &D = g_sched[d];
        if(!D.dq || !D.server_on) continue;
        cudaSetDevice(D.device);
        int one = 1;
        cudaMemcpy(&D.dq->stop, &one, sizeof(int), cudaMemcpyHostToDevice);
    }
}

void sched_thread(DeviceSched *Dp)
{
    DeviceSched &D = *Dp;
    cudaSetDevice(D.device);
    std::vector<std::pair<Job *, int>> ev;
    for(;;) {
        bool poll;
        {
            std::lock_guard<std::mutex> lk(D.mu);
            poll = !D.running.empty();
        }
        {
            std::unique_lock<std::mutex> cl(D.cb_mu);
            // while chains run the completion words are polled (the workers write them through host-mapped memory)
            if(poll) D.cv_work.wait_for(cl, std::chrono::microseconds(200), [&] { return D.wake || D.stop; });
            else D.cv_work.wait(cl, [&] { return D.wake || D.stop; });
            if(D.stop) return;
            D.wake = false;
            ev.clear();
            ev.swap(D.events);
        }
        std::lock_guard<std::mutex> lk(D.mu);
        bool any_ready = false;
        for(auto &e : ev) {   // loop filter + border expansion finished: the picture is a usable reference
            Job *j = e.first;
            j->m->ready.store(j->life);
            cc_of(j->c)->stream_busy[j->stream_no]--;
            j->t_chain[0] = D.h_times[2 * j->slot]; j->t_chain[1] = D.h_times[2 * j->slot + 1];
            D.free_slots.insert(D.free_slots.begin(), j->slot);   // reused last
            j->state.store(JOB_READY);
            any_ready = true;
            SDBG("POC %d complete", j->P.pp.poc);
        }
        if(D.server_on && !D.running.empty()) {   // a fault inside the worker grid must not look like a picture that takes forever
            const cudaError_t e = cudaStreamQuery(D.srv);
            if(e != cudaErrorNotReady) {
                fprintf(stderr, "xeve_b200: the chain server left the device (%s) with %zu pictures in flight\n",
                        e == cudaSuccess ? "grid ended" : cudaGetErrorString(e), D.running.size());
                for(Job *j : D.running) j->state.store(JOB_FAILED);
                for(Job *j : D.pending) j->state.store(JOB_FAILED);
                D.running.clear(); D.pending.clear();
                D.chains = 0; D.server_on = false;
                D.cv_done.notify_all();
                continue;
            }
        }
        for(size_t i = 0; i < D.running.size();) {   // pictures whose last chain has finished
            Job *j = D.running[i];
            if(D.h_done[j->slot] != j->seq) { i++; continue; }
            D.running[i] = D.running.back();
            D.running.pop_back();
            D.chains -= j->n_chain;
            j->state.store(JOB_CHAIN_DONE);
            if(!finish_job(D, j)) { j->state.store(JOB_FAILED); any_ready = true; }
        }
        // publish, in enqueue order, every queued picture whose references are complete, while the workers can take all its chains
        // at once (a partly started picture would only spin)
        for(auto it = D.pending.begin(); it != D.pending.end() && D.drain == 0;) {
            Job *j = *it;
            bool ok = true;
            for(int d = 0; d < j->n_dep && ok; d++) ok = j->deps[d].m->ready.load() >= j->deps[d].life;
            if(!ok) { ++it; continue; }
            if(!server_init(D, j->c)) { it = D.pending.erase(it); j->state.store(JOB_FAILED); any_ready = true; continue; }
            if(j->smem > D.server_smem || !D.server_on) {
                // a bigger working set than the grid was launched with (or no grid): start a new one when nothing is in flight
                if(D.chains > 0) break;
                const size_t sm = j->smem > D.server_smem ? j->smem : D.server_smem;
                if(!server_stop(D) || !server_start(D, sm)) { it = D.pending.erase(it); j->state.store(JOB_FAILED); any_ready = true; continue; }
                j->c->launches++;   // the worker grid: one launch per busy period, counted for the context whose picture started it
            }
            if(D.chains + j->n_chain > D.workers && D.chains > 0) { ++it; continue; }
            if(D.free_slots.empty() || (unsigned)(D.chains + j->n_chain) > (unsigned)CH_Q_TASKS / 2) break;
            it = D.pending.erase(it);
            if(!publish_job(D, j)) { j->state.store(JOB_FAILED); any_ready = true; }
        }
        // nothing left to do: let the grid go, so that cudaFree & co. (device-wide synchronisation) are not held up by idle workers
        if(D.server_on && (D.pending.empty() || D.drain > 0) && D.running.empty() && D.chains == 0) { server_stop(D); any_ready = true; }
        if(any_ready) D.cv_done.notify_all();
    }
}

// wait until the job is complete (or failed); called with the CONTEXT mutex released
int wait_job(Job *j)
{
    DeviceSched &D = g_sched[j->c->device & 63];
    std::unique_lock<std::mutex> lk(D.mu);
    D.cv_done.wait(lk, [&] { const int st = j->state.load(); return st == JOB_READY || st == JOB_FAILED; });
    return j->state.load() == JOB_READY ? XB200_OK : XB200_ERR_UNEXPECTED;
}

} // namespace

void xb200_chain_free(xb200_ctx *c)
{
    if(!c || !c->chain) return;
    ChainCtx    *k = cc_of(c);
    DeviceSched &D = g_sched[c->device & 63];
    {   // pictures of this context that were never launched leave the queue; the launched ones are waited for
        std::unique_lock<std::mutex> lk(D.mu);
        for(auto it = D.pending.begin(); it != D.pending.end();)
            if((*it)->c == c) { (*it)->state.store(JOB_FAILED); it = D.pending.erase(it); }
            else ++it;
        D.cv_done.wait(lk, [&] {
            for(auto &kv : k->jobs) { const int st = kv.second->state.load(); if(st == JOB_RUNNING || st == JOB_CHAIN_DONE) return false; }
            return true;
        });
        if(k->ready) D.n_ctx--;
        // cudaFree waits for the device, and the worker grid lives as long as ANY context has work: ask the scheduler to let the grid
        // end (it publishes nothing meanwhile, so this takes as long as the chains now running), free, then let it go on
        // (xb200_chain_drain_end, called by xb200_destroy when everything of the context has been freed)
        D.drain++;
        sched_kick(D);
        D.cv_done.wait(lk, [&] { return !D.server_on; });
    }
    for(int i = 0; i < N_STREAMS; i++)
        if(k->streams[i]) cudaStreamSynchronize(k->streams[i]);
    for(auto &kv : k->jobs) { bufs_free(kv.second->b); delete kv.second; }
    for(auto &b : k->pool) bufs_free(b);
    for(PicMaps *m : k->maps) {
        if(!m) continue;
        for(void *p : {(void *)m->scu, (void *)m->ipm, (void *)m->refi, (void *)m->mv, (void *)m->flags})
            if(p) cudaFree(p);
        delete m;
    }
    if(k->zero_mv) cudaFree(k->zero_mv);
    for(int i = 0; i < N_STREAMS; i++)
        if(k->streams[i]) cudaStreamDestroy(k->streams[i]);
    if(k->copy) cudaStreamDestroy(k->copy);
    if(k->ev_span0) cudaEventDestroy(k->ev_span0);
    delete k;
    c->chain = nullptr;
}

void xb200_chain_drain_end(int device)
{
    DeviceSched &D = g_sched[device & 63];
    { std::lock_guard<std::mutex> dl(D.mu); if(D.drain > 0) D.drain--; }
    sched_kick(D);
}

extern "C" {

int xb200_chain_capacity(xb200_ctx *c)
{
    if(!c) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    int r = chain_init(c);
    if(r) return r;
    int32_t cap[4];
    for(int l2 = 3; l2 <= 6; l2++) { const int ext = (1 << l2) + 2 * 10 + 7; cap[l2 - 3] = (align_up(ext, 8) + 8) * ext + 16; }
    const size_t smem = chain_smem_bytes(cap);
    int bps = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&bps, k_chain_server<3>, CH_T, smem));
    return c->sms * (bps < 1 ? 1 : bps) - (c->sms / 8 < 8 ? 8 : c->sms / 8);
}

int xb200_picture_log_enable(xb200_ctx *c, int64_t cap_cu, int64_t cap_intra)
{
    if(!c || cap_cu < 0 || cap_intra < 0 || cap_cu > (1 << 24) || cap_intra > (1 << 24)) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    int r = chain_init(c);
    if(r) return r;
    cc_of(c)->log_cu = cap_cu; cc_of(c)->log_intra = cap_intra;
    return XB200_OK;
}

int xb200_picture_adopt(xb200_ctx *c, int32_t pic, const int16_t *map_mv)
{
    if(!c || !pic_ok(c, pic) || !c->pics[pic].padded || !map_mv) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    int r = chain_init(c);
    if(r) return r;
    PicMaps *m;
    if((r = maps_of(c, pic, &m))) return r;
    CK(cudaMemcpyAsync(m->mv, map_mv, cc_of(c)->f_scu * 8, cudaMemcpyHostToDevice, cc_of(c)->copy));   // never the legacy stream:
    CK(cudaStreamSynchronize(cc_of(c)->copy));                                                           // its copies can wait for the device
    m->ready.store(m->issued);   // nothing in flight writes it
    return XB200_OK;
}

int xb200_analyze_picture(xb200_ctx *c, const xb200_picture *pp)
{
    if(!c || !pp) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    int r = chain_init(c);
    if(r) return r;
    ChainCtx *k = cc_of(c);
    std::unique_lock<std::mutex> lk(k->mu);
    // ---- arguments ----
    if(pp->slice_type == 1) return XB200_ERR_UNSUPPORTED;   // P slices: the next CTU starts from the bitstream coder's state, not the RDO's
    if(pp->slice_type != 0 && pp->slice_type != 2) return XB200_ERR_INVALID_ARGUMENT;
    if(!pic_ok(c, pp->cur_pic) || !pic_ok(c, pp->rec_pic) || !c->pics[pp->rec_pic].padded || pp->rec_pic == pp->cur_pic)
        return XB200_ERR_INVALID_ARGUMENT;
    if(pp->unfiltered_pic >= 0 && (!pic_ok(c, pp->unfiltered_pic) || !c->pics[pp->unfiltered_pic].padded || pp->unfiltered_pic == pp->rec_pic))
        return XB200_ERR_INVALID_ARGUMENT;
    auto pow2 = [](int v) { return v > 0 && (v & (v - 1)) == 0; };
    if(!pow2(pp->max_cu_inter) || !pow2(pp->min_cu_inter) || !pow2(pp->max_cu_intra) || !pow2(pp->min_cu_intra)) return XB200_ERR_INVALID_ARGUMENT;
    if(pp->max_cu_inter > 64 || pp->min_cu_inter < 8 || pp->max_cu_intra > 64 || pp->min_cu_intra < 4 || pp->min_cu_inter > pp->max_cu_inter ||
       pp->min_cu_intra > pp->max_cu_intra)
        return XB200_ERR_UNSUPPORTED;
    if(pp->tile_qp < 0 || pp->tile_qp > 63 || pp->parallel_rows < 0 || pp->max_search_range < 1 || pp->max_search_range > 256 ||
       c->seq.merge_num < 1 || c->seq.merge_num > 4 || c->seq.gop_size < 1 || (c->seq.me_complexity > 1))
        return c->seq.me_complexity > 1 ? XB200_ERR_UNSUPPORTED : XB200_ERR_INVALID_ARGUMENT;
    for(int i = 0; i < 3; i++)
        if(pp->qp[i] < 0 || pp->qp[i] > 100) return XB200_ERR_INVALID_ARGUMENT;
    int margin = 6;   // bi search: window radius 5, +1 for the integer refinement pass
    if(pp->slice_type == 0) {
        for(int l = 0; l < 2; l++) {
            if(pp->num_refp[l] < 1 || pp->num_refp[l] > XB200_MAX_REFP) return XB200_ERR_UNSUPPORTED;
            for(int q = 0; q < pp->num_refp[l]; q++) {
                const int h = pp->ref_pic[l][q];
                if(!pic_ok(c, h) || !c->pics[h].padded || h == pp->rec_pic) return XB200_ERR_INVALID_ARGUMENT;
                int d = pp->poc - pp->ref_poc[l][q];
                d = d < 0 ? -d : d;
                int dyn = (pp->max_search_range * d + (c->seq.gop_size >> 1)) / c->seq.gop_size;
                dyn = dyn < (pp->max_search_range >> 2) ? (pp->max_search_range >> 2) : (dyn > pp->max_search_range ? pp->max_search_range : dyn);
                if(dyn + 2 > margin) margin = dyn + 2;
            }
        }
    }
    if(k->jobs.count(pp->rec_pic)) return XB200_ERR_INVALID_ARGUMENT;   // fetch the previous result of this handle first
    if((r = xb200_sync_pics(c))) return r;

    ChainPic P;
    memset(&P, 0, sizeof(P));
    P.pp = *pp;
    P.w = c->seq.w; P.h = c->seq.h; P.w_scu = k->w_scu; P.h_scu = k->h_scu; P.w_lcu = k->w_lcu; P.h_lcu = k->h_lcu;
    P.pp.df.w_scu = k->w_scu; P.pp.df.h_scu = k->h_scu;
    P.n_chain = pp->parallel_rows > 1 ? (pp->parallel_rows > k->h_lcu ? k->h_lcu : pp->parallel_rows) : 1;
    for(int l2 = 3; l2 <= 6; l2++) { const int ext = (1 << l2) + 2 * margin + 7; P.win_cap[l2 - 3] = (align_up(ext, 8) + 8) * ext + 16; }
    {
        // 4x4 / 8x8 intra CUs on a warp team: 2.4x shorter latency inside a chain than the thread-per-CU variant the batched operator
        // uses for throughput (profiles/r02s06_chain_phase_profile.txt); XB200_INTRA_SMALL_TEAM=0 selects the thread variant (same results)
        const char *e = getenv("XB200_INTRA_SMALL_TEAM");
        P.small_team = e ? atoi(e) : 3;
    }
    const size_t smem = chain_smem_bytes(P.win_cap, pp->max_cu_intra, pp->slice_type == 2 ? 8 : pp->max_cu_inter);
    if(smem > 227 * 1024) return XB200_ERR_UNSUPPORTED;
    {   // candidate modes of 8x8 / 16x16 CUs on three warps when their working sets fit (XB200_CHAIN_PAR=0: serial analysis, same results)
        const char *e = getenv("XB200_CHAIN_PAR");
        P.par_stride = (e && e[0] == '0') ? 0 : chain_par_stride(P.win_cap, smem);
    }
    {
        const Pic &p = c->pics[pp->rec_pic];
        for(int q = 0; q < 3; q++) { P.rec.p[q] = p.buf[q] + (size_t)p.pad[q] * p.s[q] + p.pad[q]; P.rec.s[q] = p.s[q]; }
        P.rec.w = p.w[0]; P.rec.h = p.h[0]; P.rec.pad_l = p.pad[0]; P.rec.pad_c = p.pad[1]; P.rec.valid = 1;
    }
    PicMaps *m;
    if((r = maps_of(c, pp->rec_pic, &m))) return r;
    P.map_scu = m->scu; P.map_ipm = m->ipm; P.map_refi = m->refi; P.map_mv = m->mv; P.df_flags = m->flags;
    P.col0 = P.col1 = k->zero_mv;
    Job *j = new Job();
    j->c = c; j->rec_pic = pp->rec_pic; j->m = m;
    // what has to be complete before the picture may start: its reference pictures (if something in flight still writes them) and
    // the previous life of its own handle.  Readers of that previous life are the caller's business, as before: a handle is reused
    // only after every picture that referenced it has been fetched.
    if(pp->slice_type == 0) {
        for(int l = 0; l < 2; l++)
            for(int q = 0; q < pp->num_refp[l]; q++) {
                const int h = pp->ref_pic[l][q];
                PicMaps  *rm = h < (int)k->maps.size() ? k->maps[h] : nullptr;
                if(rm && rm->mv) {
                    if(q == 0) (l ? P.col1 : P.col0) = rm->mv;
                    if(rm->ready.load() < rm->issued) { j->deps[j->n_dep].m = rm; j->deps[j->n_dep].life = rm->issued; j->n_dep++; }
                }
                else if(q == 0) { delete j; return XB200_ERR_INVALID_ARGUMENT; }   // a reference picture needs its motion map (decided here or adopted)
            }
    }
    if(m->ready.load() < m->issued) { j->deps[j->n_dep].m = m; j->deps[j->n_dep].life = m->issued; j->n_dep++; }
    // workers the device can hold with this picture's shared memory (registers x shared memory)
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&j->b_dense, k_chain_server<3>, CH_T, smem));
    if(j->b_dense < 1 || P.n_chain > j->b_dense * c->sms - c->sms / 8 - 8) { delete j; return XB200_ERR_UNSUPPORTED; }
    if((r = bufs_get(c, P.n_chain, &j->b))) { delete j; return r; }
    j->n_chain = P.n_chain;
    P.scu_out = j->b.scu; P.coef_out = j->b.coef; P.ctu_state = j->b.ctu_state; P.ctu_cost = j->b.ctu_cost; P.done = j->b.done;
    P.counts = j->b.counts; P.ws = j->b.ws;
    if(k->log_cu > 0) { P.cu_log = j->b.cu_log; P.cu_cap = k->log_cu; }
    if(k->log_intra > 0) { P.intra_log = j->b.intra_log; P.intra_cap = k->log_intra; }
    j->P = P; j->smem = smem;
    j->rp = c->pics[pp->rec_pic];
    j->has_up = pp->unfiltered_pic >= 0;
    if(j->has_up) j->up = c->pics[pp->unfiltered_pic];
    j->deblock = pp->deblock != 0;
    j->life = ++m->issued;
    k->jobs[pp->rec_pic] = j;
    DeviceSched &D = g_sched[c->device & 63];
    {
        std::lock_guard<std::mutex> dl(D.mu);
        D.pending.push_back(j);
    }
    sched_kick(D);
    return XB200_OK;
}

int xb200_picture_fetch(xb200_ctx *c, int32_t rec_pic, xb200_scu_rec *scu, int16_t *coef, xb200_state *ctu_states, double *ctu_cost,
                        xb200_picture_stat *stat)
{
    if(!c || !c->chain) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    ChainCtx *k = cc_of(c);
    std::unique_lock<std::mutex> lk(k->mu);
    auto it = k->jobs.find(rec_pic);
    if(it == k->jobs.end()) return XB200_ERR_INVALID_ARGUMENT;
    Job *j = it->second;
    {
        lk.unlock();
        const int wr = wait_job(j);
        lk.lock();
        it = k->jobs.find(rec_pic);
        if(it == k->jobs.end() || it->second != j) return XB200_ERR_INVALID_ARGUMENT;
        if(wr) {   // the launch failed (reported on stderr by the scheduler): nothing to fetch
            k->pool.push_back(j->b);
            k->jobs.erase(it);
            delete j;
            return wr;
        }
    }
    const size_t n = (size_t)k->n_lcu;
    if(scu) CK(cudaMemcpyAsync(scu, j->b.scu, n * 256 * sizeof(xb200_scu_rec), cudaMemcpyDeviceToHost, k->copy));
    if(coef) CK(cudaMemcpyAsync(coef, j->b.coef, n * 6144 * sizeof(int16_t), cudaMemcpyDeviceToHost, k->copy));
    if(ctu_states) CK(cudaMemcpyAsync(ctu_states, j->b.ctu_state, n * 2 * sizeof(ChState), cudaMemcpyDeviceToHost, k->copy));
    if(ctu_cost) CK(cudaMemcpyAsync(ctu_cost, j->b.ctu_cost, n * sizeof(double), cudaMemcpyDeviceToHost, k->copy));
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(cnt, j->b.counts, sizeof(cnt), cudaMemcpyDeviceToHost, k->copy));
    int err = 0;
    CK(cudaMemcpyAsync(&err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, k->copy));
    CK(cudaStreamSynchronize(k->copy));
    if(stat) {
        float a = 0.f, b = 0.f;
        a = j->t_chain[1] > j->t_chain[0] ? (float)(1e-6 * (double)(j->t_chain[1] - j->t_chain[0])) : 0.f;   // %globaltimer, ns
        cudaEventElapsedTime(&b, j->b.ev1, j->b.ev2);
        stat->n_inter = (int64_t)cnt[0]; stat->n_intra = (int64_t)cnt[1]; stat->chain_ms = a; stat->filter_ms = b;
    }
    {
        DeviceSched &D = g_sched[c->device & 63];
        std::lock_guard<std::mutex> dl(D.mu);
        if(k->span_on) {
            float sp = 0.f;
            if(cudaEventElapsedTime(&sp, k->ev_span0, j->b.ev2) == cudaSuccess && sp > k->span_ms) k->span_ms = sp;
        }
    }
    k->pool.push_back(j->b);
    k->jobs.erase(it);
    delete j;
    if(err) { fprintf(stderr, "xeve_b200: search window overflow inside the decision pass\n"); return XB200_ERR_UNEXPECTED; }
    return XB200_OK;
}

int xb200_picture_ready(xb200_ctx *c, int32_t rec_pic)
{
    if(!c || !c->chain) return XB200_ERR_INVALID_ARGUMENT;
    ChainCtx *k = cc_of(c);
    std::unique_lock<std::mutex> lk(k->mu);
    auto it = k->jobs.find(rec_pic);
    if(it == k->jobs.end()) return XB200_ERR_INVALID_ARGUMENT;
    const int st = it->second->state.load();
    return st == JOB_READY || st == JOB_FAILED ? 1 : 0;
}

int xb200_picture_log(xb200_ctx *c, int32_t rec_pic, xb200_cu_item *cu, xb200_intra_item *intra, int64_t n[2])
{
    if(!c || !c->chain || !n) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    ChainCtx *k = cc_of(c);
    std::unique_lock<std::mutex> lk(k->mu);
    auto it = k->jobs.find(rec_pic);
    if(it == k->jobs.end()) return XB200_ERR_INVALID_ARGUMENT;
    Job *j = it->second;
    lk.unlock();
    const int wr = wait_job(j);
    lk.lock();
    if(wr) return wr;
    unsigned long long cnt[2] = {0, 0};
    CK(cudaMemcpyAsync(cnt, j->b.counts, sizeof(cnt), cudaMemcpyDeviceToHost, k->copy));
    CK(cudaStreamSynchronize(k->copy));
    n[0] = (int64_t)cnt[0] < k->log_cu ? (int64_t)cnt[0] : k->log_cu;
    n[1] = (int64_t)cnt[1] < k->log_intra ? (int64_t)cnt[1] : k->log_intra;
    if(cu && n[0]) CK(cudaMemcpyAsync(cu, j->b.cu_log, (size_t)n[0] * sizeof(xb200_cu_item), cudaMemcpyDeviceToHost, k->copy));
    if(intra && n[1]) CK(cudaMemcpyAsync(intra, j->b.intra_log, (size_t)n[1] * sizeof(xb200_intra_item), cudaMemcpyDeviceToHost, k->copy));
    CK(cudaStreamSynchronize(k->copy));
    return XB200_OK;
}

/* debug builds (-DXB200_CHAIN_DEBUG): the kernel's progress words (host-mapped memory: readable while a kernel hangs) */
int xb200_chain_debug(xb200_ctx *c, int32_t out[64])
{
#ifdef XB200_CHAIN_DEBUG
    if(!c || !out) return XB200_ERR_INVALID_ARGUMENT;
    if(!h_dbg_words) return XB200_ERR;
    for(int i = 0; i < 64; i++) out[i] = h_dbg_words[i];
    return XB200_OK;
#else
    (void)c; (void)out;
    return XB200_ERR_UNSUPPORTED;
#endif
}

double xb200_chain_span_ms(xb200_ctx *c, int reset)
{
    if(!c || !c->chain) return -1.0;
    ChainCtx *k = cc_of(c);
    std::lock_guard<std::mutex> dl(g_sched[c->device & 63].mu);
    const double v = k->span_ms;
    if(reset) { k->span_on = false; k->span_ms = 0.f; }
    return v;
}

/* debug builds (-DXB200_CHAIN_PROF): cycles [0..31] and counts [32..63] per phase of the decision kernel since the last call */
int xb200_chain_prof(xb200_ctx *c, uint64_t out[64])
{
#ifdef XB200_CHAIN_PROF
    if(!c || !out) return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    CK(cudaDeviceSynchronize());
    static unsigned long long z[64];
    CK(cudaMemcpyFromSymbol(out, g_chain_prof, sizeof(z)));
    CK(cudaMemcpyToSymbol(g_chain_prof, z, sizeof(z)));
    return XB200_OK;
#else
    (void)c; (void)out;
    return XB200_ERR_UNSUPPORTED;
#endif
}

int xb200_picture_maps(xb200_ctx *c, int32_t rec_pic, uint32_t *map_scu, int8_t *map_ipm, int8_t *map_refi, int16_t *map_mv)
{
    if(!c || !c->chain || rec_pic < 0 || rec_pic >= (int)cc_of(c)->maps.size() || !cc_of(c)->maps[rec_pic] || !cc_of(c)->maps[rec_pic]->scu)
        return XB200_ERR_INVALID_ARGUMENT;
    CK(cudaSetDevice(c->device));
    ChainCtx *k = cc_of(c);
    std::unique_lock<std::mutex> lk(k->mu);
    PicMaps  &m = *k->maps[rec_pic];
    {
        auto it = k->jobs.find(rec_pic);
        if(it != k->jobs.end()) {   // still in flight: wait for it
            Job *j = it->second;
            lk.unlock();
            const int wr = wait_job(j);
            lk.lock();
            if(wr) return wr;
        }
    }
    const size_t f = k->f_scu;
    if(map_scu) CK(cudaMemcpyAsync(map_scu, m.scu, f * 4, cudaMemcpyDeviceToHost, k->copy));
    if(map_ipm) CK(cudaMemcpyAsync(map_ipm, m.ipm, f, cudaMemcpyDeviceToHost, k->copy));
    if(map_refi) CK(cudaMemcpyAsync(map_refi, m.refi, f * 2, cudaMemcpyDeviceToHost, k->copy));
    if(map_mv) CK(cudaMemcpyAsync(map_mv, m.mv, f * 8, cudaMemcpyDeviceToHost, k->copy));
    CK(cudaStreamSynchronize(k->copy));
    return XB200_OK;
}

} // extern "C"

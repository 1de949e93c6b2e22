// mkhe_api.cu -- the C ABI of include/mkhe.h: context, device-resident handles and the op schedules
// that replace the bodies of mkrlwe.KeySwitcher / mkbfv.KeySwitcher / mkckks.Evaluator.Rescale.
// No torch types, no exceptions across the boundary, no CPU fallback: every op is a sequence of
// launches of the kernels in mkhe_kernels.cuh on the context's stream.
#include "../../include/mkhe.h"
#include "mkhe_kernels.cuh"
#include "mkhe_tables.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#ifndef MKHE_EMU
#include <nvtx3/nvToolsExt.h>   // header-only NVTX: one range per C-ABI call (nsys / ncu timelines show the ops)
#include <cuda.h>            // types of the stream memory operations only: the entry point is fetched at run time (no libcuda link)
#ifdef MKHE_WITH_NCCL
#include <nccl.h>
#endif
#endif

using namespace mkhe;

namespace {

#define MKHE_MAX_LANES 4
#define MKHE_LANE_EVENTS 64
enum ObjKind { OBJ_POLY = 1, OBJ_SWK = 2 };
struct Obj {
    int kind;
    u64 *d;
    int cap_limbs;     // poly: allocated limbs
    int nlimbs;        // poly: current view
    cudaEvent_t xfer = nullptr;     // completion of the last asynchronous transfer touching the object
    bool xfer_pending = false;      // the context's stream has not been ordered after it yet
    // forked contexts (lanes): the last op of every lane that touched the object, so that a use on another lane is ordered
    // after it (read-after-write, write-after-read, write-after-write).  Unused while the root has no forks.
    // ev: the lane's last op that touched the object (a slot of the lane's event ring: a recycled slot only makes a waiter wait
    // longer); wev: the lane's last op that WROTE it (an event of its own, so later reads by the same lane do not move it --
    // otherwise every reader on another lane would wait for the writer lane's most recent op, e.g. on keys uploaded once)
    struct Use { cudaEvent_t ev = nullptr, wev = nullptr, wev_own = nullptr; bool valid = false, wrote = false; };
    Use use[MKHE_MAX_LANES];
};
enum AccessMode { ACC_WRITE = 0, ACC_READ = 1 };       // the default is the conservative one

struct Scratch {
    u64 *p = nullptr;
    size_t bytes = 0;
};

}  // namespace

struct mkhe_ctx {
    int logN = 0, N = 0, nQ = 0, nP = 0, nQMul = 0, gamma = 2, device = 0;
    int S1 = 0, dmax = 0;
    int p2_timing_n = 0;
    bool debug_ck = false;                // development: checksum every scratch buffer after every launch (MKHE_DEBUG_CK=<n-th context>)
    u64 *ck_log = nullptr;
    u64 *ck_snap = nullptr;
    size_t ck_snap_bytes = 0;
    int ck_n = 0;
    std::vector<std::string> ck_names;
    int hoist_digits = 0;                 // development: digits per hoisting launch pair (MKHE_DEBUG_HOIST_DIGITS, read once; 0 = all)
    bool l2_hints = true;                 // L2 eviction priorities on the bulk copies of k_mac_intt (MKHE_DEBUG_NO_L2_HINTS switches them off: A/B runs)
    bool debug_nodiag = false;            // development: transform every tensor operand instead of reading the hoisted diagonal (MKHE_DEBUG_NODIAG, read once)
    bool debug_sync = false;              // development: host-synchronise after every launch (MKHE_DEBUG_SYNC=<n-th context of the process>)
    int num_sms = 148;                    // persistent kernels size their grids from it
    int alpha = 1, beta_max = 0;          // alpha = #P/gamma limbs per digit, beta_max = ceil(nQ/alpha) digits (mkrlwe/params.go:63-71)
    LiftTable *d_lift = nullptr;          // alpha > 1: [beta_max][alpha-1] tables of the Decomposer
    u64 T = 0;
    std::vector<u64> mod;                 // Q | P | QMul
    std::vector<ModTables> tabs;
    cudaStream_t stream = nullptr;        // the context's stream: every op is ordered on it
    cudaStream_t h2d = nullptr, d2h = nullptr;   // copy streams of the asynchronous transfers (one DMA engine per direction)
    cudaEvent_t ev_compute = nullptr;

    ModC *d_mods = nullptr;
    ulonglong2 *d_twf = nullptr, *d_twi = nullptr;
    ulonglong2 *d_twf_tiled = nullptr;      // per (modulus, tile): the staged image pass 2 fetches with one TMA copy
    ulonglong2 *d_twi_tiled = nullptr;      // the same for the inverse pass A
    bool tables_dirty = true;
    ConvTable *d_conv_PtoQ = nullptr;      // ModDownQPtoQ of the key switch
    ConvTable *d_conv_QtoQMul = nullptr;   // BFV
    ConvTable *d_conv_QMultoQ = nullptr;   // BFV
    std::vector<u64> h_mformQMul;          // MForm(QMul mod q_i)
    std::unordered_set<Obj *> objs;       // the ROOT's registry; forks resolve handles through root->objs
    // lanes: a fork shares the root's tables, keys and ciphertext objects and owns its stream, copy streams and scratch pools
    // (another KeySwitcher over the same Parameters, mkrlwe/keyswitch.go:33-47; cf. FastBasisExtender.ShallowCopy,
    // mkrlwe/basis_extension.go:155-175)
    mkhe_ctx *root = nullptr;
    int lane = 0;
    std::vector<mkhe_ctx *> lanes;        // root only: [0] = root, then the live forks (nullptr = destroyed)
    struct Touched { Obj *o; bool write; };
    std::vector<Touched> touched;         // objects of the op being enqueued
    cudaEvent_t lane_ev[MKHE_LANE_EVENTS] = {nullptr};
    unsigned lane_ev_next = 0;
    cudaEvent_t override_ev = nullptr;    // an asynchronous transfer completes on a copy stream: its event stands for the op
    bool multi() const { return root->lanes.size() > 1; }
    std::map<std::string, Scratch> scratch;
    // CUDA graphs: an op issued again with the very same operands -- device addresses and scalars -- replays
    // the launch sequence captured at its third occurrence (one cudaGraphLaunch instead of up to 22 kernel launches).  Per lane.
    struct GraphEntry { cudaGraphExec_t exec = nullptr; u64 check = 0; uint64_t launches = 0, stamp = 0; int seen = 1; bool bad = false; };
    std::unordered_map<u64, GraphEntry> graphs;
    bool graphs_on = false;               // opt-in (mkhe_ctx_set_graphs, or MKHE_GRAPHS=1 in the environment): measured, the cache does not pay
                                          // for GPU-bound ops -- profiles/r02p_cuda_graphs.txt
    bool capturing = false;
    uint64_t graph_stamp = 0, graph_replays = 0, graph_captures = 0;
    std::string err;
    int sticky = 0;
    uint64_t launches = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    void *nccl = nullptr;
    // fused exchange of the party-sharded MulRelin over peer memory (mkhe_p2p_export / _import)
    bool p2p = false;
    u64 *p2p_stage = nullptr, *p2p_xy = nullptr, *p2p_flag = nullptr;
    u64 *peer_stage[MKHE_MAX_RANKS] = {nullptr}, *peer_xy[MKHE_MAX_RANKS] = {nullptr};
    int nranks = 1, rank = 0;
    // limb-sharded ops: the team memory of this rank / lane and the ranks' mappings of it
    struct Team {
        int n = 0, rank = 0, maxk = 0;
        u64 *mem = nullptr;                 // this rank's block (cudaMalloc)
        size_t elems = 0;
        u64 *peer[MKHE_MAX_RANKS] = {nullptr};
        bool ipc[MKHE_MAX_RANKS] = {false}; // peer[r] came from cudaIpcOpenMemHandle
        u64 epoch = 0;                      // barriers passed so far (the same on every rank: all ranks issue the same ops)
        size_t off_pp = 0, off_p = 0, off_g = 0, off_in = 0;   // + two staging areas of 2 (maxk + 1) polys for mkhe_team_allgather
        unsigned in_parity = 0;
        u64 timeout_ns = 5000000000ull;
        bool spin = false;                  // development: the spinning-kernel barrier instead of the stream memory wait
    } team;
    // per-kernel CUDA-event profiling (bench.py's roofline leg)
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> ev_pool;
    cudaEvent_t prof_event() {
        if (!ev_pool.empty()) { cudaEvent_t e = ev_pool.back(); ev_pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};

namespace {

int fail(mkhe_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    return code;
}
#define MKHE_CK_MAX_LAUNCHES 4096
#define MKHE_CK_SLOTS 16
void debug_checksums(mkhe_ctx *ctx, const char *kernel) {
    if (!ctx->ck_log) {
        cudaMalloc((void **)&ctx->ck_log, (size_t)MKHE_CK_MAX_LAUNCHES * MKHE_CK_SLOTS * 8);
        cudaMemsetAsync(ctx->ck_log, 0, (size_t)MKHE_CK_MAX_LAUNCHES * MKHE_CK_SLOTS * 8, ctx->stream);
    }
    if (ctx->ck_n >= MKHE_CK_MAX_LAUNCHES) return;
    const char *only = getenv("MKHE_DEBUG_CK_ONLY");
    if (only && !strstr(kernel, only)) return;
    if (const char *snap = getenv("MKHE_DEBUG_SNAP")) {          // keep a copy of one scratch buffer per matching launch
        auto it = ctx->scratch.find(snap);
        if (it != ctx->scratch.end() && ctx->ck_n < 64) {
            if (!ctx->ck_snap) cudaMalloc((void **)&ctx->ck_snap, it->second.bytes * 64);
            ctx->ck_snap_bytes = it->second.bytes;
            cudaMemcpyAsync((char *)ctx->ck_snap + (size_t)ctx->ck_n * it->second.bytes, it->second.p, it->second.bytes, cudaMemcpyDeviceToDevice, ctx->stream);
        }
        ctx->ck_names.push_back(kernel);
        ctx->ck_n++;
        return;                                                   // snapshots only: no checksum kernels
    }
    int idx = 0;
    std::string names = kernel;
    for (auto &kv : ctx->scratch) {
        if (idx >= MKHE_CK_SLOTS) break;
        names += " " + kv.first;
        MKHE_LAUNCH(k_checksum, dim3(296), dim3(MKHE_THREADS), 0, ctx->stream, kv.second.p, kv.second.bytes / 8,
                    ctx->ck_log + (size_t)ctx->ck_n * MKHE_CK_SLOTS + idx);
        idx++;
    }
    ctx->ck_names.push_back(names);
    ctx->ck_n++;
}
#define CU(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e_ = (call);                                                                      \
        if (e_ != cudaSuccess) {                                                                      \
            ctx->sticky = MKHE_ERR_CUDA;                                                              \
            return fail(ctx, MKHE_ERR_CUDA, "CUDA error %d (%s) at %s:%d", (int)e_, cudaGetErrorString(e_), __FILE__, __LINE__); \
        }                                                                                             \
    } while (0)
// every entry point: validate, select the device, and open the scope that publishes the op's completion event to the
// objects it touched (only when the root has forks)
struct OpScope {
    mkhe_ctx *c;
    explicit OpScope(mkhe_ctx *ctx, const char *name) : c(ctx) {
#ifndef MKHE_EMU
        nvtxRangePushA(name);
#else
        (void)name;
#endif
    }
    ~OpScope() {
#ifndef MKHE_EMU
        nvtxRangePop();
#endif
        if (c->touched.empty()) { c->override_ev = nullptr; return; }
        if (c->multi()) {
            cudaEvent_t ev = c->override_ev;
            if (!ev) {
                cudaEvent_t &slot = c->lane_ev[c->lane_ev_next++ % MKHE_LANE_EVENTS];
                if (!slot) cudaEventCreateWithFlags(&slot, cudaEventDisableTiming);
                cudaEventRecord(slot, c->stream);       // a recycled slot only makes a later waiter wait longer
                ev = slot;
            }
            for (auto &t : c->touched) {
                Obj::Use &u = t.o->use[c->lane];
                if (t.write) {
                    if (c->override_ev) {
                        u.wev = nullptr;                    // an asynchronous upload: its completion is the object's xfer event
                    } else {
                        if (!u.wev_own) cudaEventCreateWithFlags(&u.wev_own, cudaEventDisableTiming);
                        cudaEventRecord(u.wev_own, c->stream);
                        u.wev = u.wev_own;
                    }
                    u.wrote = true;
                } else if (!u.valid) {
                    u.wrote = false;
                }
                u.valid = true;
                u.ev = ev;
            }
        }
        c->touched.clear();
        c->override_ev = nullptr;
    }
};
#define CHECK_CTX()                                                                                             \
    if (!ctx) return MKHE_ERR_INVALID;                                                                          \
    if (ctx->sticky) return ctx->sticky;                                                                        \
    if (cudaSetDevice(ctx->device) != cudaSuccess) return fail(ctx, MKHE_ERR_CUDA, "cudaSetDevice failed");      \
    OpScope op_scope_(ctx, __func__)
#define TRY(expr)                  \
    do {                           \
        int rc_ = (expr);          \
        if (rc_ != MKHE_OK) return rc_; \
    } while (0)
#define LAUNCH(kernel, grid, block, smem, ...)                        \
    do {                                                              \
        cudaEvent_t pa_ = nullptr, pb_ = nullptr;                     \
        if (ctx->profiling) {                                         \
            pa_ = ctx->prof_event();                                  \
            pb_ = ctx->prof_event();                                  \
            cudaEventRecord(pa_, ctx->stream);                        \
        }                                                             \
        MKHE_LAUNCH(kernel, grid, block, smem, ctx->stream, __VA_ARGS__); \
        ctx->launches++;                                              \
        if (ctx->profiling) {                                         \
            cudaEventRecord(pb_, ctx->stream);                        \
            ctx->prof.push_back({#kernel, pa_, pb_});                 \
        }                                                             \
        if (ctx->debug_sync) cudaStreamSynchronize(ctx->stream);      \
        if (ctx->debug_ck) debug_checksums(ctx, #kernel);             \
        CU(cudaGetLastError());                                       \
    } while (0)

Obj *as_obj(mkhe_ctx *ctx, uint64_t h, int kind, int mode = ACC_WRITE) {
    Obj *o = reinterpret_cast<Obj *>(h);
    if (!o || !ctx->root->objs.count(o) || o->kind != kind) return nullptr;
    const bool multi = ctx->multi();
    if (o->xfer_pending) {          // whatever uses the object next on the context's stream waits for its transfer
        cudaStreamWaitEvent(ctx->stream, o->xfer, 0);
        if (!multi) o->xfer_pending = false;        // with forks every lane has to see it: kept until the next transfer replaces it
    }
    if (multi) {
        const bool write = mode == ACC_WRITE;
        for (int l = 0; l < MKHE_MAX_LANES; l++) {
            if (l == ctx->lane) continue;
            Obj::Use &u = o->use[l];
            if (u.valid && write) cudaStreamWaitEvent(ctx->stream, u.ev, 0);                      // after every earlier use
            else if (u.valid && u.wrote && u.wev) cudaStreamWaitEvent(ctx->stream, u.wev, 0);    // after the last write only
            if (write) u.valid = false;             // ordered before this op; later users wait for this lane instead
        }
        ctx->touched.push_back({o, write});
    }
    return o;
}
#define POLY_M(var, h, mode)                                                             \
    Obj *var = as_obj(ctx, (h), OBJ_POLY, (mode));                                       \
    if (!var) return fail(ctx, MKHE_ERR_INVALID, "invalid poly handle (%s)", #h)
#define SWK_M(var, h, mode)                                                              \
    Obj *var = as_obj(ctx, (h), OBJ_SWK, (mode));                                        \
    if (!var) return fail(ctx, MKHE_ERR_INVALID, "invalid switching-key handle (%s)", #h)
#define POLY(var, h) POLY_M(var, h, ACC_WRITE)
#define SWK(var, h) SWK_M(var, h, ACC_WRITE)
#define POLY_R(var, h) POLY_M(var, h, ACC_READ)
#define SWK_R(var, h) SWK_M(var, h, ACC_READ)

size_t swk_elems(const mkhe_ctx *ctx) { return (size_t)ctx->beta_max * ctx->dmax * ctx->N; }
int beta_of(const mkhe_ctx *ctx, int levelQ) { return (levelQ + ctx->alpha) / ctx->alpha; }      // ceil((levelQ+1)/alpha)

void drop_graphs(mkhe_ctx *ctx) {
#ifndef MKHE_EMU
    for (auto &kv : ctx->graphs)
        if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
#endif
    ctx->graphs.clear();
}
// signature of an op: two independent 64-bit hashes over its scalars and device addresses (the second one guards the first)
struct GraphKey {
    u64 h1 = 0xcbf29ce484222325ull, h2 = 0x9E3779B97F4A7C15ull;
    void add(u64 v) {
        h1 = (h1 ^ v) * 0x100000001b3ull; h1 ^= h1 >> 29;
        h2 = (h2 + v) * 0xBF58476D1CE4E5B9ull; h2 ^= h2 >> 31;
    }
    void ints(int n, const int *p) { add((u64)n); for (int i = 0; i < n; i++) add((u64)(long)p[i]); }
    template <class T> void ptrs(const std::vector<T *> &v) { add(v.size()); for (T *q : v) add((u64)(uintptr_t)q); }
    explicit GraphKey(const char *op) { for (const char *c = op; *c; c++) add((u64)*c); }
};
#define MKHE_GRAPH_CACHE 256
// run `body` (the launches of one op on ctx->stream) through the graph cache
template <class F> int run_graphed(mkhe_ctx *ctx, const GraphKey &k, F &&body) {
#ifdef MKHE_EMU
    (void)k;
    return body();
#else
    if (!ctx->graphs_on || ctx->profiling || ctx->debug_sync || ctx->debug_ck || ctx->capturing) return body();
    auto it = ctx->graphs.find(k.h1);
    if (it == ctx->graphs.end()) {
        // first sight: run eagerly (the scratch pools get their sizes) and remember the signature
        if (ctx->graphs.size() >= MKHE_GRAPH_CACHE) {
            auto old = ctx->graphs.begin();
            for (auto j = ctx->graphs.begin(); j != ctx->graphs.end(); ++j)
                if (j->second.stamp < old->second.stamp) old = j;
            if (old->second.exec) cudaGraphExecDestroy(old->second.exec);
            ctx->graphs.erase(old);
        }
        mkhe_ctx::GraphEntry e;
        e.check = k.h2;
        e.stamp = ++ctx->graph_stamp;
        ctx->graphs.emplace(k.h1, e);
        return body();
    }
    mkhe_ctx::GraphEntry &e = it->second;
    if (e.bad || e.check != k.h2) return body();
    e.stamp = ++ctx->graph_stamp;
    if (!e.exec && ++e.seen < 3) return body();          // a signature that shows up only twice is not worth an instantiation
    if (!e.exec) {
        const uint64_t l0 = ctx->launches;
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
            cudaGetLastError();
            e.bad = true;
            return body();
        }
        ctx->capturing = true;
        const int rc = body();
        ctx->capturing = false;
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(ctx->stream, &g);
        const uint64_t nl = ctx->launches - l0;
        ctx->launches = l0;
        bool ok = rc == MKHE_OK && ce == cudaSuccess && g != nullptr;
        if (ok && cudaGraphInstantiate(&e.exec, g, 0) != cudaSuccess) { ok = false; e.exec = nullptr; }
        if (g) cudaGraphDestroy(g);
        if (!ok) {                              // something in this op cannot be captured: it runs eagerly from now on
            cudaGetLastError();
            e.bad = true;
            ctx->sticky = 0;
            ctx->err.clear();
            return body();
        }
        e.launches = nl;
        ctx->graph_captures++;
    }
    CU(cudaGraphLaunch(e.exec, ctx->stream));
    ctx->launches += e.launches;
    ctx->graph_replays++;
    return MKHE_OK;
#endif
}

int get_scratch(mkhe_ctx *ctx, const std::string &name, size_t bytes, u64 **out) {
    Scratch &s = ctx->scratch[name];
    if (s.bytes < bytes) {
        // a captured graph holds the old address: never reallocate inside a capture, and forget the graphs made so far
        if (ctx->capturing) return fail(ctx, MKHE_ERR_UNSUPPORTED, "scratch '%s' would grow during a graph capture", name.c_str());
        drop_graphs(ctx);
        if (s.p) {
            CU(cudaStreamSynchronize(ctx->stream));
            CU(cudaFree(s.p));
            s.p = nullptr;
            s.bytes = 0;
        }
        void *p = nullptr;
        if (cudaMalloc(&p, bytes) != cudaSuccess) return fail(ctx, MKHE_ERR_NOMEM, "cudaMalloc(%zu) failed for scratch '%s'", bytes, name.c_str());
        s.p = (u64 *)p;
        s.bytes = bytes;
    }
    *out = s.p;
    return MKHE_OK;
}

// CUDA loads a kernel lazily at its first launch, and that load may have to wait until the device is idle.  An op of a team must
// never hit one: a rank's in-kernel barrier can be spinning on the device at that moment, waiting for work the blocked host thread
// has not enqueued yet.  So every kernel this context can launch is loaded when the context is made.
template <class K> void preload(K kernel) {
#ifndef MKHE_EMU
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kernel);
#else
    (void)kernel;
#endif
}
template <int S1> void preload_s1(int nP) {
    preload(k_bcast_ntt_pass1<S1>); preload(k_ntt_pass1<S1>); preload(k_intt_passB<S1>); preload(k_moddown_P<S1>);
    switch (nP) {
        case 1: preload(k_moddown_Q<S1, 1, false>); preload(k_moddown_Q<S1, 1, true>); break;
        case 2: preload(k_moddown_Q<S1, 2, false>); preload(k_moddown_Q<S1, 2, true>); break;
        case 3: preload(k_moddown_Q<S1, 3, false>); preload(k_moddown_Q<S1, 3, true>); break;
        default: preload(k_moddown_Q<S1, 4, false>); preload(k_moddown_Q<S1, 4, true>); break;
    }
}
void preload_kernels(const mkhe_ctx *ctx) {
    switch (ctx->S1) {
        case 1: preload_s1<1>(ctx->nP); break;
        case 2: preload_s1<2>(ctx->nP); break;
        case 3: preload_s1<3>(ctx->nP); break;
        case 4: preload_s1<4>(ctx->nP); break;
        default: preload_s1<5>(ctx->nP); break;
    }
    preload(k_tile_twiddles); preload(k_decomp_lift); preload(k_ntt_pass2); preload(k_mac_intt<1>); preload(k_mac_intt<2>);
    preload(k_intt_passA); preload(k_mac_parties); preload(k_mac_parties_scatter); preload(k_reduce_gather);
    preload(k_conv<CONV_MODUP>); preload(k_conv<CONV_MODDOWN>); preload(k_tensor); preload(k_addsub<true>); preload(k_addsub<false>);
    preload(k_reduce); preload(k_rescale); preload(k_automorph); preload(k_scale); preload(k_mul_const); preload(k_neg);
    preload(k_mul_mont); preload(k_decrypt_sum); preload(k_mul2); preload(k_checksum); preload(k_bfly_peak);
    preload(k_sample); preload(k_key_fma); preload(k_permute_ntt);
    preload(k_moddown_ov); preload(k_team_barrier); preload(k_team_signal); preload(k_team_gather); preload(k_copy_limbs);
}

int upload_tables(mkhe_ctx *ctx) {
    if (!ctx->tables_dirty) return MKHE_OK;
    const size_t nm = ctx->mod.size(), N = ctx->N;
    if (!ctx->d_mods) {
        CU(cudaMalloc((void **)&ctx->d_mods, sizeof(ModC) * 64));
        CU(cudaMalloc((void **)&ctx->d_twf, sizeof(ulonglong2) * 64 * N));
        CU(cudaMalloc((void **)&ctx->d_twi, sizeof(ulonglong2) * 64 * N));
        CU(cudaMalloc((void **)&ctx->d_twf_tiled, sizeof(ulonglong2) * 64 * N));
        CU(cudaMalloc((void **)&ctx->d_twi_tiled, sizeof(ulonglong2) * 64 * N));
    }
    for (size_t i = 0; i < nm; i++) {
        CU(cudaMemcpy(ctx->d_mods + i, &ctx->tabs[i].c, sizeof(ModC), cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_twf + i * N, ctx->tabs[i].twf.data(), sizeof(ulonglong2) * N, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_twi + i * N, ctx->tabs[i].twi.data(), sizeof(ulonglong2) * N, cudaMemcpyHostToDevice));
    }
    MKHE_LAUNCH(k_tile_twiddles, dim3(ctx->N / MKHE_TILE, (unsigned)nm), dim3(MKHE_THREADS), 0, ctx->stream, ctx->d_twf, ctx->d_twf_tiled, ctx->logN);
    MKHE_LAUNCH(k_tile_twiddles, dim3(ctx->N / MKHE_TILE, (unsigned)nm), dim3(MKHE_THREADS), 0, ctx->stream, ctx->d_twi, ctx->d_twi_tiled, ctx->logN);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->tables_dirty = false;
    return MKHE_OK;
}

// alpha > 1: one LiftTable per (digit, nsrc in 2..alpha), index digit*(alpha-1) + nsrc-2 (NewDecomposer's modUpParams)
int upload_lift_tables(mkhe_ctx *ctx) {
    const int per = ctx->alpha - 1, n = ctx->beta_max * per;
    std::vector<LiftTable> h(n);
    for (int d = 0; d < ctx->beta_max; d++)
        for (int nsrc = 2; nsrc <= ctx->alpha; nsrc++) {
            const int src0 = d * ctx->alpha;
            if (src0 + nsrc > ctx->nQ) { memset(&h[d * per + nsrc - 2], 0, sizeof(LiftTable)); continue; }
            gen_lift_table(h[d * per + nsrc - 2], ctx->mod, ctx->nQ + ctx->nP, src0, nsrc);
        }
    CU(cudaMalloc((void **)&ctx->d_lift, sizeof(LiftTable) * n));
    CU(cudaMemcpy(ctx->d_lift, h.data(), sizeof(LiftTable) * n, cudaMemcpyHostToDevice));
    return MKHE_OK;
}
int upload_conv(mkhe_ctx *ctx, ConvTable **d, const std::vector<int> &src, const std::vector<int> &dst) {
    if ((int)src.size() > MKHE_CONV_MAX || (int)dst.size() > MKHE_CONV_MAX)
        return fail(ctx, MKHE_ERR_UNSUPPORTED, "basis conversion with more than %d limbs", MKHE_CONV_MAX);
    ConvTable *h = new ConvTable();
    memset(h, 0, sizeof(ConvTable));
    gen_conv_table(*h, ctx->mod, src, dst);
    if (!*d) CU(cudaMalloc((void **)d, sizeof(ConvTable)));
    CU(cudaMemcpy(*d, h, sizeof(ConvTable), cudaMemcpyHostToDevice));
    delete h;
    return MKHE_OK;
}

// limb-slot lists ---------------------------------------------------------------------------------
struct Slots {
    int n = 0;
    int slot[MKHE_MAX_SLOTS];
    int mod[MKHE_MAX_SLOTS];
};
// Q limbs 0..level followed by every P limb: the limbs of a PolyQP at `level` (slot == modulus index)
Slots qp_slots(const mkhe_ctx *ctx, int level) {
    Slots s;
    for (int j = 0; j <= level; j++) { s.slot[s.n] = j; s.mod[s.n++] = j; }
    for (int j = 0; j < ctx->nP; j++) { s.slot[s.n] = ctx->nQ + j; s.mod[s.n++] = ctx->nQ + j; }
    return s;
}
Slots q_slots(int level) {
    Slots s;
    for (int j = 0; j <= level; j++) { s.slot[s.n] = j; s.mod[s.n++] = j; }
    return s;
}
// limbs of a poly in basis R = Q u QMul
Slots r_slots(const mkhe_ctx *ctx) {
    Slots s;
    for (int j = 0; j < ctx->nQ; j++) { s.slot[s.n] = j; s.mod[s.n++] = j; }
    for (int j = 0; j < ctx->nQMul; j++) { s.slot[s.n] = ctx->nQ + j; s.mod[s.n++] = ctx->nQ + ctx->nP + j; }
    return s;
}
void fill(LimbArgs &a, const Slots &s, int logN) {
    a.nlimbs = s.n;
    a.logN = logN;
    for (int i = 0; i < s.n; i++) { a.slot_of[i] = s.slot[i]; a.mod_of_limb[i] = s.mod[i]; }
}

template <class F>
int dispatch_s1(mkhe_ctx *ctx, F &&f) {
    switch (ctx->S1) {
        case 1: return f(std::integral_constant<int, 1>());
        case 2: return f(std::integral_constant<int, 2>());
        case 3: return f(std::integral_constant<int, 3>());
        case 4: return f(std::integral_constant<int, 4>());
        case 5: return f(std::integral_constant<int, 5>());
    }
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "logN = %d unsupported (12..16)", ctx->logN);
}

const size_t SMEM_PASS2 = MKHE_P2_SMEM;                      // twiddles + per group: exchange and landing buffers
const int COLGROUPS = MKHE_TILE / MKHE_NTT_THREADS;              // CTAs per limb in the column passes

// how many chunks to split `count` looped instances into so that the grid fills the machine (>= ~4 waves of 3 CTAs/SM)
int pick_chunks(int count, long ctas_per_chunk) {
    const long want = 148L * 4 * 4;
    int chunks = (int)std::min<long>(count, std::max<long>(1, (want + ctas_per_chunk - 1) / ctas_per_chunk));
    return std::max(chunks, 1);
}

// pass 2 is persistent: the grid is what the device holds at once and every CTA owns an equal-cost stretch of the tile list
#ifndef MKHE_P2_BIGW
#define MKHE_P2_BIGW 19          // cost of a tile of a 59/60-bit modulus in sixteenths of a normal one (the sweeps)
#endif
int launch_pass2(mkhe_ctx *ctx, const Slots &s, int np, u64 *const *bufs, int count, long inst_stride, bool lazy_out = false) {
    const int tiles = ctx->N / MKHE_TILE;
    Pass2Args b;
    memset(&b, 0, sizeof b);
    b.count = count;
    b.ninst = np * count;
    b.inst_stride = inst_stride;
    b.nslots = s.n;
    b.logN = ctx->logN;
    b.lazy_out = lazy_out ? 1 : 0;
    const long per_slot = (long)tiles * b.ninst;
    for (int i = 0; i < s.n; i++) {
        b.slots[i] = s.slot[i];
        b.mods[i] = s.mod[i];
        b.wslot[i] = ctx->tabs[s.mod[i]].c.big ? MKHE_P2_BIGW : 16;
        b.wstart[i + 1] = b.wstart[i] + per_slot * b.wslot[i];
    }
    for (int i = 0; i < np; i++) b.buf.p[i] = bufs[i];
    const long ntile_inst = per_slot * s.n;
    const long grid = std::max<long>(1, std::min<long>(ctx->num_sms, ntile_inst / std::min(b.ninst, MKHE_P2_GROUPS)));
#ifdef MKHE_P2_TIMING
    TRY(get_scratch(ctx, "p2timing", 4 * 8 * 4096, &b.timing));
    ctx->p2_timing_n = (int)grid;
#endif
    LAUNCH(k_ntt_pass2, dim3((unsigned)grid), dim3(MKHE_P2_THREADS), SMEM_PASS2, b, ctx->d_mods, ctx->d_twf_tiled);
    return MKHE_OK;
}

// ---- building blocks ------------------------------------------------------------------------------
// forward NTT of `npolys` polys over the limb list `s` (out may alias in)
int ntt_fwd(mkhe_ctx *ctx, const Slots &s, int npolys, u64 *const *in, u64 *const *out) {
    for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
        int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
        LimbArgs a;
        fill(a, s, ctx->logN);
        for (int i = 0; i < np; i++) { a.in.p[i] = in[p0 + i]; a.out.p[i] = out[p0 + i]; }
        TRY(dispatch_s1(ctx, [&](auto S) -> int {
            auto k_ntt_pass1_ = k_ntt_pass1<decltype(S)::value>;
            LAUNCH(k_ntt_pass1_, dim3(COLGROUPS, s.n, np), dim3(MKHE_NTT_THREADS), 0, a, ctx->d_mods, ctx->d_twf);
            return MKHE_OK;
        }));
        TRY(launch_pass2(ctx, s, np, out + p0, 1, 0));
    }
    return MKHE_OK;
}
// inverse pass B over limb list (in place on bufs)
int intt_passB(mkhe_ctx *ctx, const Slots &s, int npolys, u64 *const *bufs_in, u64 *const *bufs_out) {
    for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
        int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
        LimbArgs a;
        fill(a, s, ctx->logN);
        for (int i = 0; i < np; i++) { a.in.p[i] = bufs_in[p0 + i]; a.out.p[i] = bufs_out[p0 + i]; }
        TRY(dispatch_s1(ctx, [&](auto S) -> int {
            auto k_intt_passB_ = k_intt_passB<decltype(S)::value>;
            LAUNCH(k_intt_passB_, dim3(COLGROUPS, s.n, np), dim3(MKHE_NTT_THREADS), 0, a, ctx->d_mods, ctx->d_twi);
            return MKHE_OK;
        }));
    }
    return MKHE_OK;
}
// full inverse NTT (canonical in, canonical out)
int ntt_inv(mkhe_ctx *ctx, const Slots &s, int npolys, u64 *const *in, u64 *const *out) {
    const int tiles = ctx->N / MKHE_TILE;
    for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
        int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
        InvAArgs a;
        memset(&a, 0, sizeof a);
        a.nslots = s.n;
        a.logN = ctx->logN;
        a.nbatch = np;
        for (int i = 0; i < s.n; i++) { a.slots[i] = s.slot[i]; a.mods[i] = s.mod[i]; }
        for (int i = 0; i < np; i++) { a.in.p[i] = in[p0 + i]; a.out.p[i] = out[p0 + i]; }
        LAUNCH(k_intt_passA, dim3(tiles, s.n, (np + MKHE_PA_GROUPS - 1) / MKHE_PA_GROUPS), dim3(MKHE_PA_THREADS), MKHE_PA_SMEM, a, ctx->d_mods, ctx->d_twi_tiled);
    }
    return intt_passB(ctx, s, npolys, out, out);
}

// Decompose: digits in_limb0 .. in_limb0+beta-1 of each input poly -> swk-shaped outputs (NTT domain)
// lazy: the NTT outputs of the broadcast digits stay un-reduced (< 2^64, congruent) -- only for forms that never leave the op
int decompose_impl(mkhe_ctx *ctx, int levelQ, int npolys, u64 *const *in, u64 *const *out, int in_limb0, bool lazy = false,
                   const Slots *only = nullptr) {
    const int alpha = ctx->alpha, beta = beta_of(ctx, levelQ);
    if (alpha > 1 && in_limb0 != 0) return fail(ctx, MKHE_ERR_UNSUPPORTED, "DecomposeBFV relies on alpha = 1 (mkbfv/keyswitch.go:64-67)");
    // digits [0, nlift) hold more than one limb and take the exact lift; a last digit of a single limb is a broadcast
    const int nsrc_last = levelQ + 1 - alpha * (beta - 1);
    const int nlift = alpha == 1 ? 0 : (nsrc_last == 1 ? beta - 1 : beta);
    Slots s = only ? *only : qp_slots(ctx, levelQ);          // limb sharding: the target limbs this rank owns
    if (s.n == 0) return MKHE_OK;
    const long digit_elems = (long)ctx->dmax * ctx->N;
    for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
        int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
        if (nlift > 0) {
            LiftArgs a;
            memset(&a, 0, sizeof a);
            a.digit0 = 0;
            a.table_stride = alpha - 1;
            a.table0 = alpha - 2;
            a.last_digit = nlift - 1;
            a.last_table = (nlift - 1) * (alpha - 1) + (nlift == beta ? nsrc_last : alpha) - 2;
            a.dmax = ctx->dmax;
            a.nslots = s.n;
            a.logN = ctx->logN;
            for (int i = 0; i < s.n; i++) a.slots[i] = s.slot[i];
            for (int i = 0; i < np; i++) { a.in.p[i] = in[p0 + i]; a.out.p[i] = out[p0 + i]; }
            LAUNCH(k_decomp_lift, dim3(ctx->N / MKHE_THREADS, nlift, np), dim3(MKHE_THREADS), 0, a, ctx->d_lift, ctx->d_mods);
            std::vector<u64 *> ent;
            for (int i = 0; i < np; i++)
                for (int d = 0; d < nlift; d++) ent.push_back(out[p0 + i] + d * digit_elems);
            TRY(ntt_fwd(ctx, s, (int)ent.size(), ent.data(), ent.data()));
        }
        if (nlift < beta) {
            // development knob (MKHE_DEBUG_HOIST_DIGITS=n): hoist n digits per launch pair, so that pass 1's output (n * D limbs per
            // poly) is still L2 resident when pass 2 reads it back -- the measured alternative to a single-pass transform
            const int chunk = ctx->hoist_digits > 0 ? ctx->hoist_digits : beta - nlift;
            for (int d0 = nlift; d0 < beta; d0 += chunk) {
                const int nd = std::min(chunk, beta - d0);
                BcastArgs a;
                memset(&a, 0, sizeof a);
                a.in_limb0 = in_limb0;
                a.in_limb_stride = alpha;
                a.digit0 = d0;
                a.dmax = ctx->dmax;
                a.nslots = s.n;
                a.slot_groups = std::min(s.n, pick_chunks(s.n, (long)COLGROUPS * nd * np));
                a.logN = ctx->logN;
                for (int i = 0; i < s.n; i++) a.slots[i] = s.slot[i];
                for (int i = 0; i < np; i++) { a.in.p[i] = in[p0 + i]; a.out.p[i] = out[p0 + i]; }
                TRY(dispatch_s1(ctx, [&](auto S) -> int {
                    auto k_bcast_ntt_pass1_ = k_bcast_ntt_pass1<decltype(S)::value>;
                    LAUNCH(k_bcast_ntt_pass1_, dim3(COLGROUPS * a.slot_groups, nd, np), dim3(MKHE_NTT_THREADS), 0, a, ctx->d_mods, ctx->d_twf);
                    return MKHE_OK;
                }));
                std::vector<u64 *> bufs(np);
                for (int i = 0; i < np; i++) bufs[i] = out[p0 + i] + d0 * digit_elems;
                TRY(launch_pass2(ctx, s, np, bufs.data(), nd, digit_elems, lazy));
            }
        }
    }
    return MKHE_OK;
}

// One external product of a batch:  dst (+)= ModDown(INTT(sum_{set} sum_i key[set][i] (.) hst[set][i]))
//   (ExternalProductHoisted, mkrlwe/keyswitch_hoisted.go:10-40).  Products that name the same dst are summed
//   into it (exact modular adds, any order); `add` = start from dst's current contents (ringQ.AddLvl) instead of zero.
// ---- limb sharding: which limb slots a rank owns, the team's barrier ---------------------------------------------------
int owner_of_slot(const mkhe_ctx *ctx, int slot) { return slot % ctx->team.n; }       // stable over levels: keys can be stored sharded
Slots own_slots(const mkhe_ctx *ctx, const Slots &all) {
    Slots s;
    for (int i = 0; i < all.n; i++)
        if (owner_of_slot(ctx, all.slot[i]) == ctx->team.rank) { s.slot[s.n] = all.slot[i]; s.mod[s.n++] = all.mod[i]; }
    return s;
}
void fill_team(const mkhe_ctx *ctx, TeamArgs &t) {
    memset(&t, 0, sizeof t);
    t.nranks = ctx->team.n;
    t.rank = ctx->team.rank;
    for (int r = 0; r < ctx->team.n; r++) t.peer[r] = ctx->team.peer[r];
}
#ifndef MKHE_EMU
typedef CUresult (*wait_value64_fn)(CUstream, CUdeviceptr, cuuint64_t, unsigned int);
wait_value64_fn stream_wait_value64() {
    static wait_value64_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult st;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue64", &p, cudaEnableDefault, &st) == cudaSuccess && st == cudaDriverEntryPointSuccess)
            fn = (wait_value64_fn)p;
        cudaGetLastError();
    }
    return fn;
}
#endif
// every rank's earlier work on this lane (its peer stores included) is complete before any rank's later work starts.
// Signal kernel + stream memory wait; the spinning kernel k_team_barrier is the fallback where stream memory operations are
// not available (and MKHE_DEBUG_SPIN_BARRIER forces it).
int team_barrier(mkhe_ctx *ctx) {
    TeamArgs t;
    fill_team(ctx, t);
    ctx->team.epoch++;
    if (ctx->team.n <= 1) return MKHE_OK;
#ifndef MKHE_EMU
    wait_value64_fn wait = ctx->team.spin ? nullptr : stream_wait_value64();
    if (wait) {
        LAUNCH(k_team_signal, dim3(1), dim3(32), 0, t);
        // CU_STREAM_WAIT_VALUE_FLUSH (flush of outstanding remote writes) exists only where the platform can do it; without it the
        // ordering still holds: the peers' data stores precede their release-add at system scope, the kernels behind the wait start
        // with an invalidated L1 and read through L2, the point of coherence for writes that arrive over NVLink
        static bool can_flush = true;
        const CUdeviceptr addr = (CUdeviceptr)(ctx->team.mem + MKHE_TEAM_COUNTER);
        const cuuint64_t target = (cuuint64_t)(ctx->team.epoch * (u64)(ctx->team.n - 1));
        CUresult r = CUDA_ERROR_NOT_SUPPORTED;
        if (can_flush) r = wait((CUstream)ctx->stream, addr, target, CU_STREAM_WAIT_VALUE_GEQ | CU_STREAM_WAIT_VALUE_FLUSH);
        if (r == CUDA_ERROR_NOT_SUPPORTED) {
            can_flush = false;
            r = wait((CUstream)ctx->stream, addr, target, CU_STREAM_WAIT_VALUE_GEQ);
        }
        if (r == CUDA_SUCCESS) return MKHE_OK;
        return fail(ctx, MKHE_ERR_CUDA, "cuStreamWaitValue64 failed (%d)", (int)r);
    }
#endif
    LAUNCH(k_team_barrier, dim3(1), dim3(32), 0, t, ctx->team.epoch, ctx->team.timeout_ns);
    return MKHE_OK;
}

struct Prod {
    u64 *key[2], *hst[2];
    u64 *dst;
    bool add;                 // start from dst's current contents (ringQ.AddLvl) instead of zero
    const u64 *src = nullptr; // start from this poly instead (overrides add); e.g. RotateHoisted's copy of ctIn["0"]
    long team_off = -1;       // limb-sharded ops: >= 0 = the result has to be complete on every rank: stored into every rank's team
                              // memory at this element offset (dst = this rank's image of that place)
};

// `team`: limb-sharded execution -- only the limb slots this rank owns are computed; the P parts travel to every rank from the
// epilogue of k_moddown_P, followed by a barrier
int ext_products(mkhe_ctx *ctx, int levelQ, int nsets, const std::vector<Prod> &prods, u64 galEl = 0, bool team = false) {
    const int nb = (int)prods.size();
    if (nb == 0) return MKHE_OK;
    const int tiles = ctx->N / MKHE_TILE;
    const Slots s = team ? own_slots(ctx, qp_slots(ctx, levelQ)) : qp_slots(ctx, levelQ);
    if (team && nb > 2 * ctx->team.maxk) return fail(ctx, MKHE_ERR_INVALID, "team memory sized for %d parties", ctx->team.maxk);
    const int vslot = ctx->dmax;                              // nP spare limbs of every accumulator: the terms of the overflow estimate v
    const size_t qp = (size_t)(ctx->dmax + ctx->nP) * ctx->N;
    for (int b0 = 0; b0 < nb; b0 += MKHE_MD_PRODUCTS) {
        const int n = std::min(MKHE_MD_PRODUCTS, nb - b0);
        // targets of this chunk; a target must not straddle chunks (its products are summed by one CTA)
        std::vector<u64 *> tg;
        std::vector<std::vector<int>> members;
        for (int i = 0; i < n; i++) {
            size_t t = std::find(tg.begin(), tg.end(), prods[b0 + i].dst) - tg.begin();
            if (t == tg.size()) { tg.push_back(prods[b0 + i].dst); members.emplace_back(); }
            members[t].push_back(i);
        }
        {   // targets with the most products first: their CTAs run longest
            std::vector<size_t> ord(tg.size());
            for (size_t t = 0; t < ord.size(); t++) ord[t] = t;
            std::stable_sort(ord.begin(), ord.end(), [&](size_t a, size_t b) { return members[a].size() > members[b].size(); });
            std::vector<u64 *> tg2;
            std::vector<std::vector<int>> mem2;
            for (size_t t : ord) { tg2.push_back(tg[t]); mem2.push_back(members[t]); }
            tg.swap(tg2);
            members.swap(mem2);
        }
        for (int i = b0 + n; i < nb; i++)
            if (std::find(tg.begin(), tg.end(), prods[i].dst) != tg.end())
                return fail(ctx, MKHE_ERR_UNSUPPORTED, "more than %d products in one batch with a shared target", MKHE_MD_PRODUCTS);
        u64 *accqp;
        TRY(get_scratch(ctx, "accqp", qp * 8 * (size_t)std::min(nb, MKHE_MD_PRODUCTS), &accqp));
        std::vector<u64 *> bufs(n);
        for (int i = 0; i < n; i++) bufs[i] = accqp + (size_t)i * qp;

        // ---- multiply-accumulate over the digits fused with the inverse pass A (k_mac_intt); neighbours sharing one operand
        //      form a group of two, the rest run as groups of one
        std::vector<std::pair<int, int>> groups[3];          // groups[G]: (first product, shared side: 1 = key, 0 = hoisted form)
        for (int i = 0; i < n;) {
            bool sk = i + 1 < n, sh = i + 1 < n;
            for (int t = 0; t < nsets && i + 1 < n; t++) {
                sk = sk && prods[b0 + i + 1].key[t] == prods[b0 + i].key[t];
                sh = sh && prods[b0 + i + 1].hst[t] == prods[b0 + i].hst[t];
            }
            if (sk || sh) { groups[2].push_back({i, sk ? 1 : 0}); i += 2; }
            else { groups[1].push_back({i, 1}); i += 1; }
        }
        for (int G = 2; G >= 1 && s.n > 0; G--)
            for (size_t g0 = 0; g0 < groups[G].size(); g0 += MKHE_MI_GROUPS) {
                const int ng = (int)std::min<size_t>(MKHE_MI_GROUPS, groups[G].size() - g0);
                MacInttArgs a;
                memset(&a, 0, sizeof a);
                a.nsets = nsets;
                a.beta = beta_of(ctx, levelQ);
                a.ngroups = ng;
                a.digit_stride = (long)ctx->dmax * ctx->N;
                a.nslots = s.n;
                a.galEl = galEl;
                a.logN = ctx->logN;
                for (int k = 0; k < s.n; k++) { a.slots[k] = s.slot[k]; a.mods[k] = s.mod[k]; }
                for (int gi = 0; gi < ng; gi++) {
                    const int j = groups[G][g0 + gi].first;
                    const bool sk = groups[G][g0 + gi].second != 0;
                    const Prod &p0 = prods[b0 + j];
                    for (int t = 0; t < nsets; t++) {
                        a.shared[t][gi] = sk ? p0.key[t] : p0.hst[t];
                        for (int g = 0; g < G; g++) a.priv[t][gi * 2 + g] = sk ? prods[b0 + j + g].hst[t] : prods[b0 + j + g].key[t];
                    }
                    for (int g = 0; g < G; g++) a.out[gi * 2 + g] = bufs[j + g];
                }
                // an operand that several groups of this launch stream (u of every party, a rotation's a, x / y) is worth keeping
                // in L2; everything else is read exactly once
                a.l2_hints = ctx->l2_hints ? 1 : 0;
                for (int gi = 0; gi < ng; gi++)
                    for (int o = 0; o <= G; o++) {
                        const u64 *ptr = o == 0 ? a.shared[0][gi] : a.priv[0][gi * 2 + o - 1];
                        int cnt = 0;
                        for (int gj = 0; gj < ng; gj++) {
                            cnt += a.shared[0][gj] == ptr;
                            for (int g = 0; g < G; g++) cnt += a.priv[0][gj * 2 + g] == ptr;
                        }
                        if (cnt > 1) {
                            if (o == 0) a.shared_multi |= 1u << gi;
                            else a.priv_multi |= (u64)1 << (2 * gi + o - 1);
                        }
                    }
                const long nunits = (long)s.n * tiles * ng;
                const dim3 grid((unsigned)std::min<long>(ctx->num_sms, nunits));
                if (G == 2) LAUNCH(k_mac_intt<2>, grid, dim3(MKHE_MI_THREADS(2)), MKHE_MI_SMEM(2), a, ctx->d_mods, ctx->d_twi_tiled, ctx->d_twi);
                else LAUNCH(k_mac_intt<1>, grid, dim3(MKHE_MI_THREADS(1)), MKHE_MI_SMEM(1), a, ctx->d_mods, ctx->d_twi_tiled, ctx->d_twi);
            }
        // ---- pass B fused with ModDown, the accumulation and (rotations) the automorphism
        ModDownPArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.np_limbs = ctx->nP;
        pa.p_slot0 = ctx->nQ;
        pa.vslot = vslot;
        pa.logN = ctx->logN;
        const size_t pp_elems = (size_t)2 * ctx->nP * ctx->N;        // P part of one product in the team memory
        for (int i = 0; i < ctx->nP; i++)
            if (!team || owner_of_slot(ctx, ctx->nQ + i) == ctx->team.rank) pa.plist[pa.nplist++] = i;
        if (team) fill_team(ctx, pa.team);
        for (int k = 0; k < n; k++) { pa.acc[k] = bufs[k]; pa.pp_off[k] = (long)(ctx->team.off_pp + (size_t)k * pp_elems); }
        if (pa.nplist > 0)
            TRY(dispatch_s1(ctx, [&](auto S) -> int {
                auto k_moddown_P_ = k_moddown_P<decltype(S)::value>;
                LAUNCH(k_moddown_P_, dim3(COLGROUPS, n, pa.nplist), dim3(MKHE_NTT_THREADS), 0, pa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi);
                return MKHE_OK;
            }));
        if (team) {
            TRY(team_barrier(ctx));                           // every rank has y_i of every product's P part
            ModDownOvArgs oa;                                 // ... and makes the overflow estimates from them, once per coefficient
            memset(&oa, 0, sizeof oa);
            oa.np_limbs = ctx->nP; oa.p_mod0 = ctx->nQ; oa.logN = ctx->logN;
            for (int k = 0; k < n; k++) oa.pp[k] = ctx->team.mem + ctx->team.off_pp + (size_t)k * pp_elems;
            LAUNCH(k_moddown_ov, dim3(ctx->N / MKHE_THREADS, n), dim3(MKHE_THREADS), 0, oa, ctx->d_mods);
        }
        for (size_t t0 = 0; t0 < tg.size(); t0 += MKHE_MD_TARGETS) {
            const int ntg = (int)std::min<size_t>(MKHE_MD_TARGETS, tg.size() - t0);
            ModDownQArgs qa;
            memset(&qa, 0, sizeof qa);
            qa.np_limbs = ctx->nP;
            for (int j = 0; j <= levelQ; j++)
                if (!team || owner_of_slot(ctx, j) == ctx->team.rank) qa.qlist[qa.nqlist++] = j;
            if (team) fill_team(ctx, qa.team);
            qa.galEl = galEl;
            if (galEl) {                 // galEl^-1 mod 2N (2-adic Newton iteration: every step doubles the number of correct bits)
                const u64 mask = ((u64)2 << ctx->logN) - 1;
                u64 inv = 1;
                for (int it = 0; it < 6; it++) inv = (inv * (2 - galEl * inv)) & mask;
                qa.galInv = inv;
            }
            qa.logN = ctx->logN;
            int cnt = 0, max_split = 1;
            for (int t = 0; t < ntg; t++) {
                const Prod &p0 = prods[b0 + members[t0 + t][0]];
                const size_t np = members[t0 + t].size();
                qa.split[t] = np >= 8 ? 8 : (np >= 4 ? 4 : (np >= 2 ? 2 : 1));
                max_split = std::max(max_split, qa.split[t]);
                qa.dst[t] = tg[t0 + t];
                qa.src[t] = p0.src ? p0.src : (p0.add ? tg[t0 + t] : nullptr);
                qa.dst_team_off[t] = team ? p0.team_off : -1;
                qa.first[t] = cnt;
                for (int m : members[t0 + t]) {
                    qa.pp[cnt] = team ? ctx->team.mem + ctx->team.off_pp + (size_t)m * pp_elems : bufs[m] + (size_t)ctx->nQ * ctx->N;
                    qa.acc[cnt++] = bufs[m];
                }
            }
            qa.first[ntg] = cnt;
            if (qa.nqlist == 0) continue;
            TRY(dispatch_s1(ctx, [&](auto S) -> int {
                constexpr int s1 = decltype(S)::value;
                const dim3 grid(COLGROUPS * max_split, qa.nqlist, ntg);
                const size_t rowsm = MKHE_MDQ_SMEM(s1);      // partial sums of split targets, the row permutation of a rotation
                switch (ctx->nP) {
                    case 1: { auto k_moddown_Q_ = team ? k_moddown_Q<s1, 1, true> : k_moddown_Q<s1, 1, false>; LAUNCH(k_moddown_Q_, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
                    case 2: { auto k_moddown_Q_ = team ? k_moddown_Q<s1, 2, true> : k_moddown_Q<s1, 2, false>; LAUNCH(k_moddown_Q_, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
                    case 3: { auto k_moddown_Q_ = team ? k_moddown_Q<s1, 3, true> : k_moddown_Q<s1, 3, false>; LAUNCH(k_moddown_Q_, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
                    case 4: { auto k_moddown_Q_ = team ? k_moddown_Q<s1, 4, true> : k_moddown_Q<s1, 4, false>; LAUNCH(k_moddown_Q_, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
                    default: return fail(ctx, MKHE_ERR_UNSUPPORTED, "key switch with %d special primes (1..4 supported)", ctx->nP);
                }
                return MKHE_OK;
            }));
        }
    }
    return MKHE_OK;
}
// array form: product b = (key0[b], hst0[b]) [+ (key1[b], hst1[b])] -> dst[b]; acc is nullptr or == dst
int ext_products(mkhe_ctx *ctx, int levelQ, int nb, int nsets, u64 *const *key0, u64 *const *hst0, u64 *const *key1,
                 u64 *const *hst1, u64 *const *dst, u64 *const *acc, bool /*shared_target*/) {
    std::vector<Prod> v(nb);
    for (int b = 0; b < nb; b++) {
        v[b].key[0] = key0[b]; v[b].hst[0] = hst0[b];
        v[b].key[1] = nsets > 1 ? key1[b] : nullptr; v[b].hst[1] = nsets > 1 ? hst1[b] : nullptr;
        v[b].dst = dst[b];
        v[b].add = acc != nullptr;
    }
    return ext_products(ctx, levelQ, nsets, v);
}

// x_i = MForm(sum_t MRed(key_t[i], hst_t[i]))   (keyswitch_hoisted.go:79-117)
int mac_parties(mkhe_ctx *ctx, int level, int n, u64 *const *key, u64 *const *hst, u64 *out, const Slots *only = nullptr) {
    Slots s = only ? *only : qp_slots(ctx, level);
    if (s.n == 0) return MKHE_OK;
    if (n > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_UNSUPPORTED, "more than %d parties", MKHE_MAX_PARTIES_K);
    MacPartiesArgs a;
    memset(&a, 0, sizeof a);
    a.out = out;
    a.nparties = n;
    a.beta = beta_of(ctx, level);
    a.dmax = ctx->dmax;
    a.nslots = s.n;
    a.logN = ctx->logN;
    for (int i = 0; i < s.n; i++) { a.slots[i] = s.slot[i]; a.mods[i] = s.mod[i]; }
    for (int i = 0; i < n; i++) { a.key.p[i] = key[i]; a.hst.p[i] = hst[i]; }
    LAUNCH(k_mac_parties, dim3(ctx->N / (2 * MKHE_THREADS), s.n, a.beta), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
}

int find_id(int n, const int *ids, int id) {
    for (int i = 0; i < n; i++) if (ids[i] == id) return i;
    return -1;
}

int add_polys(mkhe_ctx *ctx, int level, const u64 *x, const u64 *y, u64 *out, bool sub) {
    LimbArgs a;
    fill(a, q_slots(level), ctx->logN);
    if (sub) LAUNCH(k_addsub<true>, dim3(ctx->N / MKHE_THREADS, level + 1), dim3(MKHE_THREADS), 0, x, y, out, a, ctx->d_mods);
    else LAUNCH(k_addsub<false>, dim3(ctx->N / MKHE_THREADS, level + 1), dim3(MKHE_THREADS), 0, x, y, out, a, ctx->d_mods);
    return MKHE_OK;
}

// out = MForm(A) (.) B [+ MForm(C) (.) D]  (MFormLvl + MulCoeffsMontgomery[AndAdd]Lvl)
int mul2(mkhe_ctx *ctx, const Slots &s, const u64 *A, const u64 *B, const u64 *Cc, const u64 *D, u64 *out) {
    LimbArgs a;
    fill(a, s, ctx->logN);
    LAUNCH(k_mul2, dim3(ctx->N / MKHE_THREADS, s.n), dim3(MKHE_THREADS), 0, A, B, Cc, D, out, a, ctx->d_mods);
    return MKHE_OK;
}

// resolve arrays of handles
int polys_of(mkhe_ctx *ctx, int n, const mkhe_poly *h, int min_limbs, std::vector<u64 *> &out, const char *what, int mode = ACC_WRITE) {
    out.resize(n);
    for (int i = 0; i < n; i++) {
        Obj *o = as_obj(ctx, h[i], OBJ_POLY, mode);
        if (!o) return fail(ctx, MKHE_ERR_INVALID, "invalid poly handle in %s[%d]", what, i);
        if (o->cap_limbs < min_limbs) return fail(ctx, MKHE_ERR_INVALID, "%s[%d] has %d limbs, %d needed", what, i, o->cap_limbs, min_limbs);
        out[i] = o->d;
    }
    return MKHE_OK;
}
int swks_of(mkhe_ctx *ctx, int n, const mkhe_swk *h, std::vector<u64 *> &out, const char *what, int mode = ACC_WRITE) {
    out.resize(n);
    for (int i = 0; i < n; i++) {
        Obj *o = as_obj(ctx, h[i], OBJ_SWK, mode);
        if (!o) return fail(ctx, MKHE_ERR_INVALID, "invalid switching-key handle in %s[%d]", what, i);
        out[i] = o->d;
    }
    return MKHE_OK;
}
// context-owned swk-shaped scratch pools (mirrors swkPool1..3 / rlkSet.HoistPool)
int swk_pool(mkhe_ctx *ctx, const char *name, int n, std::vector<u64 *> &out) {
    u64 *base;
    TRY(get_scratch(ctx, name, swk_elems(ctx) * 8 * (size_t)std::max(n, 1), &base));
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = base + (size_t)i * swk_elems(ctx);
    return MKHE_OK;
}
int poly_pool(mkhe_ctx *ctx, const char *name, int n, int limbs, std::vector<u64 *> &out) {
    u64 *base;
    TRY(get_scratch(ctx, name, (size_t)limbs * ctx->N * 8 * (size_t)std::max(n, 1), &base));
    out.resize(n);
    for (int i = 0; i < n; i++) out[i] = base + (size_t)i * limbs * ctx->N;
    return MKHE_OK;
}

// party sharding: which parties this rank owns; inactive = single-GPU behaviour
struct Shard {
    bool active = false;
    std::vector<int> own;
    bool owns(int id) const { return !active || std::find(own.begin(), own.end(), id) != own.end(); }
};

int allreduce_mod(mkhe_ctx *ctx, u64 *buf, size_t count, const Slots &s, int nbufs, long buf_stride) {
#if !defined(MKHE_EMU) && defined(MKHE_WITH_NCCL)
    if (!ctx->nccl) return fail(ctx, MKHE_ERR_NCCL, "sharded op without mkhe_comm_init");
    if (ncclAllReduce(buf, buf, count, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
        return fail(ctx, MKHE_ERR_NCCL, "ncclAllReduce failed");
    if (nbufs > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_UNSUPPORTED, "too many buffers");
    LimbArgs a;
    fill(a, s, ctx->logN);
    for (int i = 0; i < nbufs; i++) { a.in.p[i] = buf + (size_t)i * buf_stride; a.out.p[i] = a.in.p[i]; }
    LAUNCH(k_reduce, dim3(ctx->N / MKHE_THREADS, s.n, nbufs), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
#else
    (void)buf; (void)count; (void)s; (void)nbufs; (void)buf_stride;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "built without NCCL");
#endif
}

// ---- fused exchange over peer memory ------------------------------------------------------------------
void fill_p2p(mkhe_ctx *ctx, P2PArgs &p, int which) {
    memset(&p, 0, sizeof p);
    p.nranks = ctx->nranks;
    p.rank = ctx->rank;
    p.which = which;
    p.swk_elems = (long)swk_elems(ctx);
    for (int r = 0; r < ctx->nranks; r++) { p.peer_stage[r] = ctx->peer_stage[r]; p.peer_xy[r] = ctx->peer_xy[r]; }
}
// a tiny all-reduce as a barrier in stream order: when it completes on this rank, every rank has finished the work it had
// enqueued before its own call (kernel completion makes the peer stores visible)
int p2p_barrier(mkhe_ctx *ctx) {
#if !defined(MKHE_EMU) && defined(MKHE_WITH_NCCL)
    if (ncclAllReduce(ctx->p2p_flag, ctx->p2p_flag, 1, ncclUint64, ncclSum, (ncclComm_t)ctx->nccl, ctx->stream) != ncclSuccess)
        return fail(ctx, MKHE_ERR_NCCL, "barrier all-reduce failed");
    return MKHE_OK;
#else
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "built without NCCL");
#endif
}
int mac_parties_scatter(mkhe_ctx *ctx, int level, int n, u64 *const *key, u64 *const *hst, int which) {
    Slots s = qp_slots(ctx, level);
    if (n > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_UNSUPPORTED, "more than %d parties", MKHE_MAX_PARTIES_K);
    MacPartiesArgs a;
    memset(&a, 0, sizeof a);
    a.nparties = n;                       // 0 parties: the rank contributes zeros (it still has to write its partial)
    a.beta = beta_of(ctx, level);
    a.dmax = ctx->dmax;
    a.nslots = s.n;
    a.logN = ctx->logN;
    for (int i = 0; i < s.n; i++) { a.slots[i] = s.slot[i]; a.mods[i] = s.mod[i]; }
    for (int i = 0; i < n; i++) { a.key.p[i] = key[i]; a.hst.p[i] = hst[i]; }
    P2PArgs p;
    fill_p2p(ctx, p, which);
    LAUNCH(k_mac_parties_scatter, dim3(ctx->N / (2 * MKHE_THREADS), s.n, a.beta), dim3(MKHE_THREADS), 0, a, p, ctx->d_mods);
    return MKHE_OK;
}
int reduce_gather(mkhe_ctx *ctx, int level) {
    Slots s = qp_slots(ctx, level);
    ReduceGatherArgs a;
    memset(&a, 0, sizeof a);
    a.beta = beta_of(ctx, level);
    a.dmax = ctx->dmax;
    a.nslots = s.n;
    a.logN = ctx->logN;
    for (int i = 0; i < s.n; i++) { a.slots[i] = s.slot[i]; a.mods[i] = s.mod[i]; }
    P2PArgs p;
    fill_p2p(ctx, p, 0);
    const long seg = ctx->N / ctx->nranks;
    LAUNCH(k_reduce_gather, dim3((unsigned)(seg / (2 * MKHE_THREADS)), s.n, 2 * a.beta), dim3(MKHE_THREADS), 0, a, p, ctx->d_mods);
    return MKHE_OK;
}

// MulAndRelinHoisted on raw device pointers (mkrlwe/keyswitch_hoisted.go:44-179).
// With an active Shard only the owned parties' keys / hoisted forms are touched: the partial x, y and the c_0
// contributions are summed over ranks (exact: every accumulation of the reference is a modular add, App. A.4);
// valid outputs on a rank are component "0" and the components of the parties it owns.
// hoist0 / hoist1: when set, h0 / h1 are empty pools and the operand is hoisted here (MulRelinNew's implicit hoisting,
// mkckks/evaluator.go:419-441).
int mul_relin_hoisted_impl(mkhe_ctx *ctx, int level, int n0, const int *ids0, u64 *const *op0, u64 *const *h0,
                           int n1, const int *ids1, u64 *const *op1, u64 *const *h1, u64 *const *rlk_b,
                           u64 *const *rlk_d, u64 *const *rlk_v, u64 *u, int nOut, const int *idsOut, u64 *const *out,
                           const Shard &sh = Shard(), bool hoist0 = false, bool hoist1 = false, bool fresh0 = false, bool fresh1 = false) {
    fresh0 = fresh0 || hoist0;          // hoisted in this call: the form is Decompose(op) by construction
    fresh1 = fresh1 || hoist1;
    const int N = ctx->N;
    Slots qps = qp_slots(ctx, level);
    // owned sub-lists
    std::vector<int> o0, o1;
    for (int t = 0; t < n0; t++) if (sh.owns(ids0[t])) o0.push_back(t);
    for (int t = 0; t < n1; t++) if (sh.owns(ids1[t])) o1.push_back(t);
    auto pick = [](const std::vector<int> &idx, u64 *const *arr) {
        std::vector<u64 *> r(idx.size());
        for (size_t i = 0; i < idx.size(); i++) r[i] = arr[idx[i]];
        return r;
    };
    std::vector<u64 *> d_o = pick(o0, rlk_d), v_o = pick(o0, rlk_v), h0_o = pick(o0, h0), b_o = pick(o1, rlk_b), h1_o = pick(o1, h1);
    const int m0 = (int)o0.size(), m1 = (int)o1.size();
    std::vector<u64 *> xy;
    const bool p2p = sh.active && ctx->p2p;
    if (p2p) { xy.push_back(ctx->p2p_xy); xy.push_back(ctx->p2p_xy + swk_elems(ctx)); }
    else TRY(swk_pool(ctx, "xy", 2, xy));
    u64 *x = xy[0], *y = xy[1];
    std::vector<u64 *> p, hp, tn;
    TRY(poly_pool(ctx, "relin_p", m0, ctx->nQ, p));
    TRY(swk_pool(ctx, "relin_hp", m0, hp));
    TRY(poly_pool(ctx, "tensor_ntt", n0 + n1 + 2, ctx->nQ, tn));
    if (nOut + 1 > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_UNSUPPORTED, "too many parties");

    // Forms made and consumed inside this op skip the canonical reduction of the NTT outputs: their consumers are the 128-bit
    // multiply-accumulates (<= 14 terms of < 2^60 * 2^64 fit) and the tensor product's Montgomery multiply, all of which reduce
    // properly, so every value that leaves the op is the same canonical residue.
    const bool lazy = ctx->alpha == 1 && beta_of(ctx, level) <= 14;
    if (hoist0) TRY(decompose_impl(ctx, level, n0, op0 + 1, h0, 0, lazy));
    if (hoist1) TRY(decompose_impl(ctx, level, n1, op1 + 1, h1, 0, lazy));
    // steps 2-3 (:79-117): x = MForm(sum d_id (.) h0_id), y = MForm(sum b_id (.) h1_id)
    if (p2p) {
        // fused: the partial sums are written straight into their owners' memory, summed there and the results written
        // straight into every rank's x||y (peer stores over NVLink); two stream-ordered barriers instead of a 112 MiB all-reduce
        TRY(mac_parties_scatter(ctx, level, m0, d_o.data(), h0_o.data(), 0));
        TRY(mac_parties_scatter(ctx, level, m1, b_o.data(), h1_o.data(), 1));
        TRY(p2p_barrier(ctx));
        TRY(reduce_gather(ctx, level));
        TRY(p2p_barrier(ctx));
    } else {
        if (m0) TRY(mac_parties(ctx, level, m0, d_o.data(), h0_o.data(), x));
        else CU(cudaMemsetAsync(x, 0, swk_elems(ctx) * 8, ctx->stream));
        if (m1) TRY(mac_parties(ctx, level, m1, b_o.data(), h1_o.data(), y));
        else CU(cudaMemsetAsync(y, 0, swk_elems(ctx) * 8, ctx->stream));
        if (sh.active) TRY(allreduce_mod(ctx, x, 2 * swk_elems(ctx), qps, 2 * ctx->beta_max, (long)ctx->dmax * N));
    }

    // step 4 (:119-144): tensor product in the NTT domain, then InvNTT of every output component.
    // With alpha = 1, digit i of a hoisted form is limb i of the poly broadcast to every modulus, so its limb i IS
    // NTT_{q_i}(poly limb i): when the library has just hoisted the operand itself (fresh0 / fresh1), NTT(op_id) is read from the
    // diagonal of h_id instead of being transformed again -- the same kernel produced it from the same input, bit for bit.
    {
        Slots qs = q_slots(level);
        const bool nodiag = ctx->debug_nodiag;
        const bool dg0 = fresh0 && ctx->alpha == 1 && !nodiag, dg1 = fresh1 && ctx->alpha == 1 && !nodiag;
        std::vector<u64 *> src, dst;
        src.push_back(op0[0]); dst.push_back(tn[0]);
        src.push_back(op1[0]); dst.push_back(tn[n0 + 1]);
        if (!dg0) for (int t = 1; t <= n0; t++) { src.push_back(op0[t]); dst.push_back(tn[t]); }
        if (!dg1) for (int t = 1; t <= n1; t++) { src.push_back(op1[t]); dst.push_back(tn[n0 + 1 + t]); }
        TRY(ntt_fwd(ctx, qs, (int)src.size(), src.data(), dst.data()));
        TensorArgs ta;
        memset(&ta, 0, sizeof ta);
        ta.A0 = tn[0];
        ta.B0 = tn[n0 + 1];
        ta.strideA = dg0 ? (long)(ctx->dmax + 1) * N : N;
        ta.strideB = dg1 ? (long)(ctx->dmax + 1) * N : N;
        ta.nout = nOut;
        ta.nlimbs = level + 1;
        ta.logN = ctx->logN;
        for (int i = 0; i <= level; i++) ta.limbs[i] = i;
        ta.out.p[0] = out[0];
        for (int t = 0; t < nOut; t++) {
            int i0 = find_id(n0, ids0, idsOut[t]), i1 = find_id(n1, ids1, idsOut[t]);
            ta.A.p[t] = i0 >= 0 ? (dg0 ? h0[i0] : tn[1 + i0]) : nullptr;
            ta.B.p[t] = i1 >= 0 ? (dg1 ? h1[i1] : tn[n0 + 2 + i1]) : nullptr;
            ta.out.p[1 + t] = out[1 + t];
        }
        LAUNCH(k_tensor, dim3(N / MKHE_THREADS, level + 1), dim3(MKHE_THREADS), 0, ta, ctx->d_mods);
        if (sh.active) {        // only component "0" and the components of this rank's parties are completed here
            std::vector<u64 *> mine{out[0]};
            for (int t = 0; t < nOut; t++) if (sh.owns(idsOut[t])) mine.push_back(out[1 + t]);
            TRY(ntt_inv(ctx, qs, (int)mine.size(), mine.data(), mine.data()));
        } else {
            TRY(ntt_inv(ctx, qs, nOut + 1, out, out));
        }
        // the tensor term of c_0 is counted once across ranks
        if (sh.active && ctx->rank != 0) CU(cudaMemsetAsync(out[0], 0, (size_t)(level + 1) * N * 8, ctx->stream));
    }
    // step 5 (:147-154): c_id += x [.] h1_id   and the first half of step 6 (:161-166): p_id = y [.] h0_id, one batch
    {
        std::vector<Prod> pr;
        for (int t = 0; t < m1; t++) pr.push_back(Prod{{x, nullptr}, {h1_o[t], nullptr}, out[1 + find_id(nOut, idsOut, ids1[o1[t]])], true});
        for (int t = 0; t < m0; t++) pr.push_back(Prod{{y, nullptr}, {h0_o[t], nullptr}, p[t], false});
        TRY(ext_products(ctx, level, 1, pr));
    }
    // step 6 (:167-178): Decompose(p_id) ; c_id += u [.] p_id ; c_0 += v_id [.] p_id   (the two products of a party share p_id)
    if (m0 > 0) {
        TRY(decompose_impl(ctx, level, m0, p.data(), hp.data(), 0, lazy));
        std::vector<Prod> pr;
        for (int t = 0; t < m0; t++) {
            pr.push_back(Prod{{u, nullptr}, {hp[t], nullptr}, out[1 + find_id(nOut, idsOut, ids0[o0[t]])], true});
            pr.push_back(Prod{{v_o[t], nullptr}, {hp[t], nullptr}, out[0], true});
        }
        TRY(ext_products(ctx, level, 1, pr));
    }
    if (sh.active) TRY(allreduce_mod(ctx, out[0], (size_t)(level + 1) * N, q_slots(level), 1, 0));
    return MKHE_OK;
}

int rescale_impl(mkhe_ctx *ctx, int level, int nb, int npolys, u64 *const *in, u64 *const *out) {
    if (nb < 0 || nb > level) return fail(ctx, MKHE_ERR_INVALID, "cannot Rescale: nb_rescales = %d at level %d", nb, level);
    for (int r = 0; r < nb; r++) {
        const int l = level - r;
        for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
            int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
            RescaleArgs a;
            memset(&a, 0, sizeof a);
            a.level = l;
            a.logN = ctx->logN;
            const u64 ql = ctx->mod[l], h = (ql - 1) >> 1;
            for (int i = 0; i < l; i++) {
                const u64 qi = ctx->mod[i];
                a.rescale[i] = qi - h_mform(h_invmod(ql % qi, qi), qi);
                a.halfneg[i] = qi - (h % qi);
            }
            for (int i = 0; i < np; i++) { a.in.p[i] = (r == 0) ? in[p0 + i] : out[p0 + i]; a.out.p[i] = out[p0 + i]; }
            LAUNCH(k_rescale, dim3(ctx->N / MKHE_THREADS, np), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
        }
    }
    return MKHE_OK;
}

// Limb-sharded MulRelinNew (mkckks/evaluator.go:416-443 + mkrlwe/keyswitch_hoisted.go:44-179 + Rescale :359-398).
// Every rank holds the (replicated) operand ciphertexts and computes the limb slots it owns of every hoisted form, of x, y, of
// the tensor product and of every key-switch accumulator: steps 1-6 of MulAndRelinHoisted are limb-local except
//   (i)  ModDown needs the P limbs of every product on every rank     -> k_moddown_P stores them into every rank's team memory,
//   (ii) Decompose(p_id) needs all limbs of p_id on every rank        -> k_moddown_Q stores its limbs of p_id into every rank's,
//   (iii) Rescale needs the last limb, the caller needs a whole result -> k_team_gather stores the rank's limbs into every rank's,
// each followed by one in-kernel flag barrier (k_team_barrier): four barriers per op, no collective library on the data path.
// Bit-identical to the single-GPU op: every value is computed by the same kernels from the same inputs, only elsewhere.
int mul_relin_limbs_impl(mkhe_ctx *ctx, int level, int nb_rescales, int n0, const int *ids0, u64 *const *op0, int n1, const int *ids1,
                         u64 *const *op1, u64 *const *rlk_b, u64 *const *rlk_d, u64 *const *rlk_v, u64 *u, int nOut,
                         const int *idsOut, u64 *const *out) {
    mkhe_ctx::Team &tm = ctx->team;
    const int N = ctx->N;
    if (ctx->alpha != 1) return fail(ctx, MKHE_ERR_UNSUPPORTED, "limb-sharded MulRelin relies on alpha = 1 (the tensor step reads the hoisted diagonal)");
    if (std::max(n0, n1) > tm.maxk || nOut > tm.maxk) return fail(ctx, MKHE_ERR_INVALID, "team memory sized for %d parties", tm.maxk);
    const Slots qp_own = own_slots(ctx, qp_slots(ctx, level)), q_own = own_slots(ctx, q_slots(level));
    std::vector<u64 *> h0, h1, xy, hp, tn;
    TRY(swk_pool(ctx, "hoistpool0", n0, h0));
    TRY(swk_pool(ctx, "hoistpool1", n1, h1));
    TRY(swk_pool(ctx, "xy", 2, xy));
    TRY(swk_pool(ctx, "relin_hp", n0, hp));
    TRY(poly_pool(ctx, "tensor_ntt", 2, ctx->nQ, tn));
    const bool lazy = beta_of(ctx, level) <= 14;          // forms that never leave the op (see mul_relin_hoisted_impl)
    // hoisting and x, y: the owned target limbs of every digit (the digit limbs themselves come from the replicated operands)
    TRY(decompose_impl(ctx, level, n0, op0 + 1, h0.data(), 0, lazy, &qp_own));
    TRY(decompose_impl(ctx, level, n1, op1 + 1, h1.data(), 0, lazy, &qp_own));
    if (n0) TRY(mac_parties(ctx, level, n0, rlk_d, h0.data(), xy[0], &qp_own));
    if (n1) TRY(mac_parties(ctx, level, n1, rlk_b, h1.data(), xy[1], &qp_own));
    // tensor product on the owned Q limbs (NTT(op_id limb i) = limb i of digit i of the fresh form)
    if (q_own.n > 0) {
        u64 *src[2] = {op0[0], op1[0]}, *dst[2] = {tn[0], tn[1]};
        TRY(ntt_fwd(ctx, q_own, 2, src, dst));
        TensorArgs ta;
        memset(&ta, 0, sizeof ta);
        ta.A0 = tn[0];
        ta.B0 = tn[1];
        ta.strideA = ta.strideB = (long)(ctx->dmax + 1) * N;
        ta.nout = nOut;
        ta.nlimbs = q_own.n;
        ta.logN = ctx->logN;
        for (int i = 0; i < q_own.n; i++) ta.limbs[i] = q_own.slot[i];
        ta.out.p[0] = out[0];
        for (int t = 0; t < nOut; t++) {
            const int i0 = find_id(n0, ids0, idsOut[t]), i1 = find_id(n1, ids1, idsOut[t]);
            ta.A.p[t] = i0 >= 0 ? h0[i0] : nullptr;
            ta.B.p[t] = i1 >= 0 ? h1[i1] : nullptr;
            ta.out.p[1 + t] = out[1 + t];
        }
        LAUNCH(k_tensor, dim3(N / MKHE_THREADS, q_own.n), dim3(MKHE_THREADS), 0, ta, ctx->d_mods);
        TRY(ntt_inv(ctx, q_own, nOut + 1, out, out));
    }
    // c_id += x [.] h1_id and p_id = y [.] h0_id; the limbs of p_id go to every rank
    const size_t poly_elems = (size_t)ctx->nQ * N;
    std::vector<u64 *> p(n0);
    for (int t = 0; t < n0; t++) p[t] = tm.mem + tm.off_p + (size_t)t * poly_elems;
    {
        std::vector<Prod> pr;
        for (int t = 0; t < n1; t++) pr.push_back(Prod{{xy[0], nullptr}, {h1[t], nullptr}, out[1 + find_id(nOut, idsOut, ids1[t])], true});
        for (int t = 0; t < n0; t++) {
            Prod q{{xy[1], nullptr}, {h0[t], nullptr}, p[t], false};
            q.team_off = (long)(tm.off_p + (size_t)t * poly_elems);
            pr.push_back(q);
        }
        TRY(ext_products(ctx, level, 1, pr, 0, true));
        TRY(team_barrier(ctx));                            // every p_id is complete on every rank
    }
    // the rank's limbs of the result go to every rank's gather area: straight from the epilogue of the last ModDown for the
    // components it completes (component "0" and the parties of op0), by k_team_gather for the others
    std::vector<u64 *> g(nOut + 1);
    std::vector<bool> gathered(nOut + 1, false);
    for (int t = 0; t <= nOut; t++) g[t] = tm.mem + tm.off_g + (size_t)t * poly_elems;
    if (n0 > 0) {
        TRY(decompose_impl(ctx, level, n0, p.data(), hp.data(), 0, lazy, &qp_own));
        std::vector<Prod> pr;
        for (int t = 0; t < n0; t++) {
            const int c = 1 + find_id(nOut, idsOut, ids0[t]);
            Prod a{{u, nullptr}, {hp[t], nullptr}, out[c], true};
            a.team_off = (long)(tm.off_g + (size_t)c * poly_elems);
            Prod b{{rlk_v[t], nullptr}, {hp[t], nullptr}, out[0], true};
            b.team_off = (long)tm.off_g;
            pr.push_back(a);
            pr.push_back(b);
            gathered[c] = gathered[0] = true;
        }
        TRY(ext_products(ctx, level, 1, pr, 0, true));
    }
    if (q_own.n > 0) {
        GatherArgs ga;
        memset(&ga, 0, sizeof ga);
        ga.nslots = q_own.n;
        ga.logN = ctx->logN;
        for (int i = 0; i < q_own.n; i++) ga.slots[i] = q_own.slot[i];
        int ng = 0;
        for (int t = 0; t <= nOut; t++)
            if (!gathered[t]) { ga.src.p[ng] = out[t]; ga.dst_off[ng++] = (long)(tm.off_g + (size_t)t * poly_elems); }
        if (ng > 0) {
            TeamArgs ta;
            fill_team(ctx, ta);
            LAUNCH(k_team_gather, dim3(N / (2 * MKHE_THREADS), q_own.n, ng), dim3(MKHE_THREADS), 0, ga, ta);
        }
    }
    TRY(team_barrier(ctx));
    if (nb_rescales > 0) {
        TRY(rescale_impl(ctx, level, 1, nOut + 1, g.data(), out));
        if (nb_rescales > 1) TRY(rescale_impl(ctx, level - 1, nb_rescales - 1, nOut + 1, out, out));
    } else {
        for (int t = 0; t <= nOut; t++) CU(cudaMemcpyAsync(out[t], g[t], (size_t)(level + 1) * N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    }
    return MKHE_OK;
}

int automorph_impl(mkhe_ctx *ctx, int level, u64 galEl, int npolys, u64 *const *in, u64 *const *out) {
    for (int p0 = 0; p0 < npolys; p0 += MKHE_MAX_PARTIES_K) {
        int np = std::min(MKHE_MAX_PARTIES_K, npolys - p0);
        LimbArgs a;
        fill(a, q_slots(level), ctx->logN);
        for (int i = 0; i < np; i++) { a.in.p[i] = in[p0 + i]; a.out.p[i] = out[p0 + i]; }
        LAUNCH(k_automorph, dim3(ctx->N / MKHE_THREADS, level + 1, np), dim3(MKHE_THREADS), 0, a, galEl, ctx->d_mods);
    }
    return MKHE_OK;
}
u64 galois_for_rotation(int logN, int k) {
    u64 twoN = (u64)1 << (logN + 1), mask = twoN - 1;
    u64 e = (u64)((long long)k) & mask, r = 1, b = 5;
    while (e) { if (e & 1) r = (r * b) & mask; b = (b * b) & mask; e >>= 1; }
    return r;
}

// RotateHoisted on raw pointers (mkrlwe/keyswitch_hoisted.go:183-247)
int rotate_hoisted_impl(mkhe_ctx *ctx, int level, int rotidx, int n, u64 *const *ct_in, u64 *const *hoisted,
                        u64 *const *rk, u64 *a, u64 *const *out) {
    while (rotidx < 0) rotidx += ctx->N / 2;
    const u64 galEl = galois_for_rotation(ctx->logN, rotidx);
    bool alias = false;                   // the permuted store scatters over the whole output: it must not be an input
    for (int t = 0; t <= n; t++)
        for (int r = 0; r <= n; r++) alias = alias || out[t] == ct_in[r];
    std::vector<Prod> pr;
    if (!alias) {
        // ctOut["0"] = ctIn["0"] + sum_id rk_id [.] h_id, ctOut[id] = a [.] h_id, each stored through the automorphism by the
        // ModDown kernel itself (no copy of ctIn["0"], no separate permutation pass)
        for (int t = 0; t < n; t++) {     // the two products of a party share its hoisted form
            pr.push_back(Prod{{a, nullptr}, {hoisted[t], nullptr}, out[1 + t], false});
            pr.push_back(Prod{{rk[t], nullptr}, {hoisted[t], nullptr}, out[0], false, ct_in[0]});
        }
        if (n == 0) return automorph_impl(ctx, level, galEl, 1, ct_in, out);
        return ext_products(ctx, level, 1, pr, galEl);
    }
    std::vector<u64 *> tmp;
    TRY(poly_pool(ctx, "rot_tmp", n + 1, ctx->nQ, tmp));
    CU(cudaMemcpyAsync(tmp[0], ct_in[0], (size_t)(level + 1) * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    for (int t = 0; t < n; t++) {
        pr.push_back(Prod{{a, nullptr}, {hoisted[t], nullptr}, tmp[1 + t], false});
        pr.push_back(Prod{{rk[t], nullptr}, {hoisted[t], nullptr}, tmp[0], true});
    }
    TRY(ext_products(ctx, level, 1, pr));
    return automorph_impl(ctx, level, galEl, n + 1, tmp.data(), out);
}

}  // namespace

namespace {
// Objects live in the device's stream-ordered memory pool (cudaMallocAsync): allocating and freeing never synchronise the device
// (the reference allocates a fresh ciphertext per evaluator call, mkckks/evaluator.go:306-316, and leaves it to the GC; a cudaMalloc /
// cudaFree pair per poly would stall every lane).  An object is released on the freeing lane's stream, after the last use on
// every other lane and after its last asynchronous transfer.
int free_object_ordered(mkhe_ctx *ctx, Obj *o) {
    for (mkhe_ctx *l : ctx->root->lanes) {
        if (!l || l == ctx) continue;
        const Obj::Use &u = o->use[l->lane];
        if (u.valid && u.ev) CU(cudaStreamWaitEvent(ctx->stream, u.ev, 0));
    }
    if (o->xfer) CU(cudaStreamWaitEvent(ctx->stream, o->xfer, 0));
    CU(cudaFreeAsync(o->d, ctx->stream));
    if (o->xfer) cudaEventDestroy(o->xfer);
    for (auto &u : o->use) if (u.wev_own) cudaEventDestroy(u.wev_own);
    ctx->root->objs.erase(o);
    ctx->touched.clear();
    delete o;
    return MKHE_OK;
}
// order copy stream `cs` after everything enqueued on the context's stream so far, run `copy` on it, and make the next
// user of the object on the context's stream wait for the copy
template <class F>
int async_transfer(mkhe_ctx *ctx, Obj *o, cudaStream_t cs, F &&copy) {
    CU(cudaEventRecord(ctx->ev_compute, ctx->stream));
    CU(cudaStreamWaitEvent(cs, ctx->ev_compute, 0));
    CU(copy());
    if (!o->xfer) CU(cudaEventCreate(&o->xfer));
    CU(cudaEventRecord(o->xfer, cs));
    o->xfer_pending = true;
    ctx->override_ev = o->xfer;          // other lanes order themselves after the COPY, not after the call
    return MKHE_OK;
}
}  // namespace

// =====================================================================================================
// C ABI
// =====================================================================================================
extern "C" {

const char *mkhe_version(void) {
#ifdef MKHE_EMU
    return "mkhe-b200 0.1 (CPU EMULATION BUILD - development only)";
#else
    return "mkhe-b200 0.1 (sm_100a)";
#endif
}

int mkhe_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int mkhe_ctx_create(int logN, const uint64_t *Q, int nQ, const uint64_t *P, int nP, int gamma, int device, mkhe_ctx **out) {
    if (!out || !Q || !P) return MKHE_ERR_INVALID;
    *out = nullptr;
    if (logN < 12 || logN > 16) return MKHE_ERR_UNSUPPORTED;
    // alpha = #P/gamma limbs per digit (mkrlwe/params.go:63-65); the ModDown kernels are instantiated for 1..4 special primes
    if (gamma <= 0 || nP < 1 || nP > 4 || nP / gamma < 1 || nP / gamma > MKHE_LIFT_MAX_SRC) return MKHE_ERR_UNSUPPORTED;
    if (nQ < 1 || nQ + nP > MKHE_MAX_SLOTS || nQ > MKHE_CONV_MAX) return MKHE_ERR_UNSUPPORTED;
    int ndev = mkhe_device_count();
    if (ndev <= 0 || device < 0 || device >= ndev) return MKHE_ERR_CUDA;   // no GPU: fail loudly, there is no CPU path
    mkhe_ctx *ctx = new mkhe_ctx();
    {
        static int created = 0;
        created++;
        const char *e = getenv("MKHE_DEBUG_SYNC");
        if (e && atoi(e) == created) ctx->debug_sync = true;
        e = getenv("MKHE_DEBUG_CK");
        if (e && atoi(e) == created) ctx->debug_ck = true;
        ctx->debug_nodiag = getenv("MKHE_DEBUG_NODIAG") != nullptr;
        ctx->l2_hints = getenv("MKHE_DEBUG_NO_L2_HINTS") == nullptr;
        ctx->graphs_on = getenv("MKHE_GRAPHS") != nullptr && atoi(getenv("MKHE_GRAPHS")) != 0;
        if (const char *hd = getenv("MKHE_DEBUG_HOIST_DIGITS")) ctx->hoist_digits = atoi(hd);
    }
    ctx->root = ctx;
    ctx->lanes.push_back(ctx);
    ctx->logN = logN; ctx->N = 1 << logN; ctx->nQ = nQ; ctx->nP = nP; ctx->gamma = gamma; ctx->device = device;
    ctx->S1 = logN - 11;
    ctx->dmax = nQ + nP;
    ctx->alpha = nP / gamma;
    ctx->beta_max = (nQ + ctx->alpha - 1) / ctx->alpha;
    for (int i = 0; i < nQ; i++) ctx->mod.push_back(Q[i]);
    for (int i = 0; i < nP; i++) ctx->mod.push_back(P[i]);
    for (u64 q : ctx->mod) {
        if (q >= ((u64)1 << 60) || q < ((u64)1 << 40) || !h_is_prime(q) || (q - 1) % ((u64)2 << logN) != 0) { delete ctx; return MKHE_ERR_UNSUPPORTED; }
    }
    ctx->tabs.resize(ctx->mod.size());
    for (size_t i = 0; i < ctx->mod.size(); i++) gen_mod_tables(ctx->tabs[i], logN, ctx->mod[i]);
    if (cudaSetDevice(device) != cudaSuccess) { delete ctx; return MKHE_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MKHE_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MKHE_ERR_CUDA; }
    if (cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return MKHE_ERR_CUDA; }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && v > 0) ctx->num_sms = v; }
    cudaEventCreate(&ctx->ev_compute);
    cudaEventCreate(&ctx->ev0);
    cudaEventCreate(&ctx->ev1);
#ifndef MKHE_EMU
    {   // keep freed blocks in the device's pool instead of returning them to the driver at every synchronisation
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = UINT64_MAX;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    cudaFuncSetAttribute(k_ntt_pass2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_PASS2);
    cudaFuncSetAttribute(k_intt_passA, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MKHE_PA_SMEM);
    cudaFuncSetAttribute(k_mac_intt<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MKHE_MI_SMEM(1));
    cudaFuncSetAttribute(k_mac_intt<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MKHE_MI_SMEM(2));
#endif
    preload_kernels(ctx);
    std::vector<int> src, dst;
    for (int j = 0; j < nP; j++) src.push_back(nQ + j);
    for (int j = 0; j < nQ; j++) dst.push_back(j);
    int rc = upload_conv(ctx, &ctx->d_conv_PtoQ, src, dst);
    if (rc == MKHE_OK) rc = upload_tables(ctx);
    if (rc == MKHE_OK && ctx->alpha > 1) rc = upload_lift_tables(ctx);
    if (rc != MKHE_OK) { delete ctx; return rc; }
    *out = ctx;
    return MKHE_OK;
}

static void mkhe_ctx_destroy_lane_streams(mkhe_ctx *f) {
    if (f->stream) cudaStreamDestroy(f->stream);
    if (f->h2d) cudaStreamDestroy(f->h2d);
    if (f->d2h) cudaStreamDestroy(f->d2h);
    if (f->ev_compute) cudaEventDestroy(f->ev_compute);
    if (f->ev0) cudaEventDestroy(f->ev0);
    if (f->ev1) cudaEventDestroy(f->ev1);
}
// a lane: shares the root's moduli, tables, keys and ciphertext objects (handles are valid on every lane), owns its stream,
// copy streams and scratch pools.  Ops on different lanes run concurrently on the device; uses of one object on different lanes
// are ordered automatically (events).  A context and its forks are driven from one host thread at a time.
int mkhe_ctx_fork(mkhe_ctx *parent, mkhe_ctx **out) {
    mkhe_ctx *ctx = parent;
    CHECK_CTX();
    if (!out) return MKHE_ERR_INVALID;
    mkhe_ctx *root = parent->root;
    int lane = -1;
    for (int l = 1; l < MKHE_MAX_LANES; l++)
        if (l >= (int)root->lanes.size() || !root->lanes[l]) { lane = l; break; }
    if (lane < 0) return fail(parent, MKHE_ERR_UNSUPPORTED, "at most %d lanes per root context", MKHE_MAX_LANES);
    TRY(upload_tables(root));
    mkhe_ctx *f = new mkhe_ctx();
    f->logN = root->logN; f->N = root->N; f->nQ = root->nQ; f->nP = root->nP; f->nQMul = root->nQMul; f->gamma = root->gamma;
    f->device = root->device; f->num_sms = root->num_sms; f->S1 = root->S1; f->dmax = root->dmax; f->alpha = root->alpha;
    f->beta_max = root->beta_max; f->d_lift = root->d_lift; f->T = root->T; f->mod = root->mod; f->tabs = root->tabs;
    f->d_mods = root->d_mods; f->d_twf = root->d_twf; f->d_twi = root->d_twi; f->d_twf_tiled = root->d_twf_tiled; f->d_twi_tiled = root->d_twi_tiled;
    f->tables_dirty = false;
    f->debug_nodiag = root->debug_nodiag;
    f->l2_hints = root->l2_hints;
    f->graphs_on = root->graphs_on;
    f->hoist_digits = root->hoist_digits;
    f->d_conv_PtoQ = root->d_conv_PtoQ; f->d_conv_QtoQMul = root->d_conv_QtoQMul; f->d_conv_QMultoQ = root->d_conv_QMultoQ;
    f->h_mformQMul = root->h_mformQMul;
    f->root = root;
    f->lane = lane;
    if (cudaStreamCreateWithFlags(&f->stream, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&f->h2d, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&f->d2h, cudaStreamNonBlocking) != cudaSuccess) { delete f; return fail(parent, MKHE_ERR_CUDA, "stream creation failed"); }
    cudaEventCreate(&f->ev_compute);
    cudaEventCreate(&f->ev0);
    cudaEventCreate(&f->ev1);
    // Work queued on the live lanes BEFORE this fork left no per-object use records when the root was alone (the bookkeeping is
    // skipped with a single lane): the new lane's stream starts after everything enqueued so far on every live lane, so its
    // first use of any existing object (pending outputs, the memsets of fresh allocations) is ordered after it.
    for (mkhe_ctx *l : root->lanes) {
        if (!l) continue;
        if (cudaEventRecord(l->ev_compute, l->stream) != cudaSuccess || cudaStreamWaitEvent(f->stream, l->ev_compute, 0) != cudaSuccess) {
            mkhe_ctx_destroy_lane_streams(f);
            delete f;
            return fail(parent, MKHE_ERR_CUDA, "fork: ordering the new lane after the live lanes failed");
        }
    }
    if ((int)root->lanes.size() <= lane) root->lanes.resize(lane + 1, nullptr);
    root->lanes[lane] = f;
    *out = f;
    return MKHE_OK;
}
// ctx's stream waits for everything enqueued so far on other's stream (both lanes of one root)
int mkhe_ctx_wait(mkhe_ctx *ctx, mkhe_ctx *other) {
    CHECK_CTX();
    if (!other || other->root != ctx->root) return fail(ctx, MKHE_ERR_INVALID, "mkhe_ctx_wait: not lanes of one root context");
    if (other == ctx) return MKHE_OK;
    CU(cudaEventRecord(other->ev_compute, other->stream));
    CU(cudaStreamWaitEvent(ctx->stream, other->ev_compute, 0));
    return MKHE_OK;
}

int mkhe_ctx_set_bfv(mkhe_ctx *ctx, const uint64_t *QMul, int nQMul, uint64_t T) {
    CHECK_CTX();
    if (!QMul || nQMul != ctx->nQ) return fail(ctx, MKHE_ERR_INVALID, "cannot NewParametersFromLiteral: length of Q & QMul is not equal");
    if (ctx->nQMul) return fail(ctx, MKHE_ERR_INVALID, "BFV parameters already set");
    if (ctx->root != ctx || ctx->multi()) return fail(ctx, MKHE_ERR_INVALID, "mkhe_ctx_set_bfv must precede mkhe_ctx_fork");
    if (ctx->alpha != 1) return fail(ctx, MKHE_ERR_UNSUPPORTED, "mkbfv relies on alpha = #P/gamma = 1 (mkbfv/keyswitch.go:64-67)");
    if (ctx->mod.size() + (size_t)nQMul > 64) return fail(ctx, MKHE_ERR_UNSUPPORTED, "more than 64 moduli");
    for (int i = 0; i < nQMul; i++) {
        u64 q = QMul[i];
        if (q >= ((u64)1 << 60) || q < ((u64)1 << 40) || !h_is_prime(q) || (q - 1) % ((u64)2 << ctx->logN) != 0) return fail(ctx, MKHE_ERR_UNSUPPORTED, "bad QMul prime");
        ctx->mod.push_back(q);
        ctx->tabs.emplace_back();
        gen_mod_tables(ctx->tabs.back(), ctx->logN, q);
    }
    ctx->nQMul = nQMul;
    ctx->T = T;
    ctx->tables_dirty = true;
    TRY(upload_tables(ctx));
    std::vector<int> q, qm;
    for (int j = 0; j < ctx->nQ; j++) { q.push_back(j); qm.push_back(ctx->nQ + ctx->nP + j); }
    TRY(upload_conv(ctx, &ctx->d_conv_QtoQMul, q, qm));
    TRY(upload_conv(ctx, &ctx->d_conv_QMultoQ, qm, q));
    ctx->h_mformQMul.resize(ctx->nQ);
    for (int i = 0; i < ctx->nQ; i++) {
        u64 qi = ctx->mod[i], r = 1;
        for (int j = 0; j < nQMul; j++) r = h_mulmod(r, QMul[j] % qi, qi);
        ctx->h_mformQMul[i] = h_mform(r, qi);
    }
    return MKHE_OK;
}

int mkhe_ctx_set_ntt_tables(mkhe_ctx *ctx, int m, const uint64_t *nttPsi, const uint64_t *nttPsiInv, uint64_t nttNInv) {
    CHECK_CTX();
    if (m < 0 || m >= (int)ctx->mod.size() || !nttPsi || !nttPsiInv) return fail(ctx, MKHE_ERR_INVALID, "bad modulus index %d", m);
    if (ctx->root != ctx || ctx->multi()) return fail(ctx, MKHE_ERR_INVALID, "mkhe_ctx_set_ntt_tables must precede mkhe_ctx_fork");
    set_mod_tables_from_mont(ctx->tabs[m], ctx->logN, ctx->mod[m], (const u64 *)nttPsi, (const u64 *)nttPsiInv, nttNInv);
    ctx->tables_dirty = true;
    CU(cudaStreamSynchronize(ctx->stream));
    return upload_tables(ctx);
}

void mkhe_ctx_destroy(mkhe_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->root == ctx) {                                  // the root takes its forks with it
        for (size_t l = 1; l < ctx->lanes.size(); l++)
            if (ctx->lanes[l]) mkhe_ctx_destroy(ctx->lanes[l]);
    }
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->h2d);
    cudaStreamSynchronize(ctx->d2h);
    mkhe_comm_destroy(ctx);
    drop_graphs(ctx);
#ifndef MKHE_EMU
    for (int r = 0; r < MKHE_MAX_RANKS; r++) {
        if (ctx->peer_stage[r] && ctx->peer_stage[r] != ctx->p2p_stage) cudaIpcCloseMemHandle(ctx->peer_stage[r]);
        if (ctx->peer_xy[r] && ctx->peer_xy[r] != ctx->p2p_xy) cudaIpcCloseMemHandle(ctx->peer_xy[r]);
    }
    if (ctx->p2p_stage) { cudaFree(ctx->p2p_stage); cudaFree(ctx->p2p_xy); cudaFree(ctx->p2p_flag); }
#endif
#ifndef MKHE_EMU
    for (int r = 0; r < MKHE_MAX_RANKS; r++)
        if (ctx->team.ipc[r] && ctx->team.peer[r]) cudaIpcCloseMemHandle(ctx->team.peer[r]);
#endif
    if (ctx->team.mem) cudaFree(ctx->team.mem);
    for (auto &kv : ctx->scratch) cudaFree(kv.second.p);
    for (cudaEvent_t e : ctx->lane_ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
    if (ctx->root == ctx) {
        for (Obj *o : ctx->objs) {
            if (o->xfer) cudaEventDestroy(o->xfer);
            for (auto &u : o->use) if (u.wev_own) cudaEventDestroy(u.wev_own);
            cudaFree(o->d);
            delete o;
        }
        cudaFree(ctx->d_mods); cudaFree(ctx->d_twf); cudaFree(ctx->d_twi); cudaFree(ctx->d_twf_tiled); cudaFree(ctx->d_twi_tiled);
        cudaFree(ctx->d_conv_PtoQ); cudaFree(ctx->d_conv_QtoQMul); cudaFree(ctx->d_conv_QMultoQ); cudaFree(ctx->d_lift);
    } else {
        for (Obj *o : ctx->root->objs) {                                      // the lane's work has completed (synchronised above)
            if (o->use[ctx->lane].wev_own) cudaEventDestroy(o->use[ctx->lane].wev_own);
            o->use[ctx->lane] = Obj::Use();
        }
        ctx->root->lanes[ctx->lane] = nullptr;
        while (!ctx->root->lanes.empty() && !ctx->root->lanes.back()) ctx->root->lanes.pop_back();
    }
    cudaEventDestroy(ctx->ev0); cudaEventDestroy(ctx->ev1);
    cudaEventDestroy(ctx->ev_compute);
    cudaStreamDestroy(ctx->stream);
    cudaStreamDestroy(ctx->h2d);
    cudaStreamDestroy(ctx->d2h);
    delete ctx;
}

const char *mkhe_last_error(const mkhe_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mkhe_sync(mkhe_ctx *ctx) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->h2d));
    CU(cudaStreamSynchronize(ctx->d2h));
    return MKHE_OK;
}
int mkhe_host_alloc(mkhe_ctx *ctx, size_t bytes, void **out) {
    CHECK_CTX();
    if (!out) return MKHE_ERR_INVALID;
    if (cudaMallocHost(out, bytes) != cudaSuccess) return fail(ctx, MKHE_ERR_NOMEM, "cudaMallocHost(%zu) failed", bytes);
    return MKHE_OK;
}
int mkhe_host_free(mkhe_ctx *ctx, void *p) {
    CHECK_CTX();
    CU(cudaFreeHost(p));
    return MKHE_OK;
}

uint64_t mkhe_launch_count(const mkhe_ctx *ctx) { return ctx ? ctx->launches : 0; }
int mkhe_ctx_set_graphs(mkhe_ctx *ctx, int on) {
    if (!ctx) return MKHE_ERR_INVALID;
    ctx->graphs_on = on != 0;
    if (!on) drop_graphs(ctx);
    return MKHE_OK;
}
int mkhe_graph_stats(const mkhe_ctx *ctx, uint64_t *captures, uint64_t *replays) {
    if (!ctx) return MKHE_ERR_INVALID;
    if (captures) *captures = ctx->graph_captures;
    if (replays) *replays = ctx->graph_replays;
    return MKHE_OK;
}

// ---- polys ------------------------------------------------------------------------------------------
int mkhe_poly_alloc(mkhe_ctx *ctx, int nlimbs, mkhe_poly *out) {
    CHECK_CTX();
    if (!out || nlimbs < 1 || nlimbs > 64) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    void *p = nullptr;
    size_t bytes = (size_t)nlimbs * ctx->N * 8;
    if (cudaMallocAsync(&p, bytes, ctx->stream) != cudaSuccess) { cudaGetLastError(); return fail(ctx, MKHE_ERR_NOMEM, "cudaMallocAsync(%zu) failed", bytes); }
    CU(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    Obj *o = new Obj();
    o->kind = OBJ_POLY; o->d = (u64 *)p; o->cap_limbs = nlimbs; o->nlimbs = nlimbs;
    ctx->root->objs.insert(o);
    if (ctx->multi()) ctx->touched.push_back({o, true});        // the memset above runs on this lane
    *out = reinterpret_cast<uint64_t>(o);
    return MKHE_OK;
}
int mkhe_poly_free(mkhe_ctx *ctx, mkhe_poly h) {
    CHECK_CTX();
    POLY(o, h);
    return free_object_ordered(ctx, o);
}
int mkhe_poly_set_nlimbs(mkhe_ctx *ctx, mkhe_poly h, int nlimbs) {
    CHECK_CTX();
    POLY(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs) return fail(ctx, MKHE_ERR_INVALID, "view of %d limbs exceeds capacity %d", nlimbs, o->cap_limbs);
    o->nlimbs = nlimbs;
    return MKHE_OK;
}
int mkhe_poly_get_nlimbs(mkhe_ctx *ctx, mkhe_poly h, int *nlimbs) {
    CHECK_CTX();
    POLY_R(o, h);
    *nlimbs = o->nlimbs;
    return MKHE_OK;
}
int mkhe_poly_upload_limb(mkhe_ctx *ctx, mkhe_poly h, int limb, const uint64_t *src) {
    CHECK_CTX();
    POLY(o, h);
    if (limb < 0 || limb >= o->cap_limbs || !src) return fail(ctx, MKHE_ERR_INVALID, "limb %d out of range", limb);
    CU(cudaMemcpyAsync(o->d + (size_t)limb * ctx->N, src, (size_t)ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_poly_download_limb(mkhe_ctx *ctx, mkhe_poly h, int limb, uint64_t *dst) {
    CHECK_CTX();
    POLY_R(o, h);
    if (limb < 0 || limb >= o->cap_limbs || !dst) return fail(ctx, MKHE_ERR_INVALID, "limb %d out of range", limb);
    CU(cudaMemcpyAsync(dst, o->d + (size_t)limb * ctx->N, (size_t)ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_poly_upload(mkhe_ctx *ctx, mkhe_poly h, const uint64_t *src, int nlimbs) {
    CHECK_CTX();
    POLY(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !src) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    CU(cudaMemcpyAsync(o->d, src, (size_t)nlimbs * ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_poly_download(mkhe_ctx *ctx, mkhe_poly h, uint64_t *dst, int nlimbs) {
    CHECK_CTX();
    POLY_R(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !dst) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    CU(cudaMemcpyAsync(dst, o->d, (size_t)nlimbs * ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_poly_upload_async(mkhe_ctx *ctx, mkhe_poly h, const uint64_t *src, int nlimbs) {
    CHECK_CTX();
    POLY(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !src) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    return async_transfer(ctx, o, ctx->h2d, [&] { return cudaMemcpyAsync(o->d, src, (size_t)nlimbs * ctx->N * 8, cudaMemcpyHostToDevice, ctx->h2d); });
}
int mkhe_poly_download_async(mkhe_ctx *ctx, mkhe_poly h, uint64_t *dst, int nlimbs) {
    CHECK_CTX();
    POLY_R(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !dst) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    return async_transfer(ctx, o, ctx->d2h, [&] { return cudaMemcpyAsync(dst, o->d, (size_t)nlimbs * ctx->N * 8, cudaMemcpyDeviceToHost, ctx->d2h); });
}
int mkhe_poly_upload_limb_async(mkhe_ctx *ctx, mkhe_poly h, int limb, const uint64_t *src) {
    CHECK_CTX();
    POLY(o, h);
    if (limb < 0 || limb >= o->cap_limbs || !src) return fail(ctx, MKHE_ERR_INVALID, "limb %d out of range", limb);
    return async_transfer(ctx, o, ctx->h2d, [&] { return cudaMemcpyAsync(o->d + (size_t)limb * ctx->N, src, (size_t)ctx->N * 8, cudaMemcpyHostToDevice, ctx->h2d); });
}
int mkhe_poly_download_limb_async(mkhe_ctx *ctx, mkhe_poly h, int limb, uint64_t *dst) {
    CHECK_CTX();
    POLY_R(o, h);
    if (limb < 0 || limb >= o->cap_limbs || !dst) return fail(ctx, MKHE_ERR_INVALID, "limb %d out of range", limb);
    return async_transfer(ctx, o, ctx->d2h, [&] { return cudaMemcpyAsync(dst, o->d + (size_t)limb * ctx->N, (size_t)ctx->N * 8, cudaMemcpyDeviceToHost, ctx->d2h); });
}
// the limbs of [0, nlimbs) this rank owns (limb mod nranks == rank; all of them without a team), between a host buffer in whole-poly
// layout and the poly: ONE strided DMA per call
int mkhe_poly_upload_owned_async(mkhe_ctx *ctx, mkhe_poly h, const uint64_t *src, int nlimbs) {
    CHECK_CTX();
    POLY(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !src) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    const int n = std::max(ctx->team.n, 1), first = ctx->team.n > 1 ? ctx->team.rank : 0;
    if (first >= nlimbs) return MKHE_OK;
    const size_t cnt = (size_t)(nlimbs - first + n - 1) / n, row = (size_t)ctx->N * 8;
    return async_transfer(ctx, o, ctx->h2d, [&] {
        return cudaMemcpy2DAsync(o->d + (size_t)first * ctx->N, row * n, src + (size_t)first * ctx->N, row * n, row, cnt, cudaMemcpyHostToDevice, ctx->h2d);
    });
}
int mkhe_poly_download_owned_async(mkhe_ctx *ctx, mkhe_poly h, uint64_t *dst, int nlimbs) {
    CHECK_CTX();
    POLY_R(o, h);
    if (nlimbs < 1 || nlimbs > o->cap_limbs || !dst) return fail(ctx, MKHE_ERR_INVALID, "bad limb count %d", nlimbs);
    const int n = std::max(ctx->team.n, 1), first = ctx->team.n > 1 ? ctx->team.rank : 0;
    if (first >= nlimbs) return MKHE_OK;
    const size_t cnt = (size_t)(nlimbs - first + n - 1) / n, row = (size_t)ctx->N * 8;
    return async_transfer(ctx, o, ctx->d2h, [&] {
        return cudaMemcpy2DAsync(dst + (size_t)first * ctx->N, row * n, o->d + (size_t)first * ctx->N, row * n, row, cnt, cudaMemcpyDeviceToHost, ctx->d2h);
    });
}
int mkhe_poly_copy(mkhe_ctx *ctx, mkhe_poly dsth, mkhe_poly srch) {
    CHECK_CTX();
    POLY(d, dsth);
    POLY_R(s, srch);
    if (d->cap_limbs < s->nlimbs) return fail(ctx, MKHE_ERR_INVALID, "copy: destination too small");
    CU(cudaMemcpyAsync(d->d, s->d, (size_t)s->nlimbs * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    d->nlimbs = s->nlimbs;
    return MKHE_OK;
}

int mkhe_poly_copy_lvl(mkhe_ctx *ctx, int level, mkhe_poly dsth, mkhe_poly srch) {
    CHECK_CTX();
    POLY(d, dsth);
    POLY_R(s, srch);
    if (level < 0 || d->cap_limbs < level + 1 || s->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "copy: level %d out of range", level);
    CU(cudaMemcpyAsync(d->d, s->d, (size_t)(level + 1) * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return MKHE_OK;
}

// ---- switching keys ---------------------------------------------------------------------------------
int mkhe_swk_alloc(mkhe_ctx *ctx, mkhe_swk *out) {
    CHECK_CTX();
    if (!out) return MKHE_ERR_INVALID;
    void *p = nullptr;
    size_t bytes = swk_elems(ctx) * 8;
    if (cudaMallocAsync(&p, bytes, ctx->stream) != cudaSuccess) { cudaGetLastError(); return fail(ctx, MKHE_ERR_NOMEM, "cudaMallocAsync(%zu) failed", bytes); }
    CU(cudaMemsetAsync(p, 0, bytes, ctx->stream));
    Obj *o = new Obj();
    o->kind = OBJ_SWK; o->d = (u64 *)p; o->cap_limbs = 0; o->nlimbs = 0;
    ctx->root->objs.insert(o);
    if (ctx->multi()) ctx->touched.push_back({o, true});
    *out = reinterpret_cast<uint64_t>(o);
    return MKHE_OK;
}
int mkhe_swk_free(mkhe_ctx *ctx, mkhe_swk h) {
    CHECK_CTX();
    SWK(o, h);
    return free_object_ordered(ctx, o);
}
static int swk_limb_off(mkhe_ctx *ctx, int digit, int is_p, int limb, size_t *off) {
    // a switching key holds beta_max = ceil(nQ / alpha) digits (keys.go:245-256), not nQ
    if (digit < 0 || digit >= ctx->beta_max) return fail(ctx, MKHE_ERR_INVALID, "digit %d out of range [0,%d)", digit, ctx->beta_max);
    if (limb < 0 || limb >= (is_p ? ctx->nP : ctx->nQ)) return fail(ctx, MKHE_ERR_INVALID, "limb %d out of range", limb);
    *off = ((size_t)digit * ctx->dmax + (is_p ? ctx->nQ : 0) + limb) * ctx->N;
    return MKHE_OK;
}
int mkhe_swk_upload_limb(mkhe_ctx *ctx, mkhe_swk h, int digit, int is_p, int limb, const uint64_t *src) {
    CHECK_CTX();
    SWK(o, h);
    size_t off;
    if (!src) return fail(ctx, MKHE_ERR_INVALID, "null source buffer");
    TRY(swk_limb_off(ctx, digit, is_p, limb, &off));
    CU(cudaMemcpyAsync(o->d + off, src, (size_t)ctx->N * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_swk_download_limb(mkhe_ctx *ctx, mkhe_swk h, int digit, int is_p, int limb, uint64_t *dst) {
    CHECK_CTX();
    SWK_R(o, h);
    size_t off;
    if (!dst) return fail(ctx, MKHE_ERR_INVALID, "null destination buffer");
    TRY(swk_limb_off(ctx, digit, is_p, limb, &off));
    CU(cudaMemcpyAsync(dst, o->d + off, (size_t)ctx->N * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_swk_upload(mkhe_ctx *ctx, mkhe_swk h, const uint64_t *src) {
    CHECK_CTX();
    SWK(o, h);
    if (!src) return fail(ctx, MKHE_ERR_INVALID, "null source buffer");
    CU(cudaMemcpyAsync(o->d, src, swk_elems(ctx) * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}
int mkhe_swk_download(mkhe_ctx *ctx, mkhe_swk h, uint64_t *dst) {
    CHECK_CTX();
    SWK_R(o, h);
    if (!dst) return fail(ctx, MKHE_ERR_INVALID, "null destination buffer");
    CU(cudaMemcpyAsync(dst, o->d, swk_elems(ctx) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    return MKHE_OK;
}

// ---- ring primitives --------------------------------------------------------------------------------
int mkhe_ntt(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out) {
    CHECK_CTX();
    POLY_R(i, in);
    POLY(o, out);
    if (level < 0 || level >= i->cap_limbs || level >= o->cap_limbs) return fail(ctx, MKHE_ERR_INVALID, "level %d out of range", level);
    // level >= nQ selects the BFV ring R = Q u QMul (2 nQ limbs): both polys must hold all of them
    if (level >= ctx->nQ && (ctx->nQMul == 0 || i->cap_limbs < 2 * ctx->nQ || o->cap_limbs < 2 * ctx->nQ))
        return fail(ctx, MKHE_ERR_INVALID, "level %d: a transform over R = Q u QMul needs BFV parameters and polys of %d limbs", level, 2 * ctx->nQ);
    Slots s = (level >= ctx->nQ) ? r_slots(ctx) : q_slots(level);
    return ntt_fwd(ctx, s, 1, &i->d, &o->d);
}
int mkhe_intt(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out) {
    CHECK_CTX();
    POLY_R(i, in);
    POLY(o, out);
    if (level < 0 || level >= i->cap_limbs || level >= o->cap_limbs) return fail(ctx, MKHE_ERR_INVALID, "level %d out of range", level);
    // level >= nQ selects the BFV ring R = Q u QMul (2 nQ limbs): both polys must hold all of them
    if (level >= ctx->nQ && (ctx->nQMul == 0 || i->cap_limbs < 2 * ctx->nQ || o->cap_limbs < 2 * ctx->nQ))
        return fail(ctx, MKHE_ERR_INVALID, "level %d: a transform over R = Q u QMul needs BFV parameters and polys of %d limbs", level, 2 * ctx->nQ);
    Slots s = (level >= ctx->nQ) ? r_slots(ctx) : q_slots(level);
    return ntt_inv(ctx, s, 1, &i->d, &o->d);
}

// ---- KeySwitcher ------------------------------------------------------------------------------------
static int check_level(mkhe_ctx *ctx, int level) {
    if (level < 0 || level >= ctx->nQ) return fail(ctx, MKHE_ERR_INVALID, "level %d out of range [0,%d)", level, ctx->nQ);
    return MKHE_OK;
}

int mkhe_decompose(mkhe_ctx *ctx, int levelQ, mkhe_poly a, mkhe_swk ad) {
    CHECK_CTX();
    TRY(check_level(ctx, levelQ));
    POLY_R(p, a);
    SWK(k, ad);
    if (p->cap_limbs < levelQ + 1) return fail(ctx, MKHE_ERR_INVALID, "Decompose: poly has %d limbs, level %d", p->cap_limbs, levelQ);
    GraphKey key("decompose");
    key.add((u64)levelQ); key.add((u64)(uintptr_t)p->d); key.add((u64)(uintptr_t)k->d);
    return run_graphed(ctx, key, [&]() -> int { return decompose_impl(ctx, levelQ, 1, &p->d, &k->d, 0); });
}

int mkhe_external_product_hoisted(mkhe_ctx *ctx, int levelQ, mkhe_swk a_hoisted, mkhe_swk bg, mkhe_poly c) {
    CHECK_CTX();
    TRY(check_level(ctx, levelQ));
    SWK_R(h, a_hoisted);
    SWK_R(k, bg);
    POLY(o, c);
    if (o->cap_limbs < levelQ + 1) return fail(ctx, MKHE_ERR_INVALID, "ExternalProduct: output has %d limbs", o->cap_limbs);
    return ext_products(ctx, levelQ, 1, 1, &k->d, &h->d, nullptr, nullptr, &o->d, nullptr, false);
}

/* FastBasisExtender.ModDownQPtoQNTT (mkrlwe/basis_extension.go:239-290): p1 = one rlwe.PolyQP in the NTT domain (Q limbs, then P
 * limbs), p2Q = round(p1 / P) in the NTT domain.  The reference brings only the P part out of the NTT domain, lifts it to Q,
 * transforms the lift and combines in the NTT domain; here the whole QP value goes through the key switch's own tail (inverse pass
 * A, k_moddown_P / k_moddown_Q) and the quotient is transformed forward: by linearity the same canonical residues. */
int mkhe_moddown_qp_to_q_ntt(mkhe_ctx *ctx, int levelQ, mkhe_poly p1, mkhe_poly p2Q) {
    CHECK_CTX();
    TRY(check_level(ctx, levelQ));
    POLY_R(in, p1);
    POLY(out, p2Q);
    if (in->cap_limbs < ctx->dmax) return fail(ctx, MKHE_ERR_INVALID, "ModDownQPtoQNTT needs a poly of nQ + nP = %d limbs (rlwe.PolyQP)", ctx->dmax);
    if (out->cap_limbs < levelQ + 1) return fail(ctx, MKHE_ERR_INVALID, "ModDownQPtoQNTT: output has %d limbs", out->cap_limbs);
    const Slots s = qp_slots(ctx, levelQ);
    u64 *acc;
    TRY(get_scratch(ctx, "accqp", (size_t)(ctx->dmax + ctx->nP) * ctx->N * 8, &acc));
    {   // inverse pass A of the Q limbs up to the level and of the P limbs, into the accumulator
        InvAArgs a;
        memset(&a, 0, sizeof a);
        a.nslots = s.n; a.logN = ctx->logN; a.nbatch = 1;
        for (int i = 0; i < s.n; i++) { a.slots[i] = s.slot[i]; a.mods[i] = s.mod[i]; }
        a.in.p[0] = in->d; a.out.p[0] = acc;
        LAUNCH(k_intt_passA, dim3(ctx->N / MKHE_TILE, s.n, 1), dim3(MKHE_PA_THREADS), MKHE_PA_SMEM, a, ctx->d_mods, ctx->d_twi_tiled);
    }
    ModDownPArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.np_limbs = ctx->nP; pa.p_slot0 = ctx->nQ; pa.vslot = ctx->dmax; pa.logN = ctx->logN;
    for (int i = 0; i < ctx->nP; i++) pa.plist[pa.nplist++] = i;
    pa.acc[0] = acc;
    TRY(dispatch_s1(ctx, [&](auto S) -> int {
        auto k_moddown_P_ = k_moddown_P<decltype(S)::value>;
        LAUNCH(k_moddown_P_, dim3(COLGROUPS, 1, pa.nplist), dim3(MKHE_NTT_THREADS), 0, pa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi);
        return MKHE_OK;
    }));
    ModDownQArgs qa;
    memset(&qa, 0, sizeof qa);
    qa.np_limbs = ctx->nP; qa.logN = ctx->logN;
    for (int j = 0; j <= levelQ; j++) qa.qlist[qa.nqlist++] = j;
    qa.split[0] = 1; qa.dst[0] = out->d; qa.src[0] = nullptr; qa.dst_team_off[0] = -1;
    qa.first[0] = 0; qa.first[1] = 1;
    qa.acc[0] = acc; qa.pp[0] = acc + (size_t)ctx->nQ * ctx->N;
    TRY(dispatch_s1(ctx, [&](auto S) -> int {
        constexpr int s1 = decltype(S)::value;
        const dim3 grid(COLGROUPS, qa.nqlist, 1);
        const size_t rowsm = MKHE_MDQ_SMEM(s1);
        switch (ctx->nP) {
            case 1: { auto k = k_moddown_Q<s1, 1, false>; LAUNCH(k, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
            case 2: { auto k = k_moddown_Q<s1, 2, false>; LAUNCH(k, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
            case 3: { auto k = k_moddown_Q<s1, 3, false>; LAUNCH(k, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
            case 4: { auto k = k_moddown_Q<s1, 4, false>; LAUNCH(k, grid, dim3(MKHE_NTT_THREADS), rowsm, qa, ctx->d_conv_PtoQ, ctx->d_mods, ctx->d_twi); break; }
            default: return fail(ctx, MKHE_ERR_UNSUPPORTED, "%d special primes (1..4 supported)", ctx->nP);
        }
        return MKHE_OK;
    }));
    out->nlimbs = levelQ + 1;
    return ntt_fwd(ctx, q_slots(levelQ), 1, &out->d, &out->d);
}

int mkhe_external_product(mkhe_ctx *ctx, int levelQ, mkhe_poly a, mkhe_swk bg, mkhe_poly c) {
    CHECK_CTX();
    TRY(check_level(ctx, levelQ));
    POLY_R(p, a);
    SWK_R(k, bg);
    POLY(o, c);
    if (p->cap_limbs < levelQ + 1 || o->cap_limbs < levelQ + 1) return fail(ctx, MKHE_ERR_INVALID, "ExternalProduct: too few limbs");
    std::vector<u64 *> hp;
    TRY(swk_pool(ctx, "extprod_h", 1, hp));
    TRY(decompose_impl(ctx, levelQ, 1, &p->d, hp.data(), 0));
    return ext_products(ctx, levelQ, 1, 1, &k->d, hp.data(), nullptr, nullptr, &o->d, nullptr, false);
}

static int check_ids(mkhe_ctx *ctx, int n0, const int *ids0, int n1, const int *ids1, int nOut, const int *idsOut) {
    if (n0 < 0 || n1 < 0 || nOut < 0 || n0 > MKHE_MAX_PARTIES || n1 > MKHE_MAX_PARTIES || nOut > MKHE_MAX_PARTIES)
        return fail(ctx, MKHE_ERR_INVALID, "party count out of range (max %d)", MKHE_MAX_PARTIES);
    for (int t = 0; t < n0; t++) if (find_id(nOut, idsOut, ids0[t]) < 0) return fail(ctx, MKHE_ERR_INVALID, "idsOut is not the union of ids0 and ids1");
    for (int t = 0; t < n1; t++) if (find_id(nOut, idsOut, ids1[t]) < 0) return fail(ctx, MKHE_ERR_INVALID, "idsOut is not the union of ids0 and ids1");
    for (int t = 0; t < nOut; t++) if (find_id(n0, ids0, idsOut[t]) < 0 && find_id(n1, ids1, idsOut[t]) < 0) return fail(ctx, MKHE_ERR_INVALID, "idsOut is not the union of ids0 and ids1");
    return MKHE_OK;
}

int mkhe_mul_relin_hoisted(mkhe_ctx *ctx, int level, int n0, const int *ids0, const mkhe_poly *op0, const mkhe_swk *h0,
                           int n1, const int *ids1, const mkhe_poly *op1, const mkhe_swk *h1, const mkhe_swk *rlk_b,
                           const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u, int nOut, const int *idsOut,
                           const mkhe_poly *out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    std::vector<u64 *> p0, p1, po, vh0, vh1, vb, vd, vv;
    // level checks mirror keyswitch_hoisted.go:48-54
    for (int t = 0; t <= n0; t++) {
        Obj *o = as_obj(ctx, op0[t], OBJ_POLY, ACC_READ);
        if (o && o->nlimbs - 1 < level) return fail(ctx, MKHE_ERR_INVALID, "Cannot MulAndRelin: op0 and op1 have different levels");
    }
    TRY(polys_of(ctx, n0 + 1, op0, level + 1, p0, "op0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, op1, level + 1, p1, "op1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, level + 1, po, "out"));
    TRY(swks_of(ctx, n1, rlk_b, vb, "rlk_b", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_d, vd, "rlk_d", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_v, vv, "rlk_v", ACC_READ));
    SWK_R(uk, u);
    if (h0) TRY(swks_of(ctx, n0, h0, vh0, "h0", ACC_READ));
    if (h1) TRY(swks_of(ctx, n1, h1, vh1, "h1", ACC_READ));
    GraphKey key("mul_relin_hoisted");
    key.add((u64)level); key.add((u64)(h0 != nullptr)); key.add((u64)(h1 != nullptr));
    key.ints(n0, ids0); key.ints(n1, ids1); key.ints(nOut, idsOut);
    key.ptrs(p0); key.ptrs(p1); key.ptrs(po); key.ptrs(vh0); key.ptrs(vh1); key.ptrs(vb); key.ptrs(vd); key.ptrs(vv);
    key.add((u64)(uintptr_t)uk->d);
    return run_graphed(ctx, key, [&]() -> int {
        if (!h0) {
            TRY(swk_pool(ctx, "nil_h0", n0, vh0));
            TRY(decompose_impl(ctx, level, n0, p0.data() + 1, vh0.data(), 0));
        }
        if (!h1) {
            TRY(swk_pool(ctx, "nil_h1", n1, vh1));
            TRY(decompose_impl(ctx, level, n1, p1.data() + 1, vh1.data(), 0));
        }
        return mul_relin_hoisted_impl(ctx, level, n0, ids0, p0.data(), vh0.data(), n1, ids1, p1.data(), vh1.data(), vb.data(),
                                      vd.data(), vv.data(), uk->d, nOut, idsOut, po.data());
    });
}

int mkhe_ckks_mul_relin(mkhe_ctx *ctx, int level, int nb_rescales, int same_operand, int n0, const int *ids0,
                        const mkhe_poly *op0, int n1, const int *ids1, const mkhe_poly *op1, const mkhe_swk *rlk_b,
                        const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u, int nOut, const int *idsOut,
                        const mkhe_poly *out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    std::vector<u64 *> p0, p1, po, vh0, vh1, vb, vd, vv;
    TRY(polys_of(ctx, n0 + 1, op0, level + 1, p0, "op0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, op1, level + 1, p1, "op1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, level + 1, po, "out"));
    TRY(swks_of(ctx, n1, rlk_b, vb, "rlk_b", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_d, vd, "rlk_d", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_v, vv, "rlk_v", ACC_READ));
    SWK_R(uk, u);
    GraphKey key("ckks_mul_relin");
    key.add((u64)level); key.add((u64)nb_rescales); key.add((u64)same_operand);
    key.ints(n0, ids0); key.ints(n1, ids1); key.ints(nOut, idsOut);
    key.ptrs(p0); key.ptrs(p1); key.ptrs(po); key.ptrs(vb); key.ptrs(vd); key.ptrs(vv); key.add((u64)(uintptr_t)uk->d);
    TRY(run_graphed(ctx, key, [&]() -> int {
        // hoisting into context pools (rlkSet.HoistPool[0|1], mkckks/evaluator.go:419-441), scheduled inside the op
        TRY(swk_pool(ctx, "hoistpool0", n0, vh0));
        if (same_operand) vh1 = vh0;
        else TRY(swk_pool(ctx, "hoistpool1", n1, vh1));
        TRY(mul_relin_hoisted_impl(ctx, level, n0, ids0, p0.data(), vh0.data(), n1, ids1, p1.data(), vh1.data(), vb.data(),
                                   vd.data(), vv.data(), uk->d, nOut, idsOut, po.data(), Shard(), true, !same_operand, true, true));
        return rescale_impl(ctx, level, nb_rescales, nOut + 1, po.data(), po.data());
    }));
    for (int t = 0; t <= nOut; t++) reinterpret_cast<Obj *>(out[t])->nlimbs = level + 1 - nb_rescales;
    return MKHE_OK;
}

int mkhe_ckks_mul_relin_sharded(mkhe_ctx *ctx, int level, int nb_rescales, int n0, const int *ids0, const mkhe_poly *op0,
                                int n1, const int *ids1, const mkhe_poly *op1, int nown, const int *own_ids,
                                const mkhe_swk *rlk_b, const mkhe_swk *rlk_d, const mkhe_swk *rlk_v, mkhe_swk u, int nOut,
                                const int *idsOut, const mkhe_poly *out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    if (!ctx->nccl) return fail(ctx, MKHE_ERR_NCCL, "mkhe_comm_init has not been called");
    Shard sh;
    sh.active = true;
    sh.own.assign(own_ids, own_ids + nown);
    std::vector<u64 *> p0, p1, po, vh0(n0, nullptr), vh1(n1, nullptr), vb(n1, nullptr), vd(n0, nullptr), vv(n0, nullptr);
    TRY(polys_of(ctx, n0 + 1, op0, level + 1, p0, "op0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, op1, level + 1, p1, "op1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, level + 1, po, "out"));
    SWK_R(uk, u);
    // keys and hoisting only for the parties this rank owns
    std::vector<u64 *> in0, in1, hp0, hp1, pool0, pool1;
    std::vector<int> t0, t1;
    for (int t = 0; t < n0; t++) if (sh.owns(ids0[t])) t0.push_back(t);
    for (int t = 0; t < n1; t++) if (sh.owns(ids1[t])) t1.push_back(t);
    TRY(swk_pool(ctx, "hoistpool0", (int)t0.size(), pool0));
    TRY(swk_pool(ctx, "hoistpool1", (int)t1.size(), pool1));
    for (size_t i = 0; i < t0.size(); i++) {
        int t = t0[i];
        Obj *d = as_obj(ctx, rlk_d[t], OBJ_SWK, ACC_READ), *v = as_obj(ctx, rlk_v[t], OBJ_SWK, ACC_READ);
        if (!d || !v) return fail(ctx, MKHE_ERR_INVALID, "missing relinearization key (d, v) of owned party %d", ids0[t]);
        vd[t] = d->d; vv[t] = v->d; vh0[t] = pool0[i];
        in0.push_back(p0[1 + t]);
    }
    for (size_t i = 0; i < t1.size(); i++) {
        int t = t1[i];
        Obj *b = as_obj(ctx, rlk_b[t], OBJ_SWK, ACC_READ);
        if (!b) return fail(ctx, MKHE_ERR_INVALID, "missing relinearization key (b) of owned party %d", ids1[t]);
        vb[t] = b->d; vh1[t] = pool1[i];
        in1.push_back(p1[1 + t]);
    }
    const bool lazy = ctx->alpha == 1 && beta_of(ctx, level) <= 14;       // forms that never leave the op (see mul_relin_hoisted_impl)
    TRY(decompose_impl(ctx, level, (int)in0.size(), in0.data(), pool0.data(), 0, lazy));
    TRY(decompose_impl(ctx, level, (int)in1.size(), in1.data(), pool1.data(), 0, lazy));
    // the forms of the owned parties are fresh (made above): the tensor step reads NTT(op_id) from their diagonals; the other
    // parties' components are not completed on this rank, so nothing else has to be transformed
    TRY(mul_relin_hoisted_impl(ctx, level, n0, ids0, p0.data(), vh0.data(), n1, ids1, p1.data(), vh1.data(), vb.data(),
                               vd.data(), vv.data(), uk->d, nOut, idsOut, po.data(), sh, false, false, true, true));
    TRY(rescale_impl(ctx, level, nb_rescales, nOut + 1, po.data(), po.data()));
    for (int t = 0; t <= nOut; t++) reinterpret_cast<Obj *>(out[t])->nlimbs = level + 1 - nb_rescales;
    return MKHE_OK;
}

int mkhe_rotate_hoisted(mkhe_ctx *ctx, int level, int rotidx, int n, const mkhe_poly *ct_in, const mkhe_swk *hoisted,
                        const mkhe_swk *rk, mkhe_swk a, const mkhe_poly *ct_out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 0 || n > MKHE_MAX_PARTIES) return fail(ctx, MKHE_ERR_INVALID, "party count out of range");
    for (int t = 0; t <= n; t++) {
        Obj *o = as_obj(ctx, ct_in[t], OBJ_POLY, ACC_READ);
        if (o && o->nlimbs - 1 < level) return fail(ctx, MKHE_ERR_INVALID, "Cannot Rotate: ctIn and ctOut have different levels");
    }
    std::vector<u64 *> pi, po, vh, vrk;
    TRY(polys_of(ctx, n + 1, ct_in, level + 1, pi, "ct_in", ACC_READ));
    TRY(polys_of(ctx, n + 1, ct_out, level + 1, po, "ct_out"));
    TRY(swks_of(ctx, n, hoisted, vh, "hoisted", ACC_READ));
    TRY(swks_of(ctx, n, rk, vrk, "rk", ACC_READ));
    SWK_R(ak, a);
    GraphKey key("rotate_hoisted");
    key.add((u64)level); key.add((u64)(long)rotidx);
    key.ptrs(pi); key.ptrs(po); key.ptrs(vh); key.ptrs(vrk); key.add((u64)(uintptr_t)ak->d);
    return run_graphed(ctx, key, [&]() -> int { return rotate_hoisted_impl(ctx, level, rotidx, n, pi.data(), vh.data(), vrk.data(), ak->d, po.data()); });
}

int mkhe_rotate(mkhe_ctx *ctx, int level, int rotidx, int n, const mkhe_poly *ct_in, const mkhe_swk *rk, mkhe_swk a,
                const mkhe_poly *ct_out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 0 || n > MKHE_MAX_PARTIES) return fail(ctx, MKHE_ERR_INVALID, "party count out of range");
    std::vector<u64 *> pi, po, vh, vrk;
    TRY(polys_of(ctx, n + 1, ct_in, level + 1, pi, "ct_in", ACC_READ));
    TRY(polys_of(ctx, n + 1, ct_out, level + 1, po, "ct_out"));
    TRY(swks_of(ctx, n, rk, vrk, "rk", ACC_READ));
    SWK_R(ak, a);
    GraphKey key("rotate");
    key.add((u64)level); key.add((u64)(long)rotidx);
    key.ptrs(pi); key.ptrs(po); key.ptrs(vrk); key.add((u64)(uintptr_t)ak->d);
    return run_graphed(ctx, key, [&]() -> int {
        TRY(swk_pool(ctx, "rot_h", n, vh));
        TRY(decompose_impl(ctx, level, n, pi.data() + 1, vh.data(), 0));
        return rotate_hoisted_impl(ctx, level, rotidx, n, pi.data(), vh.data(), vrk.data(), ak->d, po.data());
    });
}

int mkhe_conjugate(mkhe_ctx *ctx, int level, int n, const mkhe_poly *ct_in, const mkhe_swk *ck, mkhe_swk a,
                   const mkhe_poly *ct_out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 0 || n > MKHE_MAX_PARTIES) return fail(ctx, MKHE_ERR_INVALID, "party count out of range");
    std::vector<u64 *> pi, po, vh, vck, tmp;
    TRY(polys_of(ctx, n + 1, ct_in, level + 1, pi, "ct_in", ACC_READ));
    TRY(polys_of(ctx, n + 1, ct_out, level + 1, po, "ct_out"));
    TRY(swks_of(ctx, n, ck, vck, "ck", ACC_READ));
    SWK_R(ak, a);
    GraphKey key("conjugate");
    key.add((u64)level); key.ptrs(pi); key.ptrs(po); key.ptrs(vck); key.add((u64)(uintptr_t)ak->d);
    return run_graphed(ctx, key, [&]() -> int {
        // permute first (keyswitch.go:315-317), then key-switch the permuted components (:320-331)
        TRY(poly_pool(ctx, "conj_tmp", n + 1, ctx->nQ, tmp));
        TRY(automorph_impl(ctx, level, ((u64)2 << ctx->logN) - 1, n + 1, pi.data(), tmp.data()));
        TRY(swk_pool(ctx, "rot_h", n, vh));
        TRY(decompose_impl(ctx, level, n, tmp.data() + 1, vh.data(), 0));
        CU(cudaMemcpyAsync(po[0], tmp[0], (size_t)(level + 1) * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        std::vector<Prod> pr;
        for (int t = 0; t < n; t++) {
            pr.push_back(Prod{{vck[t], nullptr}, {vh[t], nullptr}, po[0], true});
            pr.push_back(Prod{{ak->d, nullptr}, {vh[t], nullptr}, po[1 + t], false});
        }
        return ext_products(ctx, level, 1, pr);
    });
}

// ---- mkckks.Evaluator -------------------------------------------------------------------------------
int mkhe_rescale(mkhe_ctx *ctx, int level, int nb_rescales, mkhe_poly in, mkhe_poly out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    POLY(i, in);
    POLY(o, out);
    if (i->cap_limbs < level + 1 || o->cap_limbs < level + 1 - nb_rescales) return fail(ctx, MKHE_ERR_INVALID, "Rescale: too few limbs");
    if (nb_rescales == 0) {
        if (i != o) CU(cudaMemcpyAsync(o->d, i->d, (size_t)(level + 1) * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        return MKHE_OK;
    }
    if (nb_rescales > 1 && i != o) {
        // lattigo runs the later divisions in a pool, leaving `in` untouched after the first one
        std::vector<u64 *> tmp;
        TRY(poly_pool(ctx, "rescale_tmp", 1, ctx->nQ, tmp));
        TRY(rescale_impl(ctx, level, 1, 1, &i->d, tmp.data()));
        TRY(rescale_impl(ctx, level - 1, nb_rescales - 1, 1, tmp.data(), tmp.data()));
        CU(cudaMemcpyAsync(o->d, tmp[0], (size_t)(level + 1 - nb_rescales) * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        TRY(rescale_impl(ctx, level, nb_rescales, 1, &i->d, &o->d));
    }
    o->nlimbs = level + 1 - nb_rescales;
    return MKHE_OK;
}
int mkhe_poly_add(mkhe_ctx *ctx, int level, mkhe_poly a, mkhe_poly b, mkhe_poly out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    POLY_R(x, a); POLY_R(y, b); POLY(o, out);
    if (x->cap_limbs <= level || y->cap_limbs <= level || o->cap_limbs <= level) return fail(ctx, MKHE_ERR_INVALID, "Add: too few limbs");
    return add_polys(ctx, level, x->d, y->d, o->d, false);
}
int mkhe_poly_sub(mkhe_ctx *ctx, int level, mkhe_poly a, mkhe_poly b, mkhe_poly out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    POLY_R(x, a); POLY_R(y, b); POLY(o, out);
    if (x->cap_limbs <= level || y->cap_limbs <= level || o->cap_limbs <= level) return fail(ctx, MKHE_ERR_INVALID, "Sub: too few limbs");
    return add_polys(ctx, level, x->d, y->d, o->d, true);
}

/* scaleUpExact (mkckks/utils.go:59-86): big.NewFloat(n*value) carries 53 bits, so the product and the + 0.5 are float64
 * operations; Int() truncates; mod q; a negative value gives q - res (q, not 0, for a multiple of q) */
static u64 scale_up_exact(double value, double n, u64 q) {
    const bool neg = value < 0;
    volatile double x = neg ? -n * value : n * value;
    x = x + 0.5;
    u64 res;
    if (x < 18446744073709551616.0) res = (u64)x % q;
    else res = (u64)(((unsigned __int128)x) % q);
    return neg ? q - res : res;
}
int mkhe_ckks_mult_by_const(mkhe_ctx *ctx, int level, int n, const mkhe_poly *in, const mkhe_poly *out, double c_real, double c_imag,
                            double scale) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 1 || n > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_INVALID, "MultByConst: bad component count %d", n);
    std::vector<u64 *> pi, po;
    TRY(polys_of(ctx, n, in, level + 1, pi, "in", ACC_READ));
    TRY(polys_of(ctx, n, out, level + 1, po, "out"));
    ConstArgs a;
    memset(&a, 0, sizeof a);
    a.nlimbs = level + 1;
    a.logN = ctx->logN;
    for (int i = 0; i <= level; i++) {
        const u64 qi = ctx->mod[i];
        a.mod_of_limb[i] = i;
        u64 sReal = 0, sImag = 0, sc = 0;
        if (c_real != 0) { sReal = scale_up_exact(c_real, scale, qi); sc = sReal; }
        if (c_imag != 0) {
            sImag = scale_up_exact(c_imag, scale, qi);
            sImag = h_mulmod(sImag % qi, ctx->tabs[i].twf[1].x, qi);          // MRed(., NttPsi[i][1])
            sc = sc + sImag;
            if (sc >= qi) sc -= qi;                                            // CRed
        }
        a.c_first[i] = h_mform(sc % qi, qi);
        if (c_imag != 0) {
            sc = sReal + (qi - sImag);
            if (sc >= qi) sc -= qi;
        }
        a.c_second[i] = h_mform(sc % qi, qi);
    }
    for (int i = 0; i < n; i++) { a.in.p[i] = pi[i]; a.out.p[i] = po[i]; }
    LAUNCH(k_mul_const, dim3(ctx->N / MKHE_THREADS, level + 1, n), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
}
int mkhe_poly_neg(mkhe_ctx *ctx, int level, mkhe_poly in, mkhe_poly out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    POLY_R(i, in); POLY(o, out);
    if (i->cap_limbs < level + 1 || o->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "neg: too few limbs");
    LimbArgs a;
    fill(a, q_slots(level), ctx->logN);
    a.in.p[0] = i->d; a.out.p[0] = o->d;
    LAUNCH(k_neg, dim3(ctx->N / MKHE_THREADS, level + 1, 1), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
}
int mkhe_ckks_mul_ptxt(mkhe_ctx *ctx, int level, mkhe_poly pt, int n, const mkhe_poly *in, const mkhe_poly *out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 1 || n + 1 > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_INVALID, "MulPtxt: bad component count %d", n);
    POLY_R(p, pt);
    if (p->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "MulPtxt: plaintext has too few limbs");
    std::vector<u64 *> pi, po, tn;
    TRY(polys_of(ctx, n, in, level + 1, pi, "in", ACC_READ));
    TRY(polys_of(ctx, n, out, level + 1, po, "out"));
    TRY(poly_pool(ctx, "mulptxt_ntt", n + 1, ctx->nQ, tn));
    const Slots qs = q_slots(level);
    std::vector<u64 *> src(pi);
    src.push_back(p->d);
    TRY(ntt_fwd(ctx, qs, n + 1, src.data(), tn.data()));
    for (int t = 0; t < n; t++) TRY(mul2(ctx, qs, tn[n], tn[t], nullptr, nullptr, tn[t]));   // MForm(pt) (.) ct, Montgomery
    return ntt_inv(ctx, qs, n, tn.data(), po.data());
}

/* Decryptor.Decrypt (mkrlwe/decryptor.go:48-66) = PartialDecrypt (:26-43) for every party, then ReduceLvl:
 * pt = Reduce(ct["0"] + sum_t InvNTT(NTT(ct[id_t]) (.) sk_t)).  The accumulation order of the reference follows a Go map; every
 * partial sum is the same residue, and the output is canonical, so the order is immaterial. */
int mkhe_decrypt(mkhe_ctx *ctx, int level, int n, const mkhe_poly *ct, const mkhe_poly *sk, mkhe_poly pt) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    if (n < 0 || n + 1 > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_INVALID, "Decrypt: bad party count %d", n);
    std::vector<u64 *> pc, ps, tn;
    TRY(polys_of(ctx, n + 1, ct, level + 1, pc, "ct", ACC_READ));
    TRY(polys_of(ctx, n, sk, level + 1, ps, "sk", ACC_READ));
    POLY(o, pt);
    if (o->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "Decrypt: plaintext has too few limbs");
    TRY(poly_pool(ctx, "decrypt_ntt", std::max(n, 1), ctx->nQ, tn));
    const Slots qs = q_slots(level);
    LimbArgs a;
    fill(a, qs, ctx->logN);
    if (n > 0) {
        TRY(ntt_fwd(ctx, qs, n, pc.data() + 1, tn.data()));
        PtrList other;
        memset(&other, 0, sizeof other);
        for (int t = 0; t < n; t++) { a.in.p[t] = tn[t]; a.out.p[t] = tn[t]; other.p[t] = ps[t]; }
        LAUNCH(k_mul_mont, dim3(ctx->N / MKHE_THREADS, level + 1, n), dim3(MKHE_THREADS), 0, a, other, ctx->d_mods);
        TRY(ntt_inv(ctx, qs, n, tn.data(), tn.data()));
    }
    for (int t = 0; t < n; t++) a.in.p[t] = tn[t];
    LAUNCH(k_decrypt_sum, dim3(ctx->N / MKHE_THREADS, level + 1), dim3(MKHE_THREADS), 0, a, (const u64 *)pc[0], n, o->d, ctx->d_mods);
    o->nlimbs = level + 1;
    return MKHE_OK;
}

// ---- mkbfv ------------------------------------------------------------------------------------------
namespace {
int need_bfv(mkhe_ctx *ctx) {
    if (!ctx->nQMul) return fail(ctx, MKHE_ERR_INVALID, "BFV parameters not set (mkhe_ctx_set_bfv)");
    return MKHE_OK;
}
int conv_launch(mkhe_ctx *ctx, int mode, const ConvTable *tab, int nb, u64 *const *src, int src_limb0, u64 *const *x,
                int x_limb0, int x_is_zero, u64 *const *dst, int dst_limb0, int n2) {
    const int cb = (ctx->N + MKHE_THREADS - 1) / MKHE_THREADS;
    for (int b0 = 0; b0 < nb; b0 += MKHE_MAX_PARTIES_K) {
        int n = std::min(MKHE_MAX_PARTIES_K, nb - b0);
        ConvArgs c;
        memset(&c, 0, sizeof c);
        c.src_limb0 = src_limb0; c.x_limb0 = x_limb0; c.dst_limb0 = dst_limb0;
        c.n2_used = n2; c.x_is_zero = x_is_zero; c.logN = ctx->logN;
        for (int i = 0; i < n; i++) { c.src.p[i] = src[b0 + i]; c.x.p[i] = x ? x[b0 + i] : nullptr; c.dst.p[i] = dst[b0 + i]; }
        if (mode == CONV_MODUP) LAUNCH(k_conv<CONV_MODUP>, dim3(cb, n), dim3(MKHE_THREADS), 0, c, tab, ctx->d_mods);
        else LAUNCH(k_conv<CONV_MODDOWN>, dim3(cb, n), dim3(MKHE_THREADS), 0, c, tab, ctx->d_mods);
    }
    return MKHE_OK;
}
// ModUpQtoR for a batch (mkbfv/basis_extension.go:49-63)
int bfv_modup_impl(mkhe_ctx *ctx, int nb, u64 *const *polyQ, u64 *const *polyR) {
    const size_t qbytes = (size_t)ctx->nQ * ctx->N * 8;
    for (int i = 0; i < nb; i++) CU(cudaMemcpyAsync(polyR[i], polyQ[i], qbytes, cudaMemcpyDeviceToDevice, ctx->stream));
    return conv_launch(ctx, CONV_MODUP, ctx->d_conv_QtoQMul, nb, polyQ, 0, nullptr, 0, 0, polyR, ctx->nQ, ctx->nQ);
}
// Rescale Q -> R for a batch (mkbfv/basis_extension.go:83-97)
int bfv_rescale_impl(mkhe_ctx *ctx, int nb, u64 *const *polyQ, u64 *const *polyR) {
    std::vector<u64 *> tq;
    TRY(poly_pool(ctx, "bfv_tq", nb, ctx->nQ, tq));
    for (int b0 = 0; b0 < nb; b0 += MKHE_MAX_PARTIES_K) {
        int n = std::min(MKHE_MAX_PARTIES_K, nb - b0);
        ScaleArgs a;
        memset(&a, 0, sizeof a);
        a.nlimbs = ctx->nQ; a.logN = ctx->logN;
        for (int i = 0; i < ctx->nQ; i++) { a.mod_of_limb[i] = i; a.cmont[i] = ctx->h_mformQMul[i]; }
        for (int i = 0; i < n; i++) { a.in.p[i] = polyQ[b0 + i]; a.out.p[i] = tq[b0 + i]; }
        LAUNCH(k_scale, dim3(ctx->N / MKHE_THREADS, ctx->nQ, n), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    }
    // ModDownQPtoP with a zero P part -> canonical QMul limbs, written straight into R's upper half
    TRY(conv_launch(ctx, CONV_MODDOWN, ctx->d_conv_QtoQMul, nb, tq.data(), 0, nullptr, 0, 1, polyR, ctx->nQ, ctx->nQ));
    // ModUpPtoQ: lazy lift of the QMul limbs back to Q -> R's lower half
    return conv_launch(ctx, CONV_MODUP, ctx->d_conv_QMultoQ, nb, polyR, ctx->nQ, nullptr, 0, 0, polyR, 0, ctx->nQ);
}
// Quantize for a batch: polyR (NTT domain, canonical) -> polyQ   (mkbfv/basis_extension.go:66-80)
int bfv_quantize_impl(mkhe_ctx *ctx, int nb, u64 *const *polyR, u64 *const *polyQ) {
    std::vector<u64 *> tr;
    TRY(poly_pool(ctx, "bfv_tr", nb, 2 * ctx->nQ, tr));
    Slots rs = r_slots(ctx);
    for (int b0 = 0; b0 < nb; b0 += MKHE_MAX_PARTIES_K) {
        int n = std::min(MKHE_MAX_PARTIES_K, nb - b0);
        ScaleArgs a;
        memset(&a, 0, sizeof a);
        a.nlimbs = rs.n; a.logN = ctx->logN;
        for (int i = 0; i < rs.n; i++) { a.mod_of_limb[i] = rs.mod[i]; a.cmont[i] = h_mform(ctx->T % ctx->mod[rs.mod[i]], ctx->mod[rs.mod[i]]); }
        for (int i = 0; i < n; i++) { a.in.p[i] = polyR[b0 + i]; a.out.p[i] = tr[b0 + i]; }
        LAUNCH(k_scale, dim3(ctx->N / MKHE_THREADS, rs.n, n), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    }
    TRY(ntt_inv(ctx, rs, nb, tr.data(), tr.data()));
    return conv_launch(ctx, CONV_MODDOWN, ctx->d_conv_QMultoQ, nb, tr.data(), ctx->nQ, tr.data(), 0, 0, polyQ, 0, ctx->nQ);
}
int bfv_decompose_impl(mkhe_ctx *ctx, int levelQ, int nb, u64 *const *aR, u64 *const *ad1, u64 *const *ad2) {
    TRY(decompose_impl(ctx, levelQ, nb, aR, ad1, 0));
    return decompose_impl(ctx, levelQ, nb, aR, ad2, levelQ + 1);
}
// MulAndRelinBFVHoisted on raw pointers (mkbfv/keyswitch_hoisted.go:39-207)
int bfv_mul_relin_hoisted_impl(mkhe_ctx *ctx, int level, int n0, const int *ids0, u64 *const *op0, u64 *const *h0a,
                               u64 *const *h0b, int n1, const int *ids1, u64 *const *op1, u64 *const *h1a,
                               u64 *const *h1b, u64 *const *b1, u64 *const *b2, u64 *const *d1, u64 *const *d2,
                               u64 *const *v, u64 *u, int nOut, const int *idsOut, u64 *const *out) {
    std::vector<u64 *> xy;
    TRY(swk_pool(ctx, "bfv_xy", 4, xy));
    u64 *x1 = xy[0], *x2 = xy[1], *y1 = xy[2], *y2 = xy[3];
    TRY(mac_parties(ctx, level, n0, d1, h0a, x1));
    TRY(mac_parties(ctx, level, n0, d2, h0b, x2));
    TRY(mac_parties(ctx, level, n1, b1, h1a, y1));
    TRY(mac_parties(ctx, level, n1, b2, h1b, y2));
    // tensor product in ring R, each component quantized back to Q (:144-181)
    Slots rs = r_slots(ctx);
    std::vector<u64 *> tn, prod, src(n0 + n1 + 2);
    TRY(poly_pool(ctx, "bfv_tensor_ntt", n0 + n1 + 2, 2 * ctx->nQ, tn));
    TRY(poly_pool(ctx, "bfv_tensor_prod", nOut + 1, 2 * ctx->nQ, prod));
    for (int t = 0; t <= n0; t++) src[t] = op0[t];
    for (int t = 0; t <= n1; t++) src[n0 + 1 + t] = op1[t];
    TRY(ntt_fwd(ctx, rs, n0 + n1 + 2, src.data(), tn.data()));
    const u64 *A0 = tn[0], *B0 = tn[n0 + 1];
    TRY(mul2(ctx, rs, A0, B0, nullptr, nullptr, prod[0]));
    for (int t = 0; t < nOut; t++) {
        int i0 = find_id(n0, ids0, idsOut[t]), i1 = find_id(n1, ids1, idsOut[t]);
        if (i0 >= 0 && i1 < 0) TRY(mul2(ctx, rs, B0, tn[1 + i0], nullptr, nullptr, prod[1 + t]));
        else if (i0 < 0 && i1 >= 0) TRY(mul2(ctx, rs, A0, tn[n0 + 2 + i1], nullptr, nullptr, prod[1 + t]));
        else TRY(mul2(ctx, rs, A0, tn[n0 + 2 + i1], B0, tn[1 + i0], prod[1 + t]));
    }
    TRY(bfv_quantize_impl(ctx, nOut + 1, prod.data(), out));
    // c_id += (x1,x2) [.] (h1a,h1b)   (:184-191)   and   p_id = (y1,y2) [.] (h0a,h0b)   (:198-203), one batch
    std::vector<u64 *> p, hp;
    TRY(poly_pool(ctx, "relin_p", n0, ctx->nQ, p));
    TRY(swk_pool(ctx, "relin_hp", n0, hp));
    {
        std::vector<Prod> pr;
        for (int t = 0; t < n1; t++) pr.push_back(Prod{{x1, x2}, {h1a[t], h1b[t]}, out[1 + find_id(nOut, idsOut, ids1[t])], true});
        for (int t = 0; t < n0; t++) pr.push_back(Prod{{y1, y2}, {h0a[t], h0b[t]}, p[t], false});
        TRY(ext_products(ctx, level, 2, pr));
    }
    // Decompose(p_id) ; c_0 += v_id [.] p_id ; c_id += u [.] p_id   (:204-216)
    if (n0 > 0) {
        TRY(decompose_impl(ctx, level, n0, p.data(), hp.data(), 0));
        std::vector<Prod> pr;
        for (int t = 0; t < n0; t++) {
            pr.push_back(Prod{{u, nullptr}, {hp[t], nullptr}, out[1 + find_id(nOut, idsOut, ids0[t])], true});
            pr.push_back(Prod{{v[t], nullptr}, {hp[t], nullptr}, out[0], true});
        }
        TRY(ext_products(ctx, level, 1, pr));
    }
    return MKHE_OK;
}
}  // namespace

int mkhe_bfv_modup_q_to_r(mkhe_ctx *ctx, mkhe_poly polyQ, mkhe_poly polyR) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    POLY(q, polyQ); POLY(r, polyR);
    if (q->cap_limbs < ctx->nQ || r->cap_limbs < 2 * ctx->nQ) return fail(ctx, MKHE_ERR_INVALID, "ModUpQtoR: too few limbs");
    return bfv_modup_impl(ctx, 1, &q->d, &r->d);
}
int mkhe_bfv_rescale_q_to_r(mkhe_ctx *ctx, mkhe_poly polyQ, mkhe_poly polyR) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    POLY(q, polyQ); POLY(r, polyR);
    if (q->cap_limbs < ctx->nQ || r->cap_limbs < 2 * ctx->nQ) return fail(ctx, MKHE_ERR_INVALID, "Rescale: too few limbs");
    return bfv_rescale_impl(ctx, 1, &q->d, &r->d);
}
int mkhe_bfv_quantize(mkhe_ctx *ctx, mkhe_poly polyR, mkhe_poly polyQ) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    POLY(q, polyQ); POLY(r, polyR);
    if (q->cap_limbs < ctx->nQ || r->cap_limbs < 2 * ctx->nQ) return fail(ctx, MKHE_ERR_INVALID, "Quantize: too few limbs");
    return bfv_quantize_impl(ctx, 1, &r->d, &q->d);
}
int mkhe_bfv_decompose(mkhe_ctx *ctx, int levelQ, mkhe_poly aR, mkhe_swk ad1, mkhe_swk ad2) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    TRY(check_level(ctx, levelQ));
    POLY_R(r, aR); SWK(k1, ad1); SWK(k2, ad2);
    if (r->cap_limbs < 2 * (levelQ + 1)) return fail(ctx, MKHE_ERR_INVALID, "DecomposeBFV: too few limbs");
    return bfv_decompose_impl(ctx, levelQ, 1, &r->d, &k1->d, &k2->d);
}
int mkhe_bfv_mul_relin_hoisted(mkhe_ctx *ctx, int level, int n0, const int *ids0, const mkhe_poly *op0,
                               const mkhe_swk *h0a, const mkhe_swk *h0b, int n1, const int *ids1, const mkhe_poly *op1,
                               const mkhe_swk *h1a, const mkhe_swk *h1b, const mkhe_swk *b1, const mkhe_swk *b2,
                               const mkhe_swk *d1, const mkhe_swk *d2, const mkhe_swk *v, mkhe_swk u, int nOut,
                               const int *idsOut, const mkhe_poly *out) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    TRY(check_level(ctx, level));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    std::vector<u64 *> p0, p1, po, a0, b0, a1, bb1, vb1, vb2, vd1, vd2, vv;
    TRY(polys_of(ctx, n0 + 1, op0, 2 * ctx->nQ, p0, "op0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, op1, 2 * ctx->nQ, p1, "op1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, level + 1, po, "out"));
    TRY(swks_of(ctx, n1, b1, vb1, "b1", ACC_READ)); TRY(swks_of(ctx, n1, b2, vb2, "b2", ACC_READ));
    TRY(swks_of(ctx, n0, d1, vd1, "d1", ACC_READ)); TRY(swks_of(ctx, n0, d2, vd2, "d2", ACC_READ));
    TRY(swks_of(ctx, n0, v, vv, "v", ACC_READ));
    SWK_R(uk, u);
    if (h0a && h0b) { TRY(swks_of(ctx, n0, h0a, a0, "h0a", ACC_READ)); TRY(swks_of(ctx, n0, h0b, b0, "h0b", ACC_READ)); }
    else {
        TRY(swk_pool(ctx, "nil_h0", n0, a0)); TRY(swk_pool(ctx, "nil_h0b", n0, b0));
        TRY(bfv_decompose_impl(ctx, level, n0, p0.data() + 1, a0.data(), b0.data()));
    }
    if (h1a && h1b) { TRY(swks_of(ctx, n1, h1a, a1, "h1a", ACC_READ)); TRY(swks_of(ctx, n1, h1b, bb1, "h1b", ACC_READ)); }
    else {
        TRY(swk_pool(ctx, "nil_h1", n1, a1)); TRY(swk_pool(ctx, "nil_h1b", n1, bb1));
        TRY(bfv_decompose_impl(ctx, level, n1, p1.data() + 1, a1.data(), bb1.data()));
    }
    return bfv_mul_relin_hoisted_impl(ctx, level, n0, ids0, p0.data(), a0.data(), b0.data(), n1, ids1, p1.data(), a1.data(),
                                      bb1.data(), vb1.data(), vb2.data(), vd1.data(), vd2.data(), vv.data(), uk->d, nOut,
                                      idsOut, po.data());
}
int mkhe_bfv_mul_relin(mkhe_ctx *ctx, int n0, const int *ids0, const mkhe_poly *ct0, int n1, const int *ids1,
                       const mkhe_poly *ct1, const mkhe_swk *b1, const mkhe_swk *b2, const mkhe_swk *d1,
                       const mkhe_swk *d2, const mkhe_swk *v, mkhe_swk u, int nOut, const int *idsOut,
                       const mkhe_poly *out) {
    CHECK_CTX();
    TRY(need_bfv(ctx));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    const int level = ctx->nQ - 1;
    std::vector<u64 *> p0, p1, po, r0, r1, a0, b0, a1, bb1, vb1, vb2, vd1, vd2, vv;
    TRY(polys_of(ctx, n0 + 1, ct0, ctx->nQ, p0, "ct0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, ct1, ctx->nQ, p1, "ct1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, ctx->nQ, po, "out"));
    TRY(swks_of(ctx, n1, b1, vb1, "b1", ACC_READ)); TRY(swks_of(ctx, n1, b2, vb2, "b2", ACC_READ));
    TRY(swks_of(ctx, n0, d1, vd1, "d1", ACC_READ)); TRY(swks_of(ctx, n0, d2, vd2, "d2", ACC_READ));
    TRY(swks_of(ctx, n0, v, vv, "v", ACC_READ));
    SWK_R(uk, u);
    // rlkSet.PolyRPool1/2 and HoistPool1/2 (mkbfv/keys.go:40-68) live in the context
    TRY(poly_pool(ctx, "bfv_polyR0", n0 + 1, 2 * ctx->nQ, r0));
    TRY(poly_pool(ctx, "bfv_polyR1", n1 + 1, 2 * ctx->nQ, r1));
    TRY(bfv_modup_impl(ctx, n0 + 1, p0.data(), r0.data()));
    TRY(bfv_rescale_impl(ctx, n1 + 1, p1.data(), r1.data()));
    TRY(swk_pool(ctx, "bfv_h0a", n0, a0)); TRY(swk_pool(ctx, "bfv_h0b", n0, b0));
    TRY(swk_pool(ctx, "bfv_h1a", n1, a1)); TRY(swk_pool(ctx, "bfv_h1b", n1, bb1));
    TRY(bfv_decompose_impl(ctx, level, n0, r0.data() + 1, a0.data(), b0.data()));
    TRY(bfv_decompose_impl(ctx, level, n1, r1.data() + 1, a1.data(), bb1.data()));
    return bfv_mul_relin_hoisted_impl(ctx, level, n0, ids0, r0.data(), a0.data(), b0.data(), n1, ids1, r1.data(), a1.data(),
                                      bb1.data(), vb1.data(), vb2.data(), vd1.data(), vd2.data(), vv.data(), uk->d, nOut,
                                      idsOut, po.data());
}

// ---- multi-GPU --------------------------------------------------------------------------------------
int mkhe_comm_unique_id(uint8_t out[128]) {
#if !defined(MKHE_EMU) && defined(MKHE_WITH_NCCL)
    ncclUniqueId id;
    if (ncclGetUniqueId(&id) != ncclSuccess) return MKHE_ERR_NCCL;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
    memcpy(out, &id, 128);
    return MKHE_OK;
#else
    memset(out, 0, 128);
    return MKHE_ERR_UNSUPPORTED;
#endif
}
int mkhe_comm_init(mkhe_ctx *ctx, int nranks, int rank, const uint8_t unique_id[128]) {
    CHECK_CTX();
#if !defined(MKHE_EMU) && defined(MKHE_WITH_NCCL)
    ncclUniqueId id;
    memcpy(&id, unique_id, 128);
    ncclComm_t comm;
    if (ncclCommInitRank(&comm, nranks, id, rank) != ncclSuccess) return fail(ctx, MKHE_ERR_NCCL, "ncclCommInitRank failed");
    ctx->nccl = comm;
    ctx->nranks = nranks;
    ctx->rank = rank;
    return MKHE_OK;
#else
    (void)nranks; (void)rank; (void)unique_id;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "built without NCCL");
#endif
}
int mkhe_comm_destroy(mkhe_ctx *ctx) {
    if (!ctx) return MKHE_ERR_INVALID;
#if !defined(MKHE_EMU) && defined(MKHE_WITH_NCCL)
    if (ctx->nccl) { ncclCommDestroy((ncclComm_t)ctx->nccl); ctx->nccl = nullptr; }
#endif
    return MKHE_OK;
}

// peer-memory exchange for the sharded MulRelin: every rank exports the IPC handles of its two exchange buffers, the host passes
// all of them to every rank (any transport: the bench uses torch.distributed), import maps the peers' buffers
int mkhe_p2p_export(mkhe_ctx *ctx, uint8_t out[128]) {
    CHECK_CTX();
#if !defined(MKHE_EMU)
    if (!out) return MKHE_ERR_INVALID;
    const size_t bytes = 2 * swk_elems(ctx) * 8;
    if (!ctx->p2p_stage) {
        CU(cudaMalloc((void **)&ctx->p2p_stage, bytes));
        CU(cudaMalloc((void **)&ctx->p2p_xy, bytes));
        CU(cudaMalloc((void **)&ctx->p2p_flag, 256));
        CU(cudaMemset(ctx->p2p_stage, 0, bytes));
        CU(cudaMemset(ctx->p2p_xy, 0, bytes));
        CU(cudaMemset(ctx->p2p_flag, 0, 256));
    }
    cudaIpcMemHandle_t h0, h1;
    CU(cudaIpcGetMemHandle(&h0, ctx->p2p_stage));
    CU(cudaIpcGetMemHandle(&h1, ctx->p2p_xy));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out, &h0, 64);
    memcpy(out + 64, &h1, 64);
    return MKHE_OK;
#else
    (void)out;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "no peer memory under emulation");
#endif
}
int mkhe_p2p_import(mkhe_ctx *ctx, int nranks, int rank, const uint8_t *all_handles) {
    CHECK_CTX();
#if !defined(MKHE_EMU)
    if (!all_handles || nranks < 2 || nranks > MKHE_MAX_RANKS || rank < 0 || rank >= nranks) return fail(ctx, MKHE_ERR_INVALID, "bad rank layout");
    if (!ctx->nccl || ctx->nranks != nranks || ctx->rank != rank) return fail(ctx, MKHE_ERR_NCCL, "mkhe_comm_init with the same layout comes first");
    if (!ctx->p2p_stage) return fail(ctx, MKHE_ERR_INVALID, "mkhe_p2p_export comes first");
    if (ctx->N % (nranks * 2 * MKHE_THREADS) != 0) return fail(ctx, MKHE_ERR_UNSUPPORTED, "N too small for %d ranks", nranks);
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { ctx->peer_stage[r] = ctx->p2p_stage; ctx->peer_xy[r] = ctx->p2p_xy; continue; }
        cudaIpcMemHandle_t h0, h1;
        memcpy(&h0, all_handles + (size_t)r * 128, 64);
        memcpy(&h1, all_handles + (size_t)r * 128 + 64, 64);
        CU(cudaIpcOpenMemHandle((void **)&ctx->peer_stage[r], h0, cudaIpcMemLazyEnablePeerAccess));
        CU(cudaIpcOpenMemHandle((void **)&ctx->peer_xy[r], h1, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->p2p = true;
    return MKHE_OK;
#else
    (void)nranks; (void)rank; (void)all_handles;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "no peer memory under emulation");
#endif
}

// ---- key generation and encryption on the device ---------------------------------------------------
namespace {
// `ninst` small polynomials (kind = ternary / gaussian), instance i from stream stream0 + i, lifted to the limb slots `s`
// (coefficient domain) at out + i * inst_stride
int sample_small(mkhe_ctx *ctx, int kind, uint64_t seed, uint64_t stream0, uint64_t thr53, int ninst, u64 *out, long inst_stride,
                 const Slots &s, const u64 *add_to = nullptr, const u64 *add_pt = nullptr) {
    SampleArgs a;
    memset(&a, 0, sizeof a);
    a.out = out; a.inst_stride = inst_stride; a.seed = seed; a.stream0 = stream0; a.stream_stride = 1; a.kind = kind; a.thr53 = thr53;
    a.add_to = add_to; a.add_pt = add_pt;
    a.nslots = s.n; a.logN = ctx->logN;
    for (int i = 0; i < s.n; i++) a.slots[i] = s.slot[i];
    LAUNCH(k_sample, dim3(ctx->N / MKHE_THREADS, ninst), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
}
// NTT(e_i) for `ninst` gaussian errors over Q u P (GenGaussianError, keygen.go:123-133), swk-shaped
int gaussian_errors_ntt(mkhe_ctx *ctx, uint64_t seed, uint64_t stream0, int ninst, u64 *buf) {
    const Slots qp = qp_slots(ctx, ctx->nQ - 1);
    const long digit_elems = (long)ctx->dmax * ctx->N;
    TRY(sample_small(ctx, SAMPLE_GAUSS, seed, stream0, 0, ninst, buf, digit_elems, qp));
    std::vector<u64 *> ent(ninst);
    for (int i = 0; i < ninst; i++) ent[i] = buf + (long)i * digit_elems;
    return ntt_fwd(ctx, qp, ninst, ent.data(), ent.data());
}
int key_fma(mkhe_ctx *ctx, int ndigits, const u64 *e, const u64 *a, const u64 *s_mul, const u64 *s_gad, u64 *out, int mform_e, int mul_mode,
            int neg, const u64 *gad = nullptr) {
    const Slots qp = qp_slots(ctx, ctx->nQ - 1);
    KeyFmaArgs k;
    memset(&k, 0, sizeof k);
    k.e = e; k.a = a; k.s_mul = s_mul; k.s_gad = s_gad; k.out = out; k.gad = gad;
    k.mform_e = mform_e; k.mul_mode = mul_mode; k.neg = neg;
    k.alpha = ctx->alpha; k.nQ = ctx->nQ; k.dmax = ctx->dmax;
    for (int j = 0; j < ctx->nQ; j++) {
        u64 pm = 1;
        for (int i = 0; i < ctx->nP; i++) pm = h_mulmod(pm, ctx->mod[ctx->nQ + i] % ctx->mod[j], ctx->mod[j]);
        k.pmont[j] = h_mform(pm, ctx->mod[j]);
    }
    k.nslots = qp.n; k.logN = ctx->logN;
    for (int i = 0; i < qp.n; i++) k.slots[i] = qp.slot[i];
    LAUNCH(k_key_fma, dim3(ctx->N / MKHE_THREADS, qp.n, ndigits), dim3(MKHE_THREADS), 0, k, ctx->d_mods);
    return MKHE_OK;
}
int need_qp_poly(mkhe_ctx *ctx, Obj *o, const char *what) {
    if (o->cap_limbs < ctx->dmax) return fail(ctx, MKHE_ERR_INVALID, "%s needs a poly of nQ + nP = %d limbs (rlwe.PolyQP)", what, ctx->dmax);
    return MKHE_OK;
}
}  // namespace

int mkhe_sample_crs(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_swk out) {
    CHECK_CTX();
    SWK(o, out);
    SampleArgs a;
    memset(&a, 0, sizeof a);
    const Slots qp = qp_slots(ctx, ctx->nQ - 1);
    a.out = o->d; a.inst_stride = (long)ctx->dmax * ctx->N; a.seed = seed; a.stream0 = stream; a.stream_stride = ctx->dmax; a.kind = SAMPLE_UNIFORM;
    a.mform = 1;                                                     // params.go:54-56: uniform, then MFormLvl
    a.nslots = qp.n; a.logN = ctx->logN;
    for (int i = 0; i < qp.n; i++) a.slots[i] = qp.slot[i];
    LAUNCH(k_sample, dim3(ctx->N / MKHE_THREADS, ctx->beta_max), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    return MKHE_OK;
}
int mkhe_keygen_secret(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, double pzero, mkhe_poly sk) {
    CHECK_CTX();
    POLY(s, sk);
    TRY(need_qp_poly(ctx, s, "a secret key"));
    if (!(pzero >= 0.0 && pzero <= 1.0)) return fail(ctx, MKHE_ERR_INVALID, "P(0) = %g out of range", pzero);
    const Slots qp = qp_slots(ctx, ctx->nQ - 1);
    // genSecretKeyFromSampler (keygen.go:44-55): ternary -> every limb of Q u P -> NTT -> MForm
    TRY(sample_small(ctx, SAMPLE_TERNARY, seed, stream, (uint64_t)(pzero * 9007199254740992.0), 1, s->d, 0, qp));
    TRY(ntt_fwd(ctx, qp, 1, &s->d, &s->d));
    ScaleArgs a;
    memset(&a, 0, sizeof a);
    a.nlimbs = ctx->dmax; a.logN = ctx->logN;
    for (int i = 0; i < ctx->dmax; i++) { a.mod_of_limb[i] = i; a.cmont[i] = ctx->tabs[i].c.r2; }
    a.in.p[0] = s->d; a.out.p[0] = s->d;
    LAUNCH(k_scale, dim3(ctx->N / MKHE_THREADS, ctx->dmax, 1), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
    s->nlimbs = ctx->dmax;
    return MKHE_OK;
}
int mkhe_keygen_public(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk crs_a, mkhe_poly pk0, mkhe_poly pk1) {
    CHECK_CTX();
    POLY_R(s, sk); SWK_R(a, crs_a); POLY(p0, pk0); POLY(p1, pk1);
    TRY(need_qp_poly(ctx, s, "a secret key")); TRY(need_qp_poly(ctx, p0, "a public key")); TRY(need_qp_poly(ctx, p1, "a public key"));
    // GenPublicKey (keygen.go:88-109): pk1 = CRS[0][0], pk0 = NTT(e) - MRed(sk, pk1)
    u64 *e;
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    TRY(gaussian_errors_ntt(ctx, seed, stream, 1, e));
    CU(cudaMemcpyAsync(p1->d, a->d, (size_t)ctx->dmax * ctx->N * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    TRY(key_fma(ctx, 1, e, a->d, s->d, nullptr, p0->d, 0, 1, 0));
    p0->nlimbs = p1->nlimbs = ctx->dmax;
    return MKHE_OK;
}
int mkhe_keygen_switching_key(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk out) {
    CHECK_CTX();
    POLY_R(s, sk); SWK(o, out);
    TRY(need_qp_poly(ctx, s, "a secret key"));
    u64 *e;
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    TRY(gaussian_errors_ntt(ctx, seed, stream, ctx->beta_max, e));
    return key_fma(ctx, ctx->beta_max, e, nullptr, nullptr, s->d, o->d, 1, 0, 0);
}
int mkhe_keygen_relin(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_poly r, mkhe_swk crs_a, mkhe_swk crs_u,
                      mkhe_swk b, mkhe_swk d, mkhe_swk v) {
    CHECK_CTX();
    POLY_R(s, sk); POLY_R(rr, r); SWK_R(a, crs_a); SWK_R(u, crs_u); SWK(kb, b); SWK(kd, d); SWK(kv, v);
    TRY(need_qp_poly(ctx, s, "a secret key")); TRY(need_qp_poly(ctx, rr, "a secret key"));
    const int beta = ctx->beta_max;
    u64 *e;
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    // GenRelinearizationKey (keygen.go:137-187):  b = -s a + e,  d = -r a + s g + e,  v = -s u - r g + e, streams b | d | v
    TRY(gaussian_errors_ntt(ctx, seed, stream, beta, e));
    TRY(key_fma(ctx, beta, e, a->d, s->d, nullptr, kb->d, 1, 1, 0));
    TRY(gaussian_errors_ntt(ctx, seed, stream + beta, beta, e));
    TRY(key_fma(ctx, beta, e, a->d, rr->d, s->d, kd->d, 1, 1, 0));
    TRY(gaussian_errors_ntt(ctx, seed, stream + 2 * (uint64_t)beta, beta, e));
    return key_fma(ctx, beta, e, u->d, s->d, rr->d, kv->d, 1, 2, 1);
}
namespace {
int permute_sk(mkhe_ctx *ctx, const u64 *in, u64 galEl, u64 **out) {
    TRY(get_scratch(ctx, "kg_sk", (size_t)ctx->dmax * ctx->N * 8, out));
    LAUNCH(k_permute_ntt, dim3(ctx->N / MKHE_THREADS, ctx->dmax), dim3(MKHE_THREADS), 0, in, *out, galEl, ctx->dmax, ctx->logN);
    return MKHE_OK;
}
}  // namespace
int mkhe_keygen_rotation(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, int rotidx, mkhe_poly sk, mkhe_swk crs_rot, mkhe_swk rk) {
    CHECK_CTX();
    POLY_R(s, sk); SWK_R(a, crs_rot); SWK(o, rk);
    TRY(need_qp_poly(ctx, s, "a secret key"));
    while (rotidx < 0) rotidx += ctx->N / 2;
    // GenRotationKey (keygen.go:190-229): rk = -sigma^-1(s) a_rot + s g + e
    const u64 mask = ((u64)2 << ctx->logN) - 1, gal = galois_for_rotation(ctx->logN, rotidx);
    u64 inv = 1;
    for (int it = 0; it < 6; it++) inv = (inv * (2 - gal * inv)) & mask;
    u64 *skOut, *e;
    TRY(permute_sk(ctx, s->d, inv, &skOut));
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    TRY(gaussian_errors_ntt(ctx, seed, stream, ctx->beta_max, e));
    return key_fma(ctx, ctx->beta_max, e, a->d, skOut, s->d, o->d, 1, 1, 0);
}
int mkhe_keygen_conjugation(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_swk crs_cj, mkhe_swk ck) {
    CHECK_CTX();
    POLY_R(s, sk); SWK_R(a, crs_cj); SWK(o, ck);
    TRY(need_qp_poly(ctx, s, "a secret key"));
    // GenConjugationKey (keygen.go:232-266): cj = -s a_cj + sigma_c(s) g + e
    u64 *skOut, *e;
    TRY(permute_sk(ctx, s->d, ((u64)2 << ctx->logN) - 1, &skOut));
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    TRY(gaussian_errors_ntt(ctx, seed, stream, ctx->beta_max, e));
    return key_fma(ctx, ctx->beta_max, e, a->d, s->d, skOut, o->d, 1, 1, 0);
}
/* mkbfv.KeyGenerator.GenRelinearizationKey (mkbfv/keygen.go:24-88) with GenBFVSwitchingKey (:91-162): the gadget of digit i is
 *   swk1: G_i = floor(R/q_i * t * [(R/q_i)^-1]_{q_i} * P / QMul) = (Q/q_i) t T_i P                      (R = Q QMul; exact)
 *   swk2: G_i = floor(R/q'_i * t * [(R/q'_i)^-1]_{q'_i} * P / QMul) = floor(Q t T_i P / q'_i)            (q'_i = QMul_i)
 * reduced here modulo every limb of Q u P without big integers: the first is a product of residues; for the second X = Q t T_i P is
 * 0 modulo every limb of Q u P, so floor(X / q') = (X - X mod q') / q' = -(X mod q') (q')^-1 there.
 * Streams: b1[i] = stream + 2i, b2[i] = stream + 2i + 1 (the reference interleaves them), then d1 | d2 | v, beta each. */
int mkhe_keygen_bfv_relin(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, mkhe_poly sk, mkhe_poly r, mkhe_swk crs_a1, mkhe_swk crs_a2,
                          mkhe_swk crs_u, mkhe_swk b1, mkhe_swk b2, mkhe_swk d1, mkhe_swk d2, mkhe_swk v) {
    CHECK_CTX();
    if (!ctx->nQMul) return fail(ctx, MKHE_ERR_INVALID, "BFV parameters not set (mkhe_ctx_set_bfv)");
    if (ctx->alpha != 1) return fail(ctx, MKHE_ERR_UNSUPPORTED, "mkbfv key generation needs alpha = 1 (mkbfv/keyswitch.go:64-67)");
    POLY_R(s, sk); POLY_R(rr, r); SWK_R(a1, crs_a1); SWK_R(a2, crs_a2); SWK_R(u, crs_u);
    SWK(kb1, b1); SWK(kb2, b2); SWK(kd1, d1); SWK(kd2, d2); SWK(kv, v);
    TRY(need_qp_poly(ctx, s, "a secret key")); TRY(need_qp_poly(ctx, rr, "a secret key"));
    const int beta = ctx->beta_max, nQ = ctx->nQ, nP = ctx->nP, D = ctx->dmax;
    const u64 *Q = ctx->mod.data(), *P = Q + nQ, *QM = Q + nQ + nP;
    // gadget tables [2][beta][D], Montgomery form
    std::vector<u64> gad((size_t)2 * beta * D);
    auto prod_mod = [&](const u64 *f, int n, int skip, u64 m) { u64 x = 1; for (int l = 0; l < n; l++) if (l != skip) x = h_mulmod(x, f[l] % m, m); return x; };
    for (int i = 0; i < beta; i++) {
        const u64 qi = Q[i], qmi = QM[i];
        const u64 T1 = h_invmod(h_mulmod(prod_mod(Q, nQ, i, qi), prod_mod(QM, nQ, -1, qi), qi), qi);           // [(R/q_i)^-1]_{q_i}
        const u64 T2 = h_invmod(h_mulmod(prod_mod(Q, nQ, -1, qmi), prod_mod(QM, nQ, i, qmi), qmi), qmi);       // [(R/q'_i)^-1]_{q'_i}
        const u64 Xq = h_mulmod(h_mulmod(prod_mod(Q, nQ, -1, qmi), ctx->T % qmi, qmi), h_mulmod(T2, prod_mod(P, nP, -1, qmi), qmi), qmi);   // X mod q'_i
        for (int j = 0; j < D; j++) {
            const u64 m = ctx->mod[j];
            u64 g1 = h_mulmod(h_mulmod(prod_mod(Q, nQ, i, m), ctx->T % m, m), h_mulmod(T1 % m, prod_mod(P, nP, -1, m), m), m);
            u64 g2 = h_mulmod((m - Xq % m) % m, h_invmod(qmi % m, m), m);
            gad[((size_t)0 * beta + i) * D + j] = h_mform(g1, m);
            gad[((size_t)1 * beta + i) * D + j] = h_mform(g2, m);
        }
    }
    u64 *d_gad, *e;
    TRY(get_scratch(ctx, "kg_gad", gad.size() * 8, &d_gad));
    CU(cudaMemcpyAsync(d_gad, gad.data(), gad.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));                          // `gad` is pageable host memory of this frame
    TRY(get_scratch(ctx, "kg_e", swk_elems(ctx) * 8, &e));
    const long digit_elems = (long)D * ctx->N;
    // b1[i] = -s a1[i] + e (stream + 2i), b2[i] = -s a2[i] + e (stream + 2i + 1)
    {
        const Slots qp = qp_slots(ctx, nQ - 1);
        for (int which = 0; which < 2; which++) {
            SampleArgs a;
            memset(&a, 0, sizeof a);
            a.out = e; a.inst_stride = digit_elems; a.seed = seed; a.stream0 = stream + which; a.stream_stride = 2; a.kind = SAMPLE_GAUSS;
            a.nslots = qp.n; a.logN = ctx->logN;
            for (int i = 0; i < qp.n; i++) a.slots[i] = qp.slot[i];
            LAUNCH(k_sample, dim3(ctx->N / MKHE_THREADS, beta), dim3(MKHE_THREADS), 0, a, ctx->d_mods);
            std::vector<u64 *> ent(beta);
            for (int i = 0; i < beta; i++) ent[i] = e + (long)i * digit_elems;
            TRY(ntt_fwd(ctx, qp, beta, ent.data(), ent.data()));
            TRY(key_fma(ctx, beta, e, which ? a2->d : a1->d, s->d, nullptr, which ? kb2->d : kb1->d, 1, 1, 0));
        }
    }
    // d_k = -r a_k + s G^(k) + e
    TRY(gaussian_errors_ntt(ctx, seed, stream + 2 * (uint64_t)beta, beta, e));
    TRY(key_fma(ctx, beta, e, a1->d, rr->d, s->d, kd1->d, 1, 1, 0, d_gad));
    TRY(gaussian_errors_ntt(ctx, seed, stream + 3 * (uint64_t)beta, beta, e));
    TRY(key_fma(ctx, beta, e, a2->d, rr->d, s->d, kd2->d, 1, 1, 0, d_gad + (size_t)beta * D));
    // v = -s u - r g + e with the plain gadget (GenSwitchingKey)
    TRY(gaussian_errors_ntt(ctx, seed, stream + 4 * (uint64_t)beta, beta, e));
    return key_fma(ctx, beta, e, u->d, s->d, rr->d, kv->d, 1, 2, 1);
}
/* Encryptor.Encrypt, the coefficient-domain branch the CKKS / BFV flows take (mkrlwe/encryptor.go:55-118):
 *   c0 = InvNTT(MForm(NTT(w)) (.) pk0) + e0 + pt,  c1 = InvNTT(MForm(NTT(w)) (.) pk1) + e1;  w ternary (stream), e0 (stream + 1), e1 (stream + 2) */
int mkhe_encrypt(mkhe_ctx *ctx, uint64_t seed, uint64_t stream, int level, mkhe_poly pt, mkhe_poly pk0, mkhe_poly pk1, mkhe_poly c0,
                 mkhe_poly c1) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    POLY_R(p0, pk0); POLY_R(p1, pk1); POLY(o0, c0); POLY(o1, c1);
    Obj *ptx = nullptr;
    if (pt) { ptx = as_obj(ctx, pt, OBJ_POLY, ACC_READ); if (!ptx || ptx->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "Encrypt: bad plaintext"); }
    if (p0->cap_limbs < level + 1 || p1->cap_limbs < level + 1 || o0->cap_limbs < level + 1 || o1->cap_limbs < level + 1) return fail(ctx, MKHE_ERR_INVALID, "Encrypt: too few limbs");
    const Slots qs = q_slots(level);
    std::vector<u64 *> w;
    TRY(poly_pool(ctx, "enc_w", 1, ctx->nQ, w));
    TRY(sample_small(ctx, SAMPLE_TERNARY, seed, stream, (uint64_t)1 << 52, 1, w[0], 0, qs));
    TRY(ntt_fwd(ctx, qs, 1, w.data(), w.data()));
    TRY(mul2(ctx, qs, w[0], p0->d, nullptr, nullptr, o0->d));          // MForm(w) (.) pk0
    TRY(mul2(ctx, qs, w[0], p1->d, nullptr, nullptr, o1->d));
    u64 *outs[2] = {o0->d, o1->d};
    TRY(ntt_inv(ctx, qs, 2, outs, outs));
    TRY(sample_small(ctx, SAMPLE_GAUSS, seed, stream + 1, 0, 1, o0->d, 0, qs, o0->d, ptx ? ptx->d : nullptr));
    TRY(sample_small(ctx, SAMPLE_GAUSS, seed, stream + 2, 0, 1, o1->d, 0, qs, o1->d, nullptr));
    o0->nlimbs = o1->nlimbs = level + 1;
    return MKHE_OK;
}

// ---- limb-sharded ops: the team of ranks ---------------------------------------------------------------
namespace {
int team_alloc(mkhe_ctx *ctx, int max_parties) {
    mkhe_ctx::Team &tm = ctx->team;
    if (tm.mem && tm.maxk >= max_parties) return MKHE_OK;
    if (tm.mem) return fail(ctx, MKHE_ERR_INVALID, "team memory already exported for %d parties", tm.maxk);
    if (max_parties < 1 || max_parties > MKHE_MAX_PARTIES) return fail(ctx, MKHE_ERR_INVALID, "max_parties out of range");
    const size_t N = ctx->N;
    tm.off_pp = MKHE_TEAM_FLAGS;
    tm.off_p = tm.off_pp + (size_t)2 * max_parties * 2 * ctx->nP * N;
    tm.off_g = tm.off_p + (size_t)max_parties * ctx->nQ * N;
    tm.off_in = tm.off_g + (size_t)(max_parties + 1) * ctx->nQ * N;
    tm.elems = tm.off_in + (size_t)2 * 2 * (max_parties + 1) * ctx->nQ * N;
    if (cudaMalloc((void **)&tm.mem, tm.elems * 8) != cudaSuccess) return fail(ctx, MKHE_ERR_NOMEM, "cudaMalloc(%zu) failed for the team memory", tm.elems * 8);
    CU(cudaMemsetAsync(tm.mem, 0, tm.elems * 8, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    tm.maxk = max_parties;
    tm.spin = getenv("MKHE_DEBUG_SPIN_BARRIER") != nullptr;
    return MKHE_OK;
}
}  // namespace

int mkhe_team_export(mkhe_ctx *ctx, int max_parties, uint8_t out[64]) {
    CHECK_CTX();
#if !defined(MKHE_EMU)
    if (!out) return MKHE_ERR_INVALID;
    TRY(team_alloc(ctx, max_parties));
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, ctx->team.mem));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(out, &h, 64);
    return MKHE_OK;
#else
    (void)max_parties; (void)out;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "no inter-process memory under emulation");
#endif
}
int mkhe_team_import(mkhe_ctx *ctx, int nranks, int rank, const uint8_t *all_handles) {
    CHECK_CTX();
#if !defined(MKHE_EMU)
    mkhe_ctx::Team &tm = ctx->team;
    if (!all_handles || nranks < 1 || nranks > MKHE_MAX_RANKS || rank < 0 || rank >= nranks) return fail(ctx, MKHE_ERR_INVALID, "bad rank layout");
    if (!tm.mem) return fail(ctx, MKHE_ERR_INVALID, "mkhe_team_export comes first");
    for (int r = 0; r < nranks; r++) {
        if (r == rank) { tm.peer[r] = tm.mem; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, all_handles + (size_t)r * 64, 64);
        CU(cudaIpcOpenMemHandle((void **)&tm.peer[r], h, cudaIpcMemLazyEnablePeerAccess));
        tm.ipc[r] = true;
    }
    tm.n = nranks;
    tm.rank = rank;
    return MKHE_OK;
#else
    (void)nranks; (void)rank; (void)all_handles;
    return fail(ctx, MKHE_ERR_UNSUPPORTED, "no inter-process memory under emulation");
#endif
}
int mkhe_team_join_local(mkhe_ctx *ctx, int max_parties, int nranks, int rank, mkhe_ctx *const *members) {
    CHECK_CTX();
    mkhe_ctx::Team &tm = ctx->team;
    if (!members || nranks < 1 || nranks > MKHE_MAX_RANKS || rank < 0 || rank >= nranks || members[rank] != ctx)
        return fail(ctx, MKHE_ERR_INVALID, "bad rank layout");
    TRY(team_alloc(ctx, max_parties));
    for (int r = 0; r < nranks; r++) {
        mkhe_ctx *m = members[r];
        if (!m || m->logN != ctx->logN || m->nQ != ctx->nQ || m->nP != ctx->nP) return fail(ctx, MKHE_ERR_INVALID, "team members must share the parameters");
        if (m != ctx) {
            if (cudaSetDevice(m->device) != cudaSuccess) return fail(ctx, MKHE_ERR_CUDA, "cudaSetDevice failed");
            int rc = team_alloc(m, max_parties);
            cudaSetDevice(ctx->device);
            if (rc != MKHE_OK) return fail(ctx, rc, "team memory of member %d: %s", r, m->err.c_str());
#ifndef MKHE_EMU
            if (m->device != ctx->device) {
                cudaError_t e = cudaDeviceEnablePeerAccess(m->device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return fail(ctx, MKHE_ERR_CUDA, "no peer access from device %d to %d", ctx->device, m->device);
                cudaGetLastError();
            }
#endif
        }
        tm.peer[r] = m->team.mem;
    }
    tm.n = nranks;
    tm.rank = rank;
    return MKHE_OK;
}
int mkhe_team_status(mkhe_ctx *ctx, int *timed_out) {
    CHECK_CTX();
    if (!timed_out) return MKHE_ERR_INVALID;
    *timed_out = 0;
    if (!ctx->team.mem) return MKHE_OK;
    u64 w = 0;
    CU(cudaMemcpyAsync(&w, ctx->team.mem + MKHE_MAX_RANKS, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    *timed_out = w != 0;
    return MKHE_OK;
}
int mkhe_team_allgather(mkhe_ctx *ctx, int level, int npolys, const mkhe_poly *polys) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    mkhe_ctx::Team &tm = ctx->team;
    if (tm.n < 1) return fail(ctx, MKHE_ERR_INVALID, "mkhe_team_import / mkhe_team_join_local has not been called");
    if (npolys < 1 || npolys > 2 * (tm.maxk + 1) || npolys > MKHE_MAX_PARTIES_K) return fail(ctx, MKHE_ERR_INVALID, "allgather of %d polys (team memory sized for %d)", npolys, 2 * (tm.maxk + 1));
    std::vector<u64 *> pp;
    TRY(polys_of(ctx, npolys, polys, level + 1, pp, "polys"));
    if (tm.n == 1) return MKHE_OK;
    const Slots own = own_slots(ctx, q_slots(level));
    const size_t poly_elems = (size_t)ctx->nQ * ctx->N;
    const size_t base = tm.off_in + (size_t)(tm.in_parity++ & 1) * 2 * (tm.maxk + 1) * poly_elems;     // two areas, alternating
    GatherArgs ga;
    memset(&ga, 0, sizeof ga);
    ga.logN = ctx->logN;
    for (int i = 0; i < npolys; i++) { ga.src.p[i] = pp[i]; ga.dst_off[i] = (long)(base + (size_t)i * poly_elems); }
    if (own.n > 0) {
        ga.nslots = own.n;
        for (int i = 0; i < own.n; i++) ga.slots[i] = own.slot[i];
        TeamArgs ta;
        fill_team(ctx, ta);
        LAUNCH(k_team_gather, dim3(ctx->N / (2 * MKHE_THREADS), own.n, npolys), dim3(MKHE_THREADS), 0, ga, ta);
    }
    TRY(team_barrier(ctx));
    // the other ranks' limbs: from this rank's staging area into the polys
    PtrList dst;
    memset(&dst, 0, sizeof dst);
    ga.nslots = 0;
    for (int j = 0; j <= level; j++)
        if (owner_of_slot(ctx, j) != tm.rank) ga.slots[ga.nslots++] = j;
    for (int i = 0; i < npolys; i++) { ga.src.p[i] = tm.mem + base + (size_t)i * poly_elems; dst.p[i] = pp[i]; }
    if (ga.nslots > 0) LAUNCH(k_copy_limbs, dim3(ctx->N / (2 * MKHE_THREADS), ga.nslots, npolys), dim3(MKHE_THREADS), 0, ga, dst);
    return MKHE_OK;
}
int mkhe_team_owns_limb(mkhe_ctx *ctx, int limb) {
    if (!ctx || ctx->team.n < 1) return 1;
    return owner_of_slot(ctx, limb) == ctx->team.rank ? 1 : 0;
}
// diagnostics: the 16 flag / status words at the base of this rank's team memory and the number of barriers it has issued
int mkhe_team_flags(mkhe_ctx *ctx, uint64_t out[17]) {
    CHECK_CTX();
    if (!out || !ctx->team.mem) return MKHE_ERR_INVALID;
    CU(cudaMemcpy(out, ctx->team.mem, 16 * 8, cudaMemcpyDeviceToHost));
    out[16] = ctx->team.epoch;
    return MKHE_OK;
}
int mkhe_ckks_mul_relin_limbs(mkhe_ctx *ctx, int level, int nb_rescales, int n0, const int *ids0, const mkhe_poly *op0, int n1,
                              const int *ids1, const mkhe_poly *op1, const mkhe_swk *rlk_b, const mkhe_swk *rlk_d,
                              const mkhe_swk *rlk_v, mkhe_swk u, int nOut, const int *idsOut, const mkhe_poly *out) {
    CHECK_CTX();
    TRY(check_level(ctx, level));
    TRY(check_ids(ctx, n0, ids0, n1, ids1, nOut, idsOut));
    if (ctx->team.n < 1) return fail(ctx, MKHE_ERR_INVALID, "mkhe_team_import / mkhe_team_join_local has not been called");
    if (nb_rescales < 0 || nb_rescales > level) return fail(ctx, MKHE_ERR_INVALID, "cannot Rescale: nb_rescales = %d at level %d", nb_rescales, level);
    std::vector<u64 *> p0, p1, po, vb, vd, vv;
    TRY(polys_of(ctx, n0 + 1, op0, level + 1, p0, "op0", ACC_READ));
    TRY(polys_of(ctx, n1 + 1, op1, level + 1, p1, "op1", ACC_READ));
    TRY(polys_of(ctx, nOut + 1, out, level + 1, po, "out"));
    TRY(swks_of(ctx, n1, rlk_b, vb, "rlk_b", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_d, vd, "rlk_d", ACC_READ));
    TRY(swks_of(ctx, n0, rlk_v, vv, "rlk_v", ACC_READ));
    SWK_R(uk, u);
    TRY(mul_relin_limbs_impl(ctx, level, nb_rescales, n0, ids0, p0.data(), n1, ids1, p1.data(), vb.data(), vd.data(), vv.data(), uk->d,
                             nOut, idsOut, po.data()));
    for (int t = 0; t <= nOut; t++) reinterpret_cast<Obj *>(out[t])->nlimbs = level + 1 - nb_rescales;
    return MKHE_OK;
}

// ---- measurement ------------------------------------------------------------------------------------
int mkhe_timer_start(mkhe_ctx *ctx) {
    CHECK_CTX();
    CU(cudaEventRecord(ctx->ev0, ctx->stream));
    return MKHE_OK;
}
int mkhe_timer_stop(mkhe_ctx *ctx, float *elapsed_ms) {
    CHECK_CTX();
    CU(cudaEventRecord(ctx->ev1, ctx->stream));
    CU(cudaEventSynchronize(ctx->ev1));
    CU(cudaEventElapsedTime(elapsed_ms, ctx->ev0, ctx->ev1));
    return MKHE_OK;
}
int mkhe_profile_begin(mkhe_ctx *ctx) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    for (auto &r : ctx->prof) { ctx->ev_pool.push_back(r.a); ctx->ev_pool.push_back(r.b); }
    ctx->prof.clear();
    ctx->profiling = true;
    return MKHE_OK;
}
int mkhe_profile_end(mkhe_ctx *ctx, char *buf, size_t cap) {
    CHECK_CTX();
    ctx->profiling = false;
    CU(cudaStreamSynchronize(ctx->stream));
    std::map<std::string, std::pair<double, long>> agg;
    for (auto &r : ctx->prof) {
        float ms = 0;
        cudaEventElapsedTime(&ms, r.a, r.b);
        auto &e = agg[r.name];
        e.first += ms;
        e.second += 1;
        ctx->ev_pool.push_back(r.a);
        ctx->ev_pool.push_back(r.b);
    }
    ctx->prof.clear();
    std::string out;
    for (auto &kv : agg) {
        char line[256];
        snprintf(line, sizeof line, "%s %ld %.6f\n", kv.first.c_str(), kv.second.second, kv.second.first);
        out += line;
    }
    if (buf && cap) { strncpy(buf, out.c_str(), cap - 1); buf[cap - 1] = 0; }
    return MKHE_OK;
}
#ifdef MKHE_P2_TIMING
// development builds only: {start ns, end ns, SM id, tiles} of every CTA of the last pass-2 launch
int mkhe_debug_p2_timing(mkhe_ctx *ctx, uint64_t *out, int cap, int *n) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    *n = std::min(cap, ctx->p2_timing_n);
    CU(cudaMemcpy(out, ctx->scratch["p2timing"].p, (size_t)*n * 32, cudaMemcpyDeviceToHost));
    return MKHE_OK;
}
#endif

int mkhe_debug_snap_read(mkhe_ctx *ctx, int row, uint64_t *out, size_t cap_bytes, size_t *bytes) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    if (!ctx->ck_snap || row < 0 || row >= 64 || cap_bytes < ctx->ck_snap_bytes) return fail(ctx, MKHE_ERR_INVALID, "no snapshot");
    *bytes = ctx->ck_snap_bytes;
    CU(cudaMemcpy(out, (char *)ctx->ck_snap + (size_t)row * ctx->ck_snap_bytes, ctx->ck_snap_bytes, cudaMemcpyDeviceToHost));
    return MKHE_OK;
}
// development: the checksum log (one row of 16 sums per launch since the last call) and the row labels "kernel scratch-names..."
int mkhe_debug_ck_fetch(mkhe_ctx *ctx, uint64_t *out, int cap_rows, int *nrows, char *names, size_t names_cap) {
    CHECK_CTX();
    CU(cudaStreamSynchronize(ctx->stream));
    *nrows = std::min(cap_rows, ctx->ck_n);
    if (*nrows > 0 && ctx->ck_log) CU(cudaMemcpy(out, ctx->ck_log, (size_t)*nrows * MKHE_CK_SLOTS * 8, cudaMemcpyDeviceToHost));
    std::string all;
    for (int i = 0; i < *nrows; i++) all += ctx->ck_names[i] + "\n";
    if (names && names_cap) snprintf(names, names_cap, "%s", all.c_str());
    if (ctx->ck_log) {
        CU(cudaMemsetAsync(ctx->ck_log, 0, (size_t)MKHE_CK_MAX_LAUNCHES * MKHE_CK_SLOTS * 8, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    ctx->ck_n = 0;
    ctx->ck_names.clear();
    return MKHE_OK;
}

int mkhe_bench_butterfly_peak(mkhe_ctx *ctx, double *butterflies_per_s) {
    CHECK_CTX();
    u64 *sink;
    TRY(get_scratch(ctx, "sink", 64, &sink));
    const int iters = 2000, blocks = 148 * 8;
    ModC m = ctx->tabs[0].c;
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        CU(cudaEventRecord(ctx->ev0, ctx->stream));
        LAUNCH(k_bfly_peak, dim3(blocks), dim3(MKHE_THREADS), 0, sink, m, iters);
        CU(cudaEventRecord(ctx->ev1, ctx->stream));
        CU(cudaEventSynchronize(ctx->ev1));
        float ms;
        CU(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
        if (rep > 0 && ms < best) best = ms;
    }
    *butterflies_per_s = (double)blocks * MKHE_THREADS * 12.0 * iters / (best * 1e-3);
    return MKHE_OK;
}

}  // extern "C"

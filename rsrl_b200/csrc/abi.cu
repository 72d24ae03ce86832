// abi.cu — the C ABI of include/rsrl_b200.h: engine lifecycle, host<->device marshalling, dispatch.
// No CPU fallback: every compute entry point needs a CUDA device.
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "launch.h"
#include "tile.cuh"
#include "fourier4.cuh"
#include "f4tc_launch.h"

using namespace rsrl;

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

#define CU_TRY(expr)                                                                                          \
    do {                                                                                                      \
        cudaError_t e__ = (expr);                                                                             \
        if (e__ != cudaSuccess)                                                                               \
            return fail(e__ == cudaErrorMemoryAllocation ? RSRL_ENOMEM : RSRL_ECUDA,                          \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                                 \
    } while (0)

static int need_device() {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(RSRL_ENODEVICE, "no CUDA device visible: rsrl_b200 has no CPU fallback");
    }
    return RSRL_OK;
}

struct DevBuf {  // RAII device allocation for the stateless entry points
    void* p = nullptr;
    cudaError_t alloc(size_t bytes) { return cudaMalloc(&p, bytes ? bytes : 1); }
    ~DevBuf() { if (p) cudaFree(p); }
    template <class T> T* as() { return static_cast<T*>(p); }
};

// ---------------------------------------------------------------------------
// config helpers
// ---------------------------------------------------------------------------
static int dom_dim(int d) { return d == RSRL_MOUNTAIN_CAR ? 2 : 4; }
static int dom_actions(int d) { return d == RSRL_CART_POLE ? 2 : 3; }
static bool algo_td_pred(int a) { return a == RSRL_TD_LAMBDA || a == RSRL_TD0; }
static bool algo_two_tables(int a) { return a == RSRL_GREEDY_GQ || a == RSRL_A2C; }
static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

static int validate(const rsrl_config_t* c) {
    if (!c) return fail(RSRL_EINVAL, "null config");
    if (c->struct_size != sizeof(rsrl_config_t)) return fail(RSRL_EINVAL, "rsrl_config_t.struct_size mismatch (ABI)");
    if (c->domain < 0 || c->domain > 2) return fail(RSRL_EINVAL, "unknown domain");
    if (c->basis < 0 || c->basis > 2) return fail(RSRL_EINVAL, "unknown basis");
    if (c->algo < 0 || c->algo > RSRL_A2C) return fail(RSRL_EINVAL, "unknown algo");
    if (c->policy < 0 || c->policy > RSRL_SOFTMAX) return fail(RSRL_EINVAL, "unknown policy");
    if (c->dtype != RSRL_F32 && c->dtype != RSRL_F64) return fail(RSRL_EINVAL, "unknown dtype");
    if (c->weight_mode != RSRL_SHARED && c->weight_mode != RSRL_PER_ENV) return fail(RSRL_EINVAL, "unknown weight_mode");
    if (c->trace_rule < RSRL_TRACE_ACCUMULATE || c->trace_rule > RSRL_TRACE_DUTCH) return fail(RSRL_EINVAL, "unknown trace_rule");
    if (c->update_scale != RSRL_SCALE_SUM && c->update_scale != RSRL_SCALE_MEAN) return fail(RSRL_EINVAL, "unknown update_scale");
    if (c->init_mode != RSRL_INIT_DEFAULT && c->init_mode != RSRL_INIT_UNIFORM) return fail(RSRL_EINVAL, "unknown init_mode");
    if (!std::isfinite(c->lr) || !std::isfinite(c->alpha) || !std::isfinite(c->gamma) || !std::isfinite(c->lambda) || !std::isfinite(c->epsilon))
        return fail(RSRL_EINVAL, "lr, alpha, gamma, lambda and epsilon must be finite");
    if (c->n_envs <= 0) return fail(RSRL_EINVAL, "n_envs must be > 0");
    if (c->n_envs_global != 0 && c->n_envs_global < c->env_offset + c->n_envs)
        return fail(RSRL_EINVAL, "n_envs_global must cover env_offset + n_envs (or be 0)");
    if (c->max_episode_steps < 0) return fail(RSRL_EINVAL, "max_episode_steps must be >= 0");
    if (c->env_offset < 0 || c->env_offset + c->n_envs > 0xFFFFFFFFll) return fail(RSRL_EINVAL, "global env ids must fit 32 bits");
    if (c->policy == RSRL_SOFTMAX) {
        if (!(c->epsilon <= -1e-7 || c->epsilon >= 1e-7)) return fail(RSRL_EINVAL, "Softmax: the temperature tau (epsilon field) must be non-zero (softmax.rs:61-64)");
    } else if (!(c->epsilon >= 0.0)) return fail(RSRL_EINVAL, "epsilon must be >= 0");
    if (c->basis != RSRL_TILE_CODING && (c->basis_order < 1 || c->basis_order > 7)) return fail(RSRL_EINVAL, "basis_order must be in 1..7");
    if (c->basis == RSRL_TILE_CODING) {
        if (c->n_tilings < 1 || c->n_tilings > kMaxTilings) return fail(RSRL_EINVAL, "n_tilings must be in 1..16");
        if (c->tiles_per_dim < 1) return fail(RSRL_EINVAL, "tiles_per_dim must be >= 1");
        if (c->memory_size < 2 || (c->memory_size & (c->memory_size - 1))) return fail(RSRL_EINVAL, "memory_size must be a power of two");
    }
    if (c->algo == RSRL_A2C && c->policy != RSRL_SOFTMAX)
        return fail(RSRL_EINVAL, "A2C: the policy is the Gibbs / Softmax policy over its own LFA (examples/a2c.rs:28): set policy = RSRL_SOFTMAX, tau in epsilon");
    if (algo_td_pred(c->algo) && c->policy != RSRL_RANDOM)
        return fail(RSRL_EINVAL, "TD(0)/TD(lambda) predict V(s): the behaviour policy must be RSRL_RANDOM");
    return RSRL_OK;
}

static int64_t n_features(const rsrl_config_t* c) {
    return c->basis == RSRL_TILE_CODING ? c->memory_size : ipow(c->basis_order + 1, dom_dim(c->domain));
}
static TileParams tile_params(const rsrl_config_t* c) {
    TileParams tp; tp.n_tilings = c->n_tilings; tp.tiles_per_dim = c->tiles_per_dim; tp.memory_mask = c->memory_size - 1;
    tp.div_magic = tile_div_magic(c->n_tilings); return tp;
}

static bool is_f4(const rsrl_config_t* c) {
    return c->basis == RSRL_FOURIER && dom_dim(c->domain) == 4 && (c->basis_order == 5 || c->basis_order == 7);
}

static BasisKey key_of(const rsrl_config_t* c) {
    BasisKey k;
    k.dtype = c->dtype; k.domain = c->domain; k.basis = c->basis; k.order = c->basis_order;
    k.aw = algo_td_pred(c->algo) ? 1 : dom_actions(c->domain);
    return k;
}

static PolicyParams policy_of(int policy, double eps, uint64_t seed) {
    PolicyParams p;
    p.policy = policy;
    p.tau = eps;
    p.eps_always = eps >= 1.0;
    p.eps_thresh = p.eps_always ? 0xFFFFFFFFu : (uint32_t)(eps * 4294967296.0);
    p.seed = seed;
    return p;
}

static cudaError_t dispatch_fused(const BasisKey& k, int mode, bool ext, const StepArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    static const fused_launch_fn table[2][3] = {{launch_fused_f32_d0, launch_fused_f32_d1, launch_fused_f32_d2},
                                                {launch_fused_f64_d0, launch_fused_f64_d1, launch_fused_f64_d2}};
    return table[k.dtype][k.domain](k, mode, ext, a, grid, block, smem, st);
}
// the (basis, order) pair exists as template instantiations; otherwise the run-time-order kernels of dyn.cuh take the engine
static bool has_static(const BasisKey& k) {
    typedef bool (*fn)(const BasisKey&);
    static const fn table[2][3] = {{has_static_f32_d0, has_static_f32_d1, has_static_f32_d2}, {has_static_f64_d0, has_static_f64_d1, has_static_f64_d2}};
    return table[k.dtype][k.domain](k);
}
static cudaError_t dispatch_eval(const BasisKey& k, const EvalArgs& e, cudaStream_t st) {
    static const eval_launch_fn table[2][3] = {{launch_eval_f32_d0, launch_eval_f32_d1, launch_eval_f32_d2},
                                               {launch_eval_f64_d0, launch_eval_f64_d1, launch_eval_f64_d2}};
    return table[k.dtype][k.domain](k, e, st);
}

static cudaError_t dispatch_two(const BasisKey& k, int mode, bool ext, const StepArgs& a, int grid, int block, size_t smem, cudaStream_t st) {
    static const two_launch_fn table[2][3] = {{launch_two_f32_d0, launch_two_f32_d1, launch_two_f32_d2},
                                              {launch_two_f64_d0, launch_two_f64_d1, launch_two_f64_d2}};
    return table[k.dtype][k.domain](k, mode, ext, a, grid, block, smem, st);
}
static cudaError_t dispatch_rollout(const BasisKey& k, const RolloutArgs& ra, cudaStream_t st) {
    static const rollout_launch_fn table[2][3] = {{launch_rollout_f32_d0, launch_rollout_f32_d1, launch_rollout_f32_d2},
                                                  {launch_rollout_f64_d0, launch_rollout_f64_d1, launch_rollout_f64_d2}};
    return table[k.dtype][k.domain](k, ra, st);
}

static cudaError_t dispatch_persist(const BasisKey& k, int mode, const StepArgs& a, int k_steps, const SyncArgs& sy, const PeerArgs& pe, int grid, int block, size_t smem, cudaStream_t st, int* max_clusters = nullptr) {
    static const persist_launch_fn table[2][3] = {{launch_persist_f32_d0, launch_persist_f32_d1, launch_persist_f32_d2},
                                                  {launch_persist_f64_d0, launch_persist_f64_d1, launch_persist_f64_d2}};
    return table[k.dtype][k.domain](k, mode, a, k_steps, sy, pe, grid, block, smem, st, max_clusters);
}

static int unsupported(const rsrl_config_t* c) {
    char buf[160];
    snprintf(buf, sizeof buf, "combination not built: domain=%d basis=%d order=%d dtype=%d (see DESIGN.md)", c->domain, c->basis, c->basis_order, c->dtype);
    return fail(RSRL_EUNSUPPORTED, buf);
}

// ---------------------------------------------------------------------------
// NCCL through dlopen (only needed for multi-GPU SHARED mode)
// ---------------------------------------------------------------------------
struct NcclApi {
    void* h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl() {
    if (g_nccl.h) return RSRL_OK;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return fail(RSRL_ECOMM, std::string("dlopen libnccl.so.2: ") + dlerror());
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(h, "ncclCommInitRank");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(h, "ncclAllReduce");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(h, "ncclCommDestroy");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce || !g_nccl.CommDestroy)
        return fail(RSRL_ECOMM, "libnccl.so.2 lacks a required symbol");
    g_nccl.h = h;
    return RSRL_OK;
}

// ---------------------------------------------------------------------------
// engine
// ---------------------------------------------------------------------------
struct rsrl_engine {
    rsrl_config_t cfg;
    BasisKey key;
    int D = 0, A = 0, AW = 0;
    int WT = 1;  // weight tables: 2 for GreedyGQ (fa_q, fa_td) and A2C (critic Q, policy LFA); table t starts at t * FA (* N for PER_ENV)
    int64_t N = 0, NG = 0, F = 0, FA = 0;
    bool has_trace = false;
    size_t rsz = 4;
    cudaStream_t stream = nullptr;
    // device state
    double* states = nullptr;
    int32_t *actions = nullptr, *ep_steps = nullptr, *n_ep = nullptr, *last_len = nullptr;
    unsigned long long* len_hash = nullptr;
    void *td = nullptr, *W = nullptr, *z = nullptr, *partials = nullptr, *dW = nullptr;
    Counters* counters = nullptr;
    long long* phase_prof = nullptr;  // RSRL_B200_PHASE_PROFILE=1: per-CTA phase cycle counters (development aid)
    double* stage = nullptr;  // f64 staging for import/export
    size_t stage_elems = 0;
    double* init_bounds = nullptr;  // lo[4], hi[4]
    // launch shape of the one-launch-per-step kernels
    int grid = 0, block = 0;
    size_t smem = 0;
    // persistent K-steps-per-launch kernel (persistent.cuh)
    bool persistent = false;
    int pmode = 0;  // MODE template value of the persistent kernel (SHARED / PER_ENV / kModeSharedTrace)
    int pgrid = 0, pblock = 0;
    size_t psmem = 0;
    SyncArgs sync = {nullptr, 1, 1, 1, 4, 0, 0, 0u, 0, 1, 0, 0, 0, nullptr, nullptr};
    size_t stage_bytes = 0;
    uint32_t xepoch = 0;  // exchange epoch: counts batched steps over the engine's life, NOT reset by rsrl_engine_reset
    int pcap = 0;  // slot stride of the CTA reduce buffers
    bool dyn = false;  // (basis, order) without template instantiation: run-time-order kernels (dyn.cuh)
    uint64_t t = 0;
    int64_t launches = 0;
    double epsilon = 0.0;
    // large Fourier bases on the 4-D domains (fourier4.cuh): one env kernel + one dW kernel per batched step
    bool f4 = false;
    F4Args f4args = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int f4_nseg = 0;
    // tcgen05 path (f4tc.cuh): bit 0 = env kernel, bit 1 = dW kernel (RSRL_B200_F4TC, default 3; order 7, f32)
    int f4tc = 0, f4tc_env_grid = 0, f4tc_dw_grid = 0;
    // TileCoding engines (tile.cuh)
    bool tile = false;
    TileArgs targs;
    unsigned long long tile_steps = 0;  // batched steps (fused + handle) since reset: barrier target / table rotation
    size_t tileG_bytes = 0;
    // multi-GPU
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    // in-kernel exchange over peer memory (persistent.cuh hop 3)
    PeerArgs peer;
    uint2* inbox = nullptr;      // this rank's mailbox [2][kMaxRanks][n_clusters][FA * WPV]
    size_t inbox_bytes = 0;
    void* peer_mapped[kMaxRanks] = {nullptr};
    bool peers_attached = false;
};

static size_t wcount(const rsrl_engine* e) { return (size_t)e->FA * (e->cfg.weight_mode == RSRL_PER_ENV ? (size_t)e->N : 1); }  // one table

static int ensure_stage(rsrl_engine* e, size_t elems) {
    if (elems <= e->stage_elems) return RSRL_OK;
    if (e->stage) cudaFree(e->stage);
    e->stage = nullptr; e->stage_elems = 0;
    CU_TRY(cudaMalloc(&e->stage, elems * sizeof(double)));
    e->stage_elems = elems;
    return RSRL_OK;
}

static void choose_launch(rsrl_engine* e) {
    if (e->dyn) {  // dyn.cuh: one warp per CTA, no shared memory
        e->block = 32; e->smem = 0;
    } else if (e->cfg.weight_mode == RSRL_PER_ENV) {
        e->block = 128; e->smem = 0;
    } else if (e->WT == 2) {  // twotable.cuh: W[2][FA] + phi(s) and phi(s') rows + three coefficient planes
        int block = 256;
        for (;; block /= 2) {
            e->smem = ((((size_t)2 * e->FA + 3) & ~(size_t)3) + (size_t)2 * e->F * (block + 1) + (size_t)3 * e->AW * block) * e->rsz;
            if (e->smem <= 200 * 1024 || block == 32) break;
        }
        e->block = block;
    } else {
        const size_t rows = e->has_trace ? (size_t)e->FA : (size_t)e->F;
        int block = 256;
        for (;; block /= 2) {
            const size_t bytes = (((size_t)e->FA + 3) & ~(size_t)3) * e->rsz + rows * (block + 1) * e->rsz + (size_t)(e->has_trace ? 1 : e->AW) * block * e->rsz;
            e->smem = bytes;
            if (bytes <= 200 * 1024 || block == 32) break;
        }
        e->block = block;
    }
    e->grid = (int)((e->N + e->block - 1) / e->block);
}

// Shape of the persistent kernel: one CTA per SM, grouped in thread-block clusters (SHARED: all co-resident, they wait for
// each other every step), one env per thread when the shard fits (state stays in registers), block >= reduce rows.
static bool persistent_shape(rsrl_engine* e, int grid, int cs) {
    auto round32 = [](int64_t x) { return (int)((x + 31) / 32 * 32); };
    const int64_t per_cta = (e->N + grid - 1) / grid;
    const int64_t rows = e->has_trace ? e->FA : e->F;  // traces: z (F*A rows) lives in shared memory for the whole launch
    if (e->has_trace && per_cta > kPersistMaxBlock) return false;   // needs one env per thread
    int block = round32(per_cta);
    if (block < round32(rows)) block = round32(rows);
    if (block < 64) block = 64;
    if (block > kPersistMaxBlock) block = kPersistMaxBlock;
    if (block < rows) return false;
    // persistent.cuh shapes: lpr adjacent lanes own one row in the exchanges between cluster leaders; the CTA reduce is warp-local
    // (its summation order depends on the block size only)
    int lpr = 8;
    while ((int64_t)rows * lpr > block) lpr >>= 1;
    const int cap_static = persist_cap_static((int)rows, (int)e->rsz, e->has_trace);
    const int cap = cap_static ? cap_static : persist_cap(block, (int)e->rsz);
    const size_t ndc = e->has_trace ? 1 : (size_t)e->AW;
    const size_t nvp = (size_t)persist_nvp((int)e->FA, (int)e->rsz);
    const size_t ncl = (size_t)(grid / cs);
    const size_t elems = (size_t)(2 + cs + (!e->sync.fx && ncl > 1 ? ncl : 0)) * nvp + (size_t)e->F * 4 + (size_t)(rows + ndc) * cap + (size_t)(block / 32) * nvp;
    size_t bytes = 16 + elems * e->rsz;
    if (e->sync.fx) bytes += (size_t)4 * ((e->FA + 1) / 2 * 2) * sizeof(long long);  // running sums of the counting exchange
    if (bytes > 220 * 1024) return false;
    // small blocks: two CTAs would fit on one SM (registers and shared memory) and the cluster scheduler may pair them up while other SMs
    // stay empty; asking for more than half of an SM's shared memory keeps it at one CTA per SM.  Not for large blocks: their register
    // spills (~200 B per thread) live in L1, and a larger shared-memory carve-out pushes them out to L2 (measured: +27 % step time).
    if (grid > 1 && block <= 256 && bytes < 116 * 1024) bytes = 116 * 1024;
    e->pcap = cap;
    e->sync.lpr = lpr; e->sync.cap = cap;
    e->sync.cluster_size = cs; e->sync.n_clusters = grid / cs;
    e->pgrid = grid; e->pblock = block; e->psmem = bytes;
    return true;
}

static void choose_persistent(rsrl_engine* e) {
    e->persistent = false;
    e->pmode = e->cfg.weight_mode;
    if (e->dyn) return;                                               // any-order coverage path: per-step kernels
    if (e->has_trace && e->cfg.weight_mode == RSRL_PER_ENV) return;  // per-env W + traces: per-step kernels
    if (e->WT == 2) return;                                           // GreedyGQ / A2C: per-step kernels (twotable.cuh)
    int dev = e->cfg.device, sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (sms <= 0) return;
    if (e->cfg.weight_mode == RSRL_PER_ENV) {
        e->pblock = 128;
        e->pgrid = (int)((e->N + e->pblock - 1) / e->pblock);
        // every env's own W in shared memory (F*A x 128 columns) when at least two CTAs still fit on an SM
        const size_t wbytes = (size_t)e->FA * e->pblock * e->rsz;
        e->sync.pe_smem = wbytes <= 110 * 1024 ? 1 : 0;
        e->psmem = e->sync.pe_smem ? wbytes : 0;
        e->persistent = true;
        return;
    }
    if (e->has_trace) e->pmode = kModeSharedTrace;
    e->sync.debug_skip = getenv("RSRL_B200_DEBUG_SKIP") ? atoi(getenv("RSRL_B200_DEBUG_SKIP")) : 0;
    int g0 = (int)((e->N + 127) / 128);
    if (g0 > sms) g0 = sms;
    // fp32: the counting exchange (persistent.cuh), one L2 hop, no clusters: all SMs take part.  f64 engines: cluster + LL-line exchange.
    e->sync.fx = e->cfg.dtype == RSRL_F32 ? 1 : 0;
    e->sync.ngroups = getenv("RSRL_B200_NGROUPS") ? atoi(getenv("RSRL_B200_NGROUPS")) : kMaxGroups;  // multi-GPU: CTA groups per GPU
    if (e->sync.ngroups < 1 || e->sync.ngroups > kMaxGroups) e->sync.ngroups = kMaxGroups;
    // first poll 200 ns after the reductions were issued (plus the ~150 ns the precompute of the next transition takes): earlier polls
    // only queue in front of the reductions in L2 (measured, profiles/r02_persistent.md)
    e->sync.poll_delay_ns = getenv("RSRL_B200_POLL_DELAY") ? atoi(getenv("RSRL_B200_POLL_DELAY")) : 200;
    e->sync.poll_backoff_ns = getenv("RSRL_B200_POLL_BACKOFF") ? atoi(getenv("RSRL_B200_POLL_BACKOFF")) : 0;
    e->sync.world_backoff_ns = getenv("RSRL_B200_WORLD_BACKOFF") ? atoi(getenv("RSRL_B200_WORLD_BACKOFF")) : 0;
    if (g0 > 255) g0 = 255;  // the arrival count of one step has to fit the low byte of an accumulator word
    // cluster size 4: a B200 (8 GPCs of 16-20 SMs) holds 33 such clusters = 132 CTAs at one CTA per SM, so the 65 536 envs of
    // BASELINE configs[1] still get one thread each (497 per CTA); size 8 only packs 15 clusters = 120 CTAs (547 envs per CTA:
    // two passes per step), size 16 packs 7.  Measured (profiles/r02_persistent.md).  RSRL_B200_CLUSTER overrides.
    int cs = getenv("RSRL_B200_CLUSTER") ? atoi(getenv("RSRL_B200_CLUSTER")) : 4;
    if (cs < 1 || cs > kMaxClusterSize || (cs & (cs - 1))) cs = 4;
    if (e->sync.fx) cs = 1;
    if (g0 == 1) cs = 1;
    int grid = (g0 + cs - 1) / cs * cs;
    if (!persistent_shape(e, grid, cs)) return;
    if (grid > 1) {
        // co-residency: shrink the grid to what the device can hold at once (the envs are re-split over fewer CTAs)
        int maxc = 0;
        StepArgs a;
        memset(&a, 0, sizeof a);
        PeerArgs pe;
        memset(&pe, 0, sizeof pe);
        cudaError_t ce = dispatch_persist(e->key, e->pmode, a, 0, e->sync, pe, grid, e->pblock, e->psmem, e->stream, &maxc);
        if (ce != cudaSuccess) { cudaGetLastError(); return; }  // combination not built / cluster shape not launchable: per-step kernels
        if (maxc < 1) return;
        if (grid / cs > maxc) {
            grid = maxc * cs;
            if (!persistent_shape(e, grid, cs)) return;
            ce = dispatch_persist(e->key, e->pmode, a, 0, e->sync, pe, grid, e->pblock, e->psmem, e->stream, &maxc);
            if (ce != cudaSuccess || grid / cs > maxc) { cudaGetLastError(); return; }
        }
    }
    e->persistent = true;
}

static StepArgs make_args(rsrl_engine* e) {
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.states = e->states; a.actions = e->actions; a.ep_steps = e->ep_steps; a.n_ep = e->n_ep; a.last_len = e->last_len;
    a.len_hash = e->len_hash; a.td = e->td; a.W = e->W; a.z = e->z; a.partials = e->partials; a.counters = e->counters; a.phase_prof = e->phase_prof;
    a.n = e->N; a.env_offset = e->cfg.env_offset; a.t = e->t; a.max_ep = e->cfg.max_episode_steps;
    a.algo = e->cfg.algo; a.trace_rule = e->cfg.trace_rule; a.init_mode = e->cfg.init_mode;
    a.pol = policy_of(e->cfg.policy, e->epsilon, e->cfg.seed);
    const double scale = (e->cfg.weight_mode == RSRL_SHARED && e->cfg.update_scale == RSRL_SCALE_MEAN) ? (double)e->NG : 1.0;
    a.gamma = e->cfg.gamma; a.lr_scaled = e->cfg.lr / scale; a.alpha = e->cfg.alpha; a.inv_scale = 1.0 / scale;
    a.lambda = e->cfg.lambda; a.epsilon = e->epsilon;
    for (int d = 0; d < RSRL_MAX_DIM; ++d) { a.init_lo[d] = e->cfg.init_lo[d]; a.init_hi[d] = e->cfg.init_hi[d]; }
    return a;
}

template <typename R>
static cudaError_t launch_reduce(rsrl_engine* e, int n_blocks, bool to_dw) {
    const int fa = (int)e->FA * e->WT;
    const int threads = 128, blocks = (fa + threads - 1) / threads;
    reduce_partials_kernel<R><<<blocks, threads, 0, e->stream>>>(static_cast<const R*>(e->partials), n_blocks, fa,
                                                                  static_cast<R*>(e->W), to_dw ? static_cast<R*>(e->dW) : nullptr);
    return cudaGetLastError();
}
template <typename R>
static cudaError_t launch_add(rsrl_engine* e) {
    const int fa = (int)e->FA * e->WT;
    const int threads = 128, blocks = (fa + threads - 1) / threads;
    add_kernel<R><<<blocks, threads, 0, e->stream>>>(static_cast<R*>(e->W), static_cast<const R*>(e->dW), fa);
    return cudaGetLastError();
}

// The per-step kernels exchange dW with ncclAllReduce: that needs rsrl_engine_comm_init.  A peer-attached engine
// (rsrl_engine_peer_attach) without a communicator can only run the persistent kernel.
static int need_comm(const rsrl_engine* e) {
    if (e->world > 1 && e->cfg.weight_mode == RSRL_SHARED && !e->comm)
        return fail(RSRL_ECOMM, "this engine is peer-attached without an NCCL communicator: the per-step kernels (rsrl_engine_handle, "
                                "shapes outside the persistent kernel) need rsrl_engine_comm_init");
    return RSRL_OK;
}

// after a SHARED-mode fused launch: partials -> dW (-> allreduce) -> W
static int finish_shared_step(rsrl_engine* e, int n_blocks) {
    const bool xch = e->world > 1 && e->comm != nullptr;
    if (e->f4tc) CU_TRY(launch_f4tc_reduce(e->partials, n_blocks, (int)e->FA, e->W, xch ? e->dW : nullptr, e->stream));
    else CU_TRY(e->cfg.dtype == RSRL_F32 ? launch_reduce<float>(e, n_blocks, xch) : launch_reduce<double>(e, n_blocks, xch));
    e->launches += 1;
    if (xch) {
        ncclResult_t r = g_nccl.AllReduce(e->dW, e->dW, (size_t)e->FA * e->WT, e->cfg.dtype == RSRL_F32 ? ncclFloat32 : ncclFloat64,
                                          ncclSum, e->comm, e->stream);
        if (r != ncclSuccess) return fail(RSRL_ECOMM, std::string("ncclAllReduce: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
        CU_TRY(e->cfg.dtype == RSRL_F32 ? launch_add<float>(e) : launch_add<double>(e));
        e->launches += 1;
    }
    return RSRL_OK;
}

static int f4_step(rsrl_engine* e, const StepArgs& a, bool ext, int64_t n, const int32_t* acts /* device: actions the dW pass reads */) {
    const bool f32 = e->cfg.dtype == RSRL_F32;
    cudaError_t ce;
    if (e->f4tc & 1) {
        const int n_tiles = (int)((n + 127) / 128), n_pairs = (n_tiles + 1) / 2;  // a CTA runs two tiles at a time
        ce = launch_f4tc_env(e->cfg.domain, ext, a, e->f4args, n_tiles, n_pairs < e->f4tc_env_grid ? n_pairs : e->f4tc_env_grid, e->stream);
    } else {
        const int grid = (int)((n + e->block - 1) / e->block);
        ce = (f32 ? launch_f4_env_f32 : launch_f4_env_f64)(e->cfg.domain, e->cfg.basis_order, ext, a, e->f4args, grid, e->block, e->smem, e->stream);
    }
    if (ce == cudaErrorInvalidDeviceFunction) { cudaGetLastError(); return unsupported(&e->cfg); }
    CU_TRY(ce);
    int nseg;
    if (e->f4tc & 2) {
        const int64_t n_sub = (n + 31) / 32;
        nseg = (int)(n_sub < e->f4tc_dw_grid ? n_sub : e->f4tc_dw_grid);
        CU_TRY(launch_f4tc_dw(e->cfg.domain, n, e->f4args.tabs, e->f4args.coef, acts, nseg, e->partials, e->counters, e->phase_prof ? e->phase_prof + (size_t)e->pgrid * 8 : nullptr, e->stream));
    } else {
        nseg = e->f4_nseg;
        const int64_t max_seg = (n + 63) / 64;
        if (nseg > max_seg) nseg = (int)max_seg;
        CU_TRY((f32 ? launch_f4_dw_f32 : launch_f4_dw_f64)(e->cfg.domain, e->cfg.basis_order, n, e->f4args.from_states, e->f4args.coef, acts, nseg, e->partials, e->stream));
    }
    e->launches += e->f4tc ? 4 : 2;
    return finish_shared_step(e, nseg);
}

extern "C" {

int rsrl_version(void) { return RSRL_ABI_VERSION; }
const char* rsrl_last_error(void) { return g_err.c_str(); }

int rsrl_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int rsrl_config_default(rsrl_config_t* c) {
    if (!c) return fail(RSRL_EINVAL, "null config");
    memset(c, 0, sizeof *c);
    c->struct_size = sizeof *c;
    c->domain = RSRL_MOUNTAIN_CAR; c->basis = RSRL_FOURIER; c->basis_order = 5;
    c->n_tilings = 8; c->tiles_per_dim = 8; c->memory_size = 4096;
    c->algo = RSRL_QLEARNING; c->policy = RSRL_GREEDY; c->trace_rule = RSRL_TRACE_REPLACE;
    c->weight_mode = RSRL_SHARED; c->update_scale = RSRL_SCALE_SUM; c->dtype = RSRL_F32; c->init_mode = RSRL_INIT_DEFAULT;
    c->n_envs = 1;
    c->lr = 0.001; c->alpha = 0.01; c->gamma = 0.9; c->lambda = 0.7; c->epsilon = 0.1;
    return RSRL_OK;
}

int rsrl_config_dims(const rsrl_config_t* c, int32_t* dim, int32_t* n_actions, int64_t* n_features) {
    int rc = validate(c);
    if (rc) return rc;
    if (dim) *dim = dom_dim(c->domain);
    if (n_actions) *n_actions = dom_actions(c->domain);
    if (n_features) *n_features = ::n_features(c);
    return RSRL_OK;
}

int rsrl_engine_destroy(rsrl_engine_t* e) {
    if (!e) return RSRL_OK;
    cudaSetDevice(e->cfg.device);
    if (e->phase_prof && e->t > 0) {
        std::vector<long long> h((size_t)e->pgrid * 16);
        cudaMemcpy(h.data(), e->phase_prof, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        const char* names_d[8] = {"dW: inputs -> u, v (regs)", "dW: wait MMA (buffer free)", "dW: split + scalar stores", "dW: fence + __syncthreads",
                                  "dW: MMA issue (thread 0)", "-", "-", "-"};
        const char* names_p[8] = {"env compute (warp 0)", "warp reduce + CTA barrier", "warp partials -> CTA partial", "exchange", "final bar",
                                  "leader: wait members (hop A)", "leader: sum + hops N/B", "-"};
        const char* names_t[8] = {"load + tables", "unit compute (regs)", "wait MMA (mbarrier)", "contract (LDTM + FMA)", "store unit + sync + issue",
                                  "TD + stores + bookkeeping", "issuer: MMA issue (pipe busy)", "issuer: idle (no unit ready)"};
        const char** names = e->f4tc ? names_t : names_p;
        for (int q = 0; q < 8; ++q) {
            double sum = 0, mx = 0, mn = 1e30;
            int cntq = 0;
            for (int b2 = 0; b2 < e->pgrid; ++b2) { double v = (double)h[(size_t)b2 * 8 + q] / (double)e->t; if (v == 0) continue; ++cntq; sum += v; mx = v > mx ? v : mx; mn = v < mn ? v : mn; }
            fprintf(stderr, "[phase] %-28s cycles/step: mean %8.0f  min %8.0f  max %8.0f  (%d CTAs)\n", names[q], cntq ? sum / cntq : 0.0, mn, mx, cntq);
        }
        for (int q = 0; e->f4tc && q < 5; ++q) {
            double sum = 0;
            for (int b2 = 0; b2 < e->pgrid; ++b2) sum += (double)h[(size_t)(e->pgrid + b2) * 8 + q] / (double)e->t;
            fprintf(stderr, "[phase] %-28s cycles/step: mean %8.0f\n", names_d[q], sum / e->pgrid);
        }
        cudaFree(e->phase_prof);
    }
    if (e->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(e->comm);
    for (int r = 0; r < kMaxRanks; ++r) if (e->peer_mapped[r]) cudaIpcCloseMemHandle(e->peer_mapped[r]);
    void* bufs[] = {e->states, e->actions, e->ep_steps, e->n_ep, e->last_len, e->len_hash, e->td, e->W, e->z,
                    e->partials, e->dW, e->counters, e->stage, e->init_bounds, e->sync.stage, e->inbox, e->sync.acc, e->sync.prev, e->targs.G, e->targs.barrier, e->f4args.from_states, e->f4args.coef, e->f4args.tabs, e->f4args.q, e->f4args.aux, e->f4args.next_states};
    for (void* b : bufs) if (b) cudaFree(b);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return RSRL_OK;
}

int rsrl_engine_create(const rsrl_config_t* cfg, rsrl_engine_t** out) {
    if (!out) return fail(RSRL_EINVAL, "null out");
    *out = nullptr;
    int rc = validate(cfg);
    if (rc) return rc;
    if ((rc = need_device())) return rc;
    if (cfg->basis == RSRL_TILE_CODING && (cfg->weight_mode != RSRL_SHARED || algo_has_trace(cfg->algo)))
        return fail(RSRL_EUNSUPPORTED, "TileCoding is built for SHARED weights without eligibility traces");
    CU_TRY(cudaSetDevice(cfg->device));
    rsrl_engine* e = new rsrl_engine();
    memset(&e->targs, 0, sizeof e->targs);
    e->tile = cfg->basis == RSRL_TILE_CODING;
    e->f4 = is_f4(cfg);
    if (e->f4 && (cfg->weight_mode != RSRL_SHARED || algo_has_trace(cfg->algo) || algo_td_pred(cfg->algo))) {
        delete e;
        return fail(RSRL_EUNSUPPORTED, "order-5/7 Fourier bases on 4-D domains are built for SHARED weights, TD control without traces");
    }
    memset(&e->peer, 0, sizeof e->peer);
    e->peer.world = 1;
    e->cfg = *cfg;
    e->key = key_of(cfg);
    e->D = dom_dim(cfg->domain); e->A = dom_actions(cfg->domain); e->AW = e->key.aw;
    e->N = cfg->n_envs; e->NG = cfg->n_envs_global > 0 ? cfg->n_envs_global : cfg->n_envs;
    e->F = n_features(cfg); e->FA = e->F * e->AW;
    e->has_trace = algo_has_trace(cfg->algo);
    e->WT = algo_two_tables(cfg->algo) ? 2 : 1;
    e->rsz = cfg->dtype == RSRL_F32 ? 4 : 8;
    e->dyn = !e->tile && !e->f4 && !has_static(e->key);
    if (e->WT == 2 && (e->tile || e->f4 || e->dyn)) {
        delete e;
        return fail(RSRL_EUNSUPPORTED, "GreedyGQ / A2C are built for the Fourier / Polynomial bases of the register path");
    }
    e->epsilon = cfg->epsilon;
    {
        cudaError_t se = cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking);
        if (se != cudaSuccess) { delete e; return fail(RSRL_ECUDA, std::string("cudaStreamCreateWithFlags: ") + cudaGetErrorString(se)); }
    }
    if (e->f4) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        e->block = 128;
        e->grid = (int)((e->N + e->block - 1) / e->block);
        e->smem = ((size_t)e->F * 4 + (size_t)2 * cfg->basis_order * 2 * e->block) * e->rsz;
        e->f4_nseg = (2 * sms + cfg->basis_order) / (cfg->basis_order + 1);
        const int64_t max_seg = (e->N + 63) / 64;
        if (e->f4_nseg > max_seg) e->f4_nseg = (int)max_seg;
        if (e->f4_nseg < 1) e->f4_nseg = 1;
        if (cfg->dtype == RSRL_F32 && cfg->basis_order == 7) {
            e->f4tc = getenv("RSRL_B200_F4TC") && atoi(getenv("RSRL_B200_F4TC")) == 0 ? 0 : 3;  // 0: CUDA-core path (development aid)
            e->f4tc_env_grid = sms;
            e->f4tc_dw_grid = sms;
        }
    } else if (e->tile) {
        int sms = 0;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, cfg->device);
        e->pgrid = (int)((e->N + 127) / 128);
        if (e->pgrid > sms) e->pgrid = sms;
        const int64_t per_cta = (e->N + e->pgrid - 1) / e->pgrid;
        // dense kernel: up to 1024 threads per CTA (64 registers: the f64 RK4 spills ~700 B to L1, but 32 warps hide its latencies:
        // 61.9 us/step vs 70.5 at 512 threads on cfg3), envs split evenly over the chunks; RSRL_B200_TILE_BLOCK caps the block size
        const int max_block = getenv("RSRL_B200_TILE_BLOCK") ? atoi(getenv("RSRL_B200_TILE_BLOCK")) : 1024;
        const int64_t chunks = (per_cta + max_block - 1) / max_block;
        e->pblock = (int)(((per_cta + chunks - 1) / chunks + 31) / 32 * 32);
        if (e->pblock > max_block) e->pblock = max_block;
        if (e->pblock < 64) e->pblock = 64;
        e->psmem = (size_t)e->FA * e->rsz;
        // dense variant: W + a 64-bit accumulator table per CTA in shared memory (RSRL_B200_TILE_DENSE=0 keeps the RED-atomics kernel)
        const size_t dense_smem = (size_t)e->FA * (e->rsz + sizeof(unsigned long long));
        e->targs.dense = dense_smem <= 200 * 1024 && !(getenv("RSRL_B200_TILE_DENSE") && atoi(getenv("RSRL_B200_TILE_DENSE")) == 0);
        if (e->targs.dense) e->psmem = dense_smem;
        else if (e->pblock > 512) e->pblock = 512;  // tile_persistent_kernel: __launch_bounds__(512, 1)
        e->grid = 1; e->block = 64; e->smem = 0;
    } else {
        choose_launch(e);
        choose_persistent(e);
    }
#define E_TRY(expr)                                                                                  \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            rsrl_engine_destroy(e);                                                                  \
            return fail(e__ == cudaErrorMemoryAllocation ? RSRL_ENOMEM : RSRL_ECUDA,                 \
                        std::string(#expr) + ": " + cudaGetErrorString(e__));                        \
        }                                                                                            \
    } while (0)
    const size_t N = (size_t)e->N;
    E_TRY(cudaMalloc(&e->states, N * e->D * sizeof(double)));
    E_TRY(cudaMalloc(&e->actions, N * sizeof(int32_t)));
    E_TRY(cudaMalloc(&e->ep_steps, N * sizeof(int32_t)));
    E_TRY(cudaMalloc(&e->n_ep, N * sizeof(int32_t)));
    E_TRY(cudaMalloc(&e->last_len, N * sizeof(int32_t)));
    E_TRY(cudaMalloc(&e->len_hash, N * sizeof(unsigned long long)));
    if (cfg->record_td_error) E_TRY(cudaMalloc(&e->td, N * e->rsz));
    E_TRY(cudaMalloc(&e->W, wcount(e) * e->WT * e->rsz));
    if (e->has_trace) E_TRY(cudaMalloc(&e->z, (size_t)e->FA * N * e->rsz));
    if (e->tile) {
        e->targs.tp = tile_params(cfg);
        e->targs.fx_scale = cfg->dtype == RSRL_F32 ? 1099511627776.0 : 17592186044416.0;  // 2^40 / 2^44
        e->tileG_bytes = (size_t)(e->targs.dense ? e->pgrid + 1 : 4) * e->FA * sizeof(unsigned long long);
        E_TRY(cudaMalloc(&e->targs.G, e->tileG_bytes));
        E_TRY(cudaMalloc(&e->targs.barrier, sizeof(unsigned long long)));
    } else if (e->f4) {
        E_TRY(cudaMalloc(&e->f4args.from_states, N * 4 * sizeof(double)));
        E_TRY(cudaMalloc(&e->f4args.coef, N * e->rsz));
        if (e->f4tc) {
            E_TRY(cudaMalloc(&e->f4args.tabs, N * 56 * sizeof(float)));
            E_TRY(cudaMalloc(&e->f4args.q, N * 4 * sizeof(float)));
            E_TRY(cudaMalloc(&e->f4args.aux, N * 4 * sizeof(float)));
            E_TRY(cudaMalloc(&e->f4args.next_states, N * 4 * sizeof(double)));
        }
        E_TRY(cudaMalloc(&e->partials, (size_t)(e->f4_nseg > e->f4tc_dw_grid ? e->f4_nseg : e->f4tc_dw_grid) * e->FA * e->rsz));
        E_TRY(cudaMalloc(&e->dW, (size_t)e->FA * e->rsz));
    } else if (cfg->weight_mode == RSRL_SHARED) {
        E_TRY(cudaMalloc(&e->partials, (size_t)e->grid * e->FA * e->WT * e->rsz));
        E_TRY(cudaMalloc(&e->dW, (size_t)e->FA * e->WT * e->rsz));
    }
    if (e->persistent && cfg->weight_mode == RSRL_SHARED) {
        const size_t rows = e->has_trace ? (size_t)e->FA : (size_t)e->F;       // reduce rows
        const size_t ndc = e->has_trace ? 1 : (size_t)e->AW;                       // values per row
        const size_t nl = rows * (e->cfg.dtype == RSRL_F32 ? 1 : ndc);             // 16-byte LL lines per partial
        e->stage_bytes = (size_t)2 * e->sync.n_clusters * nl * sizeof(uint4);
        E_TRY(cudaMalloc(&e->sync.stage, e->stage_bytes));
        E_TRY(cudaMemset(e->sync.stage, 0, e->stage_bytes));                       // epoch 0 is never published
        e->inbox_bytes = (size_t)2 * kMaxRanks * e->sync.n_clusters * e->FA * (e->rsz / 4) * sizeof(uint2);
        if (e->sync.fx) {  // counting exchange: local table, running sums, and the world table in the peer-visible mailbox
            const size_t tab = (size_t)2 * e->FA * kAccStride * sizeof(unsigned long long);
            E_TRY(cudaMalloc(&e->sync.acc, tab * kMaxGroups));
            E_TRY(cudaMemset(e->sync.acc, 0, tab * kMaxGroups));
            E_TRY(cudaMalloc(&e->sync.prev, (size_t)(2 + 2 * kMaxGroups) * e->FA * sizeof(long long)));
            E_TRY(cudaMemset(e->sync.prev, 0, (size_t)(2 + 2 * kMaxGroups) * e->FA * sizeof(long long)));
            e->inbox_bytes = tab;
        }
        E_TRY(cudaMalloc(&e->inbox, e->inbox_bytes));
        E_TRY(cudaMemset(e->inbox, 0, e->inbox_bytes));
    }
    E_TRY(cudaMalloc(&e->counters, sizeof(Counters)));
    if (getenv("RSRL_B200_PHASE_PROFILE") && e->f4tc) e->pgrid = e->f4tc_env_grid;
    if (getenv("RSRL_B200_PHASE_PROFILE") && (e->persistent || e->f4tc)) {
        E_TRY(cudaMalloc(&e->phase_prof, (size_t)e->pgrid * 16 * sizeof(long long)));   // f4tc: second half = dW kernel
        E_TRY(cudaMemset(e->phase_prof, 0, (size_t)e->pgrid * 16 * sizeof(long long)));
    }
    E_TRY(cudaMalloc(&e->init_bounds, 8 * sizeof(double)));
    // probe that the combination is built (fails loudly instead of at the first step)
    if (!e->tile && !e->f4) {
        StepArgs a = make_args(e);
        a.n = 0;
        cudaError_t pe = e->WT == 2 ? dispatch_two(e->key, cfg->weight_mode, false, a, 1, e->block, e->smem, e->stream)
                                    : dispatch_fused(e->key, cfg->weight_mode, false, a, 1, e->block, e->smem, e->stream);
        if (pe == cudaErrorInvalidDeviceFunction) { cudaGetLastError(); rsrl_engine_destroy(e); return unsupported(cfg); }
        E_TRY(pe);
        E_TRY(cudaStreamSynchronize(e->stream));
    }
#undef E_TRY
    rc = rsrl_engine_reset(e, nullptr);
    if (rc) { rsrl_engine_destroy(e); return rc; }
    *out = e;
    return RSRL_OK;
}

int rsrl_engine_reset(rsrl_engine_t* e, const double* init_states) {
    if (!e) return fail(RSRL_EINVAL, "null engine");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const size_t N = (size_t)e->N;
    cudaStream_t st = e->stream;
    CU_TRY(cudaMemsetAsync(e->W, 0, wcount(e) * e->WT * e->rsz, st));  // LFA::vector => Array2::zeros (examples/q_learning.rs:25)
    if (e->z) CU_TRY(cudaMemsetAsync(e->z, 0, (size_t)e->FA * N * e->rsz, st));
    if (e->td) CU_TRY(cudaMemsetAsync(e->td, 0, N * e->rsz, st));
    CU_TRY(cudaMemsetAsync(e->actions, 0xFF, N * sizeof(int32_t), st));  // -1
    CU_TRY(cudaMemsetAsync(e->ep_steps, 0, N * sizeof(int32_t), st));
    CU_TRY(cudaMemsetAsync(e->n_ep, 0, N * sizeof(int32_t), st));
    CU_TRY(cudaMemsetAsync(e->last_len, 0, N * sizeof(int32_t), st));
    CU_TRY(cudaMemsetAsync(e->len_hash, 0, N * sizeof(unsigned long long), st));
    CU_TRY(cudaMemsetAsync(e->counters, 0, sizeof(Counters), st));
    // the exchange mailboxes and their epoch (e->xepoch) live as long as the engine: resetting them could erase or
    // alias a peer GPU's in-flight words (ranks reset without a barrier between them)
    if (e->tile) {
        CU_TRY(cudaMemsetAsync(e->targs.G, 0, e->tileG_bytes, st));
        CU_TRY(cudaMemsetAsync(e->targs.barrier, 0, sizeof(unsigned long long), st));
        e->tile_steps = 0;
    }
    e->t = 0;
    if (init_states) {
        CU_TRY(cudaMemcpyAsync(e->states, init_states, N * e->D * sizeof(double), cudaMemcpyHostToDevice, st));
    } else {
        double b[8];
        for (int d = 0; d < 4; ++d) { b[d] = e->cfg.init_lo[d]; b[4 + d] = e->cfg.init_hi[d]; }
        CU_TRY(cudaMemcpyAsync(e->init_bounds, b, sizeof b, cudaMemcpyHostToDevice, st));
        const int threads = 128, blocks = (int)((e->N + threads - 1) / threads);
        if (e->cfg.domain == RSRL_MOUNTAIN_CAR)
            init_states_kernel<RSRL_MOUNTAIN_CAR><<<blocks, threads, 0, st>>>(e->N, e->states, e->cfg.init_mode, e->init_bounds, e->init_bounds + 4, e->cfg.seed, e->cfg.env_offset);
        else if (e->cfg.domain == RSRL_CART_POLE)
            init_states_kernel<RSRL_CART_POLE><<<blocks, threads, 0, st>>>(e->N, e->states, e->cfg.init_mode, e->init_bounds, e->init_bounds + 4, e->cfg.seed, e->cfg.env_offset);
        else
            init_states_kernel<RSRL_ACROBOT><<<blocks, threads, 0, st>>>(e->N, e->states, e->cfg.init_mode, e->init_bounds, e->init_bounds + 4, e->cfg.seed, e->cfg.env_offset);
        CU_TRY(cudaGetLastError());
        e->launches += 1;
    }
    CU_TRY(cudaStreamSynchronize(st));
    return RSRL_OK;
}

int rsrl_engine_step(rsrl_engine_t* e, int64_t k_steps) {
    if (!e) return fail(RSRL_EINVAL, "null engine");
    if (k_steps < 0) return fail(RSRL_EINVAL, "k_steps < 0");
    CU_TRY(cudaSetDevice(e->cfg.device));
    if (e->f4) {
        { int rc = need_comm(e); if (rc) return rc; }
        for (int64_t k = 0; k < k_steps; ++k) {
            StepArgs a = make_args(e);
            int rc = f4_step(e, a, false, e->N, e->actions);
            if (rc) return rc;
            e->t += 1;
        }
        return RSRL_OK;
    }
    if (e->tile) {
        if (e->world > 1) return fail(RSRL_EUNSUPPORTED, "TileCoding engines are single-GPU (shard envs with PER-GPU agents instead)");
        while (k_steps > 0) {
            const int k = (int)(k_steps < 65536 ? k_steps : 65536);
            StepArgs a = make_args(e);
            e->targs.barrier_base = e->tile_steps;
            CU_TRY(cudaMemsetAsync(e->targs.barrier, 0, sizeof(unsigned long long), e->stream));  // launch-local barrier targets (both kernels)
            auto fn = e->cfg.dtype == RSRL_F32 ? launch_tile_persist_f32 : launch_tile_persist_f64;
            CU_TRY(fn(e->cfg.domain, e->AW, false, a, k, e->targs, e->pgrid, e->pblock, e->psmem, e->stream));
            e->launches += 1;
            e->t += (uint64_t)k; e->tile_steps += (unsigned long long)k;
            k_steps -= k;
        }
        return RSRL_OK;
    }
    if (e->persistent && (e->world == 1 || e->peers_attached || e->cfg.weight_mode == RSRL_PER_ENV)) {
        // K batched steps per launch; bounded so that one launch stays well under a second
        while (k_steps > 0) {
            const int k = (int)(k_steps < 65536 ? k_steps : 65536);
            StepArgs a = make_args(e);
            e->sync.epoch_base = e->xepoch;
            cudaError_t ce = dispatch_persist(e->key, e->pmode, a, k, e->sync, e->peer, e->pgrid, e->pblock, e->psmem, e->stream);
            if (ce == cudaErrorCooperativeLaunchTooLarge) {  // cannot be co-resident here: per-step kernels instead
                cudaGetLastError();
                e->persistent = false;
                return rsrl_engine_step(e, k_steps);
            }
            CU_TRY(ce);
            e->launches += 1;
            e->t += (uint64_t)k;
            e->xepoch += (uint32_t)k;
            k_steps -= k;
        }
        return RSRL_OK;
    }
    { int rc = need_comm(e); if (rc) return rc; }
    for (int64_t k = 0; k < k_steps; ++k) {
        StepArgs a = make_args(e);
        CU_TRY(e->WT == 2 ? dispatch_two(e->key, e->cfg.weight_mode, false, a, e->grid, e->block, e->smem, e->stream)
                          : dispatch_fused(e->key, e->cfg.weight_mode, false, a, e->grid, e->block, e->smem, e->stream));
        e->launches += 1;
        if (e->cfg.weight_mode == RSRL_SHARED) {
            int rc = finish_shared_step(e, e->grid);
            if (rc) return rc;
        }
        e->t += 1;
    }
    return RSRL_OK;
}

int rsrl_engine_sync(rsrl_engine_t* e) {
    if (!e) return fail(RSRL_EINVAL, "null engine");
    CU_TRY(cudaSetDevice(e->cfg.device));
    CU_TRY(cudaStreamSynchronize(e->stream));
    Counters c;
    CU_TRY(cudaMemcpy(&c, e->counters, sizeof c, cudaMemcpyDeviceToHost));
    if (c.pad) return fail(RSRL_ECUDA, "tensor-core pipeline fault: a tcgen05 completion barrier timed out (f4tc.cuh)");
    if (c.nonfinite) return fail(RSRL_ENONFINITE, "a Q vector had no valid maximum (NaN weights; the reference panics in utils.rs:76), or a weight update "
                                                  "left the range of the fp32 exchange (|per-CTA dW| >= 16384: the run is diverging)");
    return RSRL_OK;
}

void* rsrl_engine_stream(rsrl_engine_t* e) { return e ? (void*)e->stream : nullptr; }

#define GET_SIMPLE(name, field, type)                                                                         \
    int name(rsrl_engine_t* e, type* out) {                                                                   \
        if (!e || !out) return fail(RSRL_EINVAL, "null argument");                                            \
        CU_TRY(cudaSetDevice(e->cfg.device));                                                                 \
        CU_TRY(cudaMemcpyAsync(out, e->field, (size_t)e->N * sizeof(type), cudaMemcpyDeviceToHost, e->stream)); \
        CU_TRY(cudaStreamSynchronize(e->stream));                                                             \
        return RSRL_OK;                                                                                       \
    }
GET_SIMPLE(rsrl_engine_get_actions, actions, int32_t)
GET_SIMPLE(rsrl_engine_get_episode_steps, ep_steps, int32_t)

int rsrl_engine_get_states(rsrl_engine_t* e, double* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->cfg.device));
    CU_TRY(cudaMemcpyAsync(out, e->states, (size_t)e->N * e->D * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

int rsrl_engine_set_states(rsrl_engine_t* e, const double* in) {
    if (!e || !in) return fail(RSRL_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->cfg.device));
    CU_TRY(cudaMemcpyAsync(e->states, in, (size_t)e->N * e->D * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

static int export_tensor(rsrl_engine* e, const void* src, int64_t n_env, int64_t fa, int transposed, double* out) {
    const size_t elems = (size_t)n_env * fa;
    int rc = ensure_stage(e, elems);
    if (rc) return rc;
    const int threads = 256, blocks = (int)((elems + threads - 1) / threads);
    if (e->cfg.dtype == RSRL_F32) export_kernel<float><<<blocks, threads, 0, e->stream>>>(static_cast<const float*>(src), e->stage, n_env, fa, transposed);
    else export_kernel<double><<<blocks, threads, 0, e->stream>>>(static_cast<const double*>(src), e->stage, n_env, fa, transposed);
    CU_TRY(cudaGetLastError());
    e->launches += 1;
    CU_TRY(cudaMemcpyAsync(out, e->stage, elems * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

static int import_tensor(rsrl_engine* e, void* dst, int64_t n_env, int64_t fa, int transposed, const double* in) {
    const size_t elems = (size_t)n_env * fa;
    int rc = ensure_stage(e, elems);
    if (rc) return rc;
    CU_TRY(cudaMemcpyAsync(e->stage, in, elems * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    const int threads = 256, blocks = (int)((elems + threads - 1) / threads);
    if (e->cfg.dtype == RSRL_F32) import_kernel<float><<<blocks, threads, 0, e->stream>>>(e->stage, static_cast<float*>(dst), n_env, fa, transposed);
    else import_kernel<double><<<blocks, threads, 0, e->stream>>>(e->stage, static_cast<double*>(dst), n_env, fa, transposed);
    CU_TRY(cudaGetLastError());
    e->launches += 1;
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

int rsrl_engine_get_weights(rsrl_engine_t* e, double* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    return export_tensor(e, e->W, pe ? e->N : 1, e->FA, pe, out);
}
int rsrl_engine_set_weights(rsrl_engine_t* e, const double* in) {
    if (!e || !in) return fail(RSRL_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    return import_tensor(e, e->W, pe ? e->N : 1, e->FA, pe, in);
}
int rsrl_engine_get_aux_weights(rsrl_engine_t* e, double* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    if (e->WT != 2) return fail(RSRL_EINVAL, "the configured agent has one weight table (aux weights: GreedyGQ fa_td, A2C policy)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    return export_tensor(e, static_cast<const char*>(e->W) + wcount(e) * e->rsz, pe ? e->N : 1, e->FA, pe, out);
}
int rsrl_engine_set_aux_weights(rsrl_engine_t* e, const double* in) {
    if (!e || !in) return fail(RSRL_EINVAL, "null argument");
    if (e->WT != 2) return fail(RSRL_EINVAL, "the configured agent has one weight table (aux weights: GreedyGQ fa_td, A2C policy)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    return import_tensor(e, static_cast<char*>(e->W) + wcount(e) * e->rsz, pe ? e->N : 1, e->FA, pe, in);
}
int rsrl_engine_get_traces(rsrl_engine_t* e, double* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    if (!e->z) return fail(RSRL_EINVAL, "the configured algorithm has no eligibility trace");
    CU_TRY(cudaSetDevice(e->cfg.device));
    return export_tensor(e, e->z, e->N, e->FA, 1, out);
}
int rsrl_engine_set_traces(rsrl_engine_t* e, const double* in) {
    if (!e || !in) return fail(RSRL_EINVAL, "null argument");
    if (!e->z) return fail(RSRL_EINVAL, "the configured algorithm has no eligibility trace");
    CU_TRY(cudaSetDevice(e->cfg.device));
    return import_tensor(e, e->z, e->N, e->FA, 1, in);
}
int rsrl_engine_get_td_errors(rsrl_engine_t* e, double* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    if (!e->td) return fail(RSRL_EINVAL, "engine was created with record_td_error = 0");
    CU_TRY(cudaSetDevice(e->cfg.device));
    return export_tensor(e, e->td, 1, e->N, 0, out);
}

int rsrl_engine_get_stats(rsrl_engine_t* e, rsrl_stats_t* out) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    CU_TRY(cudaSetDevice(e->cfg.device));
    CU_TRY(cudaStreamSynchronize(e->stream));
    Counters c;
    CU_TRY(cudaMemcpy(&c, e->counters, sizeof c, cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof *out);
    out->total_steps = (int64_t)e->t * e->N;
    out->total_episodes = (int64_t)c.episodes;
    out->terminal_episodes = (int64_t)c.terminal_episodes;
    out->batch_steps = (int64_t)e->t;
    out->kernel_launches = e->launches;
    out->nonfinite = c.nonfinite;
    return RSRL_OK;
}

int rsrl_engine_get_env_stats(rsrl_engine_t* e, int32_t* n_episodes, int32_t* last_len, uint64_t* len_hash) {
    if (!e) return fail(RSRL_EINVAL, "null engine");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const size_t N = (size_t)e->N;
    if (n_episodes) CU_TRY(cudaMemcpyAsync(n_episodes, e->n_ep, N * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (last_len) CU_TRY(cudaMemcpyAsync(last_len, e->last_len, N * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (len_hash) CU_TRY(cudaMemcpyAsync(len_hash, e->len_hash, N * sizeof(uint64_t), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

int rsrl_engine_set_epsilon(rsrl_engine_t* e, double epsilon) {
    if (!e) return fail(RSRL_EINVAL, "null engine");
    if (e->cfg.policy == RSRL_SOFTMAX ? !(epsilon <= -1e-7 || epsilon >= 1e-7) : !(epsilon >= 0.0))
        return fail(RSRL_EINVAL, "epsilon must be >= 0 (Softmax: tau must be non-zero)");
    e->epsilon = epsilon;
    return RSRL_OK;
}

// ---- trait-level entry points on the engine's weights ----
static int engine_eval(rsrl_engine* e, int mode, int64_t n, const double* states, uint64_t draw, double* q_out, int32_t* act_out) {
    if (!e || !states || n <= 0) return fail(RSRL_EINVAL, "bad argument");
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    if (pe && n != e->N) return fail(RSRL_EINVAL, "PER_ENV weights: n must equal n_envs (state i is evaluated with agent i)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    DevBuf ds, dout;
    CU_TRY(ds.alloc((size_t)n * e->D * sizeof(double)));
    CU_TRY(dout.alloc(mode == 1 ? (size_t)n * e->AW * sizeof(double) : (size_t)n * sizeof(int32_t)));
    CU_TRY(cudaMemcpyAsync(ds.p, states, (size_t)n * e->D * sizeof(double), cudaMemcpyHostToDevice, e->stream));
    EvalArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.mode = mode; ea.n = n; ea.states = ds.as<double>(); ea.w_env_stride = pe ? 1 : 0;
    // Policy::sample / mode of the A2C agent go through the Gibbs policy's own table; evaluate() is the critic's Q
    ea.W = (e->cfg.algo == RSRL_A2C && mode != 1) ? static_cast<const char*>(e->W) + wcount(e) * e->rsz : e->W;
    ea.out = mode == 1 ? dout.as<double>() : nullptr; ea.act_out = mode == 1 ? nullptr : dout.as<int32_t>();
    ea.pol = policy_of(e->cfg.policy, e->epsilon, e->cfg.seed); ea.draw = draw; ea.env_offset = e->cfg.env_offset; ea.counters = e->counters;
    if (e->tile) CU_TRY((e->cfg.dtype == RSRL_F32 ? launch_tile_eval_f32 : launch_tile_eval_f64)(e->cfg.domain, e->AW, ea, e->targs.tp, e->stream));
    else if (e->f4) CU_TRY((e->cfg.dtype == RSRL_F32 ? launch_f4_eval_f32 : launch_f4_eval_f64)(e->cfg.domain, e->cfg.basis_order, ea, e->stream));
    else CU_TRY(dispatch_eval(e->key, ea, e->stream));
    e->launches += 1;
    if (mode == 1) CU_TRY(cudaMemcpyAsync(q_out, dout.p, (size_t)n * e->AW * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
    else CU_TRY(cudaMemcpyAsync(act_out, dout.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(cudaStreamSynchronize(e->stream));
    return RSRL_OK;
}

int rsrl_engine_evaluate(rsrl_engine_t* e, int64_t n, const double* states, double* q_out) {
    if (!q_out) return fail(RSRL_EINVAL, "null q_out");
    return engine_eval(e, 1, n, states, 0, q_out, nullptr);
}
int rsrl_engine_sample(rsrl_engine_t* e, int64_t n, const double* states, uint64_t draw, int32_t* actions_out) {
    if (!actions_out) return fail(RSRL_EINVAL, "null actions_out");
    if (e && algo_td_pred(e->cfg.algo)) return fail(RSRL_EINVAL, "TD prediction engines have no Q-based policy");
    return engine_eval(e, 2, n, states, draw, nullptr, actions_out);
}
int rsrl_engine_mode(rsrl_engine_t* e, int64_t n, const double* states, int32_t* actions_out) {
    if (!actions_out) return fail(RSRL_EINVAL, "null actions_out");
    if (e && algo_td_pred(e->cfg.algo)) return fail(RSRL_EINVAL, "TD prediction engines have no Q-based policy");
    return engine_eval(e, 3, n, states, 0, nullptr, actions_out);
}

int rsrl_engine_handle(rsrl_engine_t* e, int64_t n, const double* from_states, const int32_t* actions, const double* rewards,
                       const double* to_states, const uint8_t* terminal, uint64_t draw, double* td_out) {
    if (!e || !from_states || !actions || !rewards || !to_states || !terminal || n <= 0) return fail(RSRL_EINVAL, "bad argument");
    if ((e->has_trace || e->cfg.weight_mode == RSRL_PER_ENV) && n != e->N)
        return fail(RSRL_EINVAL, "per-env traces / weights: n must equal n_envs (transition i belongs to agent i)");
    if (n > e->N) return fail(RSRL_EINVAL, "n exceeds n_envs");
    { int rc = need_comm(e); if (rc) return rc; }
    for (int64_t i = 0; i < n; ++i)
        if (actions[i] < 0 || actions[i] >= e->A) return fail(RSRL_EINVAL, "action out of range");
    CU_TRY(cudaSetDevice(e->cfg.device));
    DevBuf dfrom, dto, dact, drew, dterm, dtd;
    const size_t sb = (size_t)n * e->D * sizeof(double);
    CU_TRY(dfrom.alloc(sb)); CU_TRY(dto.alloc(sb)); CU_TRY(dact.alloc(n * sizeof(int32_t)));
    CU_TRY(drew.alloc(n * sizeof(double))); CU_TRY(dterm.alloc(n)); CU_TRY(dtd.alloc(n * e->rsz));
    cudaStream_t st = e->stream;
    CU_TRY(cudaMemcpyAsync(dfrom.p, from_states, sb, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(dto.p, to_states, sb, cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(dact.p, actions, n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(drew.p, rewards, n * sizeof(double), cudaMemcpyHostToDevice, st));
    CU_TRY(cudaMemcpyAsync(dterm.p, terminal, n, cudaMemcpyHostToDevice, st));
    StepArgs a = make_args(e);
    a.n = n; a.t = draw; a.td = dtd.p;
    a.ext_from = dfrom.as<double>(); a.ext_to = dto.as<double>(); a.ext_actions = dact.as<int32_t>();
    a.ext_rewards = drew.as<double>(); a.ext_term = dterm.as<uint8_t>();
    if (e->f4) {
        int rc = f4_step(e, a, true, n, dact.as<int32_t>());  // the dW pass reads the caller's actions (the engine's own stay untouched)
        if (rc) return rc;
    } else if (e->tile) {
        e->targs.barrier_base = e->tile_steps;
        CU_TRY(cudaMemsetAsync(e->targs.barrier, 0, sizeof(unsigned long long), st));
        int g2 = (int)((n + 127) / 128);
        if (g2 > e->pgrid) g2 = e->pgrid;
        auto fn = e->cfg.dtype == RSRL_F32 ? launch_tile_persist_f32 : launch_tile_persist_f64;
        CU_TRY(fn(e->cfg.domain, e->AW, true, a, 1, e->targs, g2, e->pblock, e->psmem, st));
        e->launches += 1;
        e->tile_steps += 1;
    } else {
    const int grid = (int)((n + e->block - 1) / e->block);
    CU_TRY(e->WT == 2 ? dispatch_two(e->key, e->cfg.weight_mode, true, a, grid, e->block, e->smem, st)
                      : dispatch_fused(e->key, e->cfg.weight_mode, true, a, grid, e->block, e->smem, st));
    e->launches += 1;
    if (e->cfg.weight_mode == RSRL_SHARED) {
        int rc = finish_shared_step(e, grid);
        if (rc) return rc;
    }
    }
    if (td_out) {
        int rc = ensure_stage(e, (size_t)n);
        if (rc) return rc;
        const int threads = 256, blocks = (int)((n + threads - 1) / threads);
        if (e->cfg.dtype == RSRL_F32) export_kernel<float><<<blocks, threads, 0, st>>>(dtd.as<float>(), e->stage, 1, n, 0);
        else export_kernel<double><<<blocks, threads, 0, st>>>(dtd.as<double>(), e->stage, 1, n, 0);
        CU_TRY(cudaGetLastError());
        e->launches += 1;
        CU_TRY(cudaMemcpyAsync(td_out, e->stage, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(cudaStreamSynchronize(st));
    return RSRL_OK;
}

// ---- Domain::rollout ----
int rsrl_engine_rollout(rsrl_engine_t* e, int64_t n, const double* init_states, int64_t step_limit, int32_t greedy, uint64_t draw,
                        double* start_out, double* next_out, int32_t* actions_out, double* rewards_out, uint8_t* terminal_out,
                        int32_t* len_out) {
    if (!e || n <= 0 || step_limit < 1 || !start_out || !next_out || !actions_out || !rewards_out || !terminal_out || !len_out)
        return fail(RSRL_EINVAL, "bad argument (step_limit >= 1)");
    if (e->tile || e->f4) return fail(RSRL_EUNSUPPORTED, "rollout is built for the Fourier / Polynomial bases of the register path");
    if (algo_td_pred(e->cfg.algo)) return fail(RSRL_EINVAL, "TD prediction engines have no Q-based policy");
    const bool pe = e->cfg.weight_mode == RSRL_PER_ENV;
    if (pe && n != e->N) return fail(RSRL_EINVAL, "PER_ENV weights: n must equal n_envs (rollout i follows agent i)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    const int64_t t_max = step_limit - 1 > 1 ? step_limit - 1 : 1;
    DevBuf dinit, dstart, dnext, dact, drew, dterm, dlen;
    const size_t sb = (size_t)n * e->D * sizeof(double), rows = (size_t)n * t_max;
    if (init_states) { CU_TRY(dinit.alloc(sb)); CU_TRY(cudaMemcpyAsync(dinit.p, init_states, sb, cudaMemcpyHostToDevice, e->stream)); }
    CU_TRY(dstart.alloc(sb)); CU_TRY(dnext.alloc(rows * e->D * sizeof(double))); CU_TRY(dact.alloc(rows * sizeof(int32_t)));
    CU_TRY(drew.alloc(rows * sizeof(double))); CU_TRY(dterm.alloc(rows)); CU_TRY(dlen.alloc(n * sizeof(int32_t)));
    RolloutArgs ra;
    memset(&ra, 0, sizeof ra);
    ra.n = n; ra.env_offset = e->cfg.env_offset; ra.t_max = t_max; ra.take = step_limit - 1;
    ra.init = init_states ? dinit.as<double>() : nullptr;
    ra.W = e->cfg.algo == RSRL_A2C ? static_cast<const char*>(e->W) + wcount(e) * e->rsz : e->W;
    ra.w_env_stride = pe ? 1 : 0;
    ra.greedy = greedy != 0; ra.init_mode = e->cfg.init_mode; ra.draw = draw;
    ra.pol = policy_of(e->cfg.policy, e->epsilon, e->cfg.seed);
    for (int d = 0; d < RSRL_MAX_DIM; ++d) { ra.init_lo[d] = e->cfg.init_lo[d]; ra.init_hi[d] = e->cfg.init_hi[d]; }
    ra.start_out = dstart.as<double>(); ra.next_out = dnext.as<double>(); ra.actions_out = dact.as<int32_t>();
    ra.rewards_out = drew.as<double>(); ra.terminal_out = dterm.as<uint8_t>(); ra.len_out = dlen.as<int32_t>();
    ra.counters = e->counters;
    cudaError_t ce = dispatch_rollout(e->key, ra, e->stream);
    if (ce == cudaErrorInvalidDeviceFunction) { cudaGetLastError(); return unsupported(&e->cfg); }
    CU_TRY(ce);
    e->launches += 1;
    cudaStream_t st = e->stream;
    CU_TRY(cudaMemcpyAsync(start_out, dstart.p, sb, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(next_out, dnext.p, rows * e->D * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(actions_out, dact.p, rows * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(rewards_out, drew.p, rows * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(terminal_out, dterm.p, rows, cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaMemcpyAsync(len_out, dlen.p, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    CU_TRY(cudaStreamSynchronize(st));
    return RSRL_OK;
}

// ---- introspection used by the parity tests ----
int rsrl_engine_get_launch_shape(rsrl_engine_t* e, int32_t out[24]) {
    if (!e || !out) return fail(RSRL_EINVAL, "null argument");
    memset(out, 0, 24 * sizeof(int32_t));
    out[0] = e->persistent ? 1 : 0; out[1] = e->pmode; out[2] = e->pgrid; out[3] = e->sync.cluster_size; out[4] = e->sync.n_clusters;
    out[5] = e->pblock; out[6] = e->sync.lpr; out[7] = 0; out[8] = 0;  /* (the CTA reduce is warp-local: its order depends on the block size only) */ out[9] = e->sync.pe_smem;
    out[10] = e->world; out[11] = e->rank; out[12] = e->peers_attached ? 1 : 0; out[13] = (int32_t)e->psmem;
    out[14] = e->tile ? 1 : 0; out[15] = e->f4 ? (1 + e->f4tc) : 0; out[16] = e->sync.fx;
    return RSRL_OK;
}

static __global__ void math_probe_kernel(int fn, int64_t n, const double* __restrict__ x, double* __restrict__ out) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s, c;
    float sf, cf;
    switch (fn) {
        case 0: out[i] = cos64(x[i]); break;
        case 1: sincos64(x[i], &s, &c); out[i] = s; break;
        case 2: sincospi32((float)x[i], &sf, &cf); out[i] = (double)sf; break;
        case 3: sincospi32((float)x[i], &sf, &cf); out[i] = (double)cf; break;
        default: out[i] = (double)exp32((float)x[i]); break;
    }
}

int rsrl_math_probe(int32_t fn, int64_t n, const double* x, double* out) {
    if (fn < 0 || fn > 4 || n <= 0 || !x || !out) return fail(RSRL_EINVAL, "bad argument");
    int rc = need_device();
    if (rc) return rc;
    DevBuf dx, dout;
    CU_TRY(dx.alloc(n * sizeof(double))); CU_TRY(dout.alloc(n * sizeof(double)));
    CU_TRY(cudaMemcpy(dx.p, x, n * sizeof(double), cudaMemcpyHostToDevice));
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    math_probe_kernel<<<blocks, threads>>>(fn, n, dx.as<double>(), dout.as<double>());
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(out, dout.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return RSRL_OK;
}

// ---- multi-GPU ----
int rsrl_comm_unique_id(uint8_t out[128]) {
    if (!out) return fail(RSRL_EINVAL, "null out");
    int rc = load_nccl();
    if (rc) return rc;
    ncclUniqueId id;
    static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
    ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != ncclSuccess) return fail(RSRL_ECOMM, "ncclGetUniqueId failed");
    memcpy(out, &id, 128);
    return RSRL_OK;
}

int rsrl_engine_comm_init(rsrl_engine_t* e, const uint8_t id_bytes[128], int rank, int world) {
    if (!e || !id_bytes || world < 1 || rank < 0 || rank >= world) return fail(RSRL_EINVAL, "bad argument");
    if (world == 1) { e->rank = 0; e->world = 1; return RSRL_OK; }
    int rc = load_nccl();
    if (rc) return rc;
    CU_TRY(cudaSetDevice(e->cfg.device));
    ncclUniqueId id;
    memcpy(&id, id_bytes, 128);
    ncclResult_t r = g_nccl.CommInitRank(&e->comm, world, id, rank);
    if (r != ncclSuccess) return fail(RSRL_ECOMM, std::string("ncclCommInitRank: ") + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?"));
    e->rank = rank; e->world = world;
    return RSRL_OK;
}

int rsrl_engine_peer_export(rsrl_engine_t* e, uint8_t handle_out[64]) {
    if (!e || !handle_out) return fail(RSRL_EINVAL, "null argument");
    if (!e->inbox) return fail(RSRL_EINVAL, "engine has no peer mailbox (SHARED weights + persistent kernel only)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    cudaIpcMemHandle_t h;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    CU_TRY(cudaIpcGetMemHandle(&h, e->inbox));
    memcpy(handle_out, &h, 64);
    return RSRL_OK;
}

int rsrl_engine_peer_attach(rsrl_engine_t* e, const uint8_t* handles, int rank, int world) {
    if (!e || !handles || world < 1 || world > kMaxRanks || rank < 0 || rank >= world) return fail(RSRL_EINVAL, "bad argument (world <= 8)");
    if (e->xepoch != 0) return fail(RSRL_EINVAL, "attach the peers before the first step (the exchange tables count arrivals from step 0)");
    if (!e->inbox) return fail(RSRL_EINVAL, "engine has no peer mailbox (SHARED weights + persistent kernel only)");
    CU_TRY(cudaSetDevice(e->cfg.device));
    for (int r = 0; r < world; ++r) {
        if (r == rank) { e->peer.inbox[r] = e->inbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)r * 64, 64);
        void* p = nullptr;
        cudaError_t ce = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (ce != cudaSuccess) { cudaGetLastError(); return fail(RSRL_ECOMM, std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(ce)); }
        e->peer_mapped[r] = p;
        e->peer.inbox[r] = static_cast<uint2*>(p);
    }
    e->peer.rank = rank; e->peer.world = world;
    e->rank = rank; e->world = world;
    if (!getenv("RSRL_B200_NGROUPS")) {  // CTA groups per GPU: world x groups arrivals per world-table word (measured: 8 groups at 2 GPUs, 4 at 8)
        int ng = 32 / world;
        e->sync.ngroups = ng > kMaxGroups ? kMaxGroups : ng < 2 ? 2 : ng;
    }
    e->peers_attached = true;
    return RSRL_OK;
}

// ---- stateless component entry points ----
int rsrl_domain_info(int32_t domain, int32_t* dim, int32_t* n_actions, double* lo, double* hi, double* start) {
    if (domain < 0 || domain > 2) return fail(RSRL_EINVAL, "unknown domain");
    const int D = dom_dim(domain);
    if (dim) *dim = D;
    if (n_actions) *n_actions = dom_actions(domain);
    for (int d = 0; d < D; ++d) {
        double l, h, s;
        if (domain == RSRL_MOUNTAIN_CAR) { l = Domain<RSRL_MOUNTAIN_CAR>::lo(d); h = Domain<RSRL_MOUNTAIN_CAR>::hi(d); s = Domain<RSRL_MOUNTAIN_CAR>::start(d); }
        else if (domain == RSRL_CART_POLE) { l = Domain<RSRL_CART_POLE>::lo(d); h = Domain<RSRL_CART_POLE>::hi(d); s = Domain<RSRL_CART_POLE>::start(d); }
        else { l = Domain<RSRL_ACROBOT>::lo(d); h = Domain<RSRL_ACROBOT>::hi(d); s = Domain<RSRL_ACROBOT>::start(d); }
        if (lo) lo[d] = l;
        if (hi) hi[d] = h;
        if (start) start[d] = s;
    }
    return RSRL_OK;
}

static int domain_call(int32_t domain, int64_t n, double* states_inout, const double* states_in, const int32_t* actions,
                       double* rewards_out, uint8_t* terminal_out) {
    if (domain < 0 || domain > 2 || n <= 0 || !terminal_out) return fail(RSRL_EINVAL, "bad argument");
    int rc = need_device();
    if (rc) return rc;
    const int D = dom_dim(domain);
    if (actions) for (int64_t i = 0; i < n; ++i) if (actions[i] < 0 || actions[i] >= dom_actions(domain)) return fail(RSRL_EINVAL, "action out of range");
    DevBuf ds, da, dr, dt;
    const size_t sb = (size_t)n * D * sizeof(double);
    CU_TRY(ds.alloc(sb)); CU_TRY(da.alloc(n * sizeof(int32_t))); CU_TRY(dr.alloc(n * sizeof(double))); CU_TRY(dt.alloc(n));
    CU_TRY(cudaMemcpy(ds.p, actions ? states_inout : states_in, sb, cudaMemcpyHostToDevice));
    if (actions) CU_TRY(cudaMemcpy(da.p, actions, n * sizeof(int32_t), cudaMemcpyHostToDevice));
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    const int32_t* dact = actions ? da.as<int32_t>() : nullptr;
    if (domain == RSRL_MOUNTAIN_CAR) domain_step_kernel<RSRL_MOUNTAIN_CAR><<<blocks, threads>>>(n, ds.as<double>(), dact, dr.as<double>(), dt.as<uint8_t>());
    else if (domain == RSRL_CART_POLE) domain_step_kernel<RSRL_CART_POLE><<<blocks, threads>>>(n, ds.as<double>(), dact, dr.as<double>(), dt.as<uint8_t>());
    else domain_step_kernel<RSRL_ACROBOT><<<blocks, threads>>>(n, ds.as<double>(), dact, dr.as<double>(), dt.as<uint8_t>());
    CU_TRY(cudaGetLastError());
    if (actions) {
        CU_TRY(cudaMemcpy(states_inout, ds.p, sb, cudaMemcpyDeviceToHost));
        CU_TRY(cudaMemcpy(rewards_out, dr.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    }
    CU_TRY(cudaMemcpy(terminal_out, dt.p, n, cudaMemcpyDeviceToHost));
    return RSRL_OK;
}

int rsrl_domain_step(int32_t domain, int64_t n, double* states_inout, const int32_t* actions, double* rewards_out, uint8_t* terminal_out) {
    if (!states_inout || !actions || !rewards_out) return fail(RSRL_EINVAL, "null argument");
    return domain_call(domain, n, states_inout, nullptr, actions, rewards_out, terminal_out);
}
int rsrl_domain_is_terminal(int32_t domain, int64_t n, const double* states, uint8_t* terminal_out) {
    if (!states) return fail(RSRL_EINVAL, "null argument");
    return domain_call(domain, n, nullptr, states, nullptr, nullptr, terminal_out);
}

static int stateless_eval(const rsrl_config_t* cfg, int mode, int64_t n, const double* states, const double* weights, int aw, double* out) {
    int rc = validate(cfg);
    if (rc) return rc;
    if (n <= 0 || !states || !out) return fail(RSRL_EINVAL, "bad argument");
    if ((rc = need_device())) return rc;
    BasisKey k = key_of(cfg);
    k.aw = aw;
    const int D = dom_dim(cfg->domain);
    const int64_t F = n_features(cfg);
    const size_t rsz = cfg->dtype == RSRL_F32 ? 4 : 8;
    const size_t out_elems = (size_t)n * (mode == 0 ? F : aw);
    DevBuf ds, dw, dwr, dout;
    CU_TRY(ds.alloc((size_t)n * D * sizeof(double))); CU_TRY(dout.alloc(out_elems * sizeof(double)));
    CU_TRY(cudaMemcpy(ds.p, states, (size_t)n * D * sizeof(double), cudaMemcpyHostToDevice));
    if (mode == 1) {
        CU_TRY(dw.alloc((size_t)F * aw * sizeof(double))); CU_TRY(dwr.alloc((size_t)F * aw * rsz));
        CU_TRY(cudaMemcpy(dw.p, weights, (size_t)F * aw * sizeof(double), cudaMemcpyHostToDevice));
        const int threads = 256, blocks = (int)((F * aw + threads - 1) / threads);
        if (cfg->dtype == RSRL_F32) import_kernel<float><<<blocks, threads>>>(dw.as<double>(), dwr.as<float>(), 1, F * aw, 0);
        else import_kernel<double><<<blocks, threads>>>(dw.as<double>(), dwr.as<double>(), 1, F * aw, 0);
        CU_TRY(cudaGetLastError());
    }
    EvalArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.mode = mode; ea.n = n; ea.states = ds.as<double>(); ea.W = dwr.p; ea.out = dout.as<double>();
    if (mode == 0 && cfg->basis == RSRL_TILE_CODING) CU_TRY(cudaMemset(dout.p, 0, out_elems * sizeof(double)));
    cudaError_t ce = is_f4(cfg) ? (cfg->dtype == RSRL_F32 ? launch_f4_eval_f32 : launch_f4_eval_f64)(cfg->domain, cfg->basis_order, ea, 0)
                     : cfg->basis == RSRL_TILE_CODING
                         ? (cfg->dtype == RSRL_F32 ? launch_tile_eval_f32 : launch_tile_eval_f64)(cfg->domain, aw, ea, tile_params(cfg), 0)
                         : dispatch_eval(k, ea, 0);
    if (ce == cudaErrorInvalidDeviceFunction) { cudaGetLastError(); return unsupported(cfg); }
    CU_TRY(ce);
    CU_TRY(cudaMemcpy(out, dout.p, out_elems * sizeof(double), cudaMemcpyDeviceToHost));
    return RSRL_OK;
}

int rsrl_basis_project(const rsrl_config_t* cfg, int64_t n, const double* states, double* features_out) {
    return stateless_eval(cfg, 0, n, states, nullptr, cfg ? dom_actions(cfg->domain) : 0, features_out);
}

int rsrl_lfa_evaluate(const rsrl_config_t* cfg, int64_t n, const double* states, const double* weights, double* q_out) {
    if (!weights) return fail(RSRL_EINVAL, "null weights");
    return stateless_eval(cfg, 1, n, states, weights, cfg ? key_of(cfg).aw : 0, q_out);
}

int rsrl_lfa_update_index(const rsrl_config_t* cfg, int64_t n, const double* states, const int32_t* actions, const double* errors, double* weights_inout) {
    // W[:, a_i] += (lr * err_i) * phi(s_i), summed over the batch: a SHARED/SUM Q-learning engine fed
    // terminal transitions with reward = err and W = 0 for the TD part reproduces exactly that update.
    int rc = validate(cfg);
    if (rc) return rc;
    if (n <= 0 || !states || !actions || !errors || !weights_inout) return fail(RSRL_EINVAL, "bad argument");
    rsrl_config_t c = *cfg;
    c.algo = RSRL_QLEARNING; c.weight_mode = RSRL_SHARED; c.update_scale = RSRL_SCALE_SUM; c.n_envs = n; c.n_envs_global = n;
    c.record_td_error = 0; c.env_offset = 0;
    rsrl_engine_t* e = nullptr;
    if ((rc = rsrl_engine_create(&c, &e))) return rc;
    // zero weights => qsa = 0 => delta = reward = err_i
    std::vector<uint8_t> term((size_t)n, 1);
    rc = rsrl_engine_handle(e, n, states, actions, errors, states, term.data(), 0, nullptr);
    std::vector<double> dw((size_t)e->FA);
    if (!rc) rc = rsrl_engine_get_weights(e, dw.data());
    if (!rc) for (int64_t j = 0; j < e->FA; ++j) weights_inout[j] += dw[(size_t)j];
    rsrl_engine_destroy(e);
    return rc;
}

}  // extern "C"

template <int A>
static cudaError_t run_policy(int dtype, int mode, int64_t n, const double* dq, PolicyParams pol, double eps, uint64_t draw,
                              int64_t env_offset, int32_t* act, double* probs, Counters* cnt) {
    const int threads = 128, blocks = (int)((n + threads - 1) / threads);
    if (dtype == RSRL_F32) policy_kernel<float, A><<<blocks, threads>>>(mode, n, dq, pol, eps, draw, env_offset, act, probs, cnt);
    else policy_kernel<double, A><<<blocks, threads>>>(mode, n, dq, pol, eps, draw, env_offset, act, probs, cnt);
    return cudaGetLastError();
}

extern "C" {

static int policy_call(int mode, int32_t policy, double eps, uint64_t seed, uint64_t draw, int64_t env_offset, int64_t n,
                       int32_t A, const double* q, int32_t* act_out, double* probs_out) {
    if (n <= 0 || !q || A < 1 || A > 8) return fail(RSRL_EINVAL, "bad argument (1 <= n_actions <= 8)");
    if (policy < 0 || policy > RSRL_SOFTMAX || (policy == RSRL_SOFTMAX ? !(eps <= -1e-7 || eps >= 1e-7) : !(eps >= 0.0)))
        return fail(RSRL_EINVAL, "bad policy / epsilon (Softmax: tau must be non-zero)");
    int rc = need_device();
    if (rc) return rc;
    DevBuf dq, dout, dc;
    CU_TRY(dq.alloc((size_t)n * A * sizeof(double)));
    CU_TRY(dout.alloc(mode == 1 ? (size_t)n * A * sizeof(double) : (size_t)n * sizeof(int32_t)));
    CU_TRY(dc.alloc(sizeof(Counters)));
    CU_TRY(cudaMemset(dc.p, 0, sizeof(Counters)));
    CU_TRY(cudaMemcpy(dq.p, q, (size_t)n * A * sizeof(double), cudaMemcpyHostToDevice));
    PolicyParams pol = policy_of(policy, eps, seed);
    int32_t* da = mode == 1 ? nullptr : dout.as<int32_t>();
    double* dp = mode == 1 ? dout.as<double>() : nullptr;
    // explicit Q vectors are compared in f64: these entry points mirror the reference's f64 MockQ tests
    cudaError_t ce;
    switch (A) {
        case 1: ce = run_policy<1>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 2: ce = run_policy<2>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 3: ce = run_policy<3>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 4: ce = run_policy<4>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 5: ce = run_policy<5>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 6: ce = run_policy<6>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        case 7: ce = run_policy<7>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
        default: ce = run_policy<8>(RSRL_F64, mode, n, dq.as<double>(), pol, eps, draw, env_offset, da, dp, dc.as<Counters>()); break;
    }
    CU_TRY(ce);
    if (mode == 1) CU_TRY(cudaMemcpy(probs_out, dout.p, (size_t)n * A * sizeof(double), cudaMemcpyDeviceToHost));
    else CU_TRY(cudaMemcpy(act_out, dout.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
    Counters c;
    CU_TRY(cudaMemcpy(&c, dc.p, sizeof c, cudaMemcpyDeviceToHost));
    if (c.pad) return fail(RSRL_ECUDA, "tensor-core pipeline fault: a tcgen05 completion barrier timed out (f4tc.cuh)");
    if (c.nonfinite) return fail(RSRL_ENONFINITE, "a Q vector had no valid maximum; the reference panics in utils.rs:76");
    return RSRL_OK;
}

int rsrl_policy_sample(int32_t policy, double epsilon, uint64_t seed, uint64_t draw, int64_t env_offset, int64_t n,
                       int32_t n_actions, const double* q, int32_t* actions_out) {
    if (!actions_out) return fail(RSRL_EINVAL, "null actions_out");
    return policy_call(0, policy, epsilon, seed, draw, env_offset, n, n_actions, q, actions_out, nullptr);
}
int rsrl_policy_probs(int32_t policy, double epsilon, int64_t n, int32_t n_actions, const double* q, double* probs_out) {
    if (!probs_out) return fail(RSRL_EINVAL, "null probs_out");
    return policy_call(1, policy, epsilon, 0, 0, 0, n, n_actions, q, nullptr, probs_out);
}
int rsrl_policy_mode(int64_t n, int32_t n_actions, const double* q, int32_t* actions_out) {
    if (!actions_out) return fail(RSRL_EINVAL, "null actions_out");
    return policy_call(2, RSRL_GREEDY, 0.0, 0, 0, 0, n, n_actions, q, actions_out, nullptr);
}

int rsrl_trace_update(int32_t rule, double gamma, double lambda, double alpha, int64_t n, double* z_inout, const double* grad) {
    if (rule < 0 || rule > 2 || n <= 0 || !z_inout || !grad) return fail(RSRL_EINVAL, "bad argument");
    int rc = need_device();
    if (rc) return rc;
    DevBuf dz, dg;
    CU_TRY(dz.alloc(n * sizeof(double))); CU_TRY(dg.alloc(n * sizeof(double)));
    CU_TRY(cudaMemcpy(dz.p, z_inout, n * sizeof(double), cudaMemcpyHostToDevice));
    CU_TRY(cudaMemcpy(dg.p, grad, n * sizeof(double), cudaMemcpyHostToDevice));
    const double rate = rule == RSRL_TRACE_DUTCH ? gamma * lambda * (1.0 - alpha) : gamma * lambda;
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    trace_update_kernel<double><<<blocks, threads>>>(rule, rate, n, dz.as<double>(), dg.as<double>());
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(z_inout, dz.p, n * sizeof(double), cudaMemcpyDeviceToHost));
    return RSRL_OK;
}

int rsrl_philox(uint64_t seed, uint64_t draw, uint32_t stream, int64_t env_offset, int64_t n, uint32_t* out) {
    if (n <= 0 || !out) return fail(RSRL_EINVAL, "bad argument");
    int rc = need_device();
    if (rc) return rc;
    DevBuf d;
    CU_TRY(d.alloc((size_t)n * 16));
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    philox_kernel<<<blocks, threads>>>(seed, draw, stream, env_offset, n, d.as<uint32_t>());
    CU_TRY(cudaGetLastError());
    CU_TRY(cudaMemcpy(out, d.p, (size_t)n * 16, cudaMemcpyDeviceToHost));
    return RSRL_OK;
}

}  // extern "C"

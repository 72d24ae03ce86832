/*
 * oracle32.cpp — "oracle32": the DEVICE arithmetic of the bench dtype (fp32 features / Q / weights / traces, f64
 * physics) replayed on the host, bit for bit.  TEST INFRASTRUCTURE ONLY (same rules as rsrl_oracle.h: only tests/,
 * __graft_entry__.smoke() and bench.py's checker legs load it; the product never does).
 *
 * How it gets the same bits as the GPU:
 *   - the per-env arithmetic is not restated: this file compiles the product's own headers
 *     rsrl_b200/csrc/{hostdev.h, device.cuh, core.cuh} with g++ (-ffp-contract=off; the .cu units are built with
 *     --fmad=false), in which every rounding is explicit and the elementary functions are the project's own
 *     (device.cuh "rsrl math"), not libm / the CUDA math library;
 *   - the ORDER of the fp32 sums of the SHARED-weights update is replayed from persistent.cuh: slot segments summed
 *     sequentially by the reducer lanes, butterfly over the lanes; then either the counting exchange (CTA partials on the
 *     2^-40 fixed-point grid, integer sums) or cluster members in rank order, ranks and clusters lane-strided + butterfly.  That order is a function of the launch shape (rsrl_engine_get_launch_shape), which the
 *     caller passes in.
 * What this pins: actions, episode step counts and length hashes, states, TD errors, weights and traces of a free-running
 * fp32 engine — exactly (tests/test_gpu_parity.py: test_f32_*_bit_exact_vs_oracle32).  How far fp32 is from the
 * reference's f64 arithmetic is the separate, tolerance-based statement against oracle/rsrl_oracle.c.
 *
 * Reference citations for the algorithm itself are in core.cuh / device.cuh (the code compiled here) and
 * rsrl_oracle.c (the f64 restatement).
 */
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../rsrl_b200/csrc/core.cuh"

using namespace rsrl;

namespace {

constexpr int kModeTrace = 2;  // persistent.cuh kModeSharedTrace

struct Shape {
    int persistent, mode, grid, cs, ncl, block, lpr, unused7, unused8, pe_smem;
    int fx;  // counting (fixed-point) exchange instead of the cluster + LL-line exchange
};

// persistent.cuh: fx_from_float / fx_to_float / kFxScale / kFxLimit (that header needs nvcc; the three lines are restated here)
constexpr double kFxScale = 1099511627776.0;  // 2^40
constexpr float kFxLimit = 16384.0f;
long long fx_from_float(float x) { return (long long)llrint((double)x * kFxScale); }
float fx_to_float(long long q) { return (float)((double)q * (1.0 / kFxScale)); }

struct RankState {
    std::vector<double> states;
    std::vector<int32_t> actions, ep, n_ep, last_len;
    std::vector<unsigned long long> len_hash;
    std::vector<float> td, z /* [N][FA] */, Wpe /* [N][FA] */;
    unsigned long long episodes = 0, terminal_episodes = 0;
    int nonfinite = 0;
};

struct Engine {
    rsrl_config_t cfg;
    Shape sh;
    int world, D, A, AW, threads;
    int64_t N, NG, F, FA;
    bool has_trace;
    std::vector<RankState> ranks;
    std::vector<float> W;  // SHARED: [F*AW], k*AW + a
    uint64_t t = 0;
    double epsilon;
};

int dom_dim(int d) { return d == RSRL_MOUNTAIN_CAR ? 2 : 4; }
int dom_actions(int d) { return d == RSRL_CART_POLE ? 2 : 3; }
bool td_pred(int a) { return a == RSRL_TD_LAMBDA || a == RSRL_TD0; }
int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

PolicyParams policy_of(int policy, double eps, uint64_t seed) {  // abi.cu policy_of
    PolicyParams p;
    p.policy = policy;
    p.tau = eps;
    p.eps_always = eps >= 1.0;
    p.eps_thresh = p.eps_always ? 0xFFFFFFFFu : (uint32_t)(eps * 4294967296.0);
    p.seed = seed;
    return p;
}

StepArgs make_args(const Engine& e, int rank) {  // abi.cu make_args
    StepArgs a;
    memset(&a, 0, sizeof a);
    a.n = e.N; a.env_offset = e.cfg.env_offset + (int64_t)rank * e.N; a.t = e.t; a.max_ep = e.cfg.max_episode_steps;
    a.algo = e.cfg.algo; a.trace_rule = e.cfg.trace_rule; a.init_mode = e.cfg.init_mode;
    a.pol = policy_of(e.cfg.policy, e.epsilon, e.cfg.seed);
    const double scale = (e.cfg.weight_mode == RSRL_SHARED && e.cfg.update_scale == RSRL_SCALE_MEAN) ? (double)e.NG : 1.0;
    a.gamma = e.cfg.gamma; a.lr_scaled = e.cfg.lr / scale; a.alpha = e.cfg.alpha; a.inv_scale = 1.0 / scale;
    a.lambda = e.cfg.lambda; a.epsilon = e.epsilon;
    for (int d = 0; d < RSRL_MAX_DIM; ++d) { a.init_lo[d] = e.cfg.init_lo[d]; a.init_hi[d] = e.cfg.init_hi[d]; }
    return a;
}

// v[lane] += v[lane ^ off] for off = 1, 2, 4, ... < n  (the __shfl_xor butterflies of persistent.cuh); returns lane 0
float butterfly(float* v, int n) {
    float tmp[64];
    for (int off = 1; off < n; off <<= 1) {
        for (int l = 0; l < n; ++l) tmp[l] = v[l] + v[l ^ off];
        for (int l = 0; l < n; ++l) v[l] = tmp[l];
    }
    return v[0];
}

// lane l sums items l, l + lanes, ... in order starting from 0.0f, then the butterfly (ll_gather + row_butterfly)
template <class Get>
float lane_strided_sum(int cnt, int lanes, Get get) {
    float v[64];
    for (int l = 0; l < lanes; ++l) {
        float acc = 0.0f;
        for (int m = l; m < cnt; m += lanes) acc += get(m);
        v[l] = acc;
    }
    return butterfly(v, lanes);
}

template <int DOM, int BASIS, int P, int AW, int MODE>
void cta_step(Engine& e, int rank, int b, const StepArgs& a_rank, uint64_t t, float* part /* [NV] or nullptr */, Counters* cnt) {
    using R = float;
    using Dom = Domain<DOM>;
    using GB = GridBasis<R, Dom::D, P, BASIS>;
    using O = RealOps<R>;
    constexpr int D = Dom::D, F = GB::F, FA = F * AW;
    constexpr bool TDPRED = AW == 1;
    constexpr bool SHAREDW = MODE != RSRL_PER_ENV;
    constexpr bool TRACE = MODE == kModeTrace;
    constexpr int ROWS = TRACE ? FA : F;
    constexpr int NDC = TRACE ? 1 : AW;
    RankState& rs = e.ranks[rank];
    const Shape& sh = e.sh;
    const int G = sh.grid, BLOCK = sh.block;
    const int64_t N = e.N;
    const int64_t per_cta = (N + G - 1) / G;
    const int64_t base = (int64_t)b * per_cta < N ? (int64_t)b * per_cta : N;
    const int64_t end = base + per_cta < N ? base + per_cta : N;
    const int n_chunks = SHAREDW ? (int)((per_cta + BLOCK - 1) / BLOCK) : 1;
    const int NW = BLOCK / 32;  // warps: the CTA reduce is warp-local (persistent.cuh)

    StepArgs a = a_rank;
    a.counters = cnt;
    a.n_ep = rs.n_ep.data(); a.last_len = rs.last_len.data(); a.len_hash = rs.len_hash.data();

    std::vector<float> wpart;  // [NW][ROWS][NDC] warp partials
    std::vector<float> red, dcs;
    if (SHAREDW) {
        wpart.assign((size_t)NW * ROWS * NDC, 0.0f);
        red.assign((size_t)ROWS * BLOCK, 0.0f);
        dcs.assign((size_t)NDC * BLOCK, 0.0f);
    }
    const float* Wsh = e.W.data();
    for (int chunk = 0; chunk < n_chunks; ++chunk) {
        std::vector<uint8_t> term_flag;
        if (TRACE) term_flag.assign(BLOCK, 0);
        const int64_t lo = SHAREDW ? base + (int64_t)chunk * BLOCK : base;
        const int64_t hi = SHAREDW ? (lo + BLOCK < end ? lo + BLOCK : end) : end;
        if (SHAREDW) { std::fill(dcs.begin(), dcs.end(), 0.0f); if (!TRACE) std::fill(red.begin(), red.end(), 0.0f); }
        for (int64_t i = lo; i < hi; ++i) {
            const int slot = (int)(i - lo);
            const uint64_t g = (uint64_t)(a.env_offset + i);
            double* s = rs.states.data() + i * D;
            float* Wenv = SHAREDW ? nullptr : rs.Wpe.data() + (size_t)i * FA;
            float* zenv = TRACE ? rs.z.data() + (size_t)i * FA : nullptr;
            float phi_row[F];
            auto wrow = [&](int k, R* w) {
                for (int c = 0; c < AW; ++c) w[c] = SHAREDW ? Wsh[k * AW + c] : Wenv[k * AW + c];
            };
            auto evalS = [&](const typename GB::Tab& tab, R* q) {
                for (int c = 0; c < AW; ++c) q[c] = (R)0;
                GB::for_each(tab, [&](int k, R phi) {
                    phi_row[k] = phi;
                    R w[AW];
                    wrow(k, w);
                    for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, w[c], q[c]);
                });
            };
            auto evalN = [&](const typename GB::Tab& tab, R* q) {
                for (int c = 0; c < AW; ++c) q[c] = (R)0;
                GB::for_each(tab, [&](int k, R phi) {
                    R w[AW];
                    wrow(k, w);
                    for (int c = 0; c < AW; ++c) q[c] = O::mac(phi, w[c], q[c]);
                });
            };
            auto prep = [](const double* st, typename GB::Tab& tb) { grid_prepare<R, Dom, P, BASIS>(st, tb); };
            typename GB::Tab tab_s, tab_n;
            CoreOut<R> o;
            o.coef = (R)0; o.act = 0; o.terminated = false;
            env_core<R, DOM, AW, false>(a, t, g, s, prep, evalS, evalN, tab_s, tab_n, false, o, 0, 0.0, false, nullptr);
            if (!rs.td.empty()) rs.td[i] = o.residual;
            if (o.nonfinite) cnt->nonfinite = 1;
            if (MODE == RSRL_PER_ENV) {
                for (int k = 0; k < F; ++k) {
                    const int col = k * AW + (TDPRED ? 0 : o.act);
                    Wenv[col] = O::mul_add_unfused(o.coef, phi_row[k], Wenv[col]);
                }
            }
            if (TRACE) {
                const R rate = a.trace_rule == RSRL_TRACE_DUTCH ? (R)(a.gamma * a.lambda * (1.0 - a.alpha)) : (R)(a.gamma * a.lambda);
                for (int k = 0; k < F; ++k)
                    for (int c = 0; c < AW; ++c) {
                        float* zp = zenv + k * AW + c;
                        const R zv = o.reset_before ? (R)0 : *zp;
                        *zp = trace_rule<R>(a.trace_rule, rate, zv, (TDPRED || c == o.act) ? phi_row[k] : (R)0);
                    }
            }
            rs.ep[i] = env_bookkeeping<Dom>(a, t, i, g, s, rs.ep[i], o.terminated, nullptr);
            rs.actions[i] = o.act;
            if (SHAREDW) {
                if (TRACE) {
                    dcs[slot] = o.coef;
                    for (int j = 0; j < FA; ++j) red[(size_t)j * BLOCK + slot] = zenv[j];
                    term_flag[slot] = o.terminated ? 1 : 0;
                } else {
                    for (int c = 0; c < AW; ++c) dcs[(size_t)c * BLOCK + slot] = (TDPRED || c == o.act) ? o.coef : (R)0;
                    for (int k = 0; k < F; ++k) red[(size_t)k * BLOCK + slot] = phi_row[k];
                }
            }
        }
        if (SHAREDW) {
            // warp-local reduce (persistent.cuh): warp w owns slots [32 w, 32 w + 32).  Pass r0: a row gets S adjacent lanes
            // (S = 1 in a full pass of 32 rows), lane `sub` sums its 32 / S slots — even slots into accumulator 0, odd slots
            // into accumulator 1 (the halves of an FFMA2), in slot order — then even + odd and a butterfly over the S lanes.
            for (int w = 0; w < NW; ++w)
                for (int r0 = 0; r0 < ROWS; r0 += 32) {
                    const int nrow = ROWS - r0 < 32 ? ROWS - r0 : 32;
                    int S = 1;
                    while (2 * S * nrow <= 32 && 32 / (2 * S) >= 4) S *= 2;  // persist_tail_split(nrow, 4)
                    const int NSL = 32 / S;
                    for (int rl = 0; rl < nrow; ++rl)
                        for (int c = 0; c < NDC; ++c) {
                            const int row = r0 + rl;
                            float v[32];
                            for (int sub = 0; sub < S; ++sub) {
                                float a2[2] = {0.0f, 0.0f};
                                for (int q = 0; q < NSL; ++q) {
                                    const int slot = 32 * w + sub * NSL + q;
                                    a2[q & 1] = O::fma(red[(size_t)row * BLOCK + slot], dcs[(size_t)c * BLOCK + slot], a2[q & 1]);
                                }
                                v[sub] = a2[0] + a2[1];
                            }
                            const float cp = butterfly(v, S);
                            float* pp = &wpart[((size_t)w * ROWS + row) * NDC + c];
                            *pp = chunk == 0 ? cp : *pp + cp;  // chunks in order
                        }
                }
            if (TRACE)
                for (int64_t i = lo; i < hi; ++i)
                    if (term_flag[(int)(i - lo)]) std::fill(rs.z.begin() + (size_t)i * FA, rs.z.begin() + (size_t)(i + 1) * FA, 0.0f);
        }
    }
    if (SHAREDW)  // CTA partial: the warp partials in warp order
        for (int j = 0; j < ROWS * NDC; ++j) {
            float acc = wpart[j];
            for (int w = 1; w < NW; ++w) acc += wpart[(size_t)w * ROWS * NDC + j];
            part[j] = acc;
        }
}

template <int DOM, int BASIS, int P, int AW, int MODE>
void engine_step(Engine& e, int64_t k_steps) {
    using Dom = Domain<DOM>;
    using GB = GridBasis<float, Dom::D, P, BASIS>;
    constexpr int FA = GB::F * AW;
    constexpr bool SHAREDW = MODE != RSRL_PER_ENV;
    const Shape& sh = e.sh;
    const int G = sh.grid, CS = sh.cs, NCL = SHAREDW ? G / CS : 1;
    std::vector<float> parts(SHAREDW ? (size_t)e.world * G * FA : 0);
    std::vector<Counters> cnts((size_t)e.world * G);
    for (int64_t step = 0; step < k_steps; ++step) {
        const uint64_t t = e.t;
        std::vector<StepArgs> args;
        for (int r = 0; r < e.world; ++r) args.push_back(make_args(e, r));
        memset(cnts.data(), 0, cnts.size() * sizeof(Counters));
        // all CTAs of all ranks are independent within a step
        const int total = e.world * G;
        std::atomic<int> next(0);
        auto work = [&]() {
            for (;;) {
                const int w = next.fetch_add(1);
                if (w >= total) break;
                const int r = w / G, b = w % G;
                cta_step<DOM, BASIS, P, AW, MODE>(e, r, b, args[r], t, SHAREDW ? &parts[((size_t)r * G + b) * FA] : nullptr, &cnts[w]);
            }
        };
        std::vector<std::thread> th;
        const int nth = e.threads < total ? e.threads : total;
        for (int i = 1; i < nth; ++i) th.emplace_back(work);
        work();
        for (auto& x : th) x.join();
        for (int w = 0; w < total; ++w) {
            RankState& rs = e.ranks[w / G];
            rs.episodes += cnts[w].episodes; rs.terminal_episodes += cnts[w].terminal_episodes; rs.nonfinite |= cnts[w].nonfinite;
        }
        if (SHAREDW) {
            std::vector<float> total_dw(FA);
            if (G == 1 && e.world == 1) {
                for (int j = 0; j < FA; ++j) total_dw[j] = parts[j];
            } else if (sh.fx) {
                // counting exchange: every CTA partial goes to the 2^-40 grid (one rounding), integer sums over CTAs and ranks
                // (order independent), one rounding back to fp32
                for (int j = 0; j < FA; ++j) {
                    long long acc = 0;
                    for (int w = 0; w < e.world * G; ++w) {
                        float x = parts[(size_t)w * FA + j];
                        if (!(fabsf(x) < kFxLimit)) { e.ranks[w / G].nonfinite = 1; x = 0.0f; }
                        acc += fx_from_float(x);
                    }
                    total_dw[j] = fx_to_float(acc);
                }
            } else {
                // hop A: cluster partial = members in rank order
                std::vector<float> cp((size_t)e.world * NCL * FA);
                for (int r = 0; r < e.world; ++r)
                    for (int c = 0; c < NCL; ++c)
                        for (int j = 0; j < FA; ++j) {
                            float acc = parts[((size_t)r * G + (size_t)c * CS) * FA + j];
                            for (int m = 1; m < CS; ++m) acc += parts[((size_t)r * G + (size_t)c * CS + m) * FA + j];
                            cp[((size_t)r * NCL + c) * FA + j] = acc;
                        }
                for (int j = 0; j < FA; ++j) {
                    // hop N: leader c sums slot c of every rank (lane-strided + butterfly); hop B: the clusters
                    float wc[256];
                    for (int c = 0; c < NCL; ++c)
                        wc[c] = e.world > 1 ? lane_strided_sum(e.world, sh.lpr, [&](int r) { return cp[((size_t)r * NCL + c) * FA + j]; })
                                            : cp[(size_t)c * FA + j];
                    total_dw[j] = NCL > 1 ? lane_strided_sum(NCL, sh.lpr, [&](int c) { return wc[c]; }) : wc[0];
                }
            }
            for (int j = 0; j < FA; ++j) e.W[j] += total_dw[j];
        }
        e.t += 1;
    }
}

typedef void (*step_fn)(Engine&, int64_t);

template <int DOM, int BASIS, int P>
step_fn pick_mode(int aw, int mode) {
    constexpr int A = Domain<DOM>::A;
    if (aw == A) {
        if (mode == RSRL_SHARED) return engine_step<DOM, BASIS, P, A, RSRL_SHARED>;
        if (mode == RSRL_PER_ENV) return engine_step<DOM, BASIS, P, A, RSRL_PER_ENV>;
        return engine_step<DOM, BASIS, P, A, kModeTrace>;
    }
    if (mode == RSRL_SHARED) return engine_step<DOM, BASIS, P, 1, RSRL_SHARED>;
    if (mode == RSRL_PER_ENV) return engine_step<DOM, BASIS, P, 1, RSRL_PER_ENV>;
    return engine_step<DOM, BASIS, P, 1, kModeTrace>;
}

// the (domain, basis, order) combinations of rsrl_b200/csrc/inst.cu
step_fn pick(int dom, int basis, int order, int aw, int mode) {
#define X(DOMV, B, PV) if (dom == DOMV && basis == B && order == PV) return pick_mode<DOMV, B, PV>(aw, mode);
    X(0, RSRL_FOURIER, 1) X(0, RSRL_FOURIER, 2) X(0, RSRL_FOURIER, 3) X(0, RSRL_FOURIER, 5) X(0, RSRL_FOURIER, 7)
    X(0, RSRL_POLYNOMIAL, 2) X(0, RSRL_POLYNOMIAL, 3)
    X(1, RSRL_FOURIER, 2) X(1, RSRL_FOURIER, 3) X(1, RSRL_POLYNOMIAL, 2)
    X(2, RSRL_FOURIER, 2) X(2, RSRL_FOURIER, 3) X(2, RSRL_POLYNOMIAL, 2)
#undef X
    return nullptr;
}

template <int DOM>
void init_states(Engine& e, int rank, const double* init) {
    using Dom = Domain<DOM>;
    RankState& rs = e.ranks[rank];
    for (int64_t i = 0; i < e.N; ++i) {
        if (init) { for (int d = 0; d < Dom::D; ++d) rs.states[i * Dom::D + d] = init[i * Dom::D + d]; continue; }
        double lo[RSRL_MAX_DIM], hi[RSRL_MAX_DIM], s[Dom::D];
        for (int d = 0; d < RSRL_MAX_DIM; ++d) { lo[d] = e.cfg.init_lo[d]; hi[d] = e.cfg.init_hi[d]; }
        fresh_state<Dom>(s, e.cfg.init_mode, lo, hi, e.cfg.seed, (uint64_t)(e.cfg.env_offset + (int64_t)rank * e.N + i), 0);
        for (int d = 0; d < Dom::D; ++d) rs.states[i * Dom::D + d] = s[d];
    }
}

}  // namespace

extern "C" {

typedef struct Engine o32_engine_t;

/* cfg: rank 0's config (n_envs = envs per rank, env_offset = rank 0's); shape: rsrl_engine_get_launch_shape of the GPU engine;
 * world: ranks simulated (rank r owns envs [env_offset + r*n_envs, ...)); threads: host threads (results do not depend on it) */
o32_engine_t* o32_engine_create(const rsrl_config_t* cfg, const int32_t* shape, int world, int threads) {
    if (!cfg || !shape || world < 1 || cfg->dtype != RSRL_F32 || cfg->basis == RSRL_TILE_CODING || !shape[0]) return nullptr;
    Engine* e = new Engine();
    e->cfg = *cfg;
    e->sh = Shape{shape[0], shape[1], shape[2], shape[3], shape[4], shape[5], shape[6], shape[7], shape[8], shape[9], shape[16]};
    e->world = world;
    e->threads = threads > 0 ? threads : 1;
    e->D = dom_dim(cfg->domain); e->A = dom_actions(cfg->domain); e->AW = td_pred(cfg->algo) ? 1 : e->A;
    e->N = cfg->n_envs; e->NG = cfg->n_envs_global > 0 ? cfg->n_envs_global : cfg->n_envs * world;
    e->F = ipow(cfg->basis_order + 1, e->D); e->FA = e->F * e->AW;
    e->has_trace = algo_has_trace(cfg->algo);
    e->epsilon = cfg->epsilon;
    if (!pick(cfg->domain, cfg->basis, cfg->basis_order, e->AW, e->sh.mode) || e->sh.lpr > 64 || e->sh.ncl > 256) { delete e; return nullptr; }
    e->ranks.resize(world);
    for (auto& rs : e->ranks) {
        rs.states.assign((size_t)e->N * e->D, 0.0);
        rs.actions.assign(e->N, -1); rs.ep.assign(e->N, 0); rs.n_ep.assign(e->N, 0); rs.last_len.assign(e->N, 0);
        rs.len_hash.assign(e->N, 0ull);
        rs.td.assign(e->N, 0.0f);
        if (e->has_trace) rs.z.assign((size_t)e->N * e->FA, 0.0f);
        if (cfg->weight_mode == RSRL_PER_ENV) rs.Wpe.assign((size_t)e->N * e->FA, 0.0f);
    }
    e->W.assign(e->FA, 0.0f);
    for (int r = 0; r < world; ++r) {
        if (cfg->domain == RSRL_MOUNTAIN_CAR) init_states<RSRL_MOUNTAIN_CAR>(*e, r, nullptr);
        else if (cfg->domain == RSRL_CART_POLE) init_states<RSRL_CART_POLE>(*e, r, nullptr);
        else init_states<RSRL_ACROBOT>(*e, r, nullptr);
    }
    return e;
}

void o32_engine_destroy(o32_engine_t* e) { delete e; }

void o32_engine_step(o32_engine_t* e, int64_t k) {
    pick(e->cfg.domain, e->cfg.basis, e->cfg.basis_order, e->AW, e->sh.mode)(*e, k);
}

void o32_engine_set_epsilon(o32_engine_t* e, double eps) { e->epsilon = eps; }
void o32_engine_set_states(o32_engine_t* e, int rank, const double* in) { memcpy(e->ranks[rank].states.data(), in, (size_t)e->N * e->D * sizeof(double)); }
void o32_engine_get_states(o32_engine_t* e, int rank, double* out) { memcpy(out, e->ranks[rank].states.data(), (size_t)e->N * e->D * sizeof(double)); }
void o32_engine_get_actions(o32_engine_t* e, int rank, int32_t* out) { memcpy(out, e->ranks[rank].actions.data(), (size_t)e->N * sizeof(int32_t)); }
void o32_engine_get_episode_steps(o32_engine_t* e, int rank, int32_t* out) { memcpy(out, e->ranks[rank].ep.data(), (size_t)e->N * sizeof(int32_t)); }
void o32_engine_get_env_stats(o32_engine_t* e, int rank, int32_t* n_ep, int32_t* last_len, uint64_t* len_hash) {
    const RankState& rs = e->ranks[rank];
    if (n_ep) memcpy(n_ep, rs.n_ep.data(), (size_t)e->N * sizeof(int32_t));
    if (last_len) memcpy(last_len, rs.last_len.data(), (size_t)e->N * sizeof(int32_t));
    if (len_hash) memcpy(len_hash, rs.len_hash.data(), (size_t)e->N * sizeof(uint64_t));
}
void o32_engine_get_td_errors(o32_engine_t* e, int rank, double* out) { for (int64_t i = 0; i < e->N; ++i) out[i] = (double)e->ranks[rank].td[i]; }
/* SHARED: F x A (rank ignored); PER_ENV: N x F x A of `rank` */
void o32_engine_get_weights(o32_engine_t* e, int rank, double* out) {
    if (e->cfg.weight_mode == RSRL_PER_ENV) { const auto& w = e->ranks[rank].Wpe; for (size_t j = 0; j < w.size(); ++j) out[j] = (double)w[j]; }
    else for (int64_t j = 0; j < e->FA; ++j) out[j] = (double)e->W[j];
}
void o32_engine_set_weights(o32_engine_t* e, int rank, const double* in) {
    if (e->cfg.weight_mode == RSRL_PER_ENV) { auto& w = e->ranks[rank].Wpe; for (size_t j = 0; j < w.size(); ++j) w[j] = (float)in[j]; }
    else for (int64_t j = 0; j < e->FA; ++j) e->W[j] = (float)in[j];
}
void o32_engine_get_traces(o32_engine_t* e, int rank, double* out) { const auto& z = e->ranks[rank].z; for (size_t j = 0; j < z.size(); ++j) out[j] = (double)z[j]; }
void o32_engine_get_counters(o32_engine_t* e, int rank, int64_t* episodes, int64_t* terminal_episodes, int32_t* nonfinite) {
    const RankState& rs = e->ranks[rank];
    if (episodes) *episodes = (int64_t)rs.episodes;
    if (terminal_episodes) *terminal_episodes = (int64_t)rs.terminal_episodes;
    if (nonfinite) *nonfinite = rs.nonfinite;
}

/* the elementary functions of device.cuh on the host (same numbering as rsrl_math_probe) */
void o32_math(int fn, int64_t n, const double* x, double* out) {
    for (int64_t i = 0; i < n; ++i) {
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
}

/* Domain::step with the device arithmetic (states updated in place) */
void o32_domain_step(int domain, int64_t n, double* states, const int32_t* actions, double* rewards, uint8_t* terminal) {
    for (int64_t i = 0; i < n; ++i) {
        double r; bool term;
        if (domain == RSRL_MOUNTAIN_CAR) Domain<RSRL_MOUNTAIN_CAR>::step(states + i * 2, actions[i], r, term);
        else if (domain == RSRL_CART_POLE) Domain<RSRL_CART_POLE>::step(states + i * 4, actions[i], r, term);
        else Domain<RSRL_ACROBOT>::step(states + i * 4, actions[i], r, term);
        rewards[i] = r; terminal[i] = term;
    }
}

}  // extern "C"

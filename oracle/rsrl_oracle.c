/*
 * rsrl_oracle.c — CPU restatement (plain C, f64, libm) of the rsrl hot path.
 * TEST INFRASTRUCTURE ONLY: see rsrl_oracle.h for who may load it and for the
 * list of what is pinned by reference tests and what is "parity unpinned".
 *
 * Every function cites the reference file:line it restates (paths relative to
 * the reference checkout).  Arithmetic keeps the reference's association order;
 * build with -ffp-contract=off (oracle/Makefile) so no FMA contraction changes it.
 *
 * The code is deliberately "reference-shaped": every evaluate()/project()
 * heap-allocates its result like the Rust Vec/Array1 (rsrl_domains
 * mountain_car/discrete.rs:76-82, rsrl/src/fa/linear.rs:310) and every agent
 * handle() re-projects its states (q_learning.rs:53,59,64; greedy.rs:78), so
 * the CPU baseline timed from here has the reference's cost structure.
 */
#define _GNU_SOURCE
#include "rsrl_oracle.h"

#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/macros.rs:3-24                                            */
/* ------------------------------------------------------------------------- */
/* clip!(lb, x, ub) = lb.max(ub.min(x)); Rust f64::max/min return the non-NaN operand == C fmax/fmin */
static inline double clip(double lb, double x, double ub) { return fmax(lb, fmin(ub, x)); }

static inline double wrap(double lb, double x, double ub) {
    double nx = x;
    double diff = ub - lb;
    while (nx > ub) nx -= diff;
    while (nx < lb) nx += diff;
    return nx;
}

/* rsrl_domains/src/consts.rs:4-13 */
static const double G_ = 9.8;
#define PI_ 3.14159265358979323846264338327950288 /* std::f64::consts::PI */
static const double FOUR_THIRDS = 4.0 / 3.0;
static const double TWELVE_DEGREES = PI_ / 15.0;
static const double PI_OVER_2 = PI_ / 2.0;

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/ode.rs:1-43  runge_kutta4 (x is ignored by both callers)   */
/* ------------------------------------------------------------------------- */
typedef void (*grad_fn)(double u, const double* y, double* out);

static void runge_kutta4(grad_fn fx, double u, double* y, double dx) {
    double k1[4], k2[4], k3[4], k4[4], tmp[4];
    int i;
    fx(u, y, k1);
    for (i = 0; i < 4; ++i) k1[i] = k1[i] * dx;
    for (i = 0; i < 4; ++i) tmp[i] = y[i] + k1[i] / 2.0;
    fx(u, tmp, k2);
    for (i = 0; i < 4; ++i) k2[i] = k2[i] * dx;
    for (i = 0; i < 4; ++i) tmp[i] = y[i] + k2[i] / 2.0;
    fx(u, tmp, k3);
    for (i = 0; i < 4; ++i) k3[i] = k3[i] * dx;
    for (i = 0; i < 4; ++i) tmp[i] = y[i] + k3[i];
    fx(u, tmp, k4);
    for (i = 0; i < 4; ++i) k4[i] = k4[i] * dx;
    for (i = 0; i < 4; ++i) y[i] += (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
}

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/mountain_car/discrete.rs:8-22,56-102                      */
/* ------------------------------------------------------------------------- */
static const double MC_X_MIN = -1.2, MC_X_MAX = 0.6, MC_V_MIN = -0.07, MC_V_MAX = 0.07;
static const double MC_FORCE_G = -0.0025, MC_FORCE_CAR = 0.001, MC_HILL_FREQ = 3.0;
static const double MC_ACTIONS[3] = {-1.0, 0.0, 1.0};

static void mc_step(double* s, int action, double* reward, int* terminal) {
    double a = MC_ACTIONS[action];
    double dv = MC_FORCE_CAR * a + MC_FORCE_G * cos(MC_HILL_FREQ * s[0]); /* :58 */
    s[1] = clip(MC_V_MIN, s[1] + dv, MC_V_MAX);                           /* :63 */
    s[0] = clip(MC_X_MIN, s[0] + s[1], MC_X_MAX);                         /* :64 (new v) */
    *terminal = s[0] >= MC_X_MAX;                                         /* :77 */
    *reward = *terminal ? 0.0 : -1.0;                                     /* :88-92 */
}

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/cart_pole.rs:7-26,34-121                                  */
/* ------------------------------------------------------------------------- */
static const double CP_DT = 0.02, CP_CART_MASS = 1.0, CP_CART_FORCE = 10.0, CP_POLE_COM = 0.5, CP_POLE_MASS = 0.1;
#define CP_POLE_MOMENT (CP_POLE_COM * CP_POLE_MASS)
#define CP_TOTAL_MASS (CP_CART_MASS + CP_POLE_MASS)

static void cp_grad(double force, const double* b, double* out) { /* :52-72 */
    double dx = b[1], theta = b[2], dtheta = b[3];
    double cos_theta = cos(theta);
    double sin_theta = sin(theta);
    double z = (force + CP_POLE_MOMENT * dtheta * dtheta * sin_theta) / CP_TOTAL_MASS;
    double numer = G_ * sin_theta - cos_theta * z;
    double denom = FOUR_THIRDS * CP_POLE_COM - CP_POLE_MOMENT * cos_theta * cos_theta;
    out[0] = dx;
    out[2] = dtheta;
    out[3] = numer / denom;
    out[1] = z - CP_POLE_COM * out[3] * cos_theta;
}

static int cp_is_terminal(const double* s) { /* :83-97 */
    return s[0] <= -2.4 || s[0] >= 2.4 || s[2] <= -TWELVE_DEGREES || s[2] >= TWELVE_DEGREES;
}

static void cp_step(double* s, int action, double* reward, int* terminal) {
    double force = (action == 0 ? -1.0 : 1.0) * CP_CART_FORCE; /* ALL_ACTIONS :26 */
    double ns[4] = {s[0], s[1], s[2], s[3]};
    runge_kutta4(cp_grad, force, ns, CP_DT);
    s[0] = clip(-2.4, ns[0], 2.4); /* :44-49 */
    s[1] = clip(-6.0, ns[1], 6.0);
    s[2] = clip(-TWELVE_DEGREES, ns[2], TWELVE_DEGREES);
    s[3] = clip(-2.0, ns[3], 2.0);
    *terminal = cp_is_terminal(s);
    *reward = *terminal ? -1.0 : 0.0; /* :23-24,103-107 */
}

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/acrobot.rs:8-36,51-152 (quirks kept: SURVEY App. C.1)      */
/* ------------------------------------------------------------------------- */
static const double AC_M1 = 1.0, AC_M2 = 1.0, AC_L1 = 1.0, AC_LC1 = 0.5, AC_LC2 = 0.5, AC_I1 = 1.0, AC_I2 = 1.0;
static const double AC_DT = 0.2;

static void ac_grad(double torque, const double* b, double* out) { /* :81-108 */
    double theta1 = b[0], theta2 = b[1], dtheta1 = b[2], dtheta2 = b[3];
    double sin_t2 = sin(theta2);
    double cos_t2 = cos(theta2);
    double d1 = AC_M1 * AC_LC1 * AC_LC1 + AC_M2 * (AC_L1 * AC_L1 + AC_LC2 * AC_LC2 + 2.0 * AC_L1 * AC_LC2 * cos_t2) + AC_I1 + AC_I2;
    double d2 = AC_M2 * (AC_LC2 * AC_LC2 + AC_L1 * AC_LC2 * cos_t2) + AC_I2;
    double phi2 = AC_M2 * AC_LC2 * G_ * cos(theta1 + theta2 - PI_OVER_2);
    double phi1 = -1.0 * AC_L1 * AC_LC2 * dtheta2 * dtheta2 * sin_t2
                  - 2.0 * AC_M2 * AC_L1 * AC_LC2 * dtheta2 * dtheta1 * sin_t2
                  + (AC_M1 * AC_LC1 + AC_M2 * AC_L1) * G_ * cos(theta1 - PI_OVER_2)
                  + phi2;
    out[0] = dtheta1;
    out[1] = dtheta2;
    out[2] = (torque + d2 / d1 * phi1 - AC_M2 * AC_L1 * AC_LC2 * dtheta1 * dtheta1 * sin_t2 - phi2)
             / (AC_M2 * AC_LC2 * AC_LC2 + AC_I2 - d2 * d2 / d1);
    out[3] = -(d2 * out[2] + phi1) / d1;
}

static int ac_is_terminal(const double* s) { return cos(s[0]) + cos(s[0] + s[1]) < -1.0; } /* :56-58 */

static void ac_step(double* s, int action, double* reward, int* terminal) {
    double torque = (double)(action - 1); /* ALL_ACTIONS = [-1, 0, 1] :36 */
    double ns[4] = {s[0], s[1], s[2], s[3]};
    runge_kutta4(ac_grad, torque, ns, AC_DT);
    s[0] = wrap(-PI_, ns[0], PI_); /* :64-78 */
    s[1] = wrap(-PI_, ns[1], PI_);
    s[2] = clip(-4.0 * PI_, ns[2], 4.0 * PI_);
    s[3] = clip(-9.0 * PI_, ns[3], 9.0 * PI_);
    *terminal = ac_is_terminal(s);
    *reward = *terminal ? 0.0 : -1.0; /* :32-33,134-138 */
}

/* ------------------------------------------------------------------------- */
/* rsrl_domains/src/mountain_car/continuous.rs:8-85 and rsrl_domains/src/hiv.rs:6-153 (component level) */
/* ------------------------------------------------------------------------- */
int orc_domain_ex_dim(int domain) { return domain == RSRL_HIV ? 6 : 2; }

void orc_domain_ex_default(int domain, double* s) {
    if (domain == RSRL_HIV) { /* hiv.rs:105-109 */
        const double d[6] = {163573.0, 11945.0, 5.0, 46.0, 63919.0, 24.0};
        memcpy(s, d, sizeof d);
    } else { s[0] = -0.5; s[1] = 0.0; } /* continuous.rs:51-53 */
}

static void hiv_grad(const double* eps, const double* b, double* out) { /* hiv.rs:72-103 */
    const double LAMBDA1 = 1e4, LAMBDA2 = 31.98, D1 = 0.01, D2 = 0.01, F = 0.34, K1 = 8e-7, K2 = 1e-4, DELTA = 0.7, M1 = 1e-5, M2 = 1e-5,
                 NT = 100.0, C = 13.0, RHO1 = 1.0, RHO2 = 1.0, LAMBDA_E = 1.0, BE = 0.3, KB = 100.0, DE = 0.25, KD = 500.0, DELTA_E = 0.1;
    double t1 = b[0], t1s = b[1], t2 = b[2], t2s = b[3], v = b[4], e = b[5];
    double tmp1 = (1.0 - eps[0]) * K1 * v * t1;
    double tmp2 = (1.0 - F * eps[0]) * K2 * v * t2;
    double sum_ts = t1s + t2s;
    out[0] = LAMBDA1 - D1 * t1 - tmp1;
    out[1] = tmp1 - DELTA * t1s - M1 * e * t1s;
    out[2] = LAMBDA2 - D2 * t2 - tmp2;
    out[3] = tmp2 - DELTA * t2s - M2 * e * t2s;
    out[4] = (1.0 - eps[1]) * NT * DELTA * sum_ts - C * v - ((1.0 - eps[0]) * RHO1 * K1 * t1 + (1.0 - F * eps[0]) * RHO2 * K2 * t2) * v;
    out[5] = LAMBDA_E + BE * sum_ts / (sum_ts + KB) * e - DE * sum_ts / (sum_ts + KD) * e - DELTA_E * e;
}

void orc_domain_ex_emit(int domain, const double* s, double* obs, int* terminal) {
    if (domain == RSRL_HIV) { /* hiv.rs:131-135 */
        for (int d = 0; d < 6; ++d) obs[d] = clip(-5.0, log10(s[d]), 8.0);
        *terminal = 0;
    } else { /* continuous.rs:60-66 */
        obs[0] = s[0]; obs[1] = s[1];
        *terminal = s[0] >= 0.6;
    }
}

/* Domain::step: `action` is the index for HIV and the force for ContinuousMountainCar */
void orc_domain_ex_step(int domain, double* s, double action, double* obs, double* reward, int* terminal) {
    if (domain == RSRL_HIV) {
        static const double ALL[4][2] = {{0.0, 0.0}, {0.7, 0.0}, {0.0, 0.3}, {0.7, 0.3}}; /* hiv.rs:35 */
        const double* eps = ALL[(int)action];
        const double dt = 5.0 / (double)1000; /* DT_STEP :31 */
        double y[6];
        memcpy(y, s, sizeof y);
        for (int it = 0; it < 1000; ++it) { /* :61-64, runge_kutta4 of ode.rs:1-43 on 6 components */
            double k1[6], k2[6], k3[6], k4[6], tmp[6];
            int i;
            hiv_grad(eps, y, k1);
            for (i = 0; i < 6; ++i) k1[i] = k1[i] * dt;
            for (i = 0; i < 6; ++i) tmp[i] = y[i] + k1[i] / 2.0;
            hiv_grad(eps, tmp, k2);
            for (i = 0; i < 6; ++i) k2[i] = k2[i] * dt;
            for (i = 0; i < 6; ++i) tmp[i] = y[i] + k2[i] / 2.0;
            hiv_grad(eps, tmp, k3);
            for (i = 0; i < 6; ++i) k3[i] = k3[i] * dt;
            for (i = 0; i < 6; ++i) tmp[i] = y[i] + k3[i];
            hiv_grad(eps, tmp, k4);
            for (i = 0; i < 6; ++i) k4[i] = k4[i] * dt;
            for (i = 0; i < 6; ++i) y[i] += (k1[i] + 2.0 * k2[i] + 2.0 * k3[i] + k4[i]) / 6.0;
        }
        memcpy(s, y, sizeof y);
        orc_domain_ex_emit(domain, s, obs, terminal);
        *reward = (1e3 * obs[5] - 0.1 * obs[4] - 2e4 * (eps[0] * eps[0]) - 2e3 * (eps[1] * eps[1])) / 1e5; /* :141-148 */
    } else { /* continuous.rs:41-48,68-79; Interval::map_onto clips onto [-1, 1] */
        double a = clip(-1.0, action, 1.0);
        double dv = 0.0015 * a + MC_FORCE_G * cos(MC_HILL_FREQ * s[0]);
        s[1] = clip(MC_V_MIN, s[1] + dv, MC_V_MAX);
        s[0] = clip(MC_X_MIN, s[0] + s[1], MC_X_MAX);
        orc_domain_ex_emit(domain, s, obs, terminal);
        *reward = *terminal ? 0.0 : -1.0;
    }
}

/* ---- Domain trait dispatch (rsrl_domains/src/lib.rs:417-446) ---- */
int orc_domain_dim(int domain) { return domain == RSRL_MOUNTAIN_CAR ? 2 : 4; }
int orc_domain_n_actions(int domain) { return domain == RSRL_CART_POLE ? 2 : 3; }

void orc_domain_limits(int domain, double* lo, double* hi) {
    if (domain == RSRL_MOUNTAIN_CAR) { /* discrete.rs:97-99 */
        lo[0] = MC_X_MIN; hi[0] = MC_X_MAX; lo[1] = MC_V_MIN; hi[1] = MC_V_MAX;
    } else if (domain == RSRL_CART_POLE) { /* cart_pole.rs:112-118 */
        lo[0] = -2.4; hi[0] = 2.4; lo[1] = -6.0; hi[1] = 6.0;
        lo[2] = -TWELVE_DEGREES; hi[2] = TWELVE_DEGREES; lo[3] = -2.0; hi[3] = 2.0;
    } else { /* acrobot.rs:143-149 */
        lo[0] = -PI_; hi[0] = PI_; lo[1] = -PI_; hi[1] = PI_;
        lo[2] = -4.0 * PI_; hi[2] = 4.0 * PI_; lo[3] = -9.0 * PI_; hi[3] = 9.0 * PI_;
    }
}

void orc_domain_default(int domain, double* s) {
    if (domain == RSRL_MOUNTAIN_CAR) { s[0] = -0.5; s[1] = 0.0; } /* discrete.rs:68-70 */
    else { s[0] = s[1] = s[2] = s[3] = 0.0; }                     /* cart_pole.rs:76, acrobot.rs:112 */
}

int orc_domain_is_terminal(int domain, const double* s) {
    if (domain == RSRL_MOUNTAIN_CAR) return s[0] >= MC_X_MAX;
    if (domain == RSRL_CART_POLE) return cp_is_terminal(s);
    return ac_is_terminal(s);
}

void orc_domain_step(int domain, double* s, int action, double* reward, int* terminal) {
    if (domain == RSRL_MOUNTAIN_CAR) mc_step(s, action, reward, terminal);
    else if (domain == RSRL_CART_POLE) cp_step(s, action, reward, terminal);
    else ac_step(s, action, reward, terminal);
}

/* ------------------------------------------------------------------------- */
/* lfa 0.15 bases (external crate; PARITY UNPINNED — restated from the         */
/* published algorithm; call sites examples/q_learning.rs:24, fa/linear.rs:310)*/
/* ------------------------------------------------------------------------- */
static int64_t ipow(int64_t b, int e) { int64_t r = 1; while (e-- > 0) r *= b; return r; }

int64_t orc_basis_n_features(const rsrl_config_t* cfg) {
    int D = orc_domain_dim(cfg->domain);
    if (cfg->basis == RSRL_TILE_CODING) return cfg->memory_size;
    return ipow(cfg->basis_order + 1, D); /* (order+1)^D - 1 features + bias (.with_bias()) */
}

/* Fourier::from_space(order, space): coefficient vectors = {0..=order}^D minus the all-zero
 * vector, sorted descending lexicographically.  Row k therefore holds the base-(order+1)
 * digits (most significant = dim 0) of ((order+1)^D - 1 - k). */
static void coef_row(int order, int D, int64_t k, int* c) {
    int n = order + 1;
    int64_t v = ipow(n, D) - 1 - k;
    for (int d = D - 1; d >= 0; --d) { c[d] = (int)(v % n); v /= n; }
}

void orc_fourier_coefficients(int order, int D, double* coef) {
    int64_t F = ipow(order + 1, D);
    int c[RSRL_MAX_DIM];
    for (int64_t k = 0; k < F - 1; ++k) {
        coef_row(order, D, k, c);
        for (int d = 0; d < D; ++d) coef[k * D + d] = (double)c[d];
    }
}

static uint32_t tile_hash(uint32_t tiling, const int32_t* coord, int D) {
    /* project-defined: FNV-style combine + murmur3 fmix32 avalanche */
    uint32_t h = (tiling + 1u) * 0x9E3779B1u;
    for (int d = 0; d < D; ++d) h = (h ^ (uint32_t)coord[d]) * 0x85EBCA6Bu;
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

/* TileCoding (project-defined spec): x^ = (x-lo)/(hi-lo); q_d = floor(x^_d * P * T);
 * tiling t: coord_d = (q_d + t*(1+2d)) / T (tiles3 displacement); row = hash(t, coord) & (M-1).
 * Active set = unique rows (lfa's SparseActivations is a HashMap: duplicates collapse). */
int orc_tile_indices(const rsrl_config_t* cfg, const double* s, int32_t* idx) {
    int D = orc_domain_dim(cfg->domain);
    double lo[RSRL_MAX_DIM], hi[RSRL_MAX_DIM];
    int32_t q[RSRL_MAX_DIM], coord[RSRL_MAX_DIM];
    int T = cfg->n_tilings, n = 0;
    orc_domain_limits(cfg->domain, lo, hi);
    for (int d = 0; d < D; ++d) {
        double xh = (s[d] - lo[d]) / (hi[d] - lo[d]);
        q[d] = (int32_t)floor(xh * (double)cfg->tiles_per_dim * (double)T);
    }
    for (int t = 0; t < T; ++t) {
        for (int d = 0; d < D; ++d) coord[d] = (q[d] + t * (1 + 2 * d)) / T;
        int32_t row = (int32_t)(tile_hash((uint32_t)t, coord, D) & (uint32_t)(cfg->memory_size - 1));
        int dup = 0;
        for (int j = 0; j < n; ++j) dup |= idx[j] == row;
        if (!dup) idx[n++] = row;
    }
    for (int j = n; j < T; ++j) idx[j] = -1;
    return n;
}

void orc_basis_project(const rsrl_config_t* cfg, const double* s, double* phi) {
    int D = orc_domain_dim(cfg->domain);
    int64_t F = orc_basis_n_features(cfg);
    double lo[RSRL_MAX_DIM], hi[RSRL_MAX_DIM];
    int c[RSRL_MAX_DIM];
    orc_domain_limits(cfg->domain, lo, hi);
    if (cfg->basis == RSRL_FOURIER) {
        /* Fourier::project: scaled = (v - lo)/(hi - lo); phi_k = cos(PI * fold(0, acc + c*v)) */
        double scaled[RSRL_MAX_DIM];
        for (int d = 0; d < D; ++d) scaled[d] = (s[d] - lo[d]) / (hi[d] - lo[d]);
        for (int64_t k = 0; k < F - 1; ++k) {
            coef_row(cfg->basis_order, D, k, c);
            double cx = 0.0;
            for (int d = 0; d < D; ++d) cx = cx + (double)c[d] * scaled[d];
            phi[k] = cos(PI_ * cx);
        }
        phi[F - 1] = 1.0; /* Combinators::with_bias(): constant feature stacked last */
    } else if (cfg->basis == RSRL_POLYNOMIAL) {
        /* Polynomial::project: prod_d v_d.powi(e_d) on the raw state (project-defined ordering = Fourier's) */
        for (int64_t k = 0; k < F - 1; ++k) {
            coef_row(cfg->basis_order, D, k, c);
            double p = 1.0;
            for (int d = 0; d < D; ++d) { double pw = 1.0; for (int j = 0; j < c[d]; ++j) pw = pw * s[d]; p = p * pw; }
            phi[k] = p;
        }
        phi[F - 1] = 1.0;
    } else {
        int32_t idx[64];
        int n = orc_tile_indices(cfg, s, idx);
        memset(phi, 0, (size_t)F * sizeof(double));
        for (int j = 0; j < n; ++j) phi[idx[j]] = 1.0;
    }
}

static double* project_alloc(const rsrl_config_t* cfg, const double* s) {
    double* phi = (double*)malloc((size_t)orc_basis_n_features(cfg) * sizeof(double)); /* Features::Dense(Array1) */
    orc_basis_project(cfg, s, phi);
    return phi;
}

/* ------------------------------------------------------------------------- */
/* LFA (rsrl/src/fa/linear.rs:303-324,353-391 over lfa::LFA; SGD = w += (lr*err)*phi) */
/* ------------------------------------------------------------------------- */
void orc_lfa_evaluate(const rsrl_config_t* cfg, const double* W, int A, const double* s, double* q) {
    int64_t F = orc_basis_n_features(cfg);
    double* phi = project_alloc(cfg, s);
    for (int a = 0; a < A; ++a) { /* phi^T W, one dot per column, index order */
        double acc = 0.0;
        for (int64_t k = 0; k < F; ++k) acc = acc + phi[k] * W[k * A + a];
        q[a] = acc;
    }
    free(phi);
}

static double* evaluate_alloc(const rsrl_config_t* cfg, const double* W, int A, const double* s) {
    double* q = (double*)malloc((size_t)A * sizeof(double)); /* Vec<f64> (fa/linear.rs:310 into_raw_vec) */
    orc_lfa_evaluate(cfg, W, A, s, q);
    return q;
}

/* LFA::update_index -> SGD::step: W[:,a] += (lr*err) * phi(s); `lr_err` is the product */
void orc_lfa_update_index(const rsrl_config_t* cfg, double* W, int A, const double* s, int a, double lr_err) {
    int64_t F = orc_basis_n_features(cfg);
    double* phi = project_alloc(cfg, s);
    for (int64_t k = 0; k < F; ++k) W[k * A + a] = W[k * A + a] + lr_err * phi[k];
    free(phi);
}

/* ------------------------------------------------------------------------- */
/* argmax family                                                               */
/* ------------------------------------------------------------------------- */
int orc_argmaxima(const double* v, int n, int* ixs, double* max_out) { /* utils.rs:6-21 */
    double max = -DBL_MAX; /* f64::MIN */
    int cnt = 0;
    for (int i = 0; i < n; ++i) {
        if (fabs(v[i] - max) < 1e-7) ixs[cnt++] = i;       /* joining does not raise max */
        else if (v[i] > max) { max = v[i]; cnt = 0; ixs[cnt++] = i; }
    }
    if (max_out) *max_out = max;
    return cnt;
}

int orc_find_max(const double* v, int n, double* max_out) { /* core.rs:96-105 */
    int idx = 0; double x = v[0];
    for (int i = 1; i < n; ++i) { if (x > v[i]) { /* keep acc */ } else { idx = i; x = v[i]; } }
    if (max_out) *max_out = x;
    return idx;
}

int orc_argmax_first(const double* v, int n, double* max_out) { /* utils.rs:23-34 */
    int idx = 0; double x = -DBL_MAX;
    for (int j = 0; j < n; ++j) if (v[j] - x > 1e-7) { idx = j; x = v[j]; }
    if (max_out) *max_out = x;
    return idx;
}

/* ------------------------------------------------------------------------- */
/* Philox4x32-10 (Salmon et al., SC'11; Random123 constants)                   */
/* ------------------------------------------------------------------------- */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

/* key = seed (lo, hi); counter = (global env id, draw lo, draw hi, stream) */
void orc_draw(uint64_t seed, uint64_t env, uint64_t draw, uint32_t stream, uint32_t out[4]) {
    uint32_t ctr[4] = {(uint32_t)env, (uint32_t)draw, (uint32_t)(draw >> 32), stream};
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    orc_philox4x32_10(ctr, key, out);
}

enum { STREAM_INIT = 0, STREAM_BEHAVIOUR = 1, STREAM_TARGET = 2 };

/* ------------------------------------------------------------------------- */
/* policies                                                                    */
/* ------------------------------------------------------------------------- */
/* Greedy::evaluate (greedy.rs:30-44), EpsilonGreedy::evaluate (epsilon_greedy.rs:38-45), Random (random.rs) */
/* softmax.rs:15-36 softmax_stable */
static void orc_softmax(const double* q, int n, double tau, double* p) {
    double c = NAN, z = 0.0;
    for (int i = 0; i < n; ++i) c = fmax(c, q[i]);            /* fold(f64::NAN, f64::max) */
    for (int i = 0; i < n; ++i) { p[i] = exp((q[i] - c) / tau); z += p[i]; }
    for (int i = 0; i < n; ++i) p[i] = fmin(p[i] / z, DBL_MAX);
}

void orc_policy_probs(int policy, double epsilon, const double* q, int n, double* p) {
    int ixs[16];
    if (policy == RSRL_SOFTMAX) { orc_softmax(q, n, epsilon, p); return; } /* epsilon carries tau */
    if (policy == RSRL_RANDOM) { for (int i = 0; i < n; ++i) p[i] = 1.0 / (double)n; return; }
    for (int i = 0; i < n; ++i) p[i] = 0.0;
    int cnt = orc_argmaxima(q, n, ixs, NULL);
    double pg = 1.0 / (double)cnt;
    for (int j = 0; j < cnt; ++j) p[ixs[j]] = pg;
    if (policy == RSRL_EPSILON_GREEDY) {
        double pr = epsilon / (double)n;
        for (int i = 0; i < n; ++i) p[i] = pr + p[i] * (1.0 - epsilon);
    }
}

/* Policy::sample.  RNG use (project-defined counter-based stream):
 * rnd[0] -> gen_bool(eps) (epsilon_greedy.rs:75), rnd[1] -> Uniform::new(0,n) (random.rs:44),
 * rnd[2] -> SliceRandom::choose among the maxima (utils.rs:70-76, only consulted on ties). */
int orc_policy_sample(int policy, double epsilon, const double* q, int n, const uint32_t rnd[4], int* nonfinite) {
    int ixs[16];
    if (policy == RSRL_SOFTMAX) { /* softmax.rs:137-139 -> policies/mod.rs:46-61: r = gen::<f64>() (53 bits: rnd[0] high, rnd[1] low) */
        double p[16], cum = 0.0;
        double r = (double)((((uint64_t)rnd[0] << 32) | (uint64_t)rnd[1]) >> 11) * (1.0 / 9007199254740992.0);
        orc_softmax(q, n, epsilon, p);
        for (int i = 0; i < n; ++i) { cum = cum + p[i]; if (cum > r) return i; }
        return n - 1;
    }
    int explore = policy == RSRL_RANDOM;
    if (policy == RSRL_EPSILON_GREEDY)
        explore = epsilon >= 1.0 || rnd[0] < (uint32_t)(epsilon * 4294967296.0);
    if (explore) return (int)(((uint64_t)rnd[1] * (uint64_t)n) >> 32);
    int cnt = orc_argmaxima(q, n, ixs, NULL);
    if (cnt == 0) { if (nonfinite) *nonfinite = 1; return 0; } /* reference panics (utils.rs:76) */
    if (cnt == 1) return ixs[0];
    return ixs[(int)(((uint64_t)rnd[2] * (uint64_t)cnt) >> 32)];
}

/* ------------------------------------------------------------------------- */
/* traces.rs:196-240 over a dense buffer.  Columnar (params/columnar.rs:97-109) stores only
 * visited columns; since every rule maps (0,0) -> 0 a dense F x A buffer is equivalent. */
/* ------------------------------------------------------------------------- */
void orc_trace_update(int rule, double gamma, double lambda, double alpha, int64_t n, double* z, const double* grad) {
    if (rule == RSRL_TRACE_ACCUMULATE) {
        double rate = gamma * lambda;
        for (int64_t i = 0; i < n; ++i) z[i] = rate * z[i] + grad[i];
    } else if (rule == RSRL_TRACE_REPLACE) { /* Saturate */
        double rate = gamma * lambda;
        for (int64_t i = 0; i < n; ++i) z[i] = fmax(-1.0, fmin(1.0, rate * z[i] + grad[i]));
    } else { /* Dutch */
        double rate = gamma * lambda * (1.0 - alpha);
        for (int64_t i = 0; i < n; ++i) z[i] = rate * z[i] + grad[i];
    }
}

/* ------------------------------------------------------------------------- */
/* batched engine                                                              */
/* ------------------------------------------------------------------------- */
struct orc_engine {
    rsrl_config_t cfg;
    int D, A /* domain actions */, AW /* weight columns */;
    int64_t N, NG, F, t;
    int has_trace;
    double step_scale; /* 1/scale_div applied to every update (1.0 unless SHARED+MEAN) */
    double* s; int32_t* a; int32_t* ep; double* W; double* z; double* td; double* G;
    double* W2; double* G2; /* second weight table of GreedyGQ (fa_td) / A2C (the policy's LFA) and its per-step dW accumulator */
    int32_t* n_ep; int32_t* last_len; uint64_t* len_hash;
    rsrl_stats_t st;
    double min_gap;
};

static int is_td_pred(int algo) { return algo == RSRL_TD_LAMBDA || algo == RSRL_TD0; }
static int is_trace(int algo) { return algo == RSRL_SARSA_LAMBDA || algo == RSRL_Q_LAMBDA || algo == RSRL_TD_LAMBDA; }

orc_engine_t* orc_engine_create(const rsrl_config_t* cfg) {
    orc_engine_t* e = (orc_engine_t*)calloc(1, sizeof(*e));
    e->cfg = *cfg;
    e->D = orc_domain_dim(cfg->domain);
    e->A = orc_domain_n_actions(cfg->domain);
    e->AW = is_td_pred(cfg->algo) ? 1 : e->A;
    e->N = cfg->n_envs;
    e->NG = cfg->n_envs_global > 0 ? cfg->n_envs_global : cfg->n_envs;
    e->F = orc_basis_n_features(cfg);
    e->has_trace = is_trace(cfg->algo);
    e->step_scale = (cfg->weight_mode == RSRL_SHARED && cfg->update_scale == RSRL_SCALE_MEAN) ? (double)e->NG : 1.0;
    int64_t wn = e->F * e->AW * (cfg->weight_mode == RSRL_PER_ENV ? e->N : 1);
    e->s = (double*)calloc((size_t)(e->N * e->D), sizeof(double));
    e->a = (int32_t*)calloc((size_t)e->N, sizeof(int32_t));
    e->ep = (int32_t*)calloc((size_t)e->N, sizeof(int32_t));
    e->W = (double*)calloc((size_t)wn, sizeof(double));
    e->G = (double*)calloc((size_t)(e->F * e->AW), sizeof(double));
    e->W2 = (double*)calloc((size_t)wn, sizeof(double));
    e->G2 = (double*)calloc((size_t)(e->F * e->AW), sizeof(double));
    e->z = e->has_trace ? (double*)calloc((size_t)(e->N * e->F * e->AW), sizeof(double)) : NULL;
    e->td = (double*)calloc((size_t)e->N, sizeof(double));
    e->n_ep = (int32_t*)calloc((size_t)e->N, sizeof(int32_t));
    e->last_len = (int32_t*)calloc((size_t)e->N, sizeof(int32_t));
    e->len_hash = (uint64_t*)calloc((size_t)e->N, sizeof(uint64_t));
    orc_engine_reset(e, NULL);
    return e;
}

void orc_engine_destroy(orc_engine_t* e) {
    if (!e) return;
    free(e->s); free(e->a); free(e->ep); free(e->W); free(e->G); free(e->W2); free(e->G2); free(e->z); free(e->td);
    free(e->n_ep); free(e->last_len); free(e->len_hash); free(e);
}

/* start state of the episode that begins at batched step `t` for global env `g` */
static void fresh_state(const orc_engine_t* e, int64_t g, int64_t t, double* s) {
    if (e->cfg.init_mode == RSRL_INIT_DEFAULT) { orc_domain_default(e->cfg.domain, s); return; }
    uint32_t r[4];
    orc_draw(e->cfg.seed, (uint64_t)g, (uint64_t)t, STREAM_INIT, r);
    for (int d = 0; d < e->D; ++d)
        s[d] = e->cfg.init_lo[d] + (e->cfg.init_hi[d] - e->cfg.init_lo[d]) * ((double)r[d] * (1.0 / 4294967296.0));
}

void orc_engine_reset(orc_engine_t* e, const double* init_states) {
    int64_t wn = e->F * e->AW * (e->cfg.weight_mode == RSRL_PER_ENV ? e->N : 1);
    memset(e->W, 0, (size_t)wn * sizeof(double)); /* LFA::vector => Array2::zeros (examples/q_learning.rs:25) */
    memset(e->W2, 0, (size_t)wn * sizeof(double));
    if (e->z) memset(e->z, 0, (size_t)(e->N * e->F * e->AW) * sizeof(double));
    memset(e->td, 0, (size_t)e->N * sizeof(double));
    memset(e->ep, 0, (size_t)e->N * sizeof(int32_t));
    memset(e->n_ep, 0, (size_t)e->N * sizeof(int32_t));
    memset(e->last_len, 0, (size_t)e->N * sizeof(int32_t));
    memset(e->len_hash, 0, (size_t)e->N * sizeof(uint64_t));
    memset(&e->st, 0, sizeof(e->st));
    e->t = 0;
    e->min_gap = INFINITY;
    for (int64_t i = 0; i < e->N; ++i) {
        e->a[i] = -1;
        if (init_states) memcpy(e->s + i * e->D, init_states + i * e->D, (size_t)e->D * sizeof(double));
        else fresh_state(e, e->cfg.env_offset + i, 0, e->s + i * e->D);
    }
}

static void note_gap(orc_engine_t* e, const double* q, int n) {
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            double d = fabs(q[i] - q[j]);
            if (d == 0.0 || d != d) continue; /* exact ties are reproduced exactly by both sides */
            double m = fmin(d, fabs(d - 1e-7));
            if (m < e->min_gap) e->min_gap = m;
        }
}

/* apply `coef * vec` to column `a` (vec = phi) or to all columns (vec = trace z) */
static void accumulate_col(orc_engine_t* e, double* dst, int a, double coef, const double* phi) {
    for (int64_t k = 0; k < e->F; ++k) dst[k * e->AW + a] = dst[k * e->AW + a] + coef * phi[k];
}

/* One reference agent `handle(&Transition)` for env slot i (i < 0: stateless batch call, no traces).
 * W  = weights every evaluate() reads (W_t);  dst = where the update lands (W itself for PER_ENV /
 * N = 1 — exactly the reference — or the step's dW accumulator in SHARED mode). */
/* softmax.rs:15-36,75-81: the policy's probabilities from its own LFA (theta) */
static void gibbs_probs(const rsrl_config_t* c, const double* theta, int A, const double* s, double* p) {
    double* h = evaluate_alloc(c, theta, A, s);
    orc_policy_probs(RSRL_SOFTMAX, c->epsilon, h, A, p);
    free(h);
}

/* GreedyGQ::handle — control/td/greedy_gq.rs:73-141.  fa_q = W (SGD lr), fa_td = W2 (SGD alpha). */
static double greedy_gq_handle(orc_engine_t* e, const double* W, double* dst, const double* W2, double* dst2,
                               const double* from, int a, double r, const double* to, int terminated) {
    const rsrl_config_t* c = &e->cfg;
    const int A = e->AW;
    double* qs = evaluate_alloc(c, W, A, from);
    double qsa = qs[a];                              /* :77 fa_q.evaluate_index((s,), a) */
    double* vs = evaluate_alloc(c, W2, A, from);
    double td_est = vs[a];                           /* :78 fa_td.evaluate((s, a)) */
    double td_error;
    free(qs); free(vs);
    if (terminated) {
        td_error = r - qsa;                          /* :81 */
        double* phi = project_alloc(c, from);
        accumulate_col(e, dst, a, (c->lr * td_error) / e->step_scale, phi);                /* :83-89 */
        accumulate_col(e, dst2, a, (c->alpha * (td_error - td_est)) / e->step_scale, phi); /* :91-97 */
        free(phi);
    } else {
        double* nq = evaluate_alloc(c, W, A, to);
        double qnsna;
        int na = orc_find_max(nq, A, &qnsna);        /* :107 */
        free(nq);
        td_error = r + c->gamma * qnsna - qsa;       /* :109 */
        double* phi = project_alloc(c, from);
        double* nphi = project_alloc(c, to);
        accumulate_col(e, dst, a, (c->lr * td_error) / e->step_scale, phi);                /* :111-116 */
        accumulate_col(e, dst, na, (c->lr * (-c->gamma * td_est)) / e->step_scale, nphi);  /* :117-124 */
        accumulate_col(e, dst2, a, (c->alpha * (td_error - td_est)) / e->step_scale, phi); /* :127-133 */
        free(phi); free(nphi);
    }
    return td_error;
}

/* One iteration of examples/a2c.rs:55-58: eval.handle(&t) (SARSA critic, control/td/sarsa.rs:53-75, with the Gibbs policy) then
 * agent.handle(&t) (control/ac.rs:100-114 with the critic closure of a2c.rs:37-46 and Softmax::handle, softmax.rs:113-129,146-160).
 * W = the critic's Q (SGD lr), W2 = the policy's LFA theta.  W_own / W2 are what evaluate() reads; in SHARED mode the critic
 * closure sees W plus THIS transition's own critic update (for N = 1 that is exactly the reference's in-place update). */
static double a2c_handle(orc_engine_t* e, int64_t g, uint64_t draw, const double* W, double* dst, const double* W2, double* dst2,
                         const double* from, int a, double r, const double* to, int terminated) {
    const rsrl_config_t* c = &e->cfg;
    const int A = e->AW;
    const int64_t F = e->F;
    double* qs = evaluate_alloc(c, W, A, from);
    double qsa = qs[a], residual;                    /* sarsa.rs:56 */
    if (terminated) {
        residual = r - qsa;                          /* sarsa.rs:58 */
    } else {
        uint32_t rnd[4]; int nf = 0;
        double p[16];
        orc_draw(c->seed, (uint64_t)g, draw, STREAM_TARGET, rnd);
        gibbs_probs(c, W2, A, to, p);
        double* h = evaluate_alloc(c, W2, A, to);
        int na = orc_policy_sample(RSRL_SOFTMAX, c->epsilon, h, A, rnd, &nf); /* sarsa.rs:61 policy.sample(thread_rng, ns) */
        free(h);
        double* nq = evaluate_alloc(c, W, A, to);
        residual = r + c->gamma * nq[na] - qsa;      /* sarsa.rs:62-64 */
        free(nq);
    }
    double* phi = project_alloc(c, from);
    const double cq = (c->lr * residual) / e->step_scale;
    /* critic closure (a2c.rs:40-45) on the updated Q: only column a of Q(s) changed (W may alias dst: read it before the update lands) */
    {
        double acc = 0.0;
        for (int64_t k = 0; k < F; ++k) acc = acc + phi[k] * (W[k * A + a] + cq * phi[k]);
        qs[a] = acc;
    }
    accumulate_col(e, dst, a, cq, phi);              /* sarsa.rs:67-73 -> SGD */
    double ps[16], ev = 0.0;
    gibbs_probs(c, W2, A, from, ps);
    for (int j = 0; j < A; ++j) ev = ev + qs[j] * ps[j];
    const double adv = qs[a] - ev;
    const double error = (c->alpha * adv) / e->step_scale;   /* ac.rs:109-113 */
    /* Softmax::grad_log (softmax.rs:113-129): sf = pi(s); sf[a] -= 1; jac[:, col] = (-sf[col]) * phi; theta += error * jac (softmax.rs:146-160,
     * ScaledGradientUpdate bypasses the optimiser) */
    ps[a] = ps[a] - 1.0;
    for (int col = 0; col < A; ++col)
        for (int64_t k = 0; k < F; ++k) dst2[k * A + col] = dst2[k * A + col] + error * ((-ps[col]) * phi[k]);
    free(phi); free(qs);
    return residual;
}

static double agent_handle(orc_engine_t* e, int64_t i, int64_t g, uint64_t draw, const double* W, double* dst,
                           const double* from, int a, double r, const double* to, int terminated) {
    const rsrl_config_t* c = &e->cfg;
    const int A = e->AW;
    double residual, err;
    double* z = (e->has_trace && i >= 0) ? e->z + i * e->F * A : NULL;

    if (c->algo == RSRL_GREEDY_GQ || c->algo == RSRL_A2C) {
        const int per_env = c->weight_mode == RSRL_PER_ENV;
        const int64_t off = per_env ? (W - e->W) : 0;
        const double* W2 = e->W2 + off;
        double* dst2 = per_env ? e->W2 + off : e->G2;
        if (c->algo == RSRL_GREEDY_GQ) return greedy_gq_handle(e, W, dst, W2, dst2, from, a, r, to, terminated);
        return a2c_handle(e, g, draw, W, dst, W2, dst2, from, a, r, to, terminated);
    }

    if (c->algo == RSRL_TD0 || c->algo == RSRL_TD_LAMBDA) {
        /* prediction/td/td.rs:38-58, td_lambda.rs:44-77 */
        double pred, nv;
        orc_lfa_evaluate(c, W, 1, from, &pred);
        if (c->algo == RSRL_TD_LAMBDA && z) {
            double* grad = project_alloc(c, from); /* fa_theta.grad((from,)) */
            orc_trace_update(c->trace_rule, c->gamma, c->lambda, c->alpha, e->F, z, grad);
            free(grad);
        }
        if (terminated) residual = r - pred;
        else { orc_lfa_evaluate(c, W, 1, to, &nv); residual = r + c->gamma * nv - pred; }
        if (c->algo == RSRL_TD0) {
            double* phi = project_alloc(c, from); /* ScalarLFA update -> SGD */
            accumulate_col(e, dst, 0, (c->lr * residual) / e->step_scale, phi);
            free(phi);
        } else if (z) {
            /* ScaledGradientUpdate{alpha: td_error} bypasses the optimiser: W += td_error * z (td_lambda.rs:56-59) */
            accumulate_col(e, dst, 0, residual / e->step_scale, z);
            if (terminated) memset(z, 0, (size_t)e->F * sizeof(double));
        }
        return residual;
    }

    double* qs = evaluate_alloc(c, W, A, from); /* Shared<LFA>::evaluate_index = evaluate(args)[index] (core.rs:79-83) */
    double qsa = qs[a];
    int a_star = orc_argmax_first(qs, A, NULL); /* pal.rs:47 */
    double q_astar = qs[a_star];

    if (c->algo == RSRL_Q_LAMBDA && z) { /* q_lambda.rs:68 */
        if (a != orc_argmax_first(qs, A, NULL)) memset(z, 0, (size_t)(e->F * A) * sizeof(double));
    }
    free(qs);
    if ((c->algo == RSRL_SARSA_LAMBDA || c->algo == RSRL_Q_LAMBDA) && z) {
        /* trace.update(&grad(s,a)): Jacobian = Columnar::from_column(a, phi(s)) (fa/linear.rs:334-339) */
        double* phi = project_alloc(c, from);
        double* grad = (double*)calloc((size_t)(e->F * A), sizeof(double));
        for (int64_t k = 0; k < e->F; ++k) grad[k * A + a] = phi[k];
        orc_trace_update(c->trace_rule, c->gamma, c->lambda, c->alpha, e->F * A, z, grad);
        free(grad); free(phi);
    }

    if (terminated) {
        residual = r - qsa;
    } else {
        double* nq = evaluate_alloc(c, W, A, to);
        if (c->algo == RSRL_QLEARNING || c->algo == RSRL_Q_LAMBDA) {
            double mx; orc_find_max(nq, A, &mx);          /* q_learning.rs:59, q_lambda.rs:86 */
            residual = r + c->gamma * mx - qsa;
        } else if (c->algo == RSRL_SARSA || c->algo == RSRL_SARSA_LAMBDA) {
            uint32_t rnd[4]; int nf = 0;                     /* sarsa.rs:61: policy.sample(&mut thread_rng(), ns) */
            orc_draw(c->seed, (uint64_t)g, draw, STREAM_TARGET, rnd);
            double* nq2 = evaluate_alloc(c, W, A, to);       /* the policy evaluates Q(s') itself */
            int na = orc_policy_sample(c->policy, c->epsilon, nq2, A, rnd, &nf);
            if (c->policy != RSRL_RANDOM) note_gap(e, nq2, A);
            if (nf) e->st.nonfinite = 1;
            free(nq2);
            residual = r + c->gamma * nq[na] - qsa;
        } else if (c->algo == RSRL_PAL) { /* pal.rs:44-52: td_error bootstraps from nqs[a_star] (a_star = argmax_first Q(s)) */
            int na_star = orc_argmax_first(nq, A, NULL);
            double td_error = r + c->gamma * nq[a_star] - qsa;
            double al_error = td_error - c->alpha * (q_astar - qsa);
            double alt = td_error - c->alpha * (nq[na_star] - nq[a]);
            residual = fmax(al_error, alt);
        } else { /* expected_sarsa.rs:52-58 */
            double p[16], exp_nv = 0.0;
            double* nq2 = evaluate_alloc(c, W, A, to);
            orc_policy_probs(c->policy, c->epsilon, nq2, A, p);
            if (c->policy != RSRL_RANDOM) note_gap(e, nq2, A);
            free(nq2);
            for (int j = 0; j < A; ++j) exp_nv = exp_nv + nq[j] * p[j];
            residual = r + c->gamma * exp_nv - qsa;
        }
        free(nq);
    }

    if (c->algo == RSRL_SARSA_LAMBDA || c->algo == RSRL_Q_LAMBDA) {
        if (z) {
            /* ScaledGradientUpdate{alpha: alpha*residual, jacobian: &trace}: W += (alpha*residual) * z */
            double coef = (c->alpha * residual) / e->step_scale;
            for (int64_t j = 0; j < e->F * A; ++j) dst[j] = dst[j] + coef * z[j];
            if (terminated) memset(z, 0, (size_t)(e->F * A) * sizeof(double));
        }
        return residual;
    }
    err = (c->algo == RSRL_EXPECTED_SARSA || c->algo == RSRL_PAL) ? c->alpha * residual : residual; /* expected_sarsa.rs:64, pal.rs:57 */
    {
        double* phi = project_alloc(c, from); /* update_index re-projects (fa/linear.rs:389) */
        accumulate_col(e, dst, a, (c->lr * err) / e->step_scale, phi);
        free(phi);
    }
    return residual;
}

/* One iteration of the trajectory loop (examples/q_learning.rs:40-52) for env i. */
static void env_step(orc_engine_t* e, int64_t i) {
    const rsrl_config_t* c = &e->cfg;
    const int64_t g = c->env_offset + i;
    double* s = e->s + i * e->D;
    const double* W = c->weight_mode == RSRL_PER_ENV ? e->W + i * e->F * e->AW : e->W;
    double* dst = c->weight_mode == RSRL_PER_ENV ? e->W + i * e->F * e->AW : e->G;
    uint32_t rnd[4];
    int nf = 0, a, terminated;
    double r;

    /* action = policy.sample(&mut rng, state) (:38 / :47) */
    orc_draw(c->seed, (uint64_t)g, (uint64_t)e->t, STREAM_BEHAVIOUR, rnd);
    if (c->policy == RSRL_RANDOM || is_td_pred(c->algo)) {
        a = orc_policy_sample(RSRL_RANDOM, 0.0, NULL, e->A, rnd, &nf);
    } else if (c->algo == RSRL_A2C) { /* agent.policy.sample(&mut rng, s): Gibbs over the policy's own LFA (a2c.rs:50,60) */
        const double* W2 = c->weight_mode == RSRL_PER_ENV ? e->W2 + i * e->F * e->AW : e->W2;
        double* h = evaluate_alloc(c, W2, e->AW, s);
        a = orc_policy_sample(RSRL_SOFTMAX, c->epsilon, h, e->AW, rnd, &nf);
        free(h);
    } else {
        double* q = evaluate_alloc(c, W, e->AW, s);
        a = orc_policy_sample(c->policy, c->epsilon, q, e->AW, rnd, &nf);
        note_gap(e, q, e->AW);
        free(q);
    }
    if (nf) e->st.nonfinite = 1;

    /* t = env.transition(action) (:44; lib.rs:436-446: from = emit(), (to, r) = step(a)) */
    double* from = (double*)malloc((size_t)e->D * sizeof(double)); /* emit() -> vec![...] */
    memcpy(from, s, (size_t)e->D * sizeof(double));
    orc_domain_step(c->domain, s, a, &r, &terminated);
    double* to = (double*)malloc((size_t)e->D * sizeof(double));   /* emit() after the step */
    memcpy(to, s, (size_t)e->D * sizeof(double));

    /* agent.handle(&t) (:46) */
    e->td[i] = agent_handle(e, i, g, (uint64_t)e->t, W, dst, from, a, r, to, terminated);
    free(from); free(to);

    e->a[i] = a;
    e->ep[i] += 1;
    e->st.total_steps += 1;
    if (terminated || (c->max_episode_steps > 0 && e->ep[i] >= c->max_episode_steps)) {
        /* `if t.terminated() { break }` (:49-51) then `env = MountainCar::default()` (:37) */
        e->st.total_episodes += 1;
        e->st.terminal_episodes += terminated ? 1 : 0;
        e->n_ep[i] += 1;
        e->last_len[i] = e->ep[i];
        e->len_hash[i] = e->len_hash[i] * 1000003ull + (uint64_t)e->ep[i];
        e->ep[i] = 0;
        fresh_state(e, g, e->t + 1, s);
    }
}

void orc_engine_step_local(orc_engine_t* e, double* dW_out) {
    memset(e->G, 0, (size_t)(e->F * e->AW) * sizeof(double));
    memset(e->G2, 0, (size_t)(e->F * e->AW) * sizeof(double));
    for (int64_t i = 0; i < e->N; ++i) env_step(e, i);
    if (dW_out) memcpy(dW_out, e->G, (size_t)(e->F * e->AW) * sizeof(double));
}

void orc_engine_step_apply(orc_engine_t* e, const double* dW_sum) {
    if (e->cfg.weight_mode == RSRL_SHARED) {
        const double* g = dW_sum ? dW_sum : e->G;
        for (int64_t j = 0; j < e->F * e->AW; ++j) e->W[j] = e->W[j] + g[j];
        for (int64_t j = 0; j < e->F * e->AW; ++j) e->W2[j] = e->W2[j] + e->G2[j];
    }
    e->t += 1;
    e->st.batch_steps += 1;
}

void orc_engine_step(orc_engine_t* e, int64_t k) {
    for (int64_t j = 0; j < k; ++j) { orc_engine_step_local(e, NULL); orc_engine_step_apply(e, NULL); }
}

void orc_engine_handle(orc_engine_t* e, int64_t n, const double* from, const int32_t* actions, const double* rewards,
                       const double* to, const uint8_t* terminal, uint64_t draw, double* td_out) {
    memset(e->G, 0, (size_t)(e->F * e->AW) * sizeof(double));
    memset(e->G2, 0, (size_t)(e->F * e->AW) * sizeof(double));
    for (int64_t i = 0; i < n; ++i) {
        int per_env = e->cfg.weight_mode == RSRL_PER_ENV;
        const double* W = per_env ? e->W + i * e->F * e->AW : e->W;
        double* dst = per_env ? e->W + i * e->F * e->AW : e->G;
        double td = agent_handle(e, e->has_trace ? i : -1, e->cfg.env_offset + i, draw, W, dst, from + i * e->D,
                                 actions[i], rewards[i], to + i * e->D, terminal[i]);
        if (td_out) td_out[i] = td;
    }
    if (e->cfg.weight_mode == RSRL_SHARED) {
        for (int64_t j = 0; j < e->F * e->AW; ++j) e->W[j] = e->W[j] + e->G[j];
        for (int64_t j = 0; j < e->F * e->AW; ++j) e->W2[j] = e->W2[j] + e->G2[j];
    }
}

void orc_engine_get_aux_weights(orc_engine_t* e, double* out) {
    memcpy(out, e->W2, (size_t)(e->F * e->AW * (e->cfg.weight_mode == RSRL_PER_ENV ? e->N : 1)) * sizeof(double));
}
void orc_engine_set_aux_weights(orc_engine_t* e, const double* in) {
    memcpy(e->W2, in, (size_t)(e->F * e->AW * (e->cfg.weight_mode == RSRL_PER_ENV ? e->N : 1)) * sizeof(double));
}

/* Domain::rollout (rsrl_domains/src/lib.rs:448-479) for env slot i of the engine's config with the engine's current weights held
 * fixed.  pi = Policy::mode when greedy (greedy.rs:83: find_max; softmax.rs:141: argmax_first over the probabilities; A2C: the
 * policy table) else Policy::sample with draw index draw + j.  Returns Trajectory::n_transitions(). */
int64_t orc_engine_rollout(orc_engine_t* e, int64_t i, const double* init_state, int64_t step_limit, int greedy, uint64_t draw,
                           double* start_out, double* next_out, int32_t* actions_out, double* rewards_out, uint8_t* terminal_out) {
    const rsrl_config_t* c = &e->cfg;
    const int64_t g = c->env_offset + i;
    const double* Wq = c->weight_mode == RSRL_PER_ENV ? e->W + i * e->F * e->AW : e->W;
    const double* Wp = c->algo == RSRL_A2C ? (c->weight_mode == RSRL_PER_ENV ? e->W2 + i * e->F * e->AW : e->W2) : Wq;
    const int policy = c->algo == RSRL_A2C ? RSRL_SOFTMAX : c->policy;
    const int64_t take = step_limit - 1 > 0 ? step_limit - 1 : 0; /* iter.take(sl.saturating_sub(1)), lib.rs:472-476 */
    double s[RSRL_MAX_DIM];
    int64_t n = 0;
    int terminated = 0;
    if (init_state) memcpy(s, init_state, (size_t)e->D * sizeof(double));
    else fresh_state(e, g, (int64_t)draw, s);
    memcpy(start_out, s, (size_t)e->D * sizeof(double)); /* let start = self.emit() */
    for (int64_t j = 0;; ++j) {
        if (j > 0 && (terminated || j >= take)) break;  /* successors: Observation::Terminal => None; .take(sl - 1) */
        double* q = evaluate_alloc(c, Wp, e->AW, s);
        int a, nf = 0;
        if (greedy) {
            if (policy == RSRL_SOFTMAX) { double p[16]; orc_policy_probs(RSRL_SOFTMAX, c->epsilon, q, e->AW, p); a = orc_argmax_first(p, e->AW, NULL); }
            else a = orc_find_max(q, e->AW, NULL);
        } else {
            uint32_t rnd[4];
            orc_draw(c->seed, (uint64_t)g, draw + (uint64_t)j, STREAM_BEHAVIOUR, rnd);
            a = orc_policy_sample(policy, c->epsilon, q, e->AW, rnd, &nf);
        }
        free(q);
        double r;
        orc_domain_step(c->domain, s, a, &r, &terminated);   /* the first step is executed even when nothing is recorded (lib.rs:460-462) */
        if (j >= take) break;
        memcpy(next_out + j * e->D, s, (size_t)e->D * sizeof(double));
        actions_out[j] = a; rewards_out[j] = r; terminal_out[j] = (uint8_t)terminated;
        n = j + 1;
    }
    return n;
}

void orc_engine_get_states(orc_engine_t* e, double* out) { memcpy(out, e->s, (size_t)(e->N * e->D) * sizeof(double)); }
void orc_engine_set_states(orc_engine_t* e, const double* in) { memcpy(e->s, in, (size_t)(e->N * e->D) * sizeof(double)); }
void orc_engine_get_actions(orc_engine_t* e, int32_t* out) { memcpy(out, e->a, (size_t)e->N * sizeof(int32_t)); }
void orc_engine_get_episode_steps(orc_engine_t* e, int32_t* out) { memcpy(out, e->ep, (size_t)e->N * sizeof(int32_t)); }
static int64_t wcount(orc_engine_t* e) { return e->F * e->AW * (e->cfg.weight_mode == RSRL_PER_ENV ? e->N : 1); }
void orc_engine_get_weights(orc_engine_t* e, double* out) { memcpy(out, e->W, (size_t)wcount(e) * sizeof(double)); }
void orc_engine_set_weights(orc_engine_t* e, const double* in) { memcpy(e->W, in, (size_t)wcount(e) * sizeof(double)); }
void orc_engine_get_traces(orc_engine_t* e, double* out) { if (e->z) memcpy(out, e->z, (size_t)(e->N * e->F * e->AW) * sizeof(double)); }
void orc_engine_set_traces(orc_engine_t* e, const double* in) { if (e->z) memcpy(e->z, in, (size_t)(e->N * e->F * e->AW) * sizeof(double)); }
void orc_engine_get_td_errors(orc_engine_t* e, double* out) { memcpy(out, e->td, (size_t)e->N * sizeof(double)); }
void orc_engine_get_stats(orc_engine_t* e, rsrl_stats_t* out) { *out = e->st; }
void orc_engine_get_env_stats(orc_engine_t* e, int32_t* n_episodes, int32_t* last_len, uint64_t* len_hash) {
    if (n_episodes) memcpy(n_episodes, e->n_ep, (size_t)e->N * sizeof(int32_t));
    if (last_len) memcpy(last_len, e->last_len, (size_t)e->N * sizeof(int32_t));
    if (len_hash) memcpy(len_hash, e->len_hash, (size_t)e->N * sizeof(uint64_t));
}
void orc_engine_set_epsilon(orc_engine_t* e, double eps) { e->cfg.epsilon = eps; }
double orc_engine_min_gap(orc_engine_t* e) { return e->min_gap; }

/* ------------------------------------------------------------------------- */
/* CPU baseline: `threads` workers, each a batch of independent reference-shaped
 * single-env agents.  Engines are created and the threads are started BEFORE the clock starts (a barrier
 * releases them together); only stepping is timed. */
/* ------------------------------------------------------------------------- */
typedef struct {
    rsrl_config_t cfg; int64_t steps; int64_t done; int fused;
    pthread_barrier_t* start; pthread_barrier_t* stop;
    int32_t* ep_lens; int n_ep_lens; /* optional: first episode lengths of env 0 (single-env runs) */
} worker_t;

/* The best simple scalar CPU implementation of the same step (not the reference's cost structure): one projection
 * per env-step (phi(s') becomes phi(s)), Q(s_{t+1}) re-evaluated after the update like the reference does, no heap
 * allocation.  Q-learning / SARSA / ExpectedSARSA with PER_ENV weights (N independent agents). */
static void fused_steps(orc_engine_t* e, int64_t steps) {
    const rsrl_config_t* c = &e->cfg;
    const int A = e->AW;
    const int64_t F = e->F;
    double* phi = (double*)malloc((size_t)(e->N * F) * sizeof(double));
    double* nphi = (double*)malloc((size_t)F * sizeof(double));
    for (int64_t i = 0; i < e->N; ++i) orc_basis_project(c, e->s + i * e->D, phi + i * F);
    for (int64_t t = 0; t < steps; ++t) {
        for (int64_t i = 0; i < e->N; ++i) {
            const int64_t g = c->env_offset + i;
            double* W = e->W + i * F * A;
            double* s = e->s + i * e->D;
            double* ph = phi + i * F;
            double q[16], nq[16], r, target = 0.0;
            uint32_t rnd[4];
            int nf = 0, term;
            for (int a = 0; a < A; ++a) { double acc = 0.0; for (int64_t k = 0; k < F; ++k) acc = acc + ph[k] * W[k * A + a]; q[a] = acc; }
            orc_draw(c->seed, (uint64_t)g, (uint64_t)e->t, STREAM_BEHAVIOUR, rnd);
            const int act = orc_policy_sample(c->policy, c->epsilon, q, A, rnd, &nf);
            orc_domain_step(c->domain, s, act, &r, &term);
            if (!term) {
                orc_basis_project(c, s, nphi);
                for (int a = 0; a < A; ++a) { double acc = 0.0; for (int64_t k = 0; k < F; ++k) acc = acc + nphi[k] * W[k * A + a]; nq[a] = acc; }
                if (c->algo == RSRL_QLEARNING) orc_find_max(nq, A, &target);
                else if (c->algo == RSRL_SARSA) { orc_draw(c->seed, (uint64_t)g, (uint64_t)e->t, STREAM_TARGET, rnd); target = nq[orc_policy_sample(c->policy, c->epsilon, nq, A, rnd, &nf)]; }
                else { double p[16]; orc_policy_probs(c->policy, c->epsilon, nq, A, p); for (int a = 0; a < A; ++a) target = target + nq[a] * p[a]; }
            }
            const double residual = term ? r - q[act] : r + c->gamma * target - q[act];
            const double lr_err = c->lr * (c->algo == RSRL_EXPECTED_SARSA ? c->alpha * residual : residual);
            for (int64_t k = 0; k < F; ++k) W[k * A + act] = W[k * A + act] + lr_err * ph[k];
            e->a[i] = act;
            e->ep[i] += 1;
            e->st.total_steps += 1;
            if (term || (c->max_episode_steps > 0 && e->ep[i] >= c->max_episode_steps)) {
                e->st.total_episodes += 1;
                e->n_ep[i] += 1; e->last_len[i] = e->ep[i];
                e->len_hash[i] = e->len_hash[i] * 1000003ull + (uint64_t)e->ep[i];
                e->ep[i] = 0;
                fresh_state(e, g, e->t + 1, s);
                orc_basis_project(c, s, ph);
            } else {
                memcpy(ph, nphi, (size_t)F * sizeof(double));
            }
        }
        e->t += 1;
    }
    free(phi); free(nphi);
}

static void* worker_main(void* p) {
    worker_t* w = (worker_t*)p;
    orc_engine_t* e = orc_engine_create(&w->cfg);
    pthread_barrier_wait(w->start);
    if (w->fused) {
        fused_steps(e, w->steps);
    } else if (w->ep_lens) { /* single env: record the first episode lengths while stepping */
        int got = 0;
        for (int64_t t = 0; t < w->steps; ++t) {
            orc_engine_step(e, 1);
            if (got < w->n_ep_lens && e->n_ep[0] > got) w->ep_lens[got++] = e->last_len[0];
        }
        for (; got < w->n_ep_lens; ++got) w->ep_lens[got] = -1;
    } else {
        orc_engine_step(e, w->steps);
    }
    pthread_barrier_wait(w->stop);
    w->done = e->st.total_steps;
    orc_engine_destroy(e);
    return NULL;
}

static double run_workers(const rsrl_config_t* cfg, int threads, int64_t envs_per_thread, int64_t steps, int fused,
                          int32_t* ep_lens, int n_ep_lens, int64_t* out_steps) {
    pthread_t* th = (pthread_t*)calloc((size_t)threads, sizeof(pthread_t));
    worker_t* w = (worker_t*)calloc((size_t)threads, sizeof(worker_t));
    pthread_barrier_t start, stop;
    struct timespec t0, t1;
    int64_t total = 0;
    pthread_barrier_init(&start, NULL, (unsigned)threads + 1);
    pthread_barrier_init(&stop, NULL, (unsigned)threads + 1);
    for (int i = 0; i < threads; ++i) {
        w[i].cfg = *cfg;
        w[i].cfg.weight_mode = RSRL_PER_ENV; /* one agent per env: the only shape the Rc-based reference admits */
        w[i].cfg.n_envs = envs_per_thread;
        w[i].cfg.env_offset = cfg->env_offset + (int64_t)i * envs_per_thread;
        w[i].steps = steps; w[i].fused = fused; w[i].start = &start; w[i].stop = &stop;
        w[i].ep_lens = i == 0 ? ep_lens : NULL; w[i].n_ep_lens = n_ep_lens;
        pthread_create(&th[i], NULL, worker_main, &w[i]);
    }
    pthread_barrier_wait(&start); /* every engine exists */
    clock_gettime(CLOCK_MONOTONIC, &t0);
    pthread_barrier_wait(&stop);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    for (int i = 0; i < threads; ++i) total += w[i].done;
    if (out_steps) *out_steps = total;
    pthread_barrier_destroy(&start); pthread_barrier_destroy(&stop);
    free(th); free(w);
    return (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
}

double orc_baseline_run(const rsrl_config_t* cfg, int threads, int64_t envs_per_thread, int64_t steps, int64_t* out_steps) {
    return run_workers(cfg, threads, envs_per_thread, steps, 0, NULL, 0, out_steps);
}

double orc_baseline_run_fused(const rsrl_config_t* cfg, int threads, int64_t envs_per_thread, int64_t steps, int64_t* out_steps) {
    return run_workers(cfg, threads, envs_per_thread, steps, 1, NULL, 0, out_steps);
}

/* BASELINE configs[0] = examples/q_learning.rs on one core: `steps` env-steps of ONE env; the first n_ep_lens episode lengths. */
double orc_baseline_run_single(const rsrl_config_t* cfg, int64_t steps, int32_t* ep_lens, int n_ep_lens, int64_t* out_steps) {
    return run_workers(cfg, 1, 1, steps, 0, ep_lens, n_ep_lens, out_steps);
}

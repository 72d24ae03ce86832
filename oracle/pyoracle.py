"""ctypes binding of the CPU oracle (oracle/rsrl_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference).
The product package rsrl_b200 never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rsrl_b200.abi import Config, Stats  # interface structs only (rsrl_config_t / rsrl_stats_t)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librsrl_oracle.so")

_P = C.POINTER
_dp, _ip, _u8p, _u32p, _u64p = _P(C.c_double), _P(C.c_int32), _P(C.c_uint8), _P(C.c_uint32), _P(C.c_uint64)
_cfgp = _P(Config)
_lib = None


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("rsrl_oracle.c", "rsrl_oracle.h", "Makefile")]
    src.append(os.path.join(HERE, "..", "include", "rsrl_b200.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        sig = {
            "orc_domain_dim": (C.c_int, [C.c_int]),
            "orc_domain_n_actions": (C.c_int, [C.c_int]),
            "orc_domain_limits": (None, [C.c_int, _dp, _dp]),
            "orc_domain_default": (None, [C.c_int, _dp]),
            "orc_domain_is_terminal": (C.c_int, [C.c_int, _dp]),
            "orc_domain_step": (None, [C.c_int, _dp, C.c_int, _dp, _ip]),
            "orc_domain_ex_dim": (C.c_int, [C.c_int]),
            "orc_domain_ex_default": (None, [C.c_int, _dp]),
            "orc_domain_ex_emit": (None, [C.c_int, _dp, _dp, _ip]),
            "orc_domain_ex_step": (None, [C.c_int, _dp, C.c_double, _dp, _dp, _ip]),
            "orc_basis_n_features": (C.c_int64, [_cfgp]),
            "orc_fourier_coefficients": (None, [C.c_int, C.c_int, _dp]),
            "orc_basis_project": (None, [_cfgp, _dp, _dp]),
            "orc_tile_indices": (C.c_int, [_cfgp, _dp, _ip]),
            "orc_lfa_evaluate": (None, [_cfgp, _dp, C.c_int, _dp, _dp]),
            "orc_lfa_update_index": (None, [_cfgp, _dp, C.c_int, _dp, C.c_int, C.c_double]),
            "orc_argmaxima": (C.c_int, [_dp, C.c_int, _ip, _dp]),
            "orc_find_max": (C.c_int, [_dp, C.c_int, _dp]),
            "orc_argmax_first": (C.c_int, [_dp, C.c_int, _dp]),
            "orc_philox4x32_10": (None, [_u32p, _u32p, _u32p]),
            "orc_draw": (None, [C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint32, _u32p]),
            "orc_policy_probs": (None, [C.c_int, C.c_double, _dp, C.c_int, _dp]),
            "orc_policy_sample": (C.c_int, [C.c_int, C.c_double, _dp, C.c_int, _u32p, _ip]),
            "orc_trace_update": (None, [C.c_int, C.c_double, C.c_double, C.c_double, C.c_int64, _dp, _dp]),
            "orc_engine_create": (C.c_void_p, [_cfgp]),
            "orc_engine_destroy": (None, [C.c_void_p]),
            "orc_engine_reset": (None, [C.c_void_p, _dp]),
            "orc_engine_step": (None, [C.c_void_p, C.c_int64]),
            "orc_engine_get_states": (None, [C.c_void_p, _dp]),
            "orc_engine_set_states": (None, [C.c_void_p, _dp]),
            "orc_engine_get_actions": (None, [C.c_void_p, _ip]),
            "orc_engine_get_episode_steps": (None, [C.c_void_p, _ip]),
            "orc_engine_get_weights": (None, [C.c_void_p, _dp]),
            "orc_engine_set_weights": (None, [C.c_void_p, _dp]),
            "orc_engine_get_aux_weights": (None, [C.c_void_p, _dp]),
            "orc_engine_set_aux_weights": (None, [C.c_void_p, _dp]),
            "orc_engine_rollout": (C.c_int64, [C.c_void_p, C.c_int64, _dp, C.c_int64, C.c_int, C.c_uint64, _dp, _dp, _ip, _dp, _u8p]),
            "orc_engine_get_traces": (None, [C.c_void_p, _dp]),
            "orc_engine_set_traces": (None, [C.c_void_p, _dp]),
            "orc_engine_get_td_errors": (None, [C.c_void_p, _dp]),
            "orc_engine_get_stats": (None, [C.c_void_p, _P(Stats)]),
            "orc_engine_get_env_stats": (None, [C.c_void_p, _ip, _ip, _u64p]),
            "orc_engine_set_epsilon": (None, [C.c_void_p, C.c_double]),
            "orc_engine_min_gap": (C.c_double, [C.c_void_p]),
            "orc_engine_step_local": (None, [C.c_void_p, _dp]),
            "orc_engine_step_apply": (None, [C.c_void_p, _dp]),
            "orc_engine_handle": (None, [C.c_void_p, C.c_int64, _dp, _ip, _dp, _dp, _u8p, C.c_uint64, _dp]),
            "orc_baseline_run": (C.c_double, [_cfgp, C.c_int, C.c_int64, C.c_int64, _P(C.c_int64)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# ---- stateless helpers (numpy in / numpy out) ----
def domain_dims(domain):
    return lib().orc_domain_dim(domain), lib().orc_domain_n_actions(domain)


def domain_limits(domain):
    lo, hi = np.zeros(4), np.zeros(4)
    lib().orc_domain_limits(domain, _d(lo), _d(hi))
    D = lib().orc_domain_dim(domain)
    return lo[:D].copy(), hi[:D].copy()


def domain_default(domain):
    s = np.zeros(4)
    lib().orc_domain_default(domain, _d(s))
    return s[:lib().orc_domain_dim(domain)].copy()


def domain_is_terminal(domain, states):
    states = f64(states).reshape(-1, lib().orc_domain_dim(domain))
    return np.array([lib().orc_domain_is_terminal(domain, _d(s)) for s in states], dtype=np.uint8)


def domain_step(domain, states, actions):
    """Domain::step for a batch. Returns (next_states, rewards, terminal)."""
    D = lib().orc_domain_dim(domain)
    ns = f64(states).reshape(-1, D).copy()
    n = ns.shape[0]
    rewards, terminal = np.zeros(n), np.zeros(n, dtype=np.int32)
    r, t = C.c_double(), C.c_int32()
    for i in range(n):
        row = ns[i]
        lib().orc_domain_step(domain, _d(row), int(actions[i]), C.byref(r), C.byref(t))
        rewards[i], terminal[i] = r.value, t.value
    return ns, rewards, terminal.astype(np.uint8)


def domain_ex_default(domain):
    s = np.zeros(lib().orc_domain_ex_dim(domain))
    lib().orc_domain_ex_default(domain, _d(s))
    return s


def domain_ex_emit(domain, states):
    D = lib().orc_domain_ex_dim(domain)
    s = f64(states).reshape(-1, D)
    obs, term = np.zeros_like(s), np.zeros(s.shape[0], dtype=np.int32)
    t = C.c_int32()
    for i in range(s.shape[0]):
        lib().orc_domain_ex_emit(domain, _d(s[i]), _d(obs[i]), C.byref(t))
        term[i] = t.value
    return obs, term.astype(np.uint8)


def domain_ex_step(domain, states, actions):
    """Domain::step of ContinuousMountainCar (actions: forces) / HIVTreatment (actions: indices). Returns (states, obs, rewards, terminal)."""
    D = lib().orc_domain_ex_dim(domain)
    ns = f64(states).reshape(-1, D).copy()
    n = ns.shape[0]
    obs, rewards, term = np.zeros_like(ns), np.zeros(n), np.zeros(n, dtype=np.uint8)
    r, t = C.c_double(), C.c_int32()
    for i in range(n):
        lib().orc_domain_ex_step(domain, _d(ns[i]), float(actions[i]), _d(obs[i]), C.byref(r), C.byref(t))
        rewards[i], term[i] = r.value, t.value
    return ns, obs, rewards, term


def n_features(cfg):
    return lib().orc_basis_n_features(C.byref(cfg))


def project(cfg, states):
    D = lib().orc_domain_dim(cfg.domain)
    states = f64(states).reshape(-1, D)
    F = n_features(cfg)
    out = np.zeros((states.shape[0], F))
    for i in range(states.shape[0]):
        lib().orc_basis_project(C.byref(cfg), _d(states[i]), _d(out[i]))
    return out


def tile_indices(cfg, state):
    idx = np.zeros(cfg.n_tilings, dtype=np.int32)
    n = lib().orc_tile_indices(C.byref(cfg), _d(f64(state)), _i(idx))
    return idx[:n].copy()


def evaluate(cfg, W, states):
    D = lib().orc_domain_dim(cfg.domain)
    states = f64(states).reshape(-1, D)
    W = f64(W)
    A = W.shape[1]
    out = np.zeros((states.shape[0], A))
    for i in range(states.shape[0]):
        lib().orc_lfa_evaluate(C.byref(cfg), _d(W), A, _d(states[i]), _d(out[i]))
    return out


def update_index(cfg, W, state, action, lr_err):
    W = f64(W).copy()
    lib().orc_lfa_update_index(C.byref(cfg), _d(W), W.shape[1], _d(f64(state)), int(action), float(lr_err))
    return W


def argmaxima(q):
    q = f64(q)
    ixs = np.zeros(len(q), dtype=np.int32)
    mx = C.c_double()
    n = lib().orc_argmaxima(_d(q), len(q), _i(ixs), C.byref(mx))
    return ixs[:n].tolist(), mx.value


def find_max(q):
    q = f64(q)
    mx = C.c_double()
    i = lib().orc_find_max(_d(q), len(q), C.byref(mx))
    return i, mx.value


def argmax_first(q):
    q = f64(q)
    mx = C.c_double()
    i = lib().orc_argmax_first(_d(q), len(q), C.byref(mx))
    return i, mx.value


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data_as(_u32p), k.ctypes.data_as(_u32p), out.ctypes.data_as(_u32p))
    return out


def draw(seed, env, step, stream):
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_draw(seed, env, step, stream, out.ctypes.data_as(_u32p))
    return out


def policy_probs(policy, epsilon, q):
    q = f64(q)
    p = np.zeros(len(q))
    lib().orc_policy_probs(policy, epsilon, _d(q), len(q), _d(p))
    return p


def policy_sample(policy, epsilon, q, rnd):
    q = f64(q)
    rnd = np.asarray(rnd, dtype=np.uint32)
    nf = C.c_int32(0)
    a = lib().orc_policy_sample(policy, epsilon, _d(q), len(q), rnd.ctypes.data_as(_u32p), C.byref(nf))
    return a, nf.value


def policy_sample_batch(policy, epsilon, seed, draw_idx, env_offset, q, stream=1):
    q = f64(q)
    return np.array([policy_sample(policy, epsilon, q[i], draw(seed, env_offset + i, draw_idx, stream))[0]
                     for i in range(q.shape[0])], dtype=np.int32)


def trace_update(rule, gamma, lam, alpha, z, grad):
    z = f64(z).copy()
    grad = f64(grad)
    lib().orc_trace_update(rule, gamma, lam, alpha, z.size, _d(z), _d(grad))
    return z


class Engine:
    """Batched oracle engine: N-env restatement of examples/q_learning.rs:34-55 (same surface as rsrl_b200.Engine)."""

    def __init__(self, cfg):
        self.cfg = cfg
        self.h = lib().orc_engine_create(C.byref(cfg))
        self.D, self.A = domain_dims(cfg.domain)
        self.AW = 1 if cfg.algo in (5, 6) else self.A
        self.F = n_features(cfg)
        self.N = cfg.n_envs

    def close(self):
        if self.h:
            lib().orc_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def reset(self, init_states=None):
        lib().orc_engine_reset(self.h, _d(f64(init_states)) if init_states is not None else None)

    def step(self, k=1):
        lib().orc_engine_step(self.h, k)

    def step_local(self):
        g = np.zeros((self.F, self.AW))
        lib().orc_engine_step_local(self.h, _d(g))
        return g

    def step_apply(self, g=None):
        lib().orc_engine_step_apply(self.h, _d(f64(g)) if g is not None else None)

    def sync(self):
        pass

    def states(self):
        out = np.zeros((self.N, self.D))
        lib().orc_engine_get_states(self.h, _d(out))
        return out

    def set_states(self, s):
        lib().orc_engine_set_states(self.h, _d(f64(s)))

    def actions(self):
        out = np.zeros(self.N, dtype=np.int32)
        lib().orc_engine_get_actions(self.h, _i(out))
        return out

    def episode_steps(self):
        out = np.zeros(self.N, dtype=np.int32)
        lib().orc_engine_get_episode_steps(self.h, _i(out))
        return out

    def _wshape(self):
        return (self.N, self.F, self.AW) if self.cfg.weight_mode == 1 else (self.F, self.AW)

    def weights(self):
        out = np.zeros(self._wshape())
        lib().orc_engine_get_weights(self.h, _d(out))
        return out

    def set_weights(self, w):
        w = f64(w)
        assert w.shape == self._wshape()
        lib().orc_engine_set_weights(self.h, _d(w))

    def aux_weights(self):
        out = np.empty(self._wshape())
        lib().orc_engine_get_aux_weights(self.h, _d(out))
        return out

    def set_aux_weights(self, w):
        w = f64(w)
        assert w.shape == self._wshape()
        lib().orc_engine_set_aux_weights(self.h, _d(w))

    def rollout(self, n=None, init_states=None, step_limit=500, greedy=True, draw=0):
        """Domain::rollout per env (rsrl_domains/src/lib.rs:448-479); same dict layout as rsrl_b200.engine.Engine.rollout"""
        n = self.N if n is None else n
        T = max(step_limit - 1, 1)
        out = dict(start=np.zeros((n, self.D)), next=np.zeros((n, T, self.D)), actions=np.full((n, T), -1, dtype=np.int32),
                   rewards=np.zeros((n, T)), terminal=np.zeros((n, T), dtype=np.uint8), len=np.zeros(n, dtype=np.int32))
        init = None if init_states is None else f64(init_states)
        for i in range(n):
            st, nx, ac, rw, tm = out["start"][i], out["next"][i], out["actions"][i], out["rewards"][i], out["terminal"][i]
            out["len"][i] = lib().orc_engine_rollout(self.h, i, None if init is None else _d(init[i]), step_limit, 1 if greedy else 0,
                                                     draw, _d(st), _d(nx), _i(ac), _d(rw), tm.ctypes.data_as(_u8p))
        return out

    def traces(self):
        out = np.zeros((self.N, self.F, self.AW))
        lib().orc_engine_get_traces(self.h, _d(out))
        return out

    def set_traces(self, z):
        lib().orc_engine_set_traces(self.h, _d(f64(z)))

    def td_errors(self):
        out = np.zeros(self.N)
        lib().orc_engine_get_td_errors(self.h, _d(out))
        return out

    def stats(self):
        st = Stats()
        lib().orc_engine_get_stats(self.h, C.byref(st))
        return st.as_dict()

    def env_stats(self):
        n_ep, last = np.zeros(self.N, dtype=np.int32), np.zeros(self.N, dtype=np.int32)
        h = np.zeros(self.N, dtype=np.uint64)
        lib().orc_engine_get_env_stats(self.h, _i(n_ep), _i(last), h.ctypes.data_as(_u64p))
        return n_ep, last, h

    def set_epsilon(self, eps):
        lib().orc_engine_set_epsilon(self.h, eps)

    def min_gap(self):
        return lib().orc_engine_min_gap(self.h)

    def handle(self, from_states, actions, rewards, to_states, terminal, draw_idx=0):
        n = len(actions)
        td = np.zeros(n)
        a = np.ascontiguousarray(actions, dtype=np.int32)
        t = np.ascontiguousarray(terminal, dtype=np.uint8)
        lib().orc_engine_handle(self.h, n, _d(f64(from_states)), _i(a), _d(f64(rewards)), _d(f64(to_states)),
                                t.ctypes.data_as(_u8p), draw_idx, _d(td))
        return td


NATIVE_PATH = os.path.join(HERE, "_build", "librsrl_oracle_native.so")
_native = None


def native_lib():
    """rsrl_oracle.c rebuilt -O3 -march=native ON THIS MACHINE for the CPU-baseline timing;
    never shipped (a -march=native binary may not run elsewhere).  Falls back to the portable build if gcc is missing."""
    global _native
    if _native is None:
        try:
            subprocess.check_call(["make", "-C", HERE, "-s", "-B", "native"])
            L = C.CDLL(NATIVE_PATH)
            kind = "-O3 -march=native"
        except (OSError, subprocess.CalledProcessError):
            L, kind = lib(), "-O2 (portable build: native rebuild failed)"
        for name in ("orc_baseline_run", "orc_baseline_run_fused"):
            fn = getattr(L, name)
            fn.restype, fn.argtypes = C.c_double, [_cfgp, C.c_int, C.c_int64, C.c_int64, _P(C.c_int64)]
        L.orc_baseline_run_single.restype = C.c_double
        L.orc_baseline_run_single.argtypes = [_cfgp, C.c_int64, _ip, C.c_int, _P(C.c_int64)]
        _native = (L, kind)
    return _native


def baseline_run(cfg, threads, envs_per_thread, steps, fused=False):
    """(seconds of stepping, env-steps done): `threads` x envs_per_thread independent single-env agents, reference-shaped
    (4 projections + heap allocations per step) or fused=True (1 projection, no allocation)."""
    L, _ = native_lib()
    done = C.c_int64()
    fn = L.orc_baseline_run_fused if fused else L.orc_baseline_run
    secs = fn(C.byref(cfg), threads, envs_per_thread, steps, C.byref(done))
    return secs, done.value


def baseline_run_single(cfg, steps, n_ep_lens=40):
    """BASELINE configs[0]: one env on one core; returns (seconds, env-steps, first episode lengths)."""
    L, _ = native_lib()
    done = C.c_int64()
    lens = np.zeros(n_ep_lens, dtype=np.int32)
    secs = L.orc_baseline_run_single(C.byref(cfg), steps, _i(lens), n_ep_lens, C.byref(done))
    return secs, done.value, [int(x) for x in lens if x >= 0]


def baseline_build_flags():
    return native_lib()[1]

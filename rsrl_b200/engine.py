"""numpy-facing wrapper over the C ABI (include/rsrl_b200.h).  Every call goes through librsrl_b200.so."""
import ctypes as C

import numpy as np

from . import abi
from .abi import check, dp, ip, u8p, u32p, u64p


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def config_dims(cfg):
    lib = abi.load()
    d, a, f = C.c_int32(), C.c_int32(), C.c_int64()
    check(lib.rsrl_config_dims(C.byref(cfg), C.byref(d), C.byref(a), C.byref(f)))
    return d.value, a.value, f.value


class Engine:
    """rsrl_engine_t: N envs stepping transition -> handle -> sample fused on the GPU."""

    def __init__(self, cfg):
        self.lib = abi.load()
        self.cfg = cfg
        self.h = C.c_void_p()
        check(self.lib.rsrl_engine_create(C.byref(cfg), C.byref(self.h)))
        self.D, self.A, self.F = config_dims(cfg)
        self.AW = 1 if cfg.algo in (abi.TD_LAMBDA, abi.TD0) else self.A
        self.N = cfg.n_envs

    def close(self):
        if getattr(self, "h", None):
            self.lib.rsrl_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def reset(self, init_states=None):
        if init_states is None:
            check(self.lib.rsrl_engine_reset(self.h, None))
        else:
            s = _f64(init_states)
            assert s.shape == (self.N, self.D)
            check(self.lib.rsrl_engine_reset(self.h, dp(s)))

    def step(self, k=1):
        check(self.lib.rsrl_engine_step(self.h, k))

    def sync(self):
        check(self.lib.rsrl_engine_sync(self.h))

    def stream(self):
        return self.lib.rsrl_engine_stream(self.h)

    def states(self, out=None):
        out = np.empty((self.N, self.D)) if out is None else out
        check(self.lib.rsrl_engine_get_states(self.h, dp(out)))
        return out

    def set_states(self, s):
        s = _f64(s)
        assert s.shape == (self.N, self.D)
        check(self.lib.rsrl_engine_set_states(self.h, dp(s)))

    def actions(self, out=None):
        out = np.empty(self.N, dtype=np.int32) if out is None else out
        check(self.lib.rsrl_engine_get_actions(self.h, ip(out)))
        return out

    def episode_steps(self):
        out = np.empty(self.N, dtype=np.int32)
        check(self.lib.rsrl_engine_get_episode_steps(self.h, ip(out)))
        return out

    def _wshape(self):
        return (self.N, self.F, self.AW) if self.cfg.weight_mode == abi.PER_ENV else (self.F, self.AW)

    def weights(self, out=None):
        """Parameterised::weights(): F x A (SHARED) or N x F x A (PER_ENV)."""
        out = np.empty(self._wshape()) if out is None else out
        check(self.lib.rsrl_engine_get_weights(self.h, dp(out)))
        return out

    def set_weights(self, w):
        w = _f64(w)
        assert w.shape == self._wshape()
        check(self.lib.rsrl_engine_set_weights(self.h, dp(w)))

    def aux_weights(self):
        """second weight table: GreedyGQ fa_td (greedy_gq.rs:52), A2C the Gibbs policy's LFA (a2c.rs:27-31)"""
        out = np.empty(self._wshape())
        check(self.lib.rsrl_engine_get_aux_weights(self.h, dp(out)))
        return out

    def set_aux_weights(self, w):
        w = _f64(w)
        assert w.shape == self._wshape()
        check(self.lib.rsrl_engine_set_aux_weights(self.h, dp(w)))

    def rollout(self, n=None, init_states=None, step_limit=500, greedy=True, draw=0):
        """Domain::rollout for n envs with the current weights (rsrl_domains/src/lib.rs:448-479).  Returns a dict in the
        Trajectory{start, steps} layout: start (n, D), next (n, T, D), actions / rewards / terminal (n, T), len (n,)."""
        n = self.N if n is None else n
        T = max(step_limit - 1, 1)
        init = None if init_states is None else _f64(init_states)
        out = dict(start=np.zeros((n, self.D)), next=np.zeros((n, T, self.D)), actions=np.full((n, T), -1, dtype=np.int32),
                   rewards=np.zeros((n, T)), terminal=np.zeros((n, T), dtype=np.uint8), len=np.zeros(n, dtype=np.int32))
        check(self.lib.rsrl_engine_rollout(self.h, n, None if init is None else dp(init), step_limit, 1 if greedy else 0, draw,
                                           dp(out["start"]), dp(out["next"]), ip(out["actions"]), dp(out["rewards"]),
                                           u8p(out["terminal"]), ip(out["len"])))
        return out

    def traces(self):
        out = np.empty((self.N, self.F, self.AW))
        check(self.lib.rsrl_engine_get_traces(self.h, dp(out)))
        return out

    def set_traces(self, z):
        z = _f64(z)
        assert z.shape == (self.N, self.F, self.AW)
        check(self.lib.rsrl_engine_set_traces(self.h, dp(z)))

    def td_errors(self):
        out = np.empty(self.N)
        check(self.lib.rsrl_engine_get_td_errors(self.h, dp(out)))
        return out

    def stats(self):
        st = abi.Stats()
        check(self.lib.rsrl_engine_get_stats(self.h, C.byref(st)))
        return st.as_dict()

    def env_stats(self):
        n_ep, last = np.empty(self.N, dtype=np.int32), np.empty(self.N, dtype=np.int32)
        h = np.empty(self.N, dtype=np.uint64)
        check(self.lib.rsrl_engine_get_env_stats(self.h, ip(n_ep), ip(last), u64p(h)))
        return n_ep, last, h

    def launch_shape(self):
        """rsrl_engine_get_launch_shape as a dict (the fp32 summation order is a function of it: oracle/oracle32.cpp)."""
        out = np.zeros(24, dtype=np.int32)
        check(self.lib.rsrl_engine_get_launch_shape(self.h, ip(out)))
        keys = ["persistent", "mode", "grid", "cluster_size", "n_clusters", "block", "lpr", "unused7", "unused8", "pe_smem", "world",
                "rank", "peers", "smem", "tile", "f4", "fx"]
        return dict(zip(keys, (int(v) for v in out)))

    def set_epsilon(self, eps):
        check(self.lib.rsrl_engine_set_epsilon(self.h, eps))

    # ---- trait-level entry points ----
    def evaluate(self, states):
        s = _f64(states).reshape(-1, self.D)
        q = np.empty((s.shape[0], self.AW))
        check(self.lib.rsrl_engine_evaluate(self.h, s.shape[0], dp(s), dp(q)))
        return q

    def sample(self, states, draw=0):
        s = _f64(states).reshape(-1, self.D)
        a = np.empty(s.shape[0], dtype=np.int32)
        check(self.lib.rsrl_engine_sample(self.h, s.shape[0], dp(s), draw, ip(a)))
        return a

    def mode(self, states):
        s = _f64(states).reshape(-1, self.D)
        a = np.empty(s.shape[0], dtype=np.int32)
        check(self.lib.rsrl_engine_mode(self.h, s.shape[0], dp(s), ip(a)))
        return a

    def handle(self, from_states, actions, rewards, to_states, terminal, draw_idx=0):
        f, t = _f64(from_states).reshape(-1, self.D), _f64(to_states).reshape(-1, self.D)
        a, r = _i32(actions), _f64(rewards)
        term = np.ascontiguousarray(terminal, dtype=np.uint8)
        td = np.empty(len(a))
        check(self.lib.rsrl_engine_handle(self.h, len(a), dp(f), ip(a), dp(r), dp(t), u8p(term), draw_idx, dp(td)))
        return td

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        check(self.lib.rsrl_engine_comm_init(self.h, buf, rank, world))


def _peer_export(self):
    buf = (C.c_uint8 * 64)()
    check(self.lib.rsrl_engine_peer_export(self.h, buf))
    return bytes(buf)


def _peer_attach(self, handles, rank, world):
    blob = b"".join(bytes(h) for h in handles)
    assert len(blob) == 64 * world
    buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
    check(self.lib.rsrl_engine_peer_attach(self.h, buf, rank, world))


Engine.peer_export = _peer_export
Engine.peer_attach = _peer_attach


def comm_unique_id():
    buf = (C.c_uint8 * 128)()
    check(abi.load().rsrl_comm_unique_id(buf))
    return bytes(buf)


# ---- stateless component entry points ----
def domain_info(domain):
    d, a = C.c_int32(), C.c_int32()
    lo, hi, start = np.zeros(4), np.zeros(4), np.zeros(4)
    check(abi.load().rsrl_domain_info(domain, C.byref(d), C.byref(a), dp(lo), dp(hi), dp(start)))
    return d.value, a.value, lo[:d.value].copy(), hi[:d.value].copy(), start[:d.value].copy()


def domain_step(domain, states, actions):
    D = 2 if domain == abi.MOUNTAIN_CAR else 4
    s = _f64(states).reshape(-1, D).copy()
    a = _i32(actions)
    r, t = np.empty(len(a)), np.empty(len(a), dtype=np.uint8)
    check(abi.load().rsrl_domain_step(domain, len(a), dp(s), ip(a), dp(r), u8p(t)))
    return s, r, t


def domain_is_terminal(domain, states):
    D = 2 if domain == abi.MOUNTAIN_CAR else 4
    s = _f64(states).reshape(-1, D)
    t = np.empty(s.shape[0], dtype=np.uint8)
    check(abi.load().rsrl_domain_is_terminal(domain, s.shape[0], dp(s), u8p(t)))
    return t


def _check_ex(code):
    if code != abi.OK:
        msg = abi.load().rsrl_domain_ex_last_error()
        raise abi.RsrlError(code, msg.decode() if msg else "")


def domain_ex_info(domain):
    d, a = C.c_int32(), C.c_int32()
    lo, hi, st = np.zeros(6), np.zeros(6), np.zeros(6)
    _check_ex(abi.load().rsrl_domain_ex_info(domain, C.byref(d), C.byref(a), dp(lo), dp(hi), dp(st)))
    return d.value, a.value, lo[:d.value].copy(), hi[:d.value].copy(), st[:d.value].copy()


def domain_ex_step(domain, states, actions):
    """ContinuousMountainCar (actions: forces, f64) / HIVTreatment (actions: indices). Returns (states, obs, rewards, terminal)."""
    D = 6 if domain == abi.HIV else 2
    ns = _f64(states).reshape(-1, D).copy()
    n = ns.shape[0]
    obs, r, t = np.zeros_like(ns), np.zeros(n), np.zeros(n, dtype=np.uint8)
    if domain == abi.HIV:
        a = _i32(actions)
        _check_ex(abi.load().rsrl_domain_ex_step(domain, n, dp(ns), ip(a), None, dp(obs), dp(r), u8p(t)))
    else:
        a = _f64(actions)
        _check_ex(abi.load().rsrl_domain_ex_step(domain, n, dp(ns), None, dp(a), dp(obs), dp(r), u8p(t)))
    return ns, obs, r, t


def domain_ex_emit(domain, states):
    D = 6 if domain == abi.HIV else 2
    s = _f64(states).reshape(-1, D)
    obs, t = np.zeros_like(s), np.zeros(s.shape[0], dtype=np.uint8)
    _check_ex(abi.load().rsrl_domain_ex_emit(domain, s.shape[0], dp(s), dp(obs), u8p(t)))
    return obs, t


def weights_to_serde(weights):
    """The `serde` feature's on-disk form of Parameterised::weights() (rsrl/Cargo.toml:26): ndarray 0.13's serde layout of an
    Array2<f64> — {"v": 1, "dim": [rows = F, cols = A], "data": row-major} (the layout rsrl_engine_get_weights already returns)."""
    w = _f64(weights)
    assert w.ndim == 2
    return {"v": 1, "dim": [int(w.shape[0]), int(w.shape[1])], "data": [float(x) for x in w.ravel()]}


def weights_from_serde(obj):
    assert obj["v"] == 1 and len(obj["dim"]) == 2
    return np.asarray(obj["data"], dtype=np.float64).reshape(obj["dim"])


def basis_project(cfg, states):
    D, _, F = config_dims(cfg)
    s = _f64(states).reshape(-1, D)
    out = np.empty((s.shape[0], F))
    check(abi.load().rsrl_basis_project(C.byref(cfg), s.shape[0], dp(s), dp(out)))
    return out


def lfa_evaluate(cfg, weights, states):
    D, _, F = config_dims(cfg)
    s = _f64(states).reshape(-1, D)
    w = _f64(weights)
    q = np.empty((s.shape[0], w.shape[1]))
    check(abi.load().rsrl_lfa_evaluate(C.byref(cfg), s.shape[0], dp(s), dp(w), dp(q)))
    return q


def lfa_update_index(cfg, weights, states, actions, errors):
    D, _, F = config_dims(cfg)
    s = _f64(states).reshape(-1, D)
    w = _f64(weights).copy()
    a, e = _i32(actions), _f64(errors)
    check(abi.load().rsrl_lfa_update_index(C.byref(cfg), len(a), dp(s), ip(a), dp(e), dp(w)))
    return w


def policy_sample(policy, epsilon, seed, draw, env_offset, q):
    q = _f64(q)
    a = np.empty(q.shape[0], dtype=np.int32)
    check(abi.load().rsrl_policy_sample(policy, epsilon, seed, draw, env_offset, q.shape[0], q.shape[1], dp(q), ip(a)))
    return a


def policy_probs(policy, epsilon, q):
    q = _f64(q)
    p = np.empty_like(q)
    check(abi.load().rsrl_policy_probs(policy, epsilon, q.shape[0], q.shape[1], dp(q), dp(p)))
    return p


def policy_mode(q):
    q = _f64(q)
    a = np.empty(q.shape[0], dtype=np.int32)
    check(abi.load().rsrl_policy_mode(q.shape[0], q.shape[1], dp(q), ip(a)))
    return a


def trace_update(rule, gamma, lam, alpha, z, grad):
    z = _f64(z).copy()
    g = _f64(grad)
    check(abi.load().rsrl_trace_update(rule, gamma, lam, alpha, z.size, dp(z), dp(g)))
    return z


def philox(seed, draw, stream, env_offset, n):
    out = np.empty((n, 4), dtype=np.uint32)
    check(abi.load().rsrl_philox(seed, draw, stream, env_offset, n, u32p(out)))
    return out


def math_probe(fn, x):
    """fn: 0 cos64, 1 sin64, 2 sinpi32, 3 cospi32, 4 exp32 — evaluated on the GPU (csrc/device.cuh "rsrl math")."""
    x = _f64(x).ravel()
    out = np.empty_like(x)
    check(abi.load().rsrl_math_probe(fn, x.size, dp(x), dp(out)))
    return out

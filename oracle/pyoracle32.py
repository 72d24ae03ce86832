"""ctypes binding of oracle32 (oracle/oracle32.cpp): the fp32 device arithmetic replayed on the host, bit for bit.

TEST INFRASTRUCTURE ONLY — same import rules as oracle/pyoracle.py.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from rsrl_b200 import abi as _abi
from rsrl_b200.abi import Config

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "librsrl_oracle32.so")
CSRC = os.path.join(HERE, "..", "rsrl_b200", "csrc")

_P = C.POINTER
_dp, _ip, _u8p, _u64p, _i64p = _P(C.c_double), _P(C.c_int32), _P(C.c_uint8), _P(C.c_uint64), _P(C.c_int64)
_lib = None


def build(force=False):
    src = [os.path.join(HERE, f) for f in ("oracle32.cpp", "Makefile")]
    src += [os.path.join(CSRC, f) for f in ("hostdev.h", "device.cuh", "core.cuh")]
    src.append(os.path.join(HERE, "..", "include", "rsrl_b200.h"))
    if force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in src):
        subprocess.check_call(["make", "-C", HERE, "-s", "_build/librsrl_oracle32.so"])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        vp = C.c_void_p
        sig = {
            "o32_engine_create": (vp, [_P(Config), _ip, C.c_int, C.c_int]),
            "o32_engine_destroy": (None, [vp]),
            "o32_engine_step": (None, [vp, C.c_int64]),
            "o32_engine_set_epsilon": (None, [vp, C.c_double]),
            "o32_engine_set_states": (None, [vp, C.c_int, _dp]),
            "o32_engine_get_states": (None, [vp, C.c_int, _dp]),
            "o32_engine_get_actions": (None, [vp, C.c_int, _ip]),
            "o32_engine_get_episode_steps": (None, [vp, C.c_int, _ip]),
            "o32_engine_get_env_stats": (None, [vp, C.c_int, _ip, _ip, _u64p]),
            "o32_engine_get_td_errors": (None, [vp, C.c_int, _dp]),
            "o32_engine_get_weights": (None, [vp, C.c_int, _dp]),
            "o32_engine_set_weights": (None, [vp, C.c_int, _dp]),
            "o32_engine_get_traces": (None, [vp, C.c_int, _dp]),
            "o32_engine_get_counters": (None, [vp, C.c_int, _i64p, _i64p, _ip]),
            "o32_math": (None, [C.c_int, C.c_int64, _dp, _dp]),
            "o32_domain_step": (None, [C.c_int, C.c_int64, _dp, _ip, _dp, _u8p]),
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


SHAPE_KEYS = ["persistent", "mode", "grid", "cluster_size", "n_clusters", "block", "lpr", "unused7", "unused8", "pe_smem"]


def host_shape(cfg, grid=None, cluster_size=1, block=None, fx=1):
    """A launch shape computed like rsrl_b200/csrc/abi.cu:persistent_shape, for tests that have no GPU to ask
    (rsrl_engine_get_launch_shape is authoritative on a GPU box)."""
    D = 2 if cfg.domain == _abi.MOUNTAIN_CAR else 4
    A = 2 if cfg.domain == _abi.CART_POLE else 3
    aw = 1 if cfg.algo in (_abi.TD_LAMBDA, _abi.TD0) else A
    F = (cfg.basis_order + 1) ** D
    trace = cfg.algo in (_abi.SARSA_LAMBDA, _abi.Q_LAMBDA, _abi.TD_LAMBDA)
    N = cfg.n_envs
    if cfg.weight_mode == _abi.PER_ENV:
        return dict(persistent=1, mode=_abi.PER_ENV, grid=(N + 127) // 128, cluster_size=1, n_clusters=1, block=128, lpr=1, unused7=0,
                    unused8=0, pe_smem=1, fx=0)
    rows = F * aw if trace else F
    if grid is None:
        g0 = min((N + 127) // 128, 148)
        cs = 1 if g0 == 1 else cluster_size
        grid = min((g0 + cs - 1) // cs * cs, 132 if cs > 1 else 148)
    else:
        cs = 1 if grid == 1 else cluster_size
    per_cta = (N + grid - 1) // grid
    r32 = lambda x: (x + 31) // 32 * 32
    if block is None:
        block = min(512, max(64, r32(per_cta), r32(rows)))  # persistent.cuh: kPersistMaxBlock
    lpr = 8
    while rows * lpr > block:
        lpr >>= 1
    return dict(persistent=1, mode=2 if trace else _abi.SHARED, grid=grid, cluster_size=cs, n_clusters=grid // cs, block=block,
                lpr=lpr, unused7=0, unused8=0, pe_smem=0, fx=fx if cs == 1 else 0)


class Engine:
    """World of `world` ranks, each with cfg.n_envs envs, stepping exactly like the GPU engines of that launch shape."""

    def __init__(self, cfg, shape, world=1, threads=None):
        self.cfg, self.world = cfg, world
        sh = np.zeros(24, dtype=np.int32)
        for k, key in enumerate(SHAPE_KEYS):
            sh[k] = shape[key]
        sh[16] = shape.get("fx", 0)
        self.shape = dict(shape)
        self.D = 2 if cfg.domain == _abi.MOUNTAIN_CAR else 4
        self.A = 2 if cfg.domain == _abi.CART_POLE else 3
        self.AW = 1 if cfg.algo in (_abi.TD_LAMBDA, _abi.TD0) else self.A
        self.F = (cfg.basis_order + 1) ** self.D
        self.N = cfg.n_envs
        self.h = lib().o32_engine_create(C.byref(cfg), _i(sh), world, threads or (os.cpu_count() or 1))
        if not self.h:
            raise ValueError("oracle32: configuration / launch shape outside what it replays (fp32, grid bases, persistent kernel)")

    def close(self):
        if getattr(self, "h", None):
            lib().o32_engine_destroy(self.h)
            self.h = None

    __del__ = close

    def step(self, k=1):
        lib().o32_engine_step(self.h, k)

    def set_epsilon(self, eps):
        lib().o32_engine_set_epsilon(self.h, eps)

    def set_states(self, s, rank=0):
        s = np.ascontiguousarray(s, dtype=np.float64)
        assert s.shape == (self.N, self.D)
        lib().o32_engine_set_states(self.h, rank, _d(s))

    def states(self, rank=0):
        out = np.empty((self.N, self.D))
        lib().o32_engine_get_states(self.h, rank, _d(out))
        return out

    def actions(self, rank=0):
        out = np.empty(self.N, dtype=np.int32)
        lib().o32_engine_get_actions(self.h, rank, _i(out))
        return out

    def episode_steps(self, rank=0):
        out = np.empty(self.N, dtype=np.int32)
        lib().o32_engine_get_episode_steps(self.h, rank, _i(out))
        return out

    def env_stats(self, rank=0):
        n_ep, last = np.empty(self.N, dtype=np.int32), np.empty(self.N, dtype=np.int32)
        h = np.empty(self.N, dtype=np.uint64)
        lib().o32_engine_get_env_stats(self.h, rank, _i(n_ep), _i(last), h.ctypes.data_as(_u64p))
        return n_ep, last, h

    def td_errors(self, rank=0):
        out = np.empty(self.N)
        lib().o32_engine_get_td_errors(self.h, rank, _d(out))
        return out

    def _wshape(self):
        return (self.N, self.F, self.AW) if self.cfg.weight_mode == _abi.PER_ENV else (self.F, self.AW)

    def weights(self, rank=0):
        out = np.empty(self._wshape())
        lib().o32_engine_get_weights(self.h, rank, _d(out))
        return out

    def set_weights(self, w, rank=0):
        w = np.ascontiguousarray(w, dtype=np.float64)
        assert w.shape == self._wshape()
        lib().o32_engine_set_weights(self.h, rank, _d(w))

    def traces(self, rank=0):
        out = np.empty((self.N, self.F, self.AW))
        lib().o32_engine_get_traces(self.h, rank, _d(out))
        return out

    def counters(self, rank=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int32()
        lib().o32_engine_get_counters(self.h, rank, C.byref(a), C.byref(b), C.byref(c))
        return {"total_episodes": a.value, "terminal_episodes": b.value, "nonfinite": c.value}


def math(fn, x):
    """fn: 0 cos64, 1 sin64, 2 sinpi32, 3 cospi32, 4 exp32 (csrc/device.cuh "rsrl math"), host build."""
    x = np.ascontiguousarray(x, dtype=np.float64).ravel()
    out = np.empty_like(x)
    lib().o32_math(fn, x.size, _d(x), _d(out))
    return out


def domain_step(domain, states, actions):
    D = 2 if domain == _abi.MOUNTAIN_CAR else 4
    ns = np.ascontiguousarray(states, dtype=np.float64).reshape(-1, D).copy()
    n = ns.shape[0]
    a = np.ascontiguousarray(actions, dtype=np.int32)
    r, t = np.zeros(n), np.zeros(n, dtype=np.uint8)
    lib().o32_domain_step(domain, n, _d(ns), _i(a), _d(r), t.ctypes.data_as(_u8p))
    return ns, r, t

"""ctypes view of include/rsrl_b200.h (the C ABI) and the loader of librsrl_b200.so.

There is no CPU fallback: if the CUDA library has not been built, `load()` raises.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "librsrl_b200.so")

# enums (include/rsrl_b200.h)
OK, EINVAL, ECUDA, ENOMEM, EUNSUPPORTED, ENONFINITE, ECOMM, ENODEVICE = 0, -1, -2, -3, -4, -5, -6, -7
MOUNTAIN_CAR, CART_POLE, ACROBOT, CONTINUOUS_MOUNTAIN_CAR, HIV = 0, 1, 2, 3, 4
FOURIER, POLYNOMIAL, TILE_CODING = 0, 1, 2
QLEARNING, SARSA, EXPECTED_SARSA, SARSA_LAMBDA, Q_LAMBDA, TD_LAMBDA, TD0, PAL, GREEDY_GQ, A2C = range(10)
GREEDY, EPSILON_GREEDY, RANDOM, SOFTMAX = 0, 1, 2, 3
TRACE_ACCUMULATE, TRACE_REPLACE, TRACE_DUTCH = 0, 1, 2
SHARED, PER_ENV = 0, 1
SCALE_SUM, SCALE_MEAN = 0, 1
F32, F64 = 0, 1
INIT_DEFAULT, INIT_UNIFORM = 0, 1
MAX_DIM = 4


class Config(C.Structure):
    """rsrl_config_t"""
    _fields_ = [
        ("struct_size", C.c_uint32), ("domain", C.c_int32), ("basis", C.c_int32), ("basis_order", C.c_int32),
        ("n_tilings", C.c_int32), ("tiles_per_dim", C.c_int32), ("memory_size", C.c_int32), ("algo", C.c_int32),
        ("policy", C.c_int32), ("trace_rule", C.c_int32), ("weight_mode", C.c_int32), ("update_scale", C.c_int32),
        ("dtype", C.c_int32), ("init_mode", C.c_int32), ("device", C.c_int32), ("record_td_error", C.c_int32),
        ("n_envs", C.c_int64), ("env_offset", C.c_int64), ("n_envs_global", C.c_int64),
        ("max_episode_steps", C.c_int64), ("seed", C.c_uint64),
        ("lr", C.c_double), ("alpha", C.c_double), ("gamma", C.c_double), ("lambda_", C.c_double),
        ("epsilon", C.c_double), ("init_lo", C.c_double * MAX_DIM), ("init_hi", C.c_double * MAX_DIM),
    ]

    def copy(self, **kw):
        c = Config.from_buffer_copy(bytes(self))
        for k, v in kw.items():
            _set(c, k, v)
        return c


def _set(cfg, k, v):
    if k == "lambda":
        k = "lambda_"
    if k in ("init_lo", "init_hi"):
        arr = getattr(cfg, k)
        for i, x in enumerate(v):
            arr[i] = x
    else:
        if not hasattr(cfg, k):
            raise AttributeError(k)
        setattr(cfg, k, v)


def default_config(**kw):
    """examples/q_learning.rs:18-32 (MountainCar, Fourier(5)+bias, SGD(0.001), gamma 0.9, Greedy, seed 0).

    Pure-python mirror of rsrl_config_default() so that configs can be built without the CUDA library."""
    c = Config()
    c.struct_size = C.sizeof(Config)
    c.domain, c.basis, c.basis_order = MOUNTAIN_CAR, FOURIER, 5
    c.n_tilings, c.tiles_per_dim, c.memory_size = 8, 8, 4096
    c.algo, c.policy, c.trace_rule = QLEARNING, GREEDY, TRACE_REPLACE
    c.weight_mode, c.update_scale, c.dtype, c.init_mode = SHARED, SCALE_SUM, F32, INIT_DEFAULT
    c.device, c.record_td_error = 0, 0
    c.n_envs, c.env_offset, c.n_envs_global, c.max_episode_steps, c.seed = 1, 0, 0, 0, 0
    c.lr, c.alpha, c.gamma, c.lambda_, c.epsilon = 0.001, 0.01, 0.9, 0.7, 0.1
    for k, v in kw.items():
        _set(c, k, v)
    return c


class Stats(C.Structure):
    """rsrl_stats_t"""
    _fields_ = [("total_steps", C.c_int64), ("total_episodes", C.c_int64), ("terminal_episodes", C.c_int64),
                ("batch_steps", C.c_int64), ("kernel_launches", C.c_int64), ("nonfinite", C.c_int32),
                ("reserved", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_ if k != "reserved"}


_P = C.POINTER
_dp, _ip, _u8p, _u32p, _u64p = _P(C.c_double), _P(C.c_int32), _P(C.c_uint8), _P(C.c_uint32), _P(C.c_uint64)
_cfgp, _eng = _P(Config), C.c_void_p

# every symbol include/rsrl_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "rsrl_version": (C.c_int, []),
    "rsrl_last_error": (C.c_char_p, []),
    "rsrl_device_count": (C.c_int, []),
    "rsrl_config_default": (C.c_int, [_cfgp]),
    "rsrl_config_dims": (C.c_int, [_cfgp, _ip, _ip, _P(C.c_int64)]),
    "rsrl_engine_create": (C.c_int, [_cfgp, _P(_eng)]),
    "rsrl_engine_destroy": (C.c_int, [_eng]),
    "rsrl_engine_reset": (C.c_int, [_eng, _dp]),
    "rsrl_engine_step": (C.c_int, [_eng, C.c_int64]),
    "rsrl_engine_sync": (C.c_int, [_eng]),
    "rsrl_engine_stream": (C.c_void_p, [_eng]),
    "rsrl_engine_get_states": (C.c_int, [_eng, _dp]),
    "rsrl_engine_set_states": (C.c_int, [_eng, _dp]),
    "rsrl_engine_get_actions": (C.c_int, [_eng, _ip]),
    "rsrl_engine_get_episode_steps": (C.c_int, [_eng, _ip]),
    "rsrl_engine_get_weights": (C.c_int, [_eng, _dp]),
    "rsrl_engine_set_weights": (C.c_int, [_eng, _dp]),
    "rsrl_engine_get_aux_weights": (C.c_int, [_eng, _dp]),
    "rsrl_engine_set_aux_weights": (C.c_int, [_eng, _dp]),
    "rsrl_engine_rollout": (C.c_int, [_eng, C.c_int64, _dp, C.c_int64, C.c_int32, C.c_uint64, _dp, _dp, _ip, _dp, _u8p, _ip]),
    "rsrl_engine_get_traces": (C.c_int, [_eng, _dp]),
    "rsrl_engine_set_traces": (C.c_int, [_eng, _dp]),
    "rsrl_engine_get_td_errors": (C.c_int, [_eng, _dp]),
    "rsrl_engine_get_stats": (C.c_int, [_eng, _P(Stats)]),
    "rsrl_engine_get_env_stats": (C.c_int, [_eng, _ip, _ip, _u64p]),
    "rsrl_engine_set_epsilon": (C.c_int, [_eng, C.c_double]),
    "rsrl_engine_evaluate": (C.c_int, [_eng, C.c_int64, _dp, _dp]),
    "rsrl_engine_sample": (C.c_int, [_eng, C.c_int64, _dp, C.c_uint64, _ip]),
    "rsrl_engine_mode": (C.c_int, [_eng, C.c_int64, _dp, _ip]),
    "rsrl_engine_handle": (C.c_int, [_eng, C.c_int64, _dp, _ip, _dp, _dp, _u8p, C.c_uint64, _dp]),
    "rsrl_engine_get_launch_shape": (C.c_int, [_eng, _ip]),
    "rsrl_math_probe": (C.c_int, [C.c_int32, C.c_int64, _dp, _dp]),
    "rsrl_comm_unique_id": (C.c_int, [_u8p]),
    "rsrl_engine_comm_init": (C.c_int, [_eng, _u8p, C.c_int, C.c_int]),
    "rsrl_engine_peer_export": (C.c_int, [_eng, _u8p]),
    "rsrl_engine_peer_attach": (C.c_int, [_eng, _u8p, C.c_int, C.c_int]),
    "rsrl_domain_info": (C.c_int, [C.c_int32, _ip, _ip, _dp, _dp, _dp]),
    "rsrl_domain_step": (C.c_int, [C.c_int32, C.c_int64, _dp, _ip, _dp, _u8p]),
    "rsrl_domain_is_terminal": (C.c_int, [C.c_int32, C.c_int64, _dp, _u8p]),
    "rsrl_domain_ex_info": (C.c_int, [C.c_int32, _ip, _ip, _dp, _dp, _dp]),
    "rsrl_domain_ex_step": (C.c_int, [C.c_int32, C.c_int64, _dp, _ip, _dp, _dp, _dp, _u8p]),
    "rsrl_domain_ex_emit": (C.c_int, [C.c_int32, C.c_int64, _dp, _dp, _u8p]),
    "rsrl_domain_ex_last_error": (C.c_char_p, []),
    "rsrl_basis_project": (C.c_int, [_cfgp, C.c_int64, _dp, _dp]),
    "rsrl_lfa_evaluate": (C.c_int, [_cfgp, C.c_int64, _dp, _dp, _dp]),
    "rsrl_lfa_update_index": (C.c_int, [_cfgp, C.c_int64, _dp, _ip, _dp, _dp]),
    "rsrl_policy_sample": (C.c_int, [C.c_int32, C.c_double, C.c_uint64, C.c_uint64, C.c_int64, C.c_int64, C.c_int32,
                                     _dp, _ip]),
    "rsrl_policy_probs": (C.c_int, [C.c_int32, C.c_double, C.c_int64, C.c_int32, _dp, _dp]),
    "rsrl_policy_mode": (C.c_int, [C.c_int64, C.c_int32, _dp, _ip]),
    "rsrl_trace_update": (C.c_int, [C.c_int32, C.c_double, C.c_double, C.c_double, C.c_int64, _dp, _dp]),
    "rsrl_philox": (C.c_int, [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, _u32p]),
}

_lib = None


class RsrlError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rsrl_b200 error {code}: {msg}")
        self.code = code


def load():
    """dlopen librsrl_b200.so (built by __graft_entry__.build()).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback for the rsrl_b200 hot path)")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def check(code):
    if code != OK:
        msg = load().rsrl_last_error()
        raise RsrlError(code, msg.decode() if msg else "")
    return code


def dp(a):
    return a.ctypes.data_as(_dp)


def ip(a):
    return a.ctypes.data_as(_ip)


def u8p(a):
    return a.ctypes.data_as(_u8p)


def u32p(a):
    return a.ctypes.data_as(_u32p)


def u64p(a):
    return a.ctypes.data_as(_u64p)

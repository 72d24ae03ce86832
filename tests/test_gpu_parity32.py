"""GPU parity of the BENCH dtype (fp32 features / Q / weights / traces + f64 physics) — bit for bit.

oracle/oracle32.cpp compiles the product's own arithmetic headers (csrc/hostdev.h, device.cuh, core.cuh) for the host
and replays the summation order of persistent.cuh for the launch shape the engine reports
(rsrl_engine_get_launch_shape).  Free-running fp32 engines therefore have to reproduce it EXACTLY: action indices, episode
step counts and length hashes (north_star), and also states, TD errors, weights and traces.  The distance between this fp32
arithmetic and the reference's f64 is the separate, tolerance-based statement of the teacher-forced tests against
oracle/rsrl_oracle.c (here at BASELINE sizes; tests/test_gpu_parity.py at small sizes).
"""
import numpy as np
import pytest

from rsrl_b200 import abi

pytestmark = pytest.mark.gpu

MC, CP, AC = abi.MOUNTAIN_CAR, abi.CART_POLE, abi.ACROBOT


@pytest.fixture(scope="module")
def E(rsrl):
    from rsrl_b200 import engine
    assert abi.load().rsrl_device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return engine


def _cfg(**kw):
    base = dict(n_envs=65536, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN, record_td_error=1)
    base.update(kw)
    return abi.default_config(**base)


def _exact(e, o, what=""):
    """GPU engine e == oracle32 engine o, bit for bit."""
    assert (e.actions() == o.actions()).all(), f"{what}: action indices"
    assert (e.episode_steps() == o.episode_steps()).all(), f"{what}: episode step counts"
    for name, x, y in zip(("n_episodes", "last_len", "len_hash"), e.env_stats(), o.env_stats()):
        assert (x == y).all(), f"{what}: {name}"
    assert (e.states() == o.states()).all(), f"{what}: states"
    assert (e.td_errors() == o.td_errors()).all(), f"{what}: TD errors"
    assert (e.weights() == o.weights()).all(), f"{what}: weights"
    st, oc = e.stats(), o.counters()
    assert st["total_episodes"] == oc["total_episodes"] and st["terminal_episodes"] == oc["terminal_episodes"]


# ---------------------------------------------------------------------------------------------
# the elementary functions and the physics: same bits on both sides
# ---------------------------------------------------------------------------------------------
def _same_bits(a, b):
    """identical bit patterns; NaNs only have to be NaN on both sides (x86 and the GPU differ in the default NaN's sign bit)"""
    nan = np.isnan(a)
    return (nan == np.isnan(b)).all() and (a[~nan].view(np.uint64) == b[~nan].view(np.uint64)).all()


def test_math_functions_bit_exact(E, oracle32):
    rng = np.random.default_rng(0)
    xs64 = np.concatenate([rng.uniform(-40, 40, 200000), rng.uniform(-4, 4, 200000), rng.uniform(-1e6, 1e6, 50000),
                           [0.0, -0.0, 1e-300, np.pi / 2, np.pi, 1048575.9, 1048576.0, 1e7, np.inf, -np.inf, np.nan]])
    for fn in (0, 1):
        a, b = E.math_probe(fn, xs64), oracle32.math(fn, xs64)
        inside = np.abs(xs64) < 1048576.0     # beyond: the platform libm on each side (documented, never reached by the domains)
        assert _same_bits(a[inside], b[inside]), f"fn {fn}"
    xs32 = np.concatenate([rng.uniform(0, 1, 300000), rng.uniform(-8, 8, 100000), rng.uniform(-1e7, 1e7, 1000),
                           [0.0, 0.25, 0.5, 0.75, 1.0, 1e10, np.inf, np.nan]]).astype(np.float32).astype(np.float64)
    for fn in (2, 3):
        a, b = E.math_probe(fn, xs32), oracle32.math(fn, xs32)
        assert _same_bits(a, b), f"fn {fn}"
    xe = np.concatenate([rng.uniform(-110, 90, 300000), rng.uniform(-5, 5, 100000), [0.0, 88.7228, 88.8, -103.9, -104.0, np.inf, -np.inf, np.nan]])
    xe = xe.astype(np.float32).astype(np.float64)
    a, b = E.math_probe(4, xe), oracle32.math(4, xe)
    assert _same_bits(a, b)


@pytest.mark.parametrize("domain", [MC, CP, AC])
def test_domain_step_bit_exact(E, oracle32, oracle, domain):
    rng = np.random.default_rng(domain)
    lo, hi = oracle.domain_limits(domain)
    n = 20000
    s = rng.uniform(lo, hi, size=(n, len(lo)))
    A = 2 if domain == CP else 3
    for _ in range(5):
        a = rng.integers(0, A, n).astype(np.int32)
        ns, r, t = E.domain_step(domain, s, a)
        ns2, r2, t2 = oracle32.domain_step(domain, s, a)
        assert (ns.view(np.uint64) == ns2.view(np.uint64)).all() and (r == r2).all() and (t == t2).all()
        # and within a few ulp of the libm-based f64 oracle (the reference's arithmetic)
        ns3, r3, t3 = oracle.domain_step(domain, s, a)
        assert (np.abs(ns - ns3) / np.maximum(1.0, np.abs(ns3))).max() < (1e-12 if domain == AC else 2e-14) and (t == t3).mean() > 0.9999
        s = ns


# ---------------------------------------------------------------------------------------------
# BASELINE configs[1] at full size: free run, fp32, bit for bit (weights included => every step's dW sum is checked)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cap,steps", [(1000, 250), (64, 200)], ids=["bench_config", "with_resets"])
def test_cfg2_full_size_free_run_bit_exact(E, oracle32, cap, steps):
    cfg = _cfg(max_episode_steps=cap)
    with E.Engine(cfg) as e:
        sh = e.launch_shape()
        assert sh["persistent"] == 1 and sh["grid"] >= 64, sh   # the kernel bench.py times
        o = oracle32.Engine(cfg, sh)
        for k in (1, 9, steps - 10):
            e.step(k)
            o.step(k)
            e.sync()
            _exact(e, o, f"cfg2 after {k} more steps")
        assert cap > 200 or e.stats()["total_episodes"] > 65536


def test_cfg2_full_size_teacher_forced_vs_f64_oracle(E, oracle):
    """20 steps, each from the device's own state and weights: the fp32 path against the reference's f64 arithmetic.
    Tolerances: next states 1e-12 (f64 physics, <= 2 ulp trig), TD errors 2e-6 * sqrt(36) relative to the largest |TD|,
    weights 1e-6 relative to the largest weight; actions equal wherever the f64 decision margin exceeds 1e-4."""
    cfg = _cfg()
    rng = np.random.default_rng(5)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(rng.normal(size=(36, 3)) * 0.5)
        e.step(3)
        o.step(3)   # same batched-step index on both sides: it is the RNG draw counter
        for t in range(20):
            o.set_states(e.states())
            o.set_weights(e.weights())
            q = oracle.evaluate(cfg, e.weights(), e.states())
            srt = np.sort(q, axis=1)
            safe = srt[:, -1] - srt[:, -2] > 1e-4
            e.step(1)
            o.step(1)
            e.sync()
            assert safe.mean() > 0.97
            assert (e.actions()[safe] == o.actions()[safe]).all()
            same = e.actions() == o.actions()
            assert np.abs(e.states()[same] - o.states()[same]).max() < 1e-12
            scale = max(1.0, np.abs(o.td_errors()).max(), np.abs(o.weights()).max())
            assert np.abs(e.td_errors()[same] - o.td_errors()[same]).max() < 2e-6 * 6 * scale
            if same.all():
                assert np.abs(e.weights() - o.weights()).max() < 1e-6 * max(1.0, np.abs(o.weights()).max())


# ---------------------------------------------------------------------------------------------
# every agent / policy / weight mode of the persistent kernel, ragged sizes, all three domains
# ---------------------------------------------------------------------------------------------
AGENTS = [
    ("qlearning_greedy", dict(algo=abi.QLEARNING, policy=abi.GREEDY)),
    ("qlearning_eps", dict(algo=abi.QLEARNING, policy=abi.EPSILON_GREEDY, epsilon=0.1)),
    ("sarsa_eps", dict(algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.01)),
    ("expected_sarsa_eps", dict(algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, alpha=0.5, lr=0.01)),
    ("sarsa_softmax", dict(algo=abi.SARSA, policy=abi.SOFTMAX, epsilon=0.7, gamma=0.95, lr=0.01)),
    ("pal_eps", dict(algo=abi.PAL, policy=abi.EPSILON_GREEDY, epsilon=0.1, alpha=0.5, gamma=0.95, lr=0.01)),
    ("td0", dict(algo=abi.TD0, policy=abi.RANDOM, gamma=0.99, lr=0.01)),
    ("sarsa_lambda_replace", dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99,
                                  trace_rule=abi.TRACE_REPLACE)),
    ("q_lambda_accumulate", dict(algo=abi.Q_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.002, gamma=0.99,
                                 trace_rule=abi.TRACE_ACCUMULATE)),
    ("td_lambda", dict(algo=abi.TD_LAMBDA, policy=abi.RANDOM, gamma=0.99, trace_rule=abi.TRACE_ACCUMULATE, lambda_=0.5)),
    ("sarsa_lambda_dutch", dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99,
                                trace_rule=abi.TRACE_DUTCH)),
]


@pytest.mark.parametrize("name,kw", AGENTS, ids=[a[0] for a in AGENTS])
@pytest.mark.parametrize("n", [700, 5000])
def test_agents_free_run_bit_exact(E, oracle32, name, kw, n):
    cfg = _cfg(n_envs=n, max_episode_steps=90, seed=7, **kw)
    with E.Engine(cfg) as e:
        sh = e.launch_shape()
        assert sh["persistent"] == 1, sh
        o = oracle32.Engine(cfg, sh)
        for k in (1, 6, 30 if name == "td_lambda" else 160):   # TDLambda has no step size (td_lambda.rs:56-59): it diverges by design
            e.step(k)
            o.step(k)
            e.sync()
            _exact(e, o, name)
            if cfg.algo in (abi.SARSA_LAMBDA, abi.Q_LAMBDA, abi.TD_LAMBDA):
                assert (e.traces() == o.traces()).all(), "traces"


@pytest.mark.parametrize("n", [1, 33, 449, 1000, 2049, 20000])
@pytest.mark.parametrize("mode", [abi.SHARED, abi.PER_ENV], ids=["shared", "per_env"])
def test_ragged_sizes_bit_exact(E, oracle32, n, mode):
    cfg = _cfg(n_envs=n, weight_mode=mode, update_scale=abi.SCALE_SUM if mode == abi.PER_ENV else abi.SCALE_MEAN,
               max_episode_steps=70, seed=n, policy=abi.EPSILON_GREEDY, epsilon=0.05)
    with E.Engine(cfg) as e:
        o = oracle32.Engine(cfg, e.launch_shape())
        e.step(150)
        o.step(150)
        e.sync()
        _exact(e, o, f"n={n}")


@pytest.mark.parametrize("basis,order", [(abi.FOURIER, 1), (abi.FOURIER, 2), (abi.FOURIER, 3), (abi.FOURIER, 7), (abi.POLYNOMIAL, 2), (abi.POLYNOMIAL, 3)])
@pytest.mark.parametrize("n", [700, 65536])
def test_every_mountain_car_basis_bit_exact(E, oracle32, basis, order, n):
    """Every instantiated MountainCar basis on the persistent kernel: odd and even orders of the paired (FMUL2 / FFMA2) feature
    generation, odd NV (Polynomial 2: 27 values), the run-time reduce stride of the 64-feature basis, tail passes of 4 / 9 / 16 / 0 rows."""
    cfg = _cfg(basis=basis, basis_order=order, n_envs=n, policy=abi.EPSILON_GREEDY, epsilon=0.1, lr=0.01, max_episode_steps=50, seed=order)
    with E.Engine(cfg) as e:
        sh = e.launch_shape()
        assert sh["persistent"] == 1
        o = oracle32.Engine(cfg, sh)
        for k in (1, 79):
            e.step(k)
            o.step(k)
            e.sync()
            _exact(e, o, f"basis={basis} order={order} n={n}")


@pytest.mark.parametrize("domain,order,algo", [(CP, 3, abi.SARSA), (AC, 2, abi.EXPECTED_SARSA), (CP, 2, abi.QLEARNING)])
def test_d4_domains_bit_exact(E, oracle32, domain, order, algo):
    cfg = _cfg(domain=domain, basis_order=order, algo=algo, policy=abi.EPSILON_GREEDY, epsilon=0.1, n_envs=3000, lr=0.01, alpha=0.5, gamma=0.99,
               init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=60, seed=2)
    with E.Engine(cfg) as e:
        sh = e.launch_shape()
        if not sh["persistent"]:
            pytest.skip("shape runs on the per-step kernels (reduce buffers exceed shared memory)")
        o = oracle32.Engine(cfg, sh)
        e.step(120)
        o.step(120)
        e.sync()
        _exact(e, o, "d4")


# ---------------------------------------------------------------------------------------------
# BASELINE configs[4] (per-env eligibility traces) at its per-GPU shard size, fp32: the kernel bench.py's cfg5 line times
# ---------------------------------------------------------------------------------------------
CFG5 = [("sarsa_lambda", dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99)),
        ("q_lambda", dict(algo=abi.Q_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99)),
        ("td_lambda", dict(algo=abi.TD_LAMBDA, policy=abi.RANDOM, gamma=0.99, lambda_=0.3))]


@pytest.mark.parametrize("name,kw", CFG5, ids=[c[0] for c in CFG5])
def test_cfg5_shard_free_run_bit_exact(E, oracle32, name, kw):
    cfg = _cfg(n_envs=32768, **kw)
    with E.Engine(cfg) as e:
        sh = e.launch_shape()
        assert sh["persistent"] == 1 and sh["mode"] == 2, sh   # traces resident in shared memory
        o = oracle32.Engine(cfg, sh)
        for k in (1, 10, 14 if name == "td_lambda" else 90):
            e.step(k)
            o.step(k)
            e.sync()
            _exact(e, o, name)
            assert (e.traces() == o.traces()).all(), "traces"


@pytest.mark.parametrize("name,kw", CFG5, ids=[c[0] for c in CFG5])
def test_cfg5_shard_teacher_forced_vs_f64_oracle(E, oracle, name, kw):
    """fp32 traces against the reference's f64 arithmetic, 8 steps from the device's own state / weights / traces."""
    cfg = _cfg(n_envs=32768, **kw)
    rng = np.random.default_rng(6)
    aw = 1 if name == "td_lambda" else 3
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(rng.normal(size=(36, aw)) * 0.3)
        e.step(4)
        o.step(4)   # same batched-step index on both sides: it is the RNG draw counter
        for t in range(8):
            o.set_states(e.states())
            o.set_weights(e.weights())
            o.set_traces(e.traces())
            e.step(1)
            o.step(1)
            e.sync()
            same = e.actions() == o.actions()
            assert same.mean() > 0.98   # eps-greedy / softmax draws are integer work; greedy picks can flip at fp32 near-ties
            assert np.abs(e.states()[same] - o.states()[same]).max() < 1e-12
            scale = max(1.0, np.abs(o.td_errors()).max(), np.abs(o.weights()).max())
            assert np.abs(e.td_errors()[same] - o.td_errors()[same]).max() < 2e-6 * 6 * scale
            assert np.abs(e.traces()[same] - o.traces()[same]).max() < 2e-6
            if same.all():
                assert np.abs(e.weights() - o.weights()).max() < 2e-6 * max(1.0, np.abs(o.weights()).max())


# ---------------------------------------------------------------------------------------------
# the 16-byte LL lines of the exchange between cluster leaders (persistent.cuh: ld_ll / st_ll)
# ---------------------------------------------------------------------------------------------
def test_ll_exchange_stress_reproducible(E):
    """10^6 batched steps of the bench kernel, twice: a torn 16-byte line (payload of one epoch, flag of another) would make
    the two runs differ.  Together with the full-size bit-exact runs above (every step's total checked against oracle32)
    this is the evidence for treating an aligned 16-byte access as single-copy atomic on this part."""
    cfg = _cfg(record_td_error=0)
    res = []
    for _ in range(2):
        with E.Engine(cfg) as e:
            for _ in range(10):
                e.step(100000)
            e.sync()
            res.append((e.weights(), e.states(), e.env_stats()[2]))
            assert e.stats()["batch_steps"] == 1000000
    assert (res[0][0] == res[1][0]).all() and (res[0][1] == res[1][1]).all() and (res[0][2] == res[1][2]).all()
    assert np.isfinite(res[0][0]).all()

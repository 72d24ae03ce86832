"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (north_star): action indices / episode step counts bit-exact; f64 physics within a few ulp of the
libm-based oracle; Q weights within a stated tolerance — 1e-9 for dtype f64, the fp32 tolerances written
next to each f32 assertion.  Reference golden vectors (cart_pole.rs:144-183, greedy.rs:96-168, ...) are
re-checked on the device as well.
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


# ---------------------------------------------------------------------------------------------
# RNG: integer work, bit-exact
# ---------------------------------------------------------------------------------------------
def test_philox_bit_exact(E, oracle):
    got = E.philox(seed=0x1234567890ABCDEF, draw=(1 << 33) + 5, stream=2, env_offset=7, n=1000)
    want = np.array([oracle.draw(0x1234567890ABCDEF, 7 + i, (1 << 33) + 5, 2) for i in range(1000)])
    assert (got == want).all()
    kat = E.philox(seed=0, draw=0, stream=0, env_offset=0, n=1)[0]
    assert [int(x) for x in kat] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


# ---------------------------------------------------------------------------------------------
# K1 domains
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("action,sign", [(0, -1.0), (1, 1.0)])
def test_cartpole_golden_on_device(E, action, sign):
    g1 = np.array([0.0032931628891235, 0.3293940797883472, -0.0029499634056967, -0.2951522145037250]) * sign
    g2 = np.array([0.0131819582085161, 0.6597158115002169, -0.0118185373734479, -0.5921703414056713]) * sign
    s1, r1, t1 = E.domain_step(CP, np.zeros((1, 4)), [action])
    s2, r2, t2 = E.domain_step(CP, s1, [action])
    assert np.abs(s1[0] - g1).max() < 1e-7 and np.abs(s2[0] - g2).max() < 1e-7  # reference tolerance
    assert np.abs(s1[0] - g1).max() < 1e-15 and np.abs(s2[0] - g2).max() < 1e-15
    assert r1[0] == 0.0 and not t1[0]


def test_mountain_car_terminal_predicate_on_device(E):
    X = 0.6
    s = np.array([[-0.5, 0.0], [X, -0.05], [X, 0.0], [X, 0.05], [X - 1e-4 * X, 0.0], [X + 1e-4 * X, 0.0]])
    assert E.domain_is_terminal(MC, s).tolist() == [0, 1, 1, 1, 0, 1]
    assert not E.domain_is_terminal(CP, np.zeros((1, 4)))[0] and not E.domain_is_terminal(AC, np.zeros((1, 4)))[0]


@pytest.mark.parametrize("domain", [MC, CP, AC])
def test_domain_step_matches_oracle(E, oracle, domain):
    D, A = oracle.domain_dims(domain)
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(domain)
    n = 4096
    s = rng.uniform(lo, hi, size=(n, D))
    s[:64] = np.where(rng.random((64, D)) < 0.5, lo, hi)  # corners: clip / terminal edges
    a = rng.integers(0, A, n).astype(np.int32)
    got_s, got_r, got_t = E.domain_step(domain, s, a)
    want_s, want_r, want_t = oracle.domain_step(domain, s, a)
    # f64 physics; the only difference is CUDA's sin/cos vs glibc's (both < 1 ulp)
    scale = np.maximum(np.abs(want_s), 1.0)
    assert (np.abs(got_s - want_s) / scale).max() < 1e-13
    assert (got_r == want_r).all() and (got_t == want_t).all()
    assert (got_s == want_s).mean() > 0.9  # mostly bit-identical


def test_domain_step_multi_step_trajectory(E, oracle):
    # 500 steps of MountainCar bang-bang control: positions must track the oracle to 1e-12
    n = 256
    rng = np.random.default_rng(5)
    s_g = np.tile([-0.5, 0.0], (n, 1)) + rng.uniform(-0.1, 0.1, (n, 2)) * [1, 0]
    s_o = s_g.copy()
    for t in range(500):
        a = (s_o[:, 1] >= 0).astype(np.int32) * 2
        s_g, _, tg = E.domain_step(MC, s_g, a)
        s_o, _, to = oracle.domain_step(MC, s_o, a)
        assert (tg == to).all()
    assert np.abs(s_g - s_o).max() < 1e-12


# ---------------------------------------------------------------------------------------------
# K2/K3 basis projection + LFA evaluate
# ---------------------------------------------------------------------------------------------
COMBOS = [(MC, abi.FOURIER, 5), (MC, abi.FOURIER, 3), (MC, abi.FOURIER, 7), (MC, abi.POLYNOMIAL, 3),
          (CP, abi.FOURIER, 3), (AC, abi.FOURIER, 2), (CP, abi.POLYNOMIAL, 2)]


@pytest.mark.parametrize("domain,basis,order", COMBOS)
@pytest.mark.parametrize("dtype,tol", [(abi.F64, 2e-14), (abi.F32, 4e-6)])
def test_basis_project_matches_oracle(E, oracle, domain, basis, order, dtype, tol):
    cfg = abi.default_config(domain=domain, basis=basis, basis_order=order, dtype=dtype)
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(order)
    s = rng.uniform(lo, hi, size=(512, len(lo)))
    s[0], s[1] = lo, hi
    got, want = E.basis_project(cfg, s), oracle.project(cfg, s)
    assert got.shape == want.shape
    scale = np.maximum(np.abs(want), 1.0)
    assert (np.abs(got - want) / scale).max() < tol * (order if basis == abi.FOURIER else 1)
    assert (got[:, -1] == 1.0).all()


@pytest.mark.parametrize("domain,basis,order", COMBOS[:5])
@pytest.mark.parametrize("dtype,tol", [(abi.F64, 1e-12), (abi.F32, 2e-4)])
def test_lfa_evaluate_matches_oracle(E, oracle, domain, basis, order, dtype, tol):
    cfg = abi.default_config(domain=domain, basis=basis, basis_order=order, dtype=dtype)
    D, A = oracle.domain_dims(domain)
    F = oracle.n_features(cfg)
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(7)
    s = rng.uniform(lo, hi, size=(300, D))
    W = rng.normal(size=(F, A))
    got, want = E.lfa_evaluate(cfg, W, s), oracle.evaluate(cfg, W, s)
    assert np.abs(got - want).max() < tol * np.sqrt(F)


@pytest.mark.parametrize("dtype,tol", [(abi.F64, 1e-13), (abi.F32, 1e-6)])
def test_lfa_update_index_matches_oracle(E, oracle, dtype, tol):
    cfg = abi.default_config(dtype=dtype, lr=0.05)
    rng = np.random.default_rng(2)
    W0 = rng.normal(size=(36, 3))
    s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(50, 2))
    a = rng.integers(0, 3, 50)
    err = rng.normal(size=50)
    got = E.lfa_update_index(cfg, W0, s, a, err)
    want = W0.copy()
    for i in range(50):  # SGD: W[:, a] += (lr * err) * phi(s)
        want = oracle.update_index(cfg, want, s[i], a[i], 0.05 * err[i])
    assert np.abs(got - want).max() < tol


# ---------------------------------------------------------------------------------------------
# K5 policies: the reference's MockQ tests run against the device kernels
# ---------------------------------------------------------------------------------------------
def test_greedy_reference_cases_on_device(E):
    g = lambda q: int(E.policy_sample(abi.GREEDY, 0.0, 0, 0, 0, [q])[0])
    assert g([1.0]) == 0 and g([-100.0]) == 0
    assert g([10.0, 1.0]) == 0 and g([1.0, 10.0]) == 1
    assert g([-10.0, -1.0]) == 1 and g([-1.0, -10.0]) == 0
    assert g([10.0, -1.0]) == 0 and g([-10.0, 1.0]) == 1 and g([1.0, -10.0]) == 0 and g([-1.0, 10.0]) == 1
    assert g([-123.1, 123.1, 250.5, -1240.0, -4500.0, 10000.0, 20.1]) == 5
    assert g([1e-7, 2e-7]) == 1


def test_policy_probabilities_on_device(E):
    p = E.policy_probs(abi.GREEDY, 0.0, [[1e-7] * 4, [1e-7, 2e-7, 3e-7, 4e-7]])
    assert np.abs(p - [[0.25] * 4, [0, 0, 0, 1]]).max() < 1e-6
    p = E.policy_probs(abi.EPSILON_GREEDY, 0.5, [[1, 0, 0, 0, 0], [0, 0, 0, 0, 1], [1, 0, 0, 0, 1]])
    assert np.abs(p - [[0.6, 0.1, 0.1, 0.1, 0.1], [0.1, 0.1, 0.1, 0.1, 0.6], [0.35, 0.1, 0.1, 0.1, 0.35]]).max() < 1e-6
    p = E.policy_probs(abi.EPSILON_GREEDY, 1.0, [[-1.0, 0, 0, 0]])
    assert np.abs(p - 0.25).max() < 1e-6


def test_policy_sample_bit_exact_vs_oracle(E, oracle):
    rng = np.random.default_rng(3)
    q = rng.normal(size=(5000, 3))
    q[:1000] = np.round(q[:1000])          # many exact ties
    q[1000:1500, 1] = q[1000:1500, 0] + 5e-8  # inside the 1e-7 tolerance
    for policy, eps in [(abi.GREEDY, 0.0), (abi.EPSILON_GREEDY, 0.3), (abi.EPSILON_GREEDY, 1.0), (abi.RANDOM, 0.0)]:
        got = E.policy_sample(policy, eps, 99, 12, 1000, q)
        want = oracle.policy_sample_batch(policy, eps, 99, 12, 1000, q)
        assert (got == want).all()
    assert (E.policy_mode(q) == [oracle.find_max(r)[0] for r in q]).all()


def test_epsilon_greedy_frequencies_on_device(E):
    acts = E.policy_sample(abi.EPSILON_GREEDY, 0.5, 1, 0, 0, np.tile([1.0, 0.0], (10000, 1)))
    assert abs(0.75 - (acts == 0).mean()) < 0.05  # epsilon_greedy.rs:96-113
    acts = E.policy_sample(abi.RANDOM, 0.0, 1, 0, 0, np.zeros((10000, 2)))
    assert abs(0.5 - (acts == 0).mean()) < 0.05   # random.rs:58-76


def test_nonfinite_q_is_an_error(E):
    with pytest.raises(abi.RsrlError) as ei:
        E.policy_sample(abi.GREEDY, 0.0, 0, 0, 0, [[np.nan, np.nan]])
    assert ei.value.code == abi.ENONFINITE  # the reference panics (utils.rs:76)


# ---------------------------------------------------------------------------------------------
# K6 traces
# ---------------------------------------------------------------------------------------------
def test_trace_rules_on_device(E, oracle):
    z = E.trace_update(abi.TRACE_ACCUMULATE, 0.95, 0.7, 0.0, np.zeros(1), np.ones(1))
    z = E.trace_update(abi.TRACE_ACCUMULATE, 0.95, 0.7, 0.0, z, np.zeros(1))
    assert abs(z[0] - 0.665) < 1e-12  # traces.rs:121-125
    rng = np.random.default_rng(0)
    z0, g = rng.normal(size=1000), rng.normal(size=1000)
    for rule in (abi.TRACE_ACCUMULATE, abi.TRACE_REPLACE, abi.TRACE_DUTCH):
        got, want = E.trace_update(rule, 0.99, 0.7, 0.1, z0, g), oracle.trace_update(rule, 0.99, 0.7, 0.1, z0, g)
        assert np.abs(got - want).max() < 1e-15


# ---------------------------------------------------------------------------------------------
# fused engine, free running, dtype f64: bit-exact actions and step counts
# ---------------------------------------------------------------------------------------------
def _mc_cfg(**kw):
    base = dict(n_envs=64, dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                max_episode_steps=300, seed=11, record_td_error=1)
    base.update(kw)
    return abi.default_config(**base)


def _compare_engines(e, o, w_tol, s_tol=1e-9):
    assert (e.actions() == o.actions()).all()
    assert (e.episode_steps() == o.episode_steps()).all()
    for x, y in zip(e.env_stats(), o.env_stats()):
        assert (x == y).all()
    se, so = e.stats(), o.stats()
    for k in ("total_steps", "total_episodes", "terminal_episodes", "batch_steps", "nonfinite"):
        assert se[k] == so[k], k
    rel = lambda a, b: np.abs(a - b).max() / max(1.0, np.abs(b).max())
    assert rel(e.states(), o.states()) < s_tol
    assert rel(e.weights(), o.weights()) < w_tol
    assert rel(e.td_errors(), o.td_errors()) < max(w_tol * 100, 1e-9)


ALGOS = [
    ("qlearning_greedy", dict(algo=abi.QLEARNING, policy=abi.GREEDY)),
    ("qlearning_eps", dict(algo=abi.QLEARNING, policy=abi.EPSILON_GREEDY, epsilon=0.1)),
    ("sarsa_eps", dict(algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.01)),
    ("expected_sarsa_eps", dict(algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, alpha=0.5, lr=0.01)),
    ("sarsa_softmax", dict(algo=abi.SARSA, policy=abi.SOFTMAX, epsilon=0.7, gamma=0.95, lr=0.01)),  # policies/softmax.rs (epsilon field = tau)
    ("expected_sarsa_softmax", dict(algo=abi.EXPECTED_SARSA, policy=abi.SOFTMAX, epsilon=2.0, alpha=0.5, lr=0.01)),
    ("pal_eps", dict(algo=abi.PAL, policy=abi.EPSILON_GREEDY, epsilon=0.1, alpha=0.5, gamma=0.95, lr=0.01)),  # control/td/pal.rs
    ("sarsa_lambda_replace", dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99,
                                  trace_rule=abi.TRACE_REPLACE)),
    ("q_lambda_accumulate", dict(algo=abi.Q_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.002, gamma=0.99,
                                 trace_rule=abi.TRACE_ACCUMULATE)),
    ("td_lambda", dict(algo=abi.TD_LAMBDA, policy=abi.RANDOM, gamma=0.99, trace_rule=abi.TRACE_ACCUMULATE,
                       update_scale=abi.SCALE_MEAN, lambda_=0.5)),
    ("td0", dict(algo=abi.TD0, policy=abi.RANDOM, gamma=0.99, lr=0.01)),
]


@pytest.mark.parametrize("name,kw", ALGOS, ids=[a[0] for a in ALGOS])
@pytest.mark.parametrize("mode", [abi.SHARED, abi.PER_ENV], ids=["shared", "per_env"])
def test_engine_free_run_f64_bit_exact_actions(E, oracle, name, kw, mode):
    # TDLambda has no step size at all (td_lambda.rs:56-59: W += td_error * z), so with Fourier features it
    # diverges geometrically exactly like the reference would; compare (relative tolerance) over a short run.
    chunks = (1, 7, 12) if name == "td_lambda" else (1, 7, 392)
    cfg = _mc_cfg(weight_mode=mode, **kw)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        for chunk in chunks:
            e.step(chunk)
            o.step(chunk)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9)
        if cfg.algo in (abi.SARSA_LAMBDA, abi.Q_LAMBDA, abi.TD_LAMBDA):
            assert np.abs(e.traces() - o.traces()).max() < 1e-9
        assert name == "td_lambda" or o.stats()["total_episodes"] > 0


def test_engine_q_learning_example_n1(E, oracle):
    """BASELINE config 1: exactly examples/q_learning.rs (1 env, default start, SGD(0.001), gamma 0.9,
    greedy) with a step cap; episode lengths must match the oracle bit for bit."""
    cfg = abi.default_config(dtype=abi.F64, max_episode_steps=10000)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(30000)
        o.step(30000)
        e.sync()
        _n, _l, h_e = e.env_stats()
        _n2, _l2, h_o = o.env_stats()
        assert (h_e == h_o).all() and (_n == _n2).all() and _n[0] >= 3
        assert (e.actions() == o.actions()).all() and (e.episode_steps() == o.episode_steps()).all()
        assert np.abs(e.weights() - o.weights()).max() < 1e-9


def test_n1_shared_equals_per_env(E):
    """With N = 1 and SUM scaling the batched semantics reduce to the reference loop: both weight modes agree."""
    outs = []
    for mode in (abi.SHARED, abi.PER_ENV):
        cfg = abi.default_config(dtype=abi.F64, max_episode_steps=2000, weight_mode=mode)
        with E.Engine(cfg) as e:
            e.step(5000)
            e.sync()
            outs.append((e.weights().reshape(36, 3), e.states(), e.env_stats()[2]))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all() and (outs[0][2] == outs[1][2]).all()


@pytest.mark.parametrize("domain,order,algo", [(CP, 3, abi.SARSA), (AC, 2, abi.EXPECTED_SARSA)])
def test_engine_free_run_d4_domains(E, oracle, domain, order, algo):
    cfg = abi.default_config(domain=domain, basis_order=order, algo=algo, policy=abi.EPSILON_GREEDY, epsilon=0.1,
                             n_envs=33, dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.05] * 4, init_hi=[0.05] * 4,
                             max_episode_steps=100, seed=5, gamma=0.99, lr=0.01, alpha=1.0, record_td_error=1,
                             update_scale=abi.SCALE_MEAN)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(150)
        o.step(150)
        e.sync()
        _compare_engines(e, o, w_tol=1e-9, s_tol=1e-8)


# ---------------------------------------------------------------------------------------------
# dtype f32 (bench dtype): teacher-forced single steps within fp32 tolerance + short free run
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name,kw", ALGOS[:4] + ALGOS[6:7], ids=[a[0] for a in ALGOS[:4] + ALGOS[6:7]])
def test_engine_f32_teacher_forced(E, oracle, name, kw):
    """Each step starts from the same inputs on both sides (oracle state/weights copied from the device):
    next states within 1e-12 (f64 physics), TD errors within 1.2e-5 relative to the largest |TD error| / |weight|, weights within 1e-6 relative (fp32 features/Q),
    actions identical wherever the oracle's decision margin exceeds the fp32 resolution of Q."""
    cfg = _mc_cfg(dtype=abi.F32, n_envs=512, update_scale=abi.SCALE_MEAN, **kw)   # MEAN: lr * 512 envs summed would diverge within a few steps
    rng = np.random.default_rng(4)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        W = rng.normal(size=(36, 3)) * 0.5
        e.set_weights(W)
        for t in range(20):
            o.set_states(e.states())
            o.set_weights(e.weights())
            q = oracle.evaluate(cfg, e.weights(), e.states())
            srt = np.sort(q, axis=1)
            margin = srt[:, -1] - srt[:, -2]
            e.step(1)
            o.step(1)
            e.sync()
            safe = margin > 1e-4
            assert safe.mean() > 0.95
            assert (e.actions()[safe] == o.actions()[safe]).all()
            same = e.actions() == o.actions()
            assert np.abs(e.states()[same] - o.states()[same]).max() < 1e-12
            td_scale = max(1.0, np.abs(o.td_errors()).max(), np.abs(o.weights()).max())
            assert np.abs(e.td_errors()[same] - o.td_errors()[same]).max() < 2e-6 * td_scale * 36 ** 0.5
            if same.all():  # stated fp32 tolerance on Q weights: 1e-6 relative to the largest weight (~8 ulp)
                assert np.abs(e.weights() - o.weights()).max() < 1e-6 * max(1.0, np.abs(o.weights()).max())


def test_engine_f32_free_run_short_horizon(E, oracle):
    cfg = _mc_cfg(dtype=abi.F32, n_envs=256, seed=22)  # seed chosen so that no decision comes within 3e-5 of a tie
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(60)
        o.step(60)
        e.sync()
        assert o.min_gap() > 3e-5, "oracle run came too close to a tie for an fp32 comparison"
        assert (e.actions() == o.actions()).all() and (e.episode_steps() == o.episode_steps()).all()
        assert np.abs(e.weights() - o.weights()).max() < 1e-4


# ---------------------------------------------------------------------------------------------
# trait-level entry points (Function::evaluate, Policy::sample/mode, Handler::handle)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("algo,policy", [(abi.QLEARNING, abi.GREEDY), (abi.SARSA, abi.EPSILON_GREEDY),
                                         (abi.EXPECTED_SARSA, abi.EPSILON_GREEDY)])
def test_handle_entry_point_matches_oracle(E, oracle, algo, policy):
    cfg = _mc_cfg(algo=algo, policy=policy, epsilon=0.25, alpha=0.7, lr=0.02, n_envs=100)
    rng = np.random.default_rng(9)
    W = rng.normal(size=(36, 3))
    s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(100, 2))
    a = rng.integers(0, 3, 100).astype(np.int32)
    ns, r, term = oracle.domain_step(MC, s, a)
    term[:10] = 1
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(W)
        o.set_weights(W)
        assert np.abs(e.evaluate(s) - oracle.evaluate(cfg, W, s)).max() < 1e-12
        assert (e.mode(s) == [oracle.find_max(q)[0] for q in oracle.evaluate(cfg, W, s)]).all()
        assert (e.sample(s, draw=3) == oracle.policy_sample_batch(policy, 0.25, cfg.seed, 3, 0, oracle.evaluate(cfg, W, s))).all()
        td_e = e.handle(s, a, r, ns, term, draw_idx=17)
        td_o = o.handle(s, a, r, ns, term, draw_idx=17)
        assert np.abs(td_e - td_o).max() < 1e-12
        assert np.abs(e.weights() - o.weights()).max() < 1e-12


def test_drop_in_loop_equals_fused_engine(E, oracle):
    """The reference's driver loop written against the trait-level entry points
    (env.transition -> agent.handle -> policy.sample; examples/q_learning.rs:40-52) gives the same
    weights as the fused engine."""
    cfg = _mc_cfg(n_envs=16, max_episode_steps=0, seed=2)
    with E.Engine(cfg) as fused, E.Engine(cfg) as unfused:
        fused.step(50)
        fused.sync()
        s = unfused.states()
        for t in range(50):
            a = unfused.sample(s, draw=t)
            ns, r, term = E.domain_step(MC, s, a)
            unfused.handle(s, a, r, ns, term, draw_idx=t)
            assert not term.any()
            s = ns
        assert np.abs(fused.states() - s).max() == 0.0
        assert np.abs(fused.weights() - unfused.weights()).max() < 1e-15


# ---------------------------------------------------------------------------------------------
# BASELINE sizes: size-independent properties
# ---------------------------------------------------------------------------------------------
def _cfg2(**kw):
    base = dict(n_envs=65536, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
    base.update(kw)
    return abi.default_config(**base)


def test_cfg2_full_size_properties(E, oracle):
    cfg = _cfg2()
    runs = []
    for _ in range(2):
        with E.Engine(cfg) as e:
            e.step(300)
            e.sync()
            runs.append((e.weights(), e.states(), e.actions(), e.env_stats()[2], e.stats()))
    # deterministic reduction order: bit-identical run to run
    assert (runs[0][0] == runs[1][0]).all() and (runs[0][1] == runs[1][1]).all() and (runs[0][3] == runs[1][3]).all()
    W, S, A, H, st = runs[0]
    assert st["total_steps"] == 65536 * 300 and st["nonfinite"] == 0
    assert (S[:, 0] >= -1.2).all() and (S[:, 0] <= 0.6).all() and (np.abs(S[:, 1]) <= 0.07).all()
    assert ((A >= 0) & (A < 3)).all() and np.isfinite(W).all() and np.abs(W).max() > 0
    # the first 512 envs of the same job on the oracle (MEAN scale uses the global env count) agree for one step
    sub = cfg.copy(n_envs=512, n_envs_global=65536)
    with E.Engine(sub) as e:
        o = oracle.Engine(sub)
        e.step(1)
        o.step(1)
        e.sync()
        assert (e.actions() == o.actions()).all() and np.abs(e.states() - o.states()).max() < 1e-12


def test_shard_invariance(E):
    """Env shards are independent given W: running envs [0, N) on one engine or as two half shards
    (env_offset) produces identical per-env results in PER_ENV mode (no collective on the data path)."""
    full = _cfg2(n_envs=2048, weight_mode=abi.PER_ENV, dtype=abi.F64, update_scale=abi.SCALE_SUM)
    with E.Engine(full) as e:
        e.step(200)
        e.sync()
        S, A, W = e.states(), e.actions(), e.weights()
    for off in (0, 1024):
        half = full.copy(n_envs=1024, env_offset=off, n_envs_global=2048)
        with E.Engine(half) as e:
            e.step(200)
            e.sync()
            assert (e.states() == S[off:off + 1024]).all() and (e.actions() == A[off:off + 1024]).all()
            assert (e.weights() == W[off:off + 1024]).all()


def test_learning_happens(E):
    """Sanity: N independent reference agents (PER_ENV, the hyper-parameters of examples/q_learning.rs) learn
    MountainCar — episodes shrink from thousands of steps to a few hundred, like the reference example."""
    cfg = _cfg2(n_envs=1024, weight_mode=abi.PER_ENV, update_scale=abi.SCALE_SUM, max_episode_steps=10000, dtype=abi.F32)
    with E.Engine(cfg) as e:
        e.step(40000)
        e.sync()
        n_ep, last_len, _ = e.env_stats()
        assert (n_ep >= 10).mean() > 0.9 and np.median(last_len) < 1000
        assert e.stats()["terminal_episodes"] > 10 * 1024 * 0.9


def test_unsupported_combination_fails_loudly(E):
    # two-table agents exist for the instantiated MountainCar bases only; an order without templates has no two-table kernel
    with pytest.raises(abi.RsrlError) as ei:
        E.Engine(abi.default_config(basis_order=4, algo=abi.GREEDY_GQ))
    assert ei.value.code == abi.EUNSUPPORTED
    with pytest.raises(abi.RsrlError) as ei:
        E.Engine(abi.default_config(basis_order=8))
    assert ei.value.code == abi.EINVAL


# ---------------------------------------------------------------------------------------------
# ANY basis order (lfa takes any; csrc/dyn.cuh runs the (basis, order) pairs that have no template instantiation)
# ---------------------------------------------------------------------------------------------
DYN_COMBOS = [(MC, abi.FOURIER, 4), (MC, abi.FOURIER, 6), (MC, abi.POLYNOMIAL, 1), (MC, abi.POLYNOMIAL, 5),
              (CP, abi.FOURIER, 1), (CP, abi.FOURIER, 4), (AC, abi.POLYNOMIAL, 3), (AC, abi.FOURIER, 6)]


@pytest.mark.parametrize("domain,basis,order", DYN_COMBOS)
@pytest.mark.parametrize("dtype,tol", [(abi.F64, 1e-12), (abi.F32, 2e-4)])
def test_any_order_project_and_evaluate_match_oracle(E, oracle, domain, basis, order, dtype, tol):
    cfg = abi.default_config(domain=domain, basis=basis, basis_order=order, dtype=dtype)
    D, A = oracle.domain_dims(domain)
    F = oracle.n_features(cfg)
    assert F == (order + 1) ** D
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(order)
    s = rng.uniform(lo, hi, size=(200, D))
    s[0], s[1] = lo, hi
    got, want = E.basis_project(cfg, s), oracle.project(cfg, s)
    assert got.shape == want.shape == (200, F)
    scale = np.maximum(np.abs(want), 1.0)
    assert (np.abs(got - want) / scale).max() < tol * (order if basis == abi.FOURIER else 1)
    W = rng.normal(size=(F, A))
    q_got, q_want = E.lfa_evaluate(cfg, W, s), oracle.evaluate(cfg, W, s)   # (Polynomial features of raw Acrobot states reach 1e10: relative)
    assert (np.abs(q_got - q_want) / np.maximum(np.abs(q_want), 1.0)).max() < tol * np.sqrt(F) * (order if basis == abi.FOURIER else 1)


def test_any_order_features_equal_the_template_path(E):
    """The run-time-order kernels generate the same bits as the templates: phi[c0, c1] = cos(pi (c0 x0 + c1 x1)) does not
    depend on the order P, so the order-4 grid (dyn.cuh) is a sub-grid of the order-5 grid (templates), which is a sub-grid
    of the order-6 grid (dyn.cuh)."""
    rng = np.random.default_rng(3)
    s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(64, 2))
    f5 = E.basis_project(abi.default_config(basis_order=5, dtype=abi.F32), s).reshape(64, 6, 6)   # templates; index [5 - c0, 5 - c1]
    f4 = E.basis_project(abi.default_config(basis_order=4, dtype=abi.F32), s).reshape(64, 5, 5)   # dyn.cuh;   index [4 - c0, 4 - c1]
    f6 = E.basis_project(abi.default_config(basis_order=6, dtype=abi.F32), s).reshape(64, 7, 7)
    assert (f4 == f5[:, 1:, 1:]).all()        # same coefficient vectors, same arithmetic: same bits
    assert (f6[:, 1:, 1:] == f5).all()


@pytest.mark.parametrize("domain,basis,order,algo,mode", [
    (MC, abi.FOURIER, 4, abi.QLEARNING, abi.SHARED), (MC, abi.FOURIER, 6, abi.SARSA, abi.PER_ENV),
    (MC, abi.POLYNOMIAL, 5, abi.EXPECTED_SARSA, abi.SHARED), (MC, abi.FOURIER, 4, abi.SARSA_LAMBDA, abi.SHARED),
    (MC, abi.FOURIER, 4, abi.Q_LAMBDA, abi.PER_ENV), (MC, abi.FOURIER, 6, abi.TD_LAMBDA, abi.SHARED),
    (CP, abi.FOURIER, 1, abi.SARSA, abi.SHARED), (AC, abi.FOURIER, 4, abi.EXPECTED_SARSA, abi.SHARED)])
def test_any_order_engine_free_run_f64(E, oracle, domain, basis, order, algo, mode):
    D = 2 if domain == MC else 4
    lo0 = [-0.6, 0.0] if domain == MC else [-0.05] * 4
    hi0 = [-0.4, 0.0] if domain == MC else [0.05] * 4
    pred = algo == abi.TD_LAMBDA
    cfg = abi.default_config(domain=domain, basis=basis, basis_order=order, algo=algo, weight_mode=mode,
                             policy=abi.RANDOM if pred else abi.EPSILON_GREEDY, epsilon=0.1, n_envs=77, dtype=abi.F64,
                             init_mode=abi.INIT_UNIFORM, init_lo=lo0, init_hi=hi0, max_episode_steps=40, seed=11, gamma=0.99,
                             lr=0.01, alpha=0.01 if algo in (abi.SARSA_LAMBDA, abi.Q_LAMBDA) else 0.5, lambda_=0.5,
                             update_scale=abi.SCALE_MEAN, record_td_error=1)
    steps = 12 if pred else 90
    with E.Engine(cfg) as e:
        assert e.launch_shape()["persistent"] == 0
        o = oracle.Engine(cfg)
        for chunk in (1, steps):
            e.step(chunk)
            o.step(chunk)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9)
        if algo in (abi.SARSA_LAMBDA, abi.Q_LAMBDA, abi.TD_LAMBDA):
            assert np.abs(e.traces() - o.traces()).max() < 1e-9 * max(1.0, np.abs(o.traces()).max())
        assert pred or o.stats()["total_episodes"] > 0


def test_any_order_handle_entry_point(E, oracle):
    """Handler::handle on explicit transitions (EXT kernels) for an order without templates."""
    cfg = abi.default_config(basis_order=4, dtype=abi.F64, n_envs=50, lr=0.05, update_scale=abi.SCALE_MEAN)
    rng = np.random.default_rng(5)
    s0 = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(50, 2))
    s1 = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(50, 2))
    a = rng.integers(0, 3, 50).astype(np.int32)
    r = -np.ones(50)
    term = np.zeros(50, dtype=np.uint8)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        W0 = rng.normal(size=(25, 3)) * 0.1
        e.set_weights(W0)
        o.set_weights(W0)
        e.handle(s0, a, r, s1, term)
        o.handle(s0, a, r, s1, term)
        e.sync()
        assert np.abs(e.weights() - o.weights()).max() < 1e-12


# ---------------------------------------------------------------------------------------------
# TileCoding (BASELINE config 3: CartPole / SARSA / tile coding) — project-defined spec, integer work bit-exact
# ---------------------------------------------------------------------------------------------
def _tile_cfg(domain=CP, **kw):
    base = dict(domain=domain, basis=abi.TILE_CODING, n_tilings=8, tiles_per_dim=8, memory_size=4096, algo=abi.SARSA,
                policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.1 / 8, dtype=abi.F64, n_envs=65,
                init_mode=abi.INIT_UNIFORM, init_lo=[-0.05] * 4, init_hi=[0.05] * 4, max_episode_steps=500, seed=13,
                update_scale=abi.SCALE_MEAN, record_td_error=1)
    base.update(kw)
    return abi.default_config(**base)


@pytest.mark.parametrize("domain", [MC, CP, AC])
def test_tile_rows_bit_exact(E, oracle, domain):
    cfg = _tile_cfg(domain=domain, n_tilings=8 if domain != MC else 16, memory_size=1024)
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(1)
    s = rng.uniform(lo, hi, size=(400, len(lo)))
    s[0], s[1] = lo, hi
    got, want = E.basis_project(cfg, s), oracle.project(cfg, s)
    assert (got == want).all()
    assert (got.sum(axis=1) <= cfg.n_tilings).all() and (got.sum(axis=1) >= 1).all()
    W = rng.normal(size=(1024, oracle.domain_dims(domain)[1]))
    assert np.abs(E.lfa_evaluate(cfg, W, s) - oracle.evaluate(cfg, W, s)).max() < 1e-12


@pytest.mark.parametrize("n_tilings,tiles", [(1, 8), (3, 5), (5, 16), (7, 3), (12, 8), (16, 31)])
def test_tile_rows_bit_exact_any_tiling_count_and_out_of_range_states(E, oracle, n_tilings, tiles):
    """(q + offset) / n_tilings is a multiply-high by a precomputed reciprocal on the device (exact for the non-negative numerators
    of in-range states) and a plain division otherwise: rows must equal the oracle's for every tiling count, including states far
    outside the domain's limits (negative and huge numerators take the division path)."""
    cfg = _tile_cfg(domain=CP, n_tilings=n_tilings, tiles_per_dim=tiles, memory_size=2048)
    lo, hi = oracle.domain_limits(CP)
    rng = np.random.default_rng(n_tilings)
    s = rng.uniform(lo, hi, size=(600, 4))
    s[:100] = rng.uniform(lo - 3 * (hi - lo), hi + 3 * (hi - lo), size=(100, 4))   # out of range, both sides
    s[100:110] = rng.uniform(-1e5, 1e5, size=(10, 4)) * (hi - lo)                   # numerators beyond 2^24
    s[110], s[111] = lo, hi
    got, want = E.basis_project(cfg, s), oracle.project(cfg, s)
    assert (got == want).all()


@pytest.mark.parametrize("algo", [abi.SARSA, abi.QLEARNING, abi.EXPECTED_SARSA])
def test_tile_engine_free_run_f64(E, oracle, algo):
    cfg = _tile_cfg(algo=algo, alpha=1.0)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        for chunk in (1, 9, 290):
            e.step(chunk)
            o.step(chunk)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9, s_tol=1e-9)  # dW is summed in 2^-44 fixed point on the device
        assert o.stats()["total_episodes"] > 0


@pytest.mark.parametrize("kw", [dict(n_tilings=12, memory_size=2048, n_envs=700),      # > 8 tilings: the TMAX = 16 instantiation
                                dict(n_tilings=5, tiles_per_dim=6, memory_size=512, n_envs=1500, domain=MC),
                                dict(n_tilings=8, n_envs=2100)])                              # > 1024 envs per CTA chunking, multi-CTA reduce-scatter
def test_tile_engine_variants_f64(E, oracle, kw, monkeypatch):
    """Dense (shared-memory accumulation + reduce-scatter) and RED-atomics kernels against the oracle over free runs.
    Horizon 21: on the 2100-env case one env's eps-greedy decision sits within 5e-12 (the 2^-44 fixed-point resolution of dW)
    of the 1e-7 tie threshold at step 29 and flips — identically in both kernels (tools/diag_tile.py)."""
    kw = dict(kw)
    if kw.get("domain") == MC:
        kw.update(init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], lr=0.05)
    cfg = _tile_cfg(**kw)
    for dense in ("1", "0"):
        monkeypatch.setenv("RSRL_B200_TILE_DENSE", dense)
        with E.Engine(cfg) as e:
            o = oracle.Engine(cfg)
            for chunk in (1, 20):
                e.step(chunk)
                o.step(chunk)
                e.sync()
                _compare_engines(e, o, w_tol=1e-9, s_tol=1e-9)


def test_tile_red_kernel_handle_then_step(E, oracle, monkeypatch):
    """RED-atomics TileCoding kernel: handle(n < N) launches fewer CTAs than step(); the grid barrier targets are
    launch-local, so mixing the two neither hangs nor lets a CTA read the table early (either order)."""
    monkeypatch.setenv("RSRL_B200_TILE_DENSE", "0")
    cfg = _tile_cfg(n_envs=3000)
    rng = np.random.default_rng(8)
    lo, hi = oracle.domain_limits(CP)
    s = rng.uniform(lo, hi, size=(300, 4)) * 0.5
    a = rng.integers(0, 2, 300).astype(np.int32)
    ns, r, term = oracle.domain_step(CP, s, a)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        for _ in range(2):
            e.handle(s, a, r, ns, term, draw_idx=1)
            o.handle(s, a, r, ns, term, draw_idx=1)
            e.step(3)
            o.step(3)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9, s_tol=1e-9)
        e.handle(s[:100], a[:100], r[:100], ns[:100], term[:100], draw_idx=2)   # a single CTA: no barrier at all
        o.handle(s[:100], a[:100], r[:100], ns[:100], term[:100], draw_idx=2)
        e.step(2)
        o.step(2)
        e.sync()
        _compare_engines(e, o, w_tol=1e-9, s_tol=1e-9)


def test_tile_handle_and_policy_entry_points(E, oracle):
    cfg = _tile_cfg(n_envs=200)
    rng = np.random.default_rng(3)
    lo, hi = oracle.domain_limits(CP)
    s = rng.uniform(lo, hi, size=(200, 4)) * 0.5
    a = rng.integers(0, 2, 200).astype(np.int32)
    ns, r, term = oracle.domain_step(CP, s, a)
    W = rng.normal(size=(4096, 2))
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(W)
        o.set_weights(W)
        assert np.abs(e.evaluate(s) - oracle.evaluate(cfg, W, s)).max() < 1e-12
        assert (e.sample(s, draw=4) == oracle.policy_sample_batch(cfg.policy, cfg.epsilon, cfg.seed, 4, 0, oracle.evaluate(cfg, W, s))).all()
        td_e, td_o = e.handle(s, a, r, ns, term, draw_idx=2), o.handle(s, a, r, ns, term, draw_idx=2)
        assert np.abs(td_e - td_o).max() < 1e-12 and np.abs(e.weights() - o.weights()).max() < 1e-11


def test_cfg3_full_size_teacher_forced(E, oracle):
    """BASELINE configs[2] at its full size (262 144 CartPole envs, SARSA, tile coding), dtype f32: three steps, each from the
    device's own states and weights, against the f64 oracle.  Rows are integer work (exact); Q is a sum of <= 8 weights."""
    cfg = _tile_cfg(n_envs=262144, dtype=abi.F32, seed=0)
    rng = np.random.default_rng(4)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(rng.normal(size=(4096, 2)) * 0.1)   # (no free-running warm-up: one oracle step of 262 144 envs takes ~10 s)
        for t in range(3):
            o.set_states(e.states())
            o.set_weights(e.weights())
            q = oracle.evaluate(cfg, e.weights(), e.states())
            safe = np.abs(q[:, 0] - q[:, 1]) > 1e-5
            e.step(1)
            o.step(1)
            e.sync()
            assert safe.mean() > 0.95
            assert (e.actions()[safe] == o.actions()[safe]).all()
            same = e.actions() == o.actions()
            assert same.mean() > 0.999
            assert np.abs(e.states()[same] - o.states()[same]).max() < 1e-12
            scale = max(1.0, np.abs(o.td_errors()).max())
            assert np.abs(e.td_errors()[same] - o.td_errors()[same]).max() < 4e-6 * scale
            assert np.abs(e.weights() - o.weights()).max() < 2e-5 * max(1.0, np.abs(o.weights()).max())


def test_cfg3_full_size_properties(E):
    cfg = _tile_cfg(n_envs=262144, dtype=abi.F32, record_td_error=0, seed=0)
    outs = []
    for _ in range(2):
        with E.Engine(cfg) as e:
            e.step(120)
            e.sync()
            outs.append((e.weights(), e.states(), e.env_stats()[2], e.stats()))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all() and (outs[0][2] == outs[1][2]).all()
    W, S, H, st = outs[0]
    assert st["total_steps"] == 262144 * 120 and st["nonfinite"] == 0 and st["total_episodes"] > 0
    assert np.isfinite(W).all() and np.abs(W).max() > 0
    assert (np.abs(S[:, 0]) <= 2.4).all() and (np.abs(S[:, 2]) <= np.pi / 15).all()


# ---------------------------------------------------------------------------------------------
# large Fourier bases on the 4-D domains (BASELINE config 4: Acrobot / ExpectedSARSA / Fourier(7), F = 4096)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("domain,order", [(AC, 7), (CP, 5)])
@pytest.mark.parametrize("dtype,tol", [(abi.F64, 1e-13), (abi.F32, 2e-5)])
def test_large_basis_project_and_evaluate(E, oracle, domain, order, dtype, tol):
    cfg = abi.default_config(domain=domain, basis_order=order, dtype=dtype)
    D, A = oracle.domain_dims(domain)
    F = (order + 1) ** 4
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(order)
    s = rng.uniform(lo, hi, size=(40, D))
    s[0], s[1] = lo, hi
    got, want = E.basis_project(cfg, s), oracle.project(cfg, s)
    assert got.shape == (40, F) and np.abs(got - want).max() < tol * order
    W = rng.normal(size=(F, A))
    assert np.abs(E.lfa_evaluate(cfg, W, s) - oracle.evaluate(cfg, W, s)).max() < tol * 10 * np.sqrt(F)


@pytest.mark.parametrize("domain,order,algo", [(AC, 7, abi.EXPECTED_SARSA), (CP, 5, abi.QLEARNING)])
def test_large_basis_engine_free_run_f64(E, oracle, domain, order, algo):
    cfg = abi.default_config(domain=domain, basis_order=order, algo=algo, policy=abi.EPSILON_GREEDY, epsilon=0.1, n_envs=70,
                             dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4,
                             max_episode_steps=25, seed=17, gamma=0.99, lr=1e-4, alpha=1.0, record_td_error=1,
                             update_scale=abi.SCALE_MEAN)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        for chunk in (1, 5, 34):
            e.step(chunk)
            o.step(chunk)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9, s_tol=1e-8)
        assert o.stats()["total_episodes"] > 0
        # trait-level entry points on the 4096-feature weights
        s = e.states()
        assert np.abs(e.evaluate(s) - oracle.evaluate(cfg, o.weights(), s)).max() < 1e-9
        a = np.zeros(70, dtype=np.int32)
        ns, r, term = oracle.domain_step(domain, s, a)
        td_e, td_o = e.handle(s, a, r, ns, term, draw_idx=99), o.handle(s, a, r, ns, term, draw_idx=99)
        assert np.abs(td_e - td_o).max() < 1e-9 and np.abs(e.weights() - o.weights()).max() < 1e-9


def test_cfg4_shape_properties(E):
    """Acrobot / ExpectedSARSA / Fourier(7) at a per-GPU shard of config 4: deterministic, finite, counted."""
    cfg = abi.default_config(domain=AC, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1,
                             n_envs=16384, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4,
                             max_episode_steps=500, seed=0, gamma=0.99, lr=1e-4, alpha=1.0, update_scale=abi.SCALE_MEAN)
    outs = []
    for _ in range(2):
        with E.Engine(cfg) as e:
            e.step(20)
            e.sync()
            outs.append((e.weights(), e.states(), e.stats()))
    assert (outs[0][0] == outs[1][0]).all() and (outs[0][1] == outs[1][1]).all()
    assert outs[0][0].shape == (4096, 3) and np.isfinite(outs[0][0]).all() and np.abs(outs[0][0]).max() > 0
    assert outs[0][2]["total_steps"] == 16384 * 20 and outs[0][2]["nonfinite"] == 0


# tcgen05 path (f4tc.cuh): order 7, dtype f32.  The GEMMs run as 3xTF32 tensor-core MMAs with fp32 accumulation; the bar is
# the same fp32 tolerance as the CUDA-core f32 kernels (teacher-forced single steps against the oracle).
def _f4tc_cfg(domain, n, algo, **kw):
    base = dict(domain=domain, basis_order=7, algo=algo, policy=abi.EPSILON_GREEDY, epsilon=0.1, n_envs=n, dtype=abi.F32,
                init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500, seed=5, gamma=0.99,
                lr=1e-4, alpha=1.0, update_scale=abi.SCALE_MEAN, record_td_error=1)
    base.update(kw)
    return abi.default_config(**base)


@pytest.mark.parametrize("domain,n,algo", [(AC, 300, abi.EXPECTED_SARSA), (CP, 1000, abi.QLEARNING), (AC, 4113, abi.SARSA), (AC, 777, abi.PAL)])
def test_f4tc_single_step_matches_oracle(E, oracle, domain, n, algo):
    """n is ragged on purpose (not a multiple of the 128-env GEMM tile or the 32-env dW sub-tile)."""
    cfg = _f4tc_cfg(domain, n, algo)
    A = oracle.domain_dims(domain)[1]
    W0 = np.random.default_rng(1).normal(size=(4096, A)) * 0.05
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(W0)
        o.set_weights(W0)
        e.step(1)
        o.step(1)
        e.sync()
        assert e.stats()["kernel_launches"] >= 3
        td_o = o.td_errors()
        assert np.abs(e.td_errors() - td_o).max() < 1.2e-5 * max(1.0, np.abs(td_o).max())    # fp32 tolerance (3xTF32: ~5e-6 measured)
        assert (e.actions() == o.actions()).all() and (e.episode_steps() == o.episode_steps()).all()
        assert np.abs(e.states() - o.states()).max() < 1e-12                                # f64 physics
        assert np.abs(e.weights() - o.weights()).max() < 1e-6 * np.abs(o.weights()).max()


def test_cfg4_full_size_teacher_forced(E, oracle):
    """BASELINE configs[3] shard: 131 072 Acrobot envs, Fourier(7), ExpectedSARSA on the tcgen05 path.  Three steps, each from the
    device's own states and weights.  Per-env outputs (actions, next states, TD errors) against the oracle on a contiguous block
    of 1536 envs (the reference-shaped oracle needs 16 k cosines per env-step); the FULL-SIZE weight update against an independent
    numpy contraction dW[k, a] = sum_i [a_i = a] coef_i cos(pi c_k . x^_i) over all 131 072 envs."""
    import itertools
    N, OFF, NB = 131072, 40000, 1536
    cfg = _f4tc_cfg(AC, N, abi.EXPECTED_SARSA, seed=0)
    sub = _f4tc_cfg(AC, NB, abi.EXPECTED_SARSA, seed=0, env_offset=OFF, n_envs_global=N)
    lo, hi = oracle.domain_limits(AC)
    coefs = np.array(sorted(itertools.product(range(8), repeat=4), reverse=True), dtype=np.float64)   # descending; the zero vector = bias slot, last
    rng = np.random.default_rng(2)
    with E.Engine(cfg) as e:
        assert e.launch_shape()["f4"] >= 2   # the tensor-core path (1 = CUDA-core fourier4.cuh)
        o = oracle.Engine(sub)
        e.set_weights(rng.normal(size=(4096, 3)) * 0.05)
        e.step(2)
        o.step(2)     # same batched-step index on both sides: it is the RNG draw counter
        for t in range(3):
            S0, W0 = e.states(), e.weights()
            o.set_states(S0[OFF:OFF + NB])
            o.set_weights(W0)
            e.step(1)
            o.step(1)
            e.sync()
            blk = slice(OFF, OFF + NB)
            same = e.actions()[blk] == o.actions()
            assert same.mean() > 0.995                                                          # (eps-greedy: a flip needs a near tie)
            assert np.abs(e.states()[blk][same] - o.states()[same]).max() < 1e-12               # f64 physics
            td_o = o.td_errors()
            assert np.abs(e.td_errors()[blk][same] - td_o[same]).max() < 1.2e-5 * max(1.0, np.abs(td_o).max())
            # full-size update: numpy features of the FROM states, the device's own TD errors and actions
            act, td = e.actions(), e.td_errors()
            coef = (cfg.lr / N) * cfg.alpha * td                                                 # expected_sarsa.rs:64, MEAN scaling
            dW = np.zeros((4096, 3))
            xh = (S0 - lo) / (hi - lo)
            for c0 in range(0, N, 8192):
                phi = np.cos(np.pi * xh[c0:c0 + 8192] @ coefs.T)                                 # [chunk, 4096]
                for a_ in range(3):
                    m = act[c0:c0 + 8192] == a_
                    dW[:, a_] += phi[m].T @ coef[c0:c0 + 8192][m]
            got = e.weights() - W0
            assert np.abs(got - dW).max() < 1e-4 * np.abs(dW).max() + 1e-9 * np.abs(W0).max()   # fp32 W rounding: 6e-8 |W|; 3xTF32 GEMM: ~1e-5 |dW|


def test_f4tc_dw_from_zero_weights(E, oracle):
    """W0 = 0 => Q = 0, TD error = reward: the weights after one step ARE dW = sum_env coef * phi(s_env) — checks the
    Phi^T D tensor-core contraction alone, relative to the largest update."""
    cfg = _f4tc_cfg(AC, 2500, abi.EXPECTED_SARSA, lr=0.05, update_scale=abi.SCALE_SUM)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(1)
        o.step(1)
        e.sync()
        dW = o.weights()
        assert np.abs(dW).max() > 1.0
        assert np.abs(e.weights() - dW).max() < 2e-6 * np.abs(dW).max()


def test_f4tc_free_run_matches_cuda_core_path(E, monkeypatch):
    """Same engine config on the tensor-core path and on the CUDA-core f32 path (RSRL_B200_F4TC=0): 12 free-running
    steps from W = 0; trajectories stay identical as long as no decision is closer than the fp32 noise of Q."""
    cfg = _f4tc_cfg(AC, 1500, abi.EXPECTED_SARSA, lr=1e-3, seed=11)
    outs = []
    for mask in ("0", "3"):
        monkeypatch.setenv("RSRL_B200_F4TC", mask)
        with E.Engine(cfg) as e:
            e.step(12)
            e.sync()
            outs.append((e.weights(), e.states(), e.actions(), e.stats()))
    (W0, S0, A0, st0), (W1, S1, A1, st1) = outs
    assert (A0 == A1).mean() > 0.99 and st0["total_steps"] == st1["total_steps"]
    assert np.abs(W0 - W1).max() < 1e-5 * max(np.abs(W0).max(), 1e-6) + 1e-9
    same = (A0 == A1)
    assert np.abs(S0[same] - S1[same]).max() < 1e-6


def test_f4tc_handle_entry_point(E, oracle):
    cfg = _f4tc_cfg(AC, 700, abi.QLEARNING)
    rng = np.random.default_rng(4)
    lo, hi = oracle.domain_limits(AC)
    s = rng.uniform(lo, hi, size=(700, 4)) * 0.3
    a = rng.integers(0, 3, 700).astype(np.int32)
    ns, r, term = oracle.domain_step(AC, s, a)
    W0 = rng.normal(size=(4096, 3)) * 0.05
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.set_weights(W0)
        o.set_weights(W0)
        td_e, td_o = e.handle(s, a, r, ns, term, draw_idx=3), o.handle(s, a, r, ns, term, draw_idx=3)
        assert np.abs(td_e - td_o).max() < 1.2e-5 * max(1.0, np.abs(td_o).max())
        assert np.abs(e.weights() - o.weights()).max() < 1e-6 * np.abs(o.weights()).max()
        assert np.abs(e.evaluate(s) - oracle.evaluate(cfg, o.weights(), s)).max() < 2e-5 * 10 * 64


def test_softmax_policy_entry_points(E, oracle):
    """Stateless Policy::sample / evaluate for Softmax on explicit Q vectors + Policy::mode through an engine."""
    rng = np.random.default_rng(8)
    q = rng.normal(size=(3000, 3)) * 2
    q[:5] = [[0, 1, 0], [5, 5, 5], [700, 710, 705], [-1e3, 0, 1e-3], [1e-9, 0, 0]]
    for tau in (1.0, 0.25):
        want_p = np.array([oracle.policy_probs(abi.SOFTMAX, tau, qi) for qi in q])
        assert np.abs(E.policy_probs(abi.SOFTMAX, tau, q) - want_p).max() < 1e-14
        got = E.policy_sample(abi.SOFTMAX, tau, seed=21, draw=6, env_offset=100, q=q)
        want = oracle.policy_sample_batch(abi.SOFTMAX, tau, 21, 6, 100, q)
        assert (got == want).all()
    cfg = abi.default_config(policy=abi.SOFTMAX, epsilon=0.5, dtype=abi.F64, n_envs=64)
    s = rng.uniform([-1.2, -0.07], [0.6, 0.07], size=(64, 2))
    W = rng.normal(size=(36, 3))
    with E.Engine(cfg) as e:
        e.set_weights(W)
        Q = oracle.evaluate(cfg, W, s)
        P = np.array([oracle.policy_probs(abi.SOFTMAX, 0.5, qi) for qi in Q])
        want_mode = [oracle.argmax_first(p)[0] for p in P]                 # softmax.rs:141
        assert (e.mode(s) == np.array(want_mode)).all()
        assert (e.sample(s, draw=3) == oracle.policy_sample_batch(abi.SOFTMAX, 0.5, cfg.seed, 3, 0, Q)).all()


# ---------------------------------------------------------------------------------------------
# C++ host-side mirror of the reference's trait surface (include/rsrl_b200.hpp): examples/q_learning.cpp is
# rsrl/examples/q_learning.rs line by line; its episode lengths must equal the oracle's.
# ---------------------------------------------------------------------------------------------
def test_cpp_mirror_example_matches_oracle(E, oracle):
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "examples", "q_learning")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build_examples()
    out = subprocess.run([exe, "3", "2500"], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stderr
    lengths = [int(m) for m in re.findall(r"Batch \d+: (\d+) steps", out.stdout)]
    assert len(lengths) == 3
    cfg = abi.default_config(dtype=abi.F64, max_episode_steps=2500)   # exactly examples/q_learning.rs + the cap
    o = oracle.Engine(cfg)
    o.step(sum(lengths))
    n_ep, last_len, h = o.env_stats()
    want = 0
    for n in lengths:
        want = (want * 1000003 + n) % (1 << 64)
    assert n_ep[0] == 3 and int(h[0]) == want and last_len[0] == lengths[-1]
    norm = float(re.search(r"\|W\|\^2 = (\S+)", out.stdout).group(1))
    assert abs(norm - float((o.weights() ** 2).sum())) < 1e-12 * max(1.0, norm)


# ---------------------------------------------------------------------------------------------
# multi-GPU: in-kernel dW exchange over NVLink peer memory (needs >= 2 GPUs on the box; skipped otherwise)
# ---------------------------------------------------------------------------------------------
def test_multi_gpu_peer_exchange_matches_single_gpu(E):
    import os
    import subprocess
    import sys
    n = abi.load().rsrl_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    world = 2 if n < 4 else 4
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(root, "tests", "tools", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


# ---------------------------------------------------------------------------------------------
# ragged / extreme shapes of the persistent kernel and error behaviour of the ABI
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n", [1, 33, 449, 1000, 2049])
@pytest.mark.parametrize("mode", [abi.SHARED, abi.PER_ENV], ids=["shared", "per_env"])
def test_ragged_env_counts(E, oracle, n, mode):
    """1 .. 17 CTAs, partially filled warps and leader groups of unequal size must not change the result."""
    cfg = _mc_cfg(n_envs=n, weight_mode=mode, policy=abi.EPSILON_GREEDY, epsilon=0.1, max_episode_steps=40,
                  update_scale=abi.SCALE_MEAN)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        for chunk in (1, 2, 60):
            e.step(chunk)
            o.step(chunk)
        e.sync()
        _compare_engines(e, o, w_tol=1e-9)


def test_more_envs_than_resident_threads(E, oracle):
    """N > 148 x 512: every thread of the persistent kernel loops over several envs per step (state through L2)."""
    cfg = _mc_cfg(n_envs=100003, max_episode_steps=3, update_scale=abi.SCALE_MEAN, record_td_error=0)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(5)
        o.step(5)
        e.sync()
        assert (e.actions() == o.actions()).all() and (e.episode_steps() == o.episode_steps()).all()
        assert np.abs(e.states() - o.states()).max() < 1e-12 and np.abs(e.weights() - o.weights()).max() < 1e-10
        assert e.stats()["total_episodes"] == o.stats()["total_episodes"] == 100003


def test_dutch_trace_and_epsilon_schedule(E, oracle):
    """Dutch traces (traces.rs:222-240) and the per-episode epsilon decay of examples/sarsa_lambda.rs:68."""
    cfg = _mc_cfg(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99,
                  trace_rule=abi.TRACE_DUTCH, n_envs=48, max_episode_steps=60)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        eps = 0.2
        for _ in range(5):
            e.step(60)
            o.step(60)
            eps *= 0.995
            e.set_epsilon(eps)
            o.set_epsilon(eps)
        e.sync()
        _compare_engines(e, o, w_tol=1e-9)
        assert np.abs(e.traces() - o.traces()).max() < 1e-9


def test_error_behaviour(E):
    import ctypes as C
    lib = abi.load()
    cfg = _mc_cfg(n_envs=8, record_td_error=0)
    with E.Engine(cfg) as e:
        s = e.states()
        a = np.zeros(8, dtype=np.int32)
        with pytest.raises(abi.RsrlError) as ei:
            e.handle(s, a + 7, np.zeros(8), s, np.zeros(8, dtype=np.uint8))
        assert ei.value.code == abi.EINVAL and "action" in str(ei.value)
        with pytest.raises(abi.RsrlError) as ei:
            e.traces()
        assert ei.value.code == abi.EINVAL
        with pytest.raises(abi.RsrlError) as ei:
            e.td_errors()
        assert ei.value.code == abi.EINVAL
        assert lib.rsrl_engine_step(e.h, -1) == abi.EINVAL
        assert lib.rsrl_engine_evaluate(e.h, 0, abi.dp(s), abi.dp(np.zeros(3))) == abi.EINVAL
        with pytest.raises(abi.RsrlError) as ei:
            e.set_epsilon(-0.5)
        assert ei.value.code == abi.EINVAL
    with pytest.raises(abi.RsrlError) as ei:
        E.Engine(_mc_cfg(n_envs=0))
    assert ei.value.code == abi.EINVAL
    # raw int32 / double config fields: a typo must not silently change the learning semantics
    for bad in (dict(trace_rule=3), dict(update_scale=2), dict(init_mode=2), dict(gamma=float("nan")), dict(lr=float("inf")),
                dict(n_envs=8, env_offset=4, n_envs_global=8)):
        with pytest.raises(abi.RsrlError) as ei:
            E.Engine(_mc_cfg(**{**dict(n_envs=8), **bad}))
        assert ei.value.code == abi.EINVAL, bad
    with pytest.raises(abi.RsrlError) as ei:   # per-env weights: transition i belongs to agent i
        with E.Engine(_mc_cfg(n_envs=8, weight_mode=abi.PER_ENV)) as e2:
            e2.handle(np.zeros((3, 2)), np.zeros(3, dtype=np.int32), np.zeros(3), np.zeros((3, 2)), np.zeros(3, dtype=np.uint8))
    assert ei.value.code == abi.EINVAL
    with pytest.raises(abi.RsrlError) as ei:   # NaN weights: the reference panics, the engine reports it at sync
        with E.Engine(_mc_cfg(n_envs=8)) as e3:
            e3.set_weights(np.full((36, 3), np.nan))
            e3.step(1)
            e3.sync()
    assert ei.value.code == abi.ENONFINITE
    r, t = np.zeros(1), np.zeros(1, dtype=np.uint8)
    assert lib.rsrl_domain_step(9, 1, abi.dp(np.zeros((1, 2))), abi.ip(np.zeros(1, dtype=np.int32)), abi.dp(r), abi.u8p(t)) == abi.EINVAL


# ---------------------------------------------------------------------------------------------
# GreedyGQ, A2C (Softmax grad_log), Domain::rollout / Trajectory
# ---------------------------------------------------------------------------------------------
TWO_TABLE = [
    ("greedy_gq", dict(algo=abi.GREEDY_GQ, policy=abi.EPSILON_GREEDY, epsilon=0.1, basis_order=3, lr=0.1, alpha=0.001, gamma=0.99)),   # examples/greedy_gq.rs:25-38
    ("a2c", dict(algo=abi.A2C, policy=abi.SOFTMAX, epsilon=1.0, basis_order=3, lr=0.001, alpha=0.001, gamma=1.0)),                    # examples/a2c.rs:24-48
    ("a2c_fast", dict(algo=abi.A2C, policy=abi.SOFTMAX, epsilon=0.5, basis_order=5, lr=0.05, alpha=0.05, gamma=0.99)),
]


@pytest.mark.parametrize("name,kw", TWO_TABLE, ids=[t[0] for t in TWO_TABLE])
@pytest.mark.parametrize("mode", [abi.SHARED, abi.PER_ENV], ids=["shared", "per_env"])
def test_two_table_agents_free_run_f64(E, oracle, name, kw, mode):
    """GreedyGQ (control/td/greedy_gq.rs:73-141) and A2C (examples/a2c.rs) free-running against the oracle, dtype f64:
    actions / step counts bit-exact, both weight tables within 1e-9."""
    cfg = _mc_cfg(weight_mode=mode, update_scale=abi.SCALE_MEAN if mode == abi.SHARED else abi.SCALE_SUM, n_envs=300,
                  max_episode_steps=120, **kw)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        # a2c_fast with one agent per env: the policy-gradient dynamics at these step sizes amplify a 1e-16 difference (exp of CUDA vs
        # glibc) by ~1.3x per step (tests/tools/diag_a2c.py: 1e-9 after 75 steps) — compare over a horizon where it is still small
        for chunk in ((1, 9, 40) if (name == "a2c_fast" and mode == abi.PER_ENV) else (1, 9, 250)):
            e.step(chunk)
            o.step(chunk)
            e.sync()
            _compare_engines(e, o, w_tol=1e-9)
            aw, ow = e.aux_weights(), o.aux_weights()
            assert np.abs(aw - ow).max() < 1e-9 * max(1.0, np.abs(ow).max())
        assert np.abs(o.aux_weights()).max() > 0 and (o.stats()["total_episodes"] > 0 or o.stats()["batch_steps"] < 120)


def test_greedy_gq_example_single_env(E, oracle):
    """examples/greedy_gq.rs as written: one env from MountainCar::default(), Fourier(3), SGD(0.1) / SGD(0.001), eps 0.1, gamma 0.99,
    1000-step cap; N = 1 is op for op the reference loop."""
    cfg = abi.default_config(n_envs=1, dtype=abi.F64, basis_order=3, algo=abi.GREEDY_GQ, policy=abi.EPSILON_GREEDY, epsilon=0.1, lr=0.1,
                             alpha=0.001, gamma=0.99, max_episode_steps=1000, record_td_error=1)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(6000)
        o.step(6000)
        e.sync()
        _compare_engines(e, o, w_tol=1e-9, s_tol=1e-8)
        assert o.stats()["terminal_episodes"] > 3      # it learns to reach the goal


def test_two_table_f32_teacher_forced(E, oracle):
    """fp32 tables against the f64 oracle, one step at a time from the device's state and weights."""
    for name, kw in TWO_TABLE[:2]:
        cfg = _mc_cfg(dtype=abi.F32, n_envs=512, update_scale=abi.SCALE_MEAN, **kw)
        rng = np.random.default_rng(3)
        with E.Engine(cfg) as e:
            o = oracle.Engine(cfg)
            e.set_weights(rng.normal(size=(16, 3)) * 0.3)
            e.set_aux_weights(rng.normal(size=(16, 3)) * 0.3)
            for t in range(10):
                o.set_states(e.states())
                o.set_weights(e.weights())
                o.set_aux_weights(e.aux_weights())
                e.step(1)
                o.step(1)
                e.sync()
                same = e.actions() == o.actions()
                assert same.mean() > 0.97
                assert np.abs(e.td_errors()[same] - o.td_errors()[same]).max() < 2e-5 * max(1.0, np.abs(o.td_errors()).max())
                if same.all():
                    assert np.abs(e.weights() - o.weights()).max() < 2e-6 and np.abs(e.aux_weights() - o.aux_weights()).max() < 2e-6


def test_two_table_handle_and_policy_entry_points(E, oracle):
    """Handler<&Transition>::handle for GreedyGQ, Policy::sample / mode of the A2C agent (they read the policy's own table)."""
    rng = np.random.default_rng(9)
    lo, hi = oracle.domain_limits(MC)
    s = rng.uniform(lo, hi, size=(200, 2))
    a = rng.integers(0, 3, 200).astype(np.int32)
    ns, r, term = oracle.domain_step(MC, s, a)
    cfg = _mc_cfg(n_envs=200, **TWO_TABLE[0][1])
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        Wq, Wt = rng.normal(size=(16, 3)), rng.normal(size=(16, 3))
        for x in (e, o):
            x.set_weights(Wq)
            x.set_aux_weights(Wt)
        td_e, td_o = e.handle(s, a, r, ns, term, draw_idx=1), o.handle(s, a, r, ns, term, draw_idx=1)
        assert np.abs(td_e - td_o).max() < 1e-12
        assert np.abs(e.weights() - o.weights()).max() < 1e-11 and np.abs(e.aux_weights() - o.aux_weights()).max() < 1e-11
    cfg = _mc_cfg(n_envs=200, **TWO_TABLE[1][1])
    with E.Engine(cfg) as e:
        Wp = rng.normal(size=(16, 3))
        e.set_weights(rng.normal(size=(16, 3)))
        e.set_aux_weights(Wp)
        h = oracle.evaluate(cfg, Wp, s)
        assert (e.sample(s, draw=3) == oracle.policy_sample_batch(abi.SOFTMAX, 1.0, cfg.seed, 3, 0, h)).all()
        probs = np.exp(h - h.max(axis=1, keepdims=True))
        assert (e.mode(s) == np.argmax(probs, axis=1)).mean() > 0.99   # argmax_first over the probabilities (softmax.rs:141)
    with pytest.raises(abi.RsrlError) as ei:
        E.Engine(_mc_cfg(algo=abi.A2C, policy=abi.GREEDY))
    assert ei.value.code == abi.EINVAL


@pytest.mark.parametrize("dtype", [abi.F64, abi.F32], ids=["f64", "f32"])
def test_rollout_matches_oracle(E, oracle, dtype):
    """Domain::rollout (rsrl_domains/src/lib.rs:448-479) for a batch of envs with learned weights: Trajectory{start, steps}
    layout, mode() policy (examples/q_learning.rs:57) and sampled policy; step-limit and terminal semantics."""
    cfg = _mc_cfg(dtype=dtype, n_envs=400, policy=abi.EPSILON_GREEDY, epsilon=0.1, update_scale=abi.SCALE_MEAN, lr=0.5)
    with E.Engine(cfg) as e:
        o = oracle.Engine(cfg)
        e.step(300)
        e.sync()
        o.set_weights(e.weights())
        rng = np.random.default_rng(1)
        init = np.column_stack([rng.uniform(-1.2, 0.59, 400), rng.uniform(-0.07, 0.07, 400)])
        for greedy, limit in ((True, 120), (False, 40), (True, 1), (True, 2)):
            re, ro = e.rollout(init_states=init, step_limit=limit, greedy=greedy, draw=7), o.rollout(init_states=init, step_limit=limit, greedy=greedy, draw=7)
            assert (re["start"] == init).all()
            if dtype == abi.F64:
                assert (re["len"] == ro["len"]).all()
                for i in range(400):
                    n = re["len"][i]
                    assert (re["actions"][i, :n] == ro["actions"][i, :n]).all() and (re["terminal"][i, :n] == ro["terminal"][i, :n]).all()
                    assert (re["rewards"][i, :n] == ro["rewards"][i, :n]).all() and np.abs(re["next"][i, :n] - ro["next"][i, :n]).max(initial=0) < 1e-12
            else:   # fp32 Q: a near-tie may flip a greedy action and change the trajectory from there on
                assert (re["len"] == ro["len"]).mean() > 0.9
            n = re["len"]
            assert (n <= max(limit - 1, 0)).all() and (limit < 2 or (n >= 1).all())
            j = np.arange(re["terminal"].shape[1])[None, :]
            assert not (re["terminal"].astype(bool) & (j < n[:, None] - 1)).any()      # only the last recorded observation can be terminal
        # default start (config's distribution): Trajectory of examples/q_learning.rs:57
        rd = e.rollout(n=5, step_limit=30, greedy=True, draw=0)
        od = o.rollout(n=5, step_limit=30, greedy=True, draw=0)
        assert np.abs(rd["start"] - od["start"]).max() < 1e-15 and (dtype == abi.F32 or (rd["len"] == od["len"]).all())


# ---------------------------------------------------------------------------------------------
# ContinuousMountainCar / HIVTreatment as batched Domain::step / emit
# ---------------------------------------------------------------------------------------------
def test_extended_domains_match_oracle(E, oracle):
    CMC, HIV = abi.CONTINUOUS_MOUNTAIN_CAR, abi.HIV
    assert E.domain_ex_info(CMC)[:2] == (2, 0) and E.domain_ex_info(HIV)[:2] == (6, 4)
    assert (E.domain_ex_info(HIV)[4] == oracle.domain_ex_default(HIV)).all()
    rng = np.random.default_rng(2)
    # ContinuousMountainCar: reference terminal-predicate cases (continuous.rs:105-122) + random trajectories
    cases = np.array([[-0.5, 0.0], [0.6, -0.05], [0.6, 0.0], [0.6, 0.05], [0.6 - 0.0001 * 0.6, 0.0], [0.6 + 0.0001 * 0.6, 0.0]])
    assert (E.domain_ex_emit(CMC, cases)[1] == [0, 1, 1, 1, 0, 1]).all()
    s = np.column_stack([rng.uniform(-1.2, 0.6, 3000), rng.uniform(-0.07, 0.07, 3000)])
    for _ in range(20):
        a = rng.uniform(-2.0, 2.0, 3000)     # beyond [-1, 1]: clipped like Interval::map_onto
        ns, obs, r, t = E.domain_ex_step(CMC, s, a)
        ns2, obs2, r2, t2 = oracle.domain_ex_step(CMC, s, a)
        assert np.abs(ns - ns2).max() < 1e-15 and (r == r2).all() and (t == t2).all() and (obs == ns).all()
        s = ns
    # HIV: reference observation cases (hiv.rs:156-204) + steps under every action from states around the default
    obs, t = E.domain_ex_emit(HIV, [[1.0, 10.0, 100.0, 200.0, 500.0, 10000.0], [1e10, 1e-10, 1.0, 1.0, 1.0, 1.0]])
    assert np.abs(obs[0] - [0.0, 1.0, 2.0, 2.301029995663981, 2.698970004336019, 4.0]).max() < 1e-7 and t.sum() == 0
    assert np.abs(obs[1] - [8.0, -5.0, 0.0, 0.0, 0.0, 0.0]).max() < 1e-7
    s = oracle.domain_ex_default(HIV)[None, :] * rng.uniform(0.5, 2.0, size=(64, 6))
    a = rng.integers(0, 4, 64).astype(np.int32)
    for _ in range(2):
        ns, obs, r, t = E.domain_ex_step(HIV, s, a)
        ns2, obs2, r2, t2 = oracle.domain_ex_step(HIV, s, a)
        assert (np.abs(ns - ns2) / np.abs(ns2)).max() < 1e-11      # 4000 gradient evaluations per step, unfused f64 on both sides
        assert np.abs(obs - obs2).max() < 1e-11 and np.abs(r - r2).max() < 1e-12 and t.sum() == 0
        s = ns

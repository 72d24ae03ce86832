"""Pins the CPU oracle against every golden vector / known-answer test the reference holds for the
hot path.  Runs on CPU (-m "not gpu")."""
import numpy as np
import pytest

from rsrl_b200 import abi

MC, CP, AC = abi.MOUNTAIN_CAR, abi.CART_POLE, abi.ACROBOT
X_MAX = 0.6


# ---- rsrl_domains/src/cart_pole.rs:128-183 ----
def test_cartpole_initial_observation(oracle):
    s = oracle.domain_default(CP)
    assert (s == 0.0).all() and len(s) == 4
    assert not oracle.domain_is_terminal(CP, [s])[0]


@pytest.mark.parametrize("action,sign", [(0, -1.0), (1, 1.0)])
def test_cartpole_golden_steps(oracle, action, sign):
    # cart_pole.rs:144-162 (action 0) and :164-183 (action 1); reference tolerance 1e-7
    g1 = np.array([0.0032931628891235, 0.3293940797883472, -0.0029499634056967, -0.2951522145037250]) * sign
    g2 = np.array([0.0131819582085161, 0.6597158115002169, -0.0118185373734479, -0.5921703414056713]) * sign
    s = oracle.domain_default(CP)[None]
    s1, r1, t1 = oracle.domain_step(CP, s, [action])
    assert np.abs(s1[0] - g1).max() < 1e-7
    assert np.abs(s1[0] - g1).max() < 5e-16  # the printed digits are reproduced to the last place
    s2, r2, t2 = oracle.domain_step(CP, s1, [action])
    assert np.abs(s2[0] - g2).max() < 1e-7
    assert np.abs(s2[0] - g2).max() < 5e-16
    assert r1[0] == 0.0 and r2[0] == 0.0 and not t1[0] and not t2[0]


# ---- rsrl_domains/src/mountain_car/discrete.rs:109-137 ----
def test_mountain_car_initial_observation(oracle):
    s = oracle.domain_default(MC)
    assert s[0] == -0.5 and s[1] == 0.0
    assert not oracle.domain_is_terminal(MC, [s])[0]


def test_mountain_car_is_terminal(oracle):
    t = lambda x, v: bool(oracle.domain_is_terminal(MC, [[x, v]])[0])
    assert not t(-0.5, 0.0)
    assert t(X_MAX, -0.05) and t(X_MAX, 0.0) and t(X_MAX, 0.05)
    assert not t(X_MAX - 0.0001 * X_MAX, 0.0)
    assert t(X_MAX + 0.0001 * X_MAX, 0.0)


def test_mountain_car_step_semantics(oracle):
    # discrete.rs:58-65: v uses the old x, x uses the NEW v; no inelastic left wall; reward -1 / 0
    s = np.array([[-0.5, 0.0]])
    ns, r, term = oracle.domain_step(MC, s, [2])
    v = 0.0 + (0.001 * 1.0 + -0.0025 * np.cos(3.0 * -0.5))
    assert ns[0, 1] == v and ns[0, 0] == -0.5 + v and r[0] == -1.0 and not term[0]
    ns, r, term = oracle.domain_step(MC, [[-1.2, -0.07]], [0])
    assert ns[0, 0] == -1.2 and ns[0, 1] < 0.0  # velocity is NOT zeroed at the wall
    ns, r, term = oracle.domain_step(MC, [[0.59, 0.07]], [2])
    assert ns[0, 0] == 0.6 and term[0] and r[0] == 0.0


# ---- rsrl_domains/src/acrobot.rs:159-174 ----
def test_acrobot_initial_observation(oracle):
    s = oracle.domain_default(AC)
    assert (s == 0.0).all() and len(s) == 4
    assert not oracle.domain_is_terminal(AC, [s])[0]


def test_acrobot_wrap_clip_terminal(oracle):
    # macros.rs:3-24 + acrobot.rs:56-79: angles wrapped to [-pi, pi], velocities clipped, terminal iff cos sum < -1
    rng = np.random.default_rng(0)
    s = rng.uniform(-3, 3, size=(256, 4)) * [1, 1, 4, 9]
    a = rng.integers(0, 3, 256)
    ns, r, term = oracle.domain_step(AC, s, a)
    assert (np.abs(ns[:, :2]) <= np.pi).all()
    assert (np.abs(ns[:, 2]) <= 4 * np.pi).all() and (np.abs(ns[:, 3]) <= 9 * np.pi).all()
    expect = np.cos(ns[:, 0]) + np.cos(ns[:, 0] + ns[:, 1]) < -1.0
    assert (term.astype(bool) == expect).all()
    assert ((r == 0.0) == expect).all() and ((r == -1.0) == ~expect).all()
    assert oracle.domain_is_terminal(AC, [[np.pi, 0.0, 0.0, 0.0]])[0]


# ---- rsrl/src/policies/greedy.rs:96-168 (MockQ echoes the state as the Q vector) ----
def _greedy(oracle, q, rnd=(0, 0, 0, 0)):
    return oracle.policy_sample(abi.GREEDY, 0.0, q, rnd)[0]


def test_greedy_1d(oracle):
    assert _greedy(oracle, [1.0]) == 0
    assert _greedy(oracle, [-100.0]) == 0


def test_greedy_two(oracle):
    assert _greedy(oracle, [10.0, 1.0]) == 0 and _greedy(oracle, [1.0, 10.0]) == 1      # test_two_positive
    assert _greedy(oracle, [-10.0, -1.0]) == 1 and _greedy(oracle, [-1.0, -10.0]) == 0  # test_two_negative
    assert _greedy(oracle, [10.0, -1.0]) == 0 and _greedy(oracle, [-10.0, 1.0]) == 1    # test_two_alt
    assert _greedy(oracle, [1.0, -10.0]) == 0 and _greedy(oracle, [-1.0, 10.0]) == 1


def test_greedy_long_and_precision(oracle):
    assert _greedy(oracle, [-123.1, 123.1, 250.5, -1240.0, -4500.0, 10000.0, 20.1]) == 5
    assert _greedy(oracle, [1e-7, 2e-7]) == 1


def test_greedy_probabilities(oracle):
    p = oracle.policy_probs(abi.GREEDY, 0.0, [1e-7, 1e-7, 1e-7, 1e-7])
    assert np.abs(p - 0.25).max() < 1e-6
    p = oracle.policy_probs(abi.GREEDY, 0.0, [1e-7, 2e-7, 3e-7, 4e-7])
    assert np.abs(p - [0.0, 0.0, 0.0, 1.0]).max() < 1e-6


def test_argmaxima_join_does_not_raise_max(oracle):
    # utils.rs:6-21: a value within 1e-7 joins the tie set without raising `max`
    # 1.8e-7 is >= 1e-7 away from max (still 0.0 after 0.9e-7 joined) -> becomes the new strict max
    assert oracle.argmaxima([0.0, 0.9e-7, 1.8e-7]) == ([2], 1.8e-7)
    assert oracle.argmaxima([0.0, 0.5e-7, 0.9e-7]) == ([0, 1, 2], 0.0)


def test_three_argmax_rules(oracle):
    # core.rs:96-105 last exact max; utils.rs:23-34 first with 1e-7 tolerance
    assert oracle.find_max([1.0, 3.0, 3.0, 2.0]) == (2, 3.0)
    assert oracle.argmax_first([1.0, 3.0, 3.0 + 5e-8, 2.0]) == (1, 3.0)
    assert oracle.argmax_first([1.0, 3.0, 3.0 + 2e-7, 2.0])[0] == 2


# ---- rsrl/src/policies/epsilon_greedy.rs:95-145, random.rs:58-76 ----
def test_epsilon_greedy_probabilities(oracle):
    e = lambda q, eps: oracle.policy_probs(abi.EPSILON_GREEDY, eps, q)
    assert np.abs(e([1.0, 0.0, 0.0, 0.0, 0.0], 0.5) - [0.6, 0.1, 0.1, 0.1, 0.1]).max() < 1e-6
    assert np.abs(e([0.0, 0.0, 0.0, 0.0, 1.0], 0.5) - [0.1, 0.1, 0.1, 0.1, 0.6]).max() < 1e-6
    assert np.abs(e([1.0, 0.0, 0.0, 0.0, 1.0], 0.5) - [0.35, 0.1, 0.1, 0.1, 0.35]).max() < 1e-6
    assert np.abs(e([-1.0, 0.0, 0.0, 0.0], 1.0) - 0.25).max() < 1e-6  # test_probabilites_uniform


def test_epsilon_greedy_sampling_frequency(oracle):
    # epsilon_greedy.rs:96-113: Q = [1, 0], eps = 0.5 -> P(0) = 0.75 +- 0.05 over 10 000 samples
    acts = oracle.policy_sample_batch(abi.EPSILON_GREEDY, 0.5, 7, 0, 0, np.tile([1.0, 0.0], (10000, 1)))
    n0 = (acts == 0).mean()
    assert abs(0.75 - n0) < 0.05 and abs(0.25 - (1 - n0)) < 0.05


def test_random_sampling_frequency(oracle):
    acts = oracle.policy_sample_batch(abi.RANDOM, 0.0, 3, 0, 0, np.zeros((10000, 2)))
    assert abs(0.5 - (acts == 0).mean()) < 0.05


def test_greedy_tie_break_is_uniform(oracle):
    acts = oracle.policy_sample_batch(abi.GREEDY, 0.0, 11, 5, 0, np.zeros((9000, 3)))
    for a in range(3):
        assert abs((acts == a).mean() - 1 / 3) < 0.03


# ---- rsrl/src/traces.rs:112-126,135-148 ----
def test_trace_accumulate_doctest(oracle):
    z = oracle.trace_update(abi.TRACE_ACCUMULATE, 0.95, 0.7, 0.0, np.zeros(1), np.ones(1))
    assert abs(z[0] - 1.0) < 1e-12
    z = oracle.trace_update(abi.TRACE_ACCUMULATE, 0.95, 0.7, 0.0, z, np.zeros(1))
    assert abs(z[0] - 0.665) < 1e-12


def test_trace_replace_and_dutch(oracle):
    z = oracle.trace_update(abi.TRACE_REPLACE, 0.95, 0.7, 0.0, np.array([0.9, -0.9]), np.array([1.0, -1.0]))
    assert (z == [1.0, -1.0]).all()  # Saturate clamps to [-1, 1] (traces.rs:213-219)
    z = oracle.trace_update(abi.TRACE_DUTCH, 0.9, 0.5, 0.1, np.array([1.0]), np.array([0.5]))
    assert abs(z[0] - (0.9 * 0.5 * 0.9 * 1.0 + 0.5)) < 1e-15


# ---- Philox4x32-10 known-answer vectors (Random123 kat_vectors) ----
def test_philox_kat(oracle):
    assert [int(x) for x in oracle.philox([0, 0, 0, 0], [0, 0])] == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert [int(x) for x in oracle.philox([0xffffffff] * 4, [0xffffffff] * 2)] == \
        [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert [int(x) for x in oracle.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344],
                                          [0xa4093822, 0x299f31d0])] == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


# ---- lfa::basis::Fourier (parity unpinned; checked against an independent numpy statement) ----
@pytest.mark.parametrize("domain,order", [(MC, 5), (MC, 3), (CP, 2), (AC, 3)])
def test_fourier_matches_published_definition(oracle, domain, order):
    import itertools
    cfg = abi.default_config(domain=domain, basis_order=order)
    D, _ = oracle.domain_dims(domain)
    lo, hi = oracle.domain_limits(domain)
    rng = np.random.default_rng(1)
    s = rng.uniform(lo, hi, size=(16, D))
    coefs = sorted([c for c in itertools.product(range(order + 1), repeat=D) if any(c)], reverse=True)
    assert len(coefs) + 1 == oracle.n_features(cfg) == (order + 1) ** D
    xh = (s - lo) / (hi - lo)
    want = np.cos(np.pi * xh @ np.array(coefs, dtype=np.float64).T)
    got = oracle.project(cfg, s)
    assert np.abs(got[:, :-1] - want).max() < 1e-14
    assert (got[:, -1] == 1.0).all()  # .with_bias() stacks the constant last


def test_default_config_is_the_q_learning_example(oracle):
    cfg = abi.default_config()
    assert oracle.n_features(cfg) == 36 and oracle.domain_dims(cfg.domain) == (2, 3)
    assert (cfg.lr, cfg.gamma, cfg.policy, cfg.algo, cfg.seed) == (0.001, 0.9, abi.GREEDY, abi.QLEARNING, 0)


# ---- TileCoding: project-defined spec; structural properties only ----
def test_tile_coding_structure(oracle):
    cfg = abi.default_config(domain=CP, basis=abi.TILE_CODING, n_tilings=8, tiles_per_dim=8, memory_size=4096)
    lo, hi = oracle.domain_limits(CP)
    rng = np.random.default_rng(0)
    s = rng.uniform(lo, hi, size=(200, 4))
    for x in s:
        rows = oracle.tile_indices(cfg, x)
        assert 1 <= len(rows) <= 8 and len(set(rows.tolist())) == len(rows)
        assert ((rows >= 0) & (rows < 4096)).all()
        assert (rows == oracle.tile_indices(cfg, x)).all()
    # generalisation: a nearby state shares most tiles, a distant one few
    x = (lo + hi) / 2
    near = x + (hi - lo) * 0.004
    far = x + (hi - lo) * 0.4
    base = set(oracle.tile_indices(cfg, x).tolist())
    assert len(base & set(oracle.tile_indices(cfg, near).tolist())) >= 5
    assert len(base & set(oracle.tile_indices(cfg, far).tolist())) <= 1
    phi = oracle.project(cfg, [x])[0]
    assert phi.sum() == len(base) and set(np.nonzero(phi)[0].tolist()) == base


def test_oracle_pal_matches_independent_restatement(oracle):
    """control/td/pal.rs:35-59 restated in numpy (independent of the C oracle): residual = max(AL error, persistent
    AL error) with the reference's literal bootstrap from nqs[a_star]; update error = alpha * residual."""
    cfg = abi.default_config(algo=abi.PAL, policy=abi.EPSILON_GREEDY, epsilon=0.1, alpha=0.5, gamma=0.95, lr=0.01, n_envs=64,
                             dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0], seed=2,
                             record_td_error=1)
    o = oracle.Engine(cfg)
    W0 = np.random.default_rng(0).normal(size=(36, 3)) * 0.3
    o.set_weights(W0)
    s = o.states().copy()
    o.step(1)
    a, td = o.actions(), o.td_errors()
    Q = oracle.evaluate(cfg, W0, s)
    ns, r, term = oracle.domain_step(abi.MOUNTAIN_CAR, s, a)
    NQ = oracle.evaluate(cfg, W0, ns)

    def argmax_first(v):  # utils.rs:23-34
        idx, x = 0, -1.7976931348623157e308
        for j, y in enumerate(v):
            if y - x > 1e-7:
                idx, x = j, y
        return idx

    want, dW = [], np.zeros_like(W0)
    phi = oracle.project(cfg, s)
    for i in range(64):
        if term[i]:
            res = r[i] - Q[i, a[i]]
        else:
            a_star, na_star = argmax_first(Q[i]), argmax_first(NQ[i])
            td_error = r[i] + 0.95 * NQ[i, a_star] - Q[i, a[i]]
            al_error = td_error - 0.5 * (Q[i, a_star] - Q[i, a[i]])
            res = max(al_error, td_error - 0.5 * (NQ[i, na_star] - NQ[i, a[i]]))
        want.append(res)
        dW[:, a[i]] += 0.01 * (0.5 * res) * phi[i]
    assert np.abs(np.array(want) - td).max() < 1e-12
    assert np.abs(o.weights() - (W0 + dW)).max() < 1e-12


# Softmax / Gibbs (policies/softmax.rs): probability golden values of test_probabilites_1 (:253-269), test_1d (:239-246),
# the expected sampling frequencies of test_2d (:248-262), inverse-CDF sampling (policies/mod.rs:46-61)
def test_softmax_probabilities_golden(oracle):
    e = np.e
    assert np.abs(oracle.policy_probs(abi.SOFTMAX, 1.0, [0.0, 1.0]) - [1 / (1 + e), e / (1 + e)]).max() < 1e-15
    assert np.abs(oracle.policy_probs(abi.SOFTMAX, 1.0, [0.0, 2.0]) - [1 / (1 + e * e), e * e / (1 + e * e)]).max() < 1e-15
    p = oracle.policy_probs(abi.SOFTMAX, 0.5, [700.0, 710.0, 705.0])      # softmax_stable: no overflow
    assert np.isfinite(p).all() and abs(p.sum() - 1) < 1e-15 and p.argmax() == 1
    for i in range(1, 100):                                              # test_1d
        assert oracle.policy_sample(abi.SOFTMAX, 1.0, [float(i)], [i * 7919, i, 0, 0])[0] == 0


def test_softmax_sampling(oracle):
    # r = 53-bit uniform from (rnd[0], rnd[1]); index = first running sum > r, else the last
    q = [0.0, 1.0]
    p0 = 1 / (1 + np.e)
    lo = int(p0 * 2 ** 32) - 1
    assert oracle.policy_sample(abi.SOFTMAX, 1.0, q, [lo, 0, 0, 0])[0] == 0
    assert oracle.policy_sample(abi.SOFTMAX, 1.0, q, [lo + 2, 0, 0, 0])[0] == 1
    assert oracle.policy_sample(abi.SOFTMAX, 1.0, q, [0xFFFFFFFF, 0xFFFFFFFF, 0, 0])[0] == 1
    rng = np.random.default_rng(0)
    counts = np.zeros(2)
    for _ in range(20000):
        counts[oracle.policy_sample(abi.SOFTMAX, 1.0, q, rng.integers(0, 2 ** 32, 4, dtype=np.uint64).astype(np.uint32))[0]] += 1
    assert np.abs(counts / 20000 - [p0, 1 - p0]).max() < 1e-2            # test_2d

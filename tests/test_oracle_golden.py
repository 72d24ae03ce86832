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


# ---- GreedyGQ / A2C / rollout: the C oracle against independent numpy restatements of the reference text ----
def test_oracle_greedy_gq_matches_independent_restatement(oracle):
    """control/td/greedy_gq.rs:73-141 restated in numpy: three updates per transition on two weight tables."""
    cfg = abi.default_config(algo=abi.GREEDY_GQ, policy=abi.EPSILON_GREEDY, epsilon=0.1, basis_order=3, lr=0.1, alpha=0.001, gamma=0.99,
                             n_envs=96, dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[0.55, 0.07], seed=5,
                             record_td_error=1)
    rng = np.random.default_rng(0)
    o = oracle.Engine(cfg)
    Wq, Wt = rng.normal(size=(16, 3)) * 0.3, rng.normal(size=(16, 3)) * 0.3
    o.set_weights(Wq)
    o.set_aux_weights(Wt)
    s = o.states().copy()
    s[::3] = np.column_stack([rng.uniform(0.56, 0.6, 32), rng.uniform(0.05, 0.07, 32)])   # a third of the envs reach the goal in this step
    o.set_states(s)
    o.step(1)
    a, td = o.actions(), o.td_errors()
    ns, r, term = oracle.domain_step(abi.MOUNTAIN_CAR, s, a)
    assert term.any() and not term.all()      # both branches of handle()
    phi, nphi = oracle.project(cfg, s), oracle.project(cfg, ns)
    Q, V, NQ = phi @ Wq, phi @ Wt, nphi @ Wq
    dq, dt = np.zeros_like(Wq), np.zeros_like(Wt)
    for i in range(96):
        qsa, td_est = Q[i, a[i]], V[i, a[i]]
        if term[i]:
            e = r[i] - qsa
        else:
            na = max(range(3), key=lambda j: (NQ[i, j], j))   # find_max: last maximal index (core.rs:96-105)
            e = r[i] + 0.99 * NQ[i, na] - qsa
            dq[:, na] += 0.1 * (-0.99 * td_est) * nphi[i]
        dq[:, a[i]] += 0.1 * e * phi[i]
        dt[:, a[i]] += 0.001 * (e - td_est) * phi[i]
        assert abs(e - td[i]) < 1e-12
    assert np.abs(o.weights() - (Wq + dq)).max() < 1e-12 and np.abs(o.aux_weights() - (Wt + dt)).max() < 1e-12


def test_oracle_a2c_matches_independent_restatement(oracle):
    """examples/a2c.rs:37-58 with control/ac.rs:100-114 and policies/softmax.rs:113-129 restated in numpy for ONE agent per env
    (PER_ENV weights = the reference's in-place updates): SARSA critic, then the policy update with the critic closure
    evaluated on the updated Q."""
    cfg = abi.default_config(algo=abi.A2C, policy=abi.SOFTMAX, epsilon=1.0, basis_order=3, lr=0.05, alpha=0.02, gamma=1.0, n_envs=40,
                             weight_mode=abi.PER_ENV, dtype=abi.F64, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[0.55, 0.07],
                             seed=6, record_td_error=1)
    rng = np.random.default_rng(1)
    o = oracle.Engine(cfg)
    Wq, Wp = rng.normal(size=(40, 16, 3)) * 0.3, rng.normal(size=(40, 16, 3)) * 0.5
    o.set_weights(Wq)
    o.set_aux_weights(Wp)
    s = o.states().copy()
    s[::4] = np.column_stack([rng.uniform(0.56, 0.6, 10), rng.uniform(0.05, 0.07, 10)])   # some terminal transitions
    o.set_states(s)
    o.step(1)
    a, td = o.actions(), o.td_errors()
    ns, r, term = oracle.domain_step(abi.MOUNTAIN_CAR, s, a)
    assert term.any() and not term.all()
    phi, nphi = oracle.project(cfg, s), oracle.project(cfg, ns)

    def softmax(h):
        v = np.exp(h - h.max())
        return v / v.sum()
    for i in range(40):
        q, nq = phi[i] @ Wq[i], nphi[i] @ Wq[i]
        # behaviour action: inverse CDF of softmax(theta^T phi(s)) with the env's 53-bit uniform (policies/mod.rs:46-61)
        rnd = oracle.draw(cfg.seed, i, 0, 1)
        u = ((int(rnd[0]) << 32 | int(rnd[1])) >> 11) / 2.0 ** 53
        ps = softmax(phi[i] @ Wp[i])
        assert a[i] == min(int(np.searchsorted(np.cumsum(ps), u, side="right")), 2)
        if term[i]:
            res = r[i] - q[a[i]]
        else:
            rn = oracle.draw(cfg.seed, i, 0, 2)
            un = ((int(rn[0]) << 32 | int(rn[1])) >> 11) / 2.0 ** 53
            na = min(int(np.searchsorted(np.cumsum(softmax(nphi[i] @ Wp[i])), un, side="right")), 2)
            res = r[i] + 1.0 * nq[na] - q[a[i]]
        assert abs(res - td[i]) < 1e-12
        Wq_new = Wq[i].copy()
        Wq_new[:, a[i]] += 0.05 * res * phi[i]
        q2 = phi[i] @ Wq_new
        adv = q2[a[i]] - q2 @ ps
        sf = ps.copy()
        sf[a[i]] -= 1.0
        Wp_new = Wp[i] + (0.02 * adv) * np.outer(phi[i], -sf)
        assert np.abs(o.weights()[i] - Wq_new).max() < 1e-12 and np.abs(o.aux_weights()[i] - Wp_new).max() < 1e-12


def test_oracle_rollout_is_domain_rollout(oracle):
    """rsrl_domains/src/lib.rs:448-479: start = emit(); the first step is always executed; steps stop after a terminal
    observation; Some(limit) keeps limit - 1 steps (`iter.take(sl.saturating_sub(1))`)."""
    cfg = abi.default_config(n_envs=3, dtype=abi.F64)
    o = oracle.Engine(cfg)
    W = np.zeros((36, 3))
    W[-1, 2] = 1.0          # Q(s)[2] = 1 > others: mode() always pushes right
    o.set_weights(W)
    init = np.array([[-0.5, 0.0], [0.5, 0.06], [0.59, 0.07]])
    r = o.rollout(init_states=init, step_limit=500)
    assert (r["start"] == init).all()
    assert (r["len"] == [499, 2, 1]).all()                       # env 0 is still climbing when the limit cuts it
    for i in range(3):
        n = r["len"][i]
        s = init[i].copy()
        for j in range(n):
            s, rew, term = oracle.domain_step(abi.MOUNTAIN_CAR, s[None, :], [2])
            s = s[0]
            assert (r["next"][i, j] == s).all() and r["actions"][i, j] == 2 and r["rewards"][i, j] == rew[0] and r["terminal"][i, j] == term[0]
        assert r["terminal"][i, :n - 1].sum() == 0 and (i == 0 or r["terminal"][i, n - 1] == 1)
    assert (o.rollout(init_states=init, step_limit=1)["len"] == 0).all()     # take(0): the executed first step is not recorded
    assert (o.rollout(init_states=init, step_limit=2)["len"] == 1).all()
    total_reward = r["rewards"][1, :r["len"][1]].sum()                      # Trajectory::total_reward
    assert total_reward == -1.0


# ---- ContinuousMountainCar / HIVTreatment: the reference's own unit tests on the oracle ----
def test_continuous_mountain_car_reference_tests(oracle):
    CMC = abi.CONTINUOUS_MOUNTAIN_CAR
    assert (oracle.domain_ex_default(CMC) == [-0.5, 0.0]).all()                                     # continuous.rs:92-103
    X_MAX = 0.6
    cases = [([-0.5, 0.0], 0), ([X_MAX, -0.05], 1), ([X_MAX, 0.0], 1), ([X_MAX, 0.05], 1),            # continuous.rs:105-122
             ([X_MAX - 0.0001 * X_MAX, 0.0], 0), ([X_MAX + 0.0001 * X_MAX, 0.0], 1)]
    obs, term = oracle.domain_ex_emit(CMC, [c[0] for c in cases])
    assert (term == [c[1] for c in cases]).all() and (obs == np.array([c[0] for c in cases])).all()
    # the action is clipped onto [-1, 1]; FORCE_CAR = 0.0015 (continuous.rs:41-48)
    s, _, r, t = oracle.domain_ex_step(CMC, [[-0.5, 0.0]] * 3, [5.0, 1.0, -0.25])
    assert (s[0] == s[1]).all() and r[0] == -1.0 and t[0] == 0
    assert abs(s[2][1] - (0.0015 * -0.25 - 0.0025 * np.cos(3 * -0.5))) < 1e-18


def test_hiv_reference_tests(oracle):
    HIV = abi.HIV
    obs, term = oracle.domain_ex_emit(HIV, [[1.0, 10.0, 100.0, 200.0, 500.0, 10000.0]])           # hiv.rs:156-170
    assert np.abs(obs[0] - [0.0, 1.0, 2.0, 2.301029995663981, 2.698970004336019, 4.0]).max() < 1e-7 and term[0] == 0
    obs, _ = oracle.domain_ex_emit(HIV, [oracle.domain_ex_default(HIV)])                           # hiv.rs:172-187
    assert np.abs(obs[0] - [5.213711618903007, 4.077186154085897, 0.698970004336019, 1.662757831681574, 4.805629971908577,
                            1.380211241711606]).max() < 1e-7
    obs, _ = oracle.domain_ex_emit(HIV, [[1e10, 1e-10, 1.0, 1.0, 1.0, 1.0]])                       # hiv.rs:189-204
    assert np.abs(obs[0] - [8.0, -5.0, 0.0, 0.0, 0.0, 0.0]).max() < 1e-7
    # one treatment step from the default (unhealthy steady) state: stays near it without drugs, reward follows hiv.rs:141-148
    s0 = oracle.domain_ex_default(HIV)
    s, o, r, t = oracle.domain_ex_step(HIV, [s0, s0], [0, 3])
    assert np.abs(np.log10(s[0]) - np.log10(s0)).max() < 0.05 and t.sum() == 0
    assert abs(r[0] - (1e3 * o[0][5] - 0.1 * o[0][4]) / 1e5) < 1e-15
    assert abs(r[1] - (1e3 * o[1][5] - 0.1 * o[1][4] - 2e4 * 0.49 - 2e3 * 0.09) / 1e5) < 1e-12
    assert s[1][4] < s[0][4]        # both drugs: the free virus count falls


def test_serde_weight_layout():
    """rsrl/Cargo.toml:26 `serde`: Array2<f64> weights serialise as ndarray's {"v", "dim", "data"} with row-major data == the F x A
    layout of rsrl_engine_get_weights (Parameterised::weights_view, rsrl/src/params/mod.rs:116-134)."""
    import json
    from rsrl_b200.engine import weights_from_serde, weights_to_serde
    W = np.arange(12, dtype=np.float64).reshape(4, 3) * 0.5
    doc = json.loads(json.dumps(weights_to_serde(W)))
    assert doc == {"v": 1, "dim": [4, 3], "data": [0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.5, 4.0, 4.5, 5.0, 5.5]}
    assert (weights_from_serde(doc) == W).all()

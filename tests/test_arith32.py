"""CPU tests of the shared host/device arithmetic (csrc/device.cuh "rsrl math", compiled for the host by oracle32) and of
oracle32 itself: accuracy of the elementary functions against mpmath, agreement with the f64 oracle within the fp32
tolerances the GPU tests use, and the invariances the replay has to have (host threads, N = 1 == reference shape)."""
import mpmath as mp
import numpy as np
import pytest

from rsrl_b200 import abi


def _ulp_err(got, exact, ulp):
    return max(abs(float((mp.mpf(float(g)) - e) / mp.mpf(float(u)))) for g, e, u in zip(got, exact, ulp))


def test_f64_trig_accuracy_below_one_ulp(oracle32):
    mp.mp.dps = 40
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.uniform(-40, 40, 3000), rng.uniform(-4, 4, 3000), rng.uniform(-1e6, 1e6, 1000), [0.0, np.pi / 2, -3.6, 1.8]])
    for fn, f in ((0, mp.cos), (1, mp.sin)):
        got = oracle32.math(fn, xs)
        exact = [f(mp.mpf(float(x))) for x in xs]
        ref = np.array([float(e) for e in exact])
        assert _ulp_err(got, exact, np.spacing(np.abs(np.where(ref == 0, 1e-300, ref)))) < 0.8
        assert np.mean(got == ref) > 0.95   # correctly rounded most of the time (glibc, what Rust's f64::cos calls, is < 1 ulp too)


def test_f32_sincospi_and_exp_accuracy_below_one_ulp(oracle32):
    mp.mp.dps = 40
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.uniform(0, 1, 4000), rng.uniform(-3, 3, 1000)]).astype(np.float32).astype(np.float64)
    for fn, f in ((2, lambda x: mp.sin(mp.pi * x)), (3, lambda x: mp.cos(mp.pi * x))):
        got = oracle32.math(fn, xs)
        exact = [f(mp.mpf(float(x))) for x in xs]
        ref = np.array([float(e) for e in exact]).astype(np.float32)
        big = np.abs(ref) > 1e-3
        u = np.spacing(np.abs(ref)).astype(np.float64)
        assert _ulp_err(got[big], [e for e, b in zip(exact, big) if b], u[big]) < 1.0
        assert np.abs(got[~big] - ref[~big].astype(np.float64)).max() < 1e-10   # near the zeros: absolute
    # exact values at the grid points the Fourier tables hit
    for x, s, c in ((0.0, 0.0, 1.0), (0.5, 1.0, 0.0), (1.0, 0.0, -1.0), (1.5, -1.0, 0.0), (0.25, 2 ** -0.5, 2 ** -0.5)):
        assert abs(oracle32.math(2, [x])[0] - s) < 6e-8 and abs(oracle32.math(3, [x])[0] - c) < 6e-8
    xe = np.concatenate([rng.uniform(-20, 5, 4000), rng.uniform(-87, 88, 1000)]).astype(np.float32).astype(np.float64)
    got = oracle32.math(4, xe)
    exact = [mp.exp(mp.mpf(float(x))) for x in xe]
    ref = np.array([float(e) for e in exact]).astype(np.float32)
    assert _ulp_err(got, exact, np.spacing(ref).astype(np.float64)) < 1.0
    assert oracle32.math(4, [np.inf])[0] == np.inf and oracle32.math(4, [-np.inf])[0] == 0.0 and np.isnan(oracle32.math(4, [np.nan])[0])
    assert oracle32.math(4, [0.0])[0] == 1.0


def test_device_physics_within_ulps_of_the_f64_oracle(oracle, oracle32):
    """Domain::step with the shared trig (what the GPU runs) against the libm-based restatement of the reference."""
    for domain in (abi.MOUNTAIN_CAR, abi.CART_POLE, abi.ACROBOT):
        rng = np.random.default_rng(domain)
        lo, hi = oracle.domain_limits(domain)
        s = rng.uniform(lo, hi, size=(3000, len(lo)))
        a = rng.integers(0, 2 if domain == abi.CART_POLE else 3, 3000).astype(np.int32)
        n1, r1, t1 = oracle32.domain_step(domain, s, a)
        n2, r2, t2 = oracle.domain_step(domain, s, a)
        assert (np.abs(n1 - n2) / np.maximum(1.0, np.abs(n2))).max() < (1e-12 if domain == abi.ACROBOT else 2e-14) and (t1 == t2).all() and (r1 == r2).all()
        assert np.mean(np.all(n1 == n2, axis=1)) > 0.9
    # CartPole golden vectors of the reference (cart_pole.rs:144-183) through the device arithmetic
    s = np.zeros((1, 4))
    s, _, _ = oracle32.domain_step(abi.CART_POLE, s, [1])
    assert np.abs(s[0] - [0.0032931628891235, 0.3293940797883472, -0.0029499634056967, -0.2951522145037250]).max() < 1e-15
    s, _, _ = oracle32.domain_step(abi.CART_POLE, s, [1])
    assert np.abs(s[0] - [0.0131819582085161, 0.6597158115002169, -0.0118185373734479, -0.5921703414056713]).max() < 1e-15


def _cfg(**kw):
    base = dict(n_envs=1500, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                max_episode_steps=80, seed=3, update_scale=abi.SCALE_MEAN)
    base.update(kw)
    return abi.default_config(**base)


def test_oracle32_does_not_depend_on_host_threads(oracle32):
    cfg = _cfg()
    sh = oracle32.host_shape(cfg)
    a, b = oracle32.Engine(cfg, sh, threads=1), oracle32.Engine(cfg, sh, threads=5)
    a.step(100)
    b.step(100)
    assert (a.weights() == b.weights()).all() and (a.states() == b.states()).all() and (a.env_stats()[2] == b.env_stats()[2]).all()


@pytest.mark.parametrize("kw", [dict(), dict(algo=abi.SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, lr=0.01),
                                dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99),
                                dict(weight_mode=abi.PER_ENV, update_scale=abi.SCALE_SUM, n_envs=200)],
                         ids=["qlearning", "sarsa", "sarsa_lambda", "per_env"])
def test_oracle32_tracks_the_f64_oracle_teacher_forced(oracle, oracle32, kw):
    """The fp32 arithmetic against the reference's f64, one step at a time from identical inputs (tolerances of the GPU tests)."""
    cfg = _cfg(**kw)
    o32 = oracle32.Engine(cfg, oracle32.host_shape(cfg))
    o64 = oracle.Engine(cfg)
    for t in range(12):
        o64.set_states(o32.states())
        if cfg.weight_mode == abi.SHARED:
            o64.set_weights(o32.weights())
        else:
            o64.set_weights(o32.weights())
        if cfg.algo == abi.SARSA_LAMBDA:
            o64.set_traces(o32.traces())
        o32.step(1)
        o64.step(1)
        same = o32.actions() == o64.actions()
        assert same.mean() > 0.99
        assert np.abs(o32.states()[same] - o64.states()[same]).max() < 1e-12
        assert np.abs(o32.td_errors()[same] - o64.td_errors()[same]).max() < 2e-6 * 6 * max(1.0, np.abs(o64.td_errors()).max())
        if same.all() and cfg.weight_mode == abi.SHARED:
            assert np.abs(o32.weights() - o64.weights()).max() < 1e-6 * max(1.0, np.abs(o64.weights()).max())


def test_oracle32_single_env_is_the_reference_loop(oracle, oracle32):
    """N = 1 (examples/q_learning.rs): no reduction at all; fp32 follows the f64 trajectory for the first episode steps."""
    cfg = abi.default_config(n_envs=1, dtype=abi.F32)
    o32 = oracle32.Engine(cfg, oracle32.host_shape(cfg))
    o64 = oracle.Engine(cfg)
    o32.step(400)
    o64.step(400)
    assert np.abs(o32.states() - o64.states()).max() < 1e-3 and np.abs(o32.weights() - o64.weights()).max() < 1e-3


def test_oracle32_two_ranks_replay(oracle32):
    """world = 2: both ranks step with one W; the sum differs from the single-rank order only in rounding."""
    cfg = _cfg(n_envs=1024, n_envs_global=2048)
    sh = oracle32.host_shape(cfg)
    w2 = oracle32.Engine(cfg, sh, world=2)
    one = _cfg(n_envs=2048)
    w1 = oracle32.Engine(one, oracle32.host_shape(one))
    w2.step(40)
    w1.step(40)
    assert np.abs(w2.weights() - w1.weights()).max() < 1e-6
    s2 = np.concatenate([w2.states(0), w2.states(1)])
    assert np.abs(s2 - w1.states()).max() < 1e-6

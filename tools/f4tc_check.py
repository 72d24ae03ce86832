"""python tools/f4tc_check.py [time]: tcgen05 path of the order-7 basis (f4tc.cuh) against the CUDA-core path and the oracle."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
from oracle import pyoracle as O

AC, CP = abi.ACROBOT, abi.CART_POLE


def cfg_of(domain, n, **kw):
    base = dict(domain=domain, basis_order=7, algo=abi.EXPECTED_SARSA, policy=abi.EPSILON_GREEDY, epsilon=0.1, n_envs=n,
                dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.1] * 4, init_hi=[0.1] * 4, max_episode_steps=500, seed=5,
                gamma=0.99, lr=1e-4, alpha=1.0, update_scale=abi.SCALE_MEAN, record_td_error=1)
    base.update(kw)
    return abi.default_config(**base)


def run(cfg, mask, W0, steps):
    os.environ["RSRL_B200_F4TC"] = str(mask)
    with Engine(cfg) as e:
        e.set_weights(W0)
        e.step(steps)
        e.sync()
        return dict(W=e.weights(), td=e.td_errors(), a=e.actions(), s=e.states(), st=e.stats())


def check(domain, n, algo):
    cfg = cfg_of(domain, n, algo=algo)
    A = 3 if domain == AC else 2
    rng = np.random.default_rng(1)
    W0 = rng.normal(size=(4096, A)) * 0.05
    ref = run(cfg, 0, W0, 1)
    o = O.Engine(cfg)
    o.set_weights(W0)
    o.step(1)
    oW, otd = o.weights(), o.td_errors()
    print(f"domain {domain} n {n} algo {algo}: |W|max {np.abs(oW).max():.3e} |dW|max {np.abs(oW - W0).max():.3e} |td|max {np.abs(otd).max():.3e}")
    print(f"   cuda-core vs oracle: dW err {np.abs(ref['W'] - oW).max():.3e}  td err {np.abs(ref['td'] - otd).max():.3e}  actions differ {(ref['a'] != o.actions()).sum()}")
    for mask in (3,):
        got = run(cfg, mask, W0, 1)
        print(f"   F4TC={mask} vs oracle: dW err {np.abs(got['W'] - oW).max():.3e}  td err {np.abs(got['td'] - otd).max():.3e}  "
              f"actions differ {(got['a'] != o.actions()).sum()}  | vs cuda-core: dW {np.abs(got['W'] - ref['W']).max():.3e} td {np.abs(got['td'] - ref['td']).max():.3e} "
              f"states equal {(got['s'] == ref['s']).all()}", flush=True)


def timing(n=131072, k=20):
    cfg = cfg_of(AC, n, record_td_error=0, seed=0)
    for mask in (0, 3):
        os.environ["RSRL_B200_F4TC"] = str(mask)
        with Engine(cfg) as e:
            e.step(5); e.sync()
            t0 = time.perf_counter(); e.step(k); e.sync(); dt = time.perf_counter() - t0
            print(f"F4TC={mask} N={n}: {1e6 * dt / k:9.2f} us/step  {n * k / dt / 1e9:7.3f} G env-steps/s", flush=True)


if __name__ == "__main__":
    O.build()
    if len(sys.argv) > 1 and sys.argv[1] == "time":
        timing()
    else:
        check(AC, 300, abi.EXPECTED_SARSA)
        check(CP, 1000, abi.QLEARNING)
        check(AC, 4096 + 17, abi.SARSA)
        timing()

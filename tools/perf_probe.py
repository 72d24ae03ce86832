"""Timing probe: us per batched step for a few engine shapes (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from rsrl_b200 import abi
from rsrl_b200.engine import Engine


def run(label, k=2000, **kw):
    base = dict(n_envs=65536, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                max_episode_steps=1000, seed=0, update_scale=abi.SCALE_MEAN)
    base.update(kw)
    cfg = abi.default_config(**base)
    with Engine(cfg) as e:
        e.step(200); e.sync()
        t0 = time.perf_counter(); e.step(k); e.sync(); dt = time.perf_counter() - t0
        print(f"{label:42s} {1e6*dt/k:9.2f} us/step  {cfg.n_envs*k/dt/1e9:8.2f} G env-steps/s", flush=True)


if __name__ == "__main__":
    run("shared f32 N=65536 (148 CTAs)")
    run("shared f32 N=443 (1 CTA, no grid sync)", n_envs=443)
    run("shared f32 N=886 (2 CTAs)", n_envs=886)
    run("shared f32 N=443*16", n_envs=443 * 16)
    run("shared f32 N=128*148 (128 thr/CTA)", n_envs=128 * 148)
    run("per_env f32 N=65536", weight_mode=abi.PER_ENV, update_scale=abi.SCALE_SUM, k=500)
    run("shared f64 N=65536", dtype=abi.F64)
    run("shared f32 eps-greedy", policy=abi.EPSILON_GREEDY)
    run("shared f32 sarsa eps", policy=abi.EPSILON_GREEDY, algo=abi.SARSA)
    run("shared f32 sarsa(lambda) (per-step kernels)", policy=abi.EPSILON_GREEDY, algo=abi.SARSA_LAMBDA, k=200)
    run("cfg5 sarsa(lambda) N=32768 (smem traces)", n_envs=32768, policy=abi.EPSILON_GREEDY, epsilon=0.2, algo=abi.SARSA_LAMBDA,
        alpha=0.01, gamma=0.99, k=1000)
    run("cfg5 td(lambda) N=32768 (smem traces)", n_envs=32768, policy=abi.RANDOM, algo=abi.TD_LAMBDA, gamma=0.99, k=1000)
    run("cfg3 cartpole sarsa tile N=262144", n_envs=262144, domain=abi.CART_POLE, basis=abi.TILE_CODING, algo=abi.SARSA,
        policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=0.1 / 8, init_lo=[-0.05] * 4, init_hi=[0.05] * 4,
        max_episode_steps=500, k=300)
    run("cfg4 acrobot esarsa fourier7 N=131072", n_envs=131072, domain=abi.ACROBOT, basis_order=7, algo=abi.EXPECTED_SARSA,
        policy=abi.EPSILON_GREEDY, epsilon=0.1, gamma=0.99, lr=1e-4, alpha=1.0, init_lo=[-0.1] * 4, init_hi=[0.1] * 4,
        max_episode_steps=500, k=20)

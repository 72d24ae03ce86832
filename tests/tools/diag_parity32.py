"""Diagnostic (run on the GPU box): where does a fp32 engine first differ from oracle32?  python tests/tools/diag_parity32.py [n_envs] [steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from rsrl_b200 import abi
from rsrl_b200.engine import Engine
from oracle import pyoracle32 as O32

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
kw = {}
if os.environ.get("RSRL_ALGO") == "sarsa_lambda":
    kw = dict(algo=abi.SARSA_LAMBDA, policy=abi.EPSILON_GREEDY, epsilon=0.2, alpha=0.01, gamma=0.99)
cfg = abi.default_config(n_envs=n, dtype=abi.F32, init_mode=abi.INIT_UNIFORM, init_lo=[-0.6, 0.0], init_hi=[-0.4, 0.0],
                         max_episode_steps=64, seed=0, update_scale=abi.SCALE_MEAN, record_td_error=1, **kw)
with Engine(cfg) as e:
    sh = e.launch_shape()
    print("launch shape", sh, flush=True)
    o = O32.Engine(cfg, sh)
    for t in range(steps):
        e.step(1); o.step(1); e.sync()
        da = (e.actions() != o.actions()).sum()
        ds = (e.states() != o.states()).any(axis=1).sum()
        dtd = (e.td_errors() != o.td_errors()).sum()
        We, Wo = e.weights(), o.weights()
        dw = (We != Wo).sum()
        print(f"step {t}: actions differ {da}  states differ {ds}  td differ {dtd}  W entries differ {dw}  max|dW| {np.abs(We - Wo).max():.3e}", flush=True)
        if da or ds or dtd or dw:
            if dtd:
                i = int(np.nonzero(e.td_errors() != o.td_errors())[0][0])
                print("  first td mismatch env", i, e.td_errors()[i], o.td_errors()[i], "state", e.states()[i], o.states()[i])
            break
    else:
        print("bit-exact over", steps, "steps")

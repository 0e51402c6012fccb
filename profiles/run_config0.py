import sys, time, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, 'oracle'); sys.path.insert(0,'tests')
import numpy as np
from revs_admm_b200.revs_fixture import REVS
from revs_admm_b200.lpsolver import solve_ADMM, compute_voltage
import revs_oracle as O
fx = REVS(data_path='tests/golden/input', out_path='/tmp/o', grb_path='/tmp/g', fig_path='/tmp/f', comunityID=2, optimizer_mode='distributed')
tariff, homes, dist, saved = fx.read_inputs(adoption=90, rating=4800, seed=1234)
kw = dict(kappa=5.0, iter_max=15, vset=1.03, vlow=0.95, vhigh=1.05)
for i in range(3):
    t0 = time.perf_counter(); diff, P, S, C, st = solve_ADMM(homes, dist, tariff, None, return_stats=True, **kw); t1 = time.perf_counter()
    print(f'GPU solve_ADMM (graph -> dicts, incl. setup) {1e3*(t1-t0):.1f} ms; device total {st["total_ms"]:.2f} ms; launches {st["kernel_launches"]}; max_ws {st["max_working_set"]}')
t0 = time.perf_counter(); do, Po, So, Co = O.solve_ADMM(homes, dist, tariff, None, **kw); t1 = time.perf_counter()
print(f'CPU oracle solve_ADMM {t1-t0:.2f} s')
print('max |dP|', max(np.abs(P[h]-Po[h]).max() for h in Po), 'hours identical', all(np.array_equal(S[h],So[h]) for h in So))

import sys, types, numpy as np, collections
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
src = open('tests/test_gpu_robustness.py').read().replace('pytestmark = pytest.mark.gpu', '')
m = types.ModuleType('r'); exec(compile(src, 'r', 'exec'), m.__dict__)
from oracle import oracle_py as O
import nauyaca_b200 as nb
spec = m._spec(3, ftime=120.0, dt=0.4); osys = O.OracleSystem(spec); eng = nb.BatchEngine(spec)
rng = np.random.default_rng(99); B = 600
flat = np.empty((B, 21))
for i in range(3):
    flat[:, 7*i+0] = 10**rng.uniform(-1, 4.2, B); flat[:, 7*i+1] = rng.uniform(1.5, 9.0, B); flat[:, 7*i+2] = rng.uniform(0.0, 0.95, B)
    flat[:, 7*i+3] = rng.uniform(60, 120, B); flat[:, 7*i+4:7*i+7] = rng.uniform(-400, 800, (B, 3))
flat[0, 2] = 1.5; flat[1, 1] = 0.0; flat[2, 0] = np.nan
chi2, logl, st = eng.loglike(flat, mode=2)
oc, ol, ost = osys.loglike(flat, mode=2)
bad = np.where((st != ost) | ~((chi2 == oc) | (np.isnan(chi2) & np.isnan(oc))))[0]
print('n bad', len(bad), 'of', B)
print('status pairs (gpu, oracle):', collections.Counter(zip(st[bad].tolist(), ost[bad].tolist())))
for b in bad[:8]:
    print(b, st[b], ost[b], chi2[b], oc[b])
    n, pl, ep, tm, rs, vs, s2 = eng.ephemeris(flat[b:b+1], mode=2, cap=400)
    e = O.ephemeris(flat[b], spec['mstar'], spec['t0'], spec['dt'], spec['ftime'], cap=400)
    k = min(n[0], len(e[3]))
    d = np.where(tm[0, :k] != e[3][:k])[0]
    print('   nev', n[0], len(e[3]), 'first diff idx', d[:3], tm[0, d[:2]], e[3][d[:2]], 'planet', pl[0, d[:2]], 'rs', rs[0, d[:2]], e[4][d[:2]])

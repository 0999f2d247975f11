"""Bisect of the multi-output notebook deviation (container only, needs /root/reference): today's model with ONE prior changed at a time,
max relative deviation of mean / variance from the executed cell (Multioutput_Regression.ipynb:270-274).  Result recorded in
tests/test_notebook_parity.py; the winner (ls ~ Gamma(2,1), GP.py:408) is a committed test."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import notebook_multioutput_probe as P  # noqa: E402
import gumbi_b200.map as M
from scipy import stats, optimize
from scipy.spatial.distance import pdist
orig_build = M.build_priors
orig_find = M.find_constrained_invgamma

def run(tag, **kw):
    try:
        gp,mu,s2=P.main(**kw)
        print('%-34s obj %.5f  mu %.2e  s2 %.2e   ls %s sigma %.4f' % (tag, gp.map_result.fun, np.abs(mu/P.NB_MU-1).max(), np.abs(s2/P.NB_S2-1).max(), np.round(gp.MAP['ls_total'],4), gp.MAP['σ']), flush=True)
    except Exception as e:
        print(tag, 'FAILED', repr(e)[:200], flush=True)

run('baseline')
# C: exact constrained prior
M.find_constrained_invgamma = lambda lo, hi, mass=0.98, exact=False: orig_find(lo, hi, mass, exact=True)
run('exact find_constrained_prior')
M.find_constrained_invgamma = orig_find
# D: old find_constrained_prior: least_squares on the mass error only, from init (alpha=lower, beta=upper)
def old_fcp(lo, hi, mass=0.98, exact=False):
    f = lambda p: stats.invgamma.cdf(hi, p[0], scale=p[1]) - stats.invgamma.cdf(lo, p[0], scale=p[1]) - mass
    opt = optimize.least_squares(lambda p: [f(p)], x0=[lo, hi])
    return {"alpha": float(opt.x[0]), "beta": float(opt.x[1])}
M.find_constrained_invgamma = old_fcp
run('old find_constrained (lsq, mass only)')
M.find_constrained_invgamma = orig_find
def variant(mod):
    def bp(gp):
        pri = orig_build(gp)
        mod(gp, pri)
        return pri
    return bp
# A: ls ~ Gamma(2,1)
M.build_priors = variant(lambda gp, pri: pri.update({k: M._gamma(2.0,1.0) for k in pri if k.startswith('ls_')}))
run('ls ~ Gamma(2,1)')
# B: ls ~ Gamma(mu, sigma) from the commented-out get_ls_prior
def modB(gp, pri):
    X = gp._X[:, gp._layout['idx_s']]
    d = pdist(X); dd = d[d != 0]
    l, u = dd.min(), dd.max(); sg = max(0.1, (u-l)/6); mu = l + 3*sg
    a = mu**2/sg**2; b = mu/sg**2
    for k in pri:
        if k.startswith('ls_'): pri[k] = M._gamma(a, b)
M.build_priors = variant(modB)
run('ls ~ Gamma(mu=l+3s, sigma=s)')
for sdW in (1.0, 2.0, 5.0):
    def modW(gp, pri, sdW=sdW):
        for k in list(pri):
            if k.startswith('W_'):
                lp, dlp, _ = M._normal(0.0, sdW); pri[k] = (lp, dlp, pri[k][2])
    M.build_priors = variant(modW); run('W ~ Normal(0,%g)'%sdW)
def modK(gp, pri):
    for k in list(pri):
        if k.startswith('κ_'): pri[k] = M._gamma(2.0, 1.0)
M.build_priors = variant(modK); run('kappa ~ Gamma(2,1)')
def modS(gp, pri):
    pri['σ'] = M._gamma(2.0, 1.0)   # placeholder variant
M.build_priors = variant(modS); run('sigma ~ Gamma(2,1)')
M.build_priors = orig_build

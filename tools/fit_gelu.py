import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as C
# e(t) = 0.5*erfc(t/sqrt2) = Phi(-t), t>=0.  want q(t) ~ log2(e(t)) with abs err in e <= ~3e-8
T = 6.6
print("t*e at T:", T*0.5*erfc(T/np.sqrt(2)))
def target(t): 
    from scipy.special import log_ndtr
    return log_ndtr(-t)/np.log(2)
best=None
for deg in range(6,13):
    # weighted Remez-like: iteratively reweighted least squares on dense grid, weight = e(t)*ln2*(1+t) (error in t*e and e)
    t = np.linspace(0,T,20001)
    y = target(t)
    e = np.exp2(y)
    wgt = e*np.log(2)*np.maximum(1.0,t)
    w = wgt.copy()
    x = 2*t/T-1
    for it in range(60):
        c = C.chebfit(x,y,deg,w=w)
        r = (C.chebval(x,c)-y)*wgt
        # Lawson update
        w = w*(np.abs(r)/np.abs(r).max()+1e-3)**0.5 if it<59 else w
    q = C.chebval(x,c)
    err_e = np.abs(np.exp2(q)-e)
    err_g = t*err_e
    print(deg, "max abs err e: %.3e  t*e: %.3e"%(err_e.max(), err_g.max()))
    if best is None and max(err_e.max(),err_g.max())<2e-8: best=(deg,c)  # fp32 rounding of x*Phi is 6e-8 |x|: 1.4e-8 (degree 8) is below it
deg,c=best
# convert to monomial in t
p = C.cheb2poly(c)  # in x
# x = 2t/T - 1 -> compose
from numpy.polynomial import polynomial as P
xs = np.array([-1.0, 2.0/T])
mono = np.zeros(1)
powx = np.ones(1)
for k in range(len(p)):
    mono = P.polyadd(mono, p[k]*powx)
    powx = P.polymul(powx, xs)
print("deg",deg,"coeffs (t^0..):")
for k,v in enumerate(mono): print(k, repr(float(v)))
# fp32 evaluation check of gelu
def gelu_new(x):
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x), np.float32(T)).astype(np.float32)
    q = np.float32(mono[-1])*np.ones_like(t)
    for k in range(len(mono)-2,-1,-1):
        q = (q*t + np.float32(mono[k])).astype(np.float32)   # not fused but close
    e = np.exp2(q.astype(np.float32)).astype(np.float32)
    # the kernels' tail: x * Phi(x) = max(x, 0) - |x| * Phi(-|x|)  (one FMA; float64 product rounded once ~ fmaf)
    return (np.maximum(x, 0).astype(np.float64) - np.abs(x).astype(np.float64) * e.astype(np.float64)).astype(np.float32)
xx = np.linspace(-8,8,2000001)
xx = xx.astype(np.float32).astype(np.float64)
ref = xx*0.5*(1+erf(xx/np.sqrt(2)))
g = gelu_new(xx)
print("max abs err gelu (fp32 eval):", np.abs(g-ref).max(), "at", xx[np.abs(g-ref).argmax()])
import torch
gt = torch.nn.functional.gelu(torch.tensor(xx,dtype=torch.float32)).numpy()
print("torch fp32 gelu max abs err:", np.abs(gt-ref).max())
rel = np.abs(g-ref)/np.maximum(np.abs(ref),1e-3)
print("max rel (floor 1e-3):", rel.max(), " torch:", (np.abs(gt-ref)/np.maximum(np.abs(ref),1e-3)).max())

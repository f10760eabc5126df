"""Developer check: host-compiled device math vs SciPy golden values (no GPU needed)."""
import ctypes, subprocess, os, sys
import numpy as np
HERE=os.path.dirname(os.path.abspath(__file__))
so=os.path.join(HERE,'_hostsim_special.so')
subprocess.check_call(['g++','-O2','-std=c++17','-ffp-contract=fast','-shared','-fPIC','-o',so,os.path.join(HERE,'hostsim_special.cpp')])
L=ctypes.CDLL(so)
g=np.load(os.path.join(HERE,'..','golden','special.npz'))
P=ctypes.POINTER(ctypes.c_double)
def call(fn,*arrs,nout=1):
    arrs=[np.ascontiguousarray(a,dtype=float) for a in arrs]; n=len(arrs[0])
    outs=[np.empty(n) for _ in range(nout)]
    getattr(L,fn)(ctypes.c_int(n),*[a.ctypes.data_as(P) for a in arrs+outs])
    return outs[0] if nout==1 else outs
def rel(a,b,floor=0):
    ok=np.isfinite(b); d=np.abs(a-b)[ok]/np.maximum(np.abs(b[ok]),floor if floor else 1e-300); return d.max(), int((np.isnan(a)!=np.isnan(b)).sum())
print('rf',rel(call('hs_rf',g['c_x'],g['c_y'],g['c_z']),g['rf']))
print('rd',rel(call('hs_rd',g['c_x'],g['c_y'],g['c_z']),g['rd']))
print('rj',rel(call('hs_rj',g['c_x'],g['c_y'],g['c_z'],g['c_p']),g['rj']))
print('rc',rel(call('hs_rc',g['c_x'],g['c_y']),g['rc']))
print('F',rel(call('hs_F',g['fe_phi'],g['fe_m']),g['F']))
print('E',rel(call('hs_E',g['fe_phi'],g['fe_m']),g['E'],1e-2))
print('K',rel(call('hs_K',g['fe_m']),g['K']))
print('Pi',rel(call('hs_Pi',g['pi_n'],g['pi_phi'],g['pi_m']),g['pi'],1e-1))
print('Pi list',rel(call('hs_Pi',np.full(18,.8),g['pi_list_phi'],np.full(18,-.33)),g['pi_list']))
with np.errstate(all='ignore'): print('Pi sp',call('hs_Pi',g['pisp_n'],g['pisp_phi'],g['pisp_m']),g['pisp'])
sn,cn,dn,am=call('hs_jacobi',g['j_u'],g['j_m'],nout=4)
for k,v in (('sn',sn),('cn',cn),('dn',dn),('am',am)): print(k, np.abs(v-g[k]).max())
import mpmath as mp
mp.mp.dps=40
i=np.argmax(np.abs(dn-g['dn'])); u,m=g['j_u'][i],g['j_m'][i]
ex=mp.ellipfun('dn',mp.mpf(u),m=mp.mpf(m)); print('dn worst',u,m,dn[i],g['dn'][i],float(ex), 'mine err',float(abs(dn[i]-ex)),'scipy err',float(abs(g['dn'][i]-ex)))
pi=call('hs_Pi',g['pi_n'],g['pi_phi'],g['pi_m']); e=np.abs(pi-g['pi'])/np.maximum(np.abs(g['pi']),.1)
for i in np.argsort(-e)[:4]:
    n,phi,m=g['pi_n'][i],g['pi_phi'][i],g['pi_m'][i]
    try:
        ex=mp.ellippi(mp.mpf(n),mp.mpf(phi),mp.mpf(m)); ex=float(mp.re(ex))
    except Exception as ex_: ex=np.nan
    print('Pi',n,phi,m,pi[i],g['pi'][i],ex)
rc=call('hs_rc',g['c_x'],g['c_y']); e=np.abs(rc-g['rc'])/np.abs(g['rc'])
for i in np.argsort(-e)[:3]:
    x,y=g['c_x'][i],g['c_y'][i]; ex=float(mp.elliprc(mp.mpf(x),mp.mpf(y))); print('rc',x,y,rc[i],g['rc'][i],ex)

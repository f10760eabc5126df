import sys, os
import numpy as np
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,HERE); sys.path.insert(0,os.path.join(HERE,'..','..'))
from hostsim import Geo
np.seterr(all='ignore')
G=Geo()
S=np.load(os.path.join(HERE,'..','golden','setup.npz'))
SLOT_TAUTOT=2
for i in range(int(S['ncases'])):
    a,th_o,r_o=S['c%d_par'%i]; A,B=S['c%d_alpha'%i],S['c%d_beta'%i]
    w,fl=G.setup(a,r_o,th_o,A,B,mode=3)
    region=(fl>>8)&7
    tm=S['c%d_tmax'%i]
    ok=np.isfinite(tm)
    print('case',i,'a=%g th=%.0f ro=%g'%(a,th_o*180/np.pi,r_o),'region mismatch',(region!=S['c%d_rclass'%i]).sum(),
          'tau_tot relerr',np.nanmax(np.abs(w[:,SLOT_TAUTOT]-tm)[ok]/np.abs(tm[ok])), 'nan mismatch',(np.isnan(w[:,SLOT_TAUTOT])!=np.isnan(tm)).sum(),
          'outside mismatch', (((fl>>12)&1)!=S['c%d_outside'%i]).sum())
for name in ['default','grid17','classII','extremal','spin0','south','midspin']:
    g=np.load(os.path.join(HERE,'..','golden','tracks_%s.npz'%name))
    a,th_o,r_o,ngeo=g['par']
    eta=(g['alpha']**2-a*a)*np.cos(th_o)**2+g['beta']**2
    mode=(1 if np.any(eta>0) else 0)|(2 if np.any(eta<=0) else 0)|4
    t,r,th,ph,sig=G.points(a,r_o,th_o,g['alpha'],g['beta'],g['tau'],mode=mode)
    print(name)
    for key,got in (('r',r),('th',th),('ph',ph),('t',t),('sig',sig)):
        ref=g[key]; ok=np.isfinite(ref)&np.isfinite(got)
        scale=np.maximum(np.abs(ref),1e-3)
        e=np.abs(got-ref)/scale
        rayscale=np.maximum(np.nanmax(np.abs(ref),axis=0,keepdims=True),1.0)
        e2=np.abs(got-ref)/rayscale
        print('   %-3s nanmis %d  max rel %.2e  (excluding last sample %.2e)  max err/rayscale %.2e  (excl last %.2e)'%(key,(np.isfinite(ref)!=np.isfinite(got)).sum(), np.nanmax(np.where(ok,e,0)), np.nanmax(np.where(ok,e,0)[:-1]), np.nanmax(np.where(ok,e2,0)), np.nanmax(np.where(ok,e2,0)[:-1])))

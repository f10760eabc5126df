"""Developer report (GPU): per-array error statistics of the CUDA path against the golden fixtures."""
import sys, os
import numpy as np
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,os.path.join(HERE,'..','..'))
import kgeo_b200 as kg
from kgeo_b200 import _engine
np.seterr(all='ignore')
G=os.path.join(HERE,'..','golden')
S=np.load(os.path.join(G,'setup.npz'))
for i in range(int(S['ncases'])):
    a,th_o,r_o=[float(x) for x in S['c%d_par'%i]]; A,B=S['c%d_alpha'%i],S['c%d_beta'%i]
    out={k:v.cpu().numpy() for k,v in _engine.ray_summary(a,r_o,th_o,A,B,want=("roots","rclass","tau_tot","uplus","uminus","uratio","a2uminus","outside")).items()}
    tm=S['c%d_tmax'%i]; ok=np.isfinite(tm)
    roots=S['c%d_roots'%i]; scale=np.max(np.abs(roots),axis=0)
    print('case',i,'a=%g th=%.0f ro=%g'%(a,th_o*180/np.pi,r_o),'region mismatch',(out['rclass']!=S['c%d_rclass'%i]).sum(),
      'tau_tot relerr',np.nanmax(np.abs(out['tau_tot']-tm)[ok]/np.abs(tm[ok])),'nanmis',(np.isnan(out['tau_tot'])!=np.isnan(tm)).sum(),
      'outside mis',(out['outside']!=S['c%d_outside'%i]).sum(),
      'roots re',max(np.max(np.abs(out['roots'][k]-roots[k].real)/scale) for k in range(4)),'im',max(np.max(np.abs(out['roots'][4+k]-roots[k].imag)/scale) for k in range(4)),
      'up',np.max(np.abs(out['uplus']-S['c%d_up'%i])),'um',np.max(np.abs(out['uminus']/S['c%d_um'%i]-1)))
for name in ['default','grid17','classII','extremal','spin0','south','midspin']:
    g=np.load(os.path.join(G,'tracks_%s.npz'%name))
    a,th_o,r_o,ngeo=[float(x) for x in g['par']]
    geos=kg.raytrace_ana(a,[0,r_o,th_o,0],[g['alpha'],g['beta']],ngeo=int(ngeo),plotdata=False)
    t,r,th,ph=geos.geo_coords
    print(name,'tau maxrel',np.nanmax(np.abs(geos.mino_times-g['tau'])/np.maximum(np.abs(g['tau']),1e-300)))
    for key,got in (('r',r),('th',th),('ph',ph),('t',t),('sig',geos.affine_times)):
        ref=g[key]; ok=np.isfinite(ref)&np.isfinite(got)
        e=np.abs(got-ref)/np.maximum(np.abs(ref),1.0); e=np.where(ok,e,0)
        k,i=np.unravel_index(np.argmax(e[:-1]),e[:-1].shape)
        print('   %-3s nanmis %d  max scaled %.2e (excl last %.2e) worst k=%d pix=%d alpha=%.6f beta=%.6f got=%.15g ref=%.15g'%(key,(np.isfinite(ref)!=np.isfinite(got)).sum(),e.max(),e[:-1].max(),k,i,g['alpha'][i],g['beta'][i],got[k,i],ref[k,i]))

"""Developer check (no GPU): host-compiled image / lensing / model math vs the golden fixtures."""
import sys, os, json, ctypes
import numpy as np
HERE=os.path.dirname(os.path.abspath(__file__)); sys.path.insert(0,HERE); sys.path.insert(0,os.path.join(HERE,'..','..')); sys.path.insert(0,os.path.join(HERE,'..'))
from hostsim import build, dp
import kgeo_b200
from kgeo_b200 import _lib
from kgeo_b200.models import pack_models
np.seterr(all='ignore')
L=build('hostsim_image')
G=os.path.join(HERE,'..','golden')
D=ctypes.c_double
def mk(kind,t,kw):
    return {'em':kgeo_b200.Emissivity,'vel':kgeo_b200.Velocity,'bf':kgeo_b200.Bfield}[kind](t,**kw)
def stats(got,ref,floor=1e-300):
    nanmis=(np.isnan(got)!=np.isnan(ref)).sum()
    ok=np.isfinite(got)&np.isfinite(ref)
    zm=((got==0)!=(ref==0))&ok
    e=np.abs(got-ref)/np.maximum(np.abs(ref),floor); e=np.where(ok&~zm,e,0)
    return 'nanmis %d zeromis %d maxrel %.2e'%(nanmis,zm.sum(),e.max())
for name in ['default','testpol','subkep_pol','kep_vert_path','thermal_genpower','infall_tor','gelles_simple','simfit_rm1','spin0_default']:
    g=np.load(os.path.join(G,'image_%s.npz'%name))
    a,r_o,th_o,mbar_max,half,npx=g['par']; mbar_max=int(mbar_max)
    spec=json.loads(str(g['spec'])); kw=spec['kw']
    em,vel,bf=mk('em',*spec['em']),mk('vel',*spec['vel']),mk('bf',*spec['bf'])
    A,B=kgeo_b200.equatorial_images.pixel_grid(-half,half,-half,half,2*half/npx)
    n=len(A); M=pack_models(em,vel,bf)
    opt=_lib.ImageOpts(0,mbar_max,int(kw.get('polarization',False)),int(kw.get('pathlength',False)),float(kw.get('specind',1)),float(kw.get('diskangle',0.1)),float(kw.get('nu_obs',230e9)))
    outs={k:np.zeros((mbar_max+1,n)) for k in ('I','Q','U','r','Ir','g','sinthb','lp','Ie')}
    o1={k:np.zeros(n) for k in ('nmax','npol','imax')}
    L.hs_image(D(a),D(r_o),D(th_o),ctypes.byref(M),ctypes.byref(opt),n,dp(A),dp(B),*[dp(outs[k]) for k in ('I','Q','U','r','Ir','g','sinthb')],dp(o1['nmax']),dp(o1['npol']),dp(o1['imax']),dp(outs['lp']),dp(outs['Ie']))
    print(name)
    for k,gk in (('I','I'),('Q','Q'),('U','U'),('r','r'),('Ir','t'),('g','g'),('sinthb','sinthb'),('lp','lp'),('Ie','Ie')):
        if gk in g.files: print('   %-6s'%k, stats(outs[k].T,g[gk]))
    print('   n     ',stats(o1['nmax'],g['n']),' np',stats(o1['npol'],g['np']))
g=np.load(os.path.join(G,'lensing.npz'))
for i in range(int(g['ncases'])):
    a,th_o,r_o=g['c%d_par'%i]; A,B=g['c%d_alpha'%i],g['c%d_beta'%i]; n=len(A)
    rs,Ir,Nm=[np.zeros((3,n)) for _ in range(3)]; Im,npol=np.zeros(n),np.zeros(n)
    L.hs_requatorial(D(a),D(r_o),D(th_o),n,dp(A),dp(B),0,2,dp(rs),dp(Ir),dp(Nm),dp(Im),dp(npol))
    for m in range(3):
        ref=g['c%d_m%d'%(i,m)]
        print('lensing case',i,'m',m,'rs',stats(rs[m],ref[0],1e-6),'Ir',stats(Ir[m],ref[1]),'Imax',stats(Im,ref[2]),'Nmax',stats(Nm[m],ref[3]))
g=np.load(os.path.join(G,'counts.npz'))
a,th_o=g['par']; A,B,tau=g['alpha'],g['beta'],np.ascontiguousarray(g['tau']); n=len(A)
neq,npol=np.zeros(tau.shape),np.zeros(tau.shape)
L.hs_counts(D(a),D(th_o),n,tau.shape[0],dp(A),dp(B),dp(tau),dp(neq),dp(npol))
print('counts neq mismatches',(neq!=g['neq']).sum(),'of',neq.size,' npol',stats(npol,g['npol']))
S=np.load(os.path.join(G,'setup.npz'))
for i in range(int(S['ncases'])):
    a,th_o,r_o=S['c%d_par'%i]; roots=S['c%d_roots'%i]
    R=np.ascontiguousarray(np.concatenate([roots.real,roots.imag])); n=roots.shape[1]; out=np.zeros(n)
    L.hs_mino_total(D(a),D(r_o),n,dp(R),dp(out)); print('mino_total case',i,stats(out,S['c%d_tmax'%i]))
g=np.load(os.path.join(G,'models.npz')); spec=json.loads(str(g['spec']))
genvel=kgeo_b200.Velocity('general',fac_subkep=0.6,beta_phi=0.8,beta_r=0.7)
for ia,a in enumerate(g['spins']):
    a=float(a); r=np.ascontiguousarray(g['a%d_r'%ia]); n=len(r)
    for iv,(vt,kw) in enumerate(spec['vels']):
        M=pack_models(velocity=kgeo_b200.Velocity(vt,**kw)); u=np.zeros((4,n)); L.hs_velocity(ctypes.byref(M),D(a),n,dp(r),dp(u))
        print('a=%g vel %-8s'%(a,vt),stats(u,g['a%d_v%d'%(ia,iv)]))
    for ib,(bt,kw) in enumerate(spec['bfs']):
        M=pack_models(velocity=genvel,bfield=kgeo_b200.Bfield(bt,**kw)); Bv=np.zeros((3,n)); bsq=np.zeros(n); L.hs_bfield(ctypes.byref(M),D(a),n,dp(r),dp(Bv),dp(bsq))
        print('a=%g bf %-10s'%(a,bt),stats(Bv,g['a%d_b%d'%(ia,ib)]),' bsq',stats(bsq,g['a%d_bsq%d'%(ia,ib)]))
    for ie,(et,kw) in enumerate(spec['ems']):
        M=pack_models(emissivity=kgeo_b200.Emissivity(et,**kw)); j=np.zeros(n); gg=0.7+0*r; ss=0.6+0*r
        L.hs_emis(ctypes.byref(M),D(a),n,dp(r),dp(gg),dp(ss),None,D(230e9),D(1.0),dp(j)); print('a=%g em %-8s'%(a,et),stats(j,g['a%d_e%d'%(ia,ie)]))
    M=pack_models(emissivity=kgeo_b200.Emissivity('thermal',ne0=1.e5,Te0=1.e11,B0=5.)); j=np.zeros(n); gg=0.7+0*r; ss=0.6+0*r; bm=3.0+0*r
    L.hs_emis(ctypes.byref(M),D(a),n,dp(r),dp(gg),dp(ss),dp(bm),D(230e9),D(0.0),dp(j)); print('a=%g em thermal'%a,stats(j,g['a%d_ethermal'%ia]))

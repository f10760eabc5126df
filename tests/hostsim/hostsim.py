"""Developer-only loader for the host-compiled device math (algorithm checks without a GPU)."""
import ctypes, os, subprocess
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
P = ctypes.POINTER(ctypes.c_double)
PI = ctypes.POINTER(ctypes.c_int)

def build(name):
    so = os.path.join(HERE, '_%s.so' % name)
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-ffp-contract=fast', *os.environ.get('HOSTSIM_FLAGS', '').split(), '-shared', '-fPIC', '-o', so,
                           os.path.join(HERE, name + '.cpp')])
    return ctypes.CDLL(so)

def dp(a):
    return a.ctypes.data_as(P)

class Geo:
    def __init__(self):
        self.L = build('hostsim_geo')
        self.L.hs_nw.restype = ctypes.c_int
        self.NW = self.L.hs_nw()
    def setup(self, a, r_o, th_o, alpha, beta, mode=7):
        alpha = np.ascontiguousarray(alpha, float); beta = np.ascontiguousarray(beta, float)
        n = len(alpha); w = np.zeros((n, self.NW)); fl = np.zeros(n, np.int32)
        self.L.hs_setup(ctypes.c_double(a), ctypes.c_double(r_o), ctypes.c_double(th_o), mode, n, dp(alpha), dp(beta), dp(w), fl.ctypes.data_as(PI))
        return w, fl
    def points(self, a, r_o, th_o, alpha, beta, tau, mode=7, full=True):
        alpha = np.ascontiguousarray(alpha, float); beta = np.ascontiguousarray(beta, float)
        tau = np.ascontiguousarray(tau, float).reshape(-1, len(alpha)); n = len(alpha); nt = tau.shape[0]
        outs = [np.zeros((nt, n)) for _ in range(5)]
        self.L.hs_points(ctypes.c_double(a), ctypes.c_double(r_o), ctypes.c_double(th_o), mode, int(full), n, nt, dp(alpha), dp(beta), dp(tau), *[dp(o) for o in outs])
        return outs  # t, r, th, ph, sig

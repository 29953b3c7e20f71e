import sys, os, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from lapx_b200 import api, microstructure as ms
from common import make_polycrystal
lib = api.load_product()
grid = tuple(int(v) for v in sys.argv[1].split('x'))
rng = np.random.default_rng(0)
s, ids, grot = make_polycrystal(lib, lib, grid, 10, seed=1)
s.set_loading(api.Loading.strain_rate(np.diag([-0.5,-0.5,1.0])))
nx,ny,nz = grid
sig = rng.normal(size=(6,nz,ny,nx))*20; e = rng.normal(size=(6,nz,ny,nx))*2e-4
s.set_field(api.FIELD_STRESS, sig); s.set_field(api.FIELD_STRAIN, e)
s.begin_increment(2e-4)
s.op_green()
e1 = s.get_field(api.FIELD_STRAIN)
r = s.op_constitutive()
s1 = s.get_field(api.FIELD_STRESS)
np.savez(sys.argv[2], e1=e1, s1=s1, savg=np.array(r.savg[:]))

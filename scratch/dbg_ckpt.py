import sys, numpy as np
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
from lapx_b200 import api
from common import make_polycrystal
lib = api.load_product()
def prep(tex):
    s, ids, grot = make_polycrystal(lib, lib, (16,16,16), 10, seed=6)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=6, itmin=1, tol_newton=1e-9, newton_itmax=100, update_texture=tex)
    s.set_loading(api.Loading.uniaxial_tension(1.0)); return s
for tex in (0, 1):
    a = prep(tex); a.step(2e-4); a.step(2e-4); a.save_state('/tmp/s.ckpt')
    b = prep(tex); b.load_state('/tmp/s.ckpt')
    names = {0:'sig',1:'e',2:'epsp',4:'crss',5:'rot',8:'gacc',11:'wrot'}
    print('tex', tex, 'state equal after load:', {n: bool(np.array_equal(a.get_field(f), b.get_field(f))) for f,n in names.items()})
    print(' macro', a.get_macro()[0], b.get_macro()[0])
    a.begin_increment(2e-4); b.begin_increment(2e-4)
    print(' macro after begin', np.array_equal(a.get_macro()[0], b.get_macro()[0]), a.get_macro()[0]-b.get_macro()[0])
    for it in range(3):
        a.op_green(); b.op_green()
        de = np.abs(a.get_field(1)-b.get_field(1)).max()
        ra = a.op_constitutive(); rb = b.op_constitutive()
        ds = np.abs(a.get_field(0)-b.get_field(0)).max()
        print('  it', it, 'e diff', de, 'sig diff', ds, 'savg diff', np.abs(np.array(ra.savg[:])-np.array(rb.savg[:])).max(), 'E diff', np.abs(np.array(ra.emacro[:])-np.array(rb.emacro[:])).max())

"""Small deck for compute-sanitizer: exercises k_constitutive_p (FCC table + 24-system variant), k_zfused2<128,2>,
the x/y passes and the reductions.  usage: compute-sanitizer --tool memcheck|racecheck python scratch/sanitize_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from common import make_polycrystal
from lapx_b200 import api

lib = api.load_product()
for hcp in (False, True):
    s, ids, grot = make_polycrystal(lib, lib, (16, 8, 128), 6, seed=4, hcp=hcp)
    s.set_control(tol_stress=1e-30, tol_strain=1e-30, itmax=10**6, tol_newton=1e-9, newton_itmax=100)
    s.set_loading(api.Loading.uniaxial_tension(1.0))
    s.begin_increment(2e-4)
    for it in range(2):
        r = s.equilibrium_iter()
    s.end_increment()
    print("hcp" if hcp else "fcc", r.err_stress, r.newton_max, float(np.abs(s.get_field(api.FIELD_STRESS)).max()))
    s.close()
print("SANITIZE_DECK_OK")

"""Material tables shared by the golden generator and the tests (inputs, not algorithm)."""
from lapx_b200 import microstructure as ms


def twin_phase(lib):
    """HCP Zr-like phase with an easy {10-12} tensile twin mode and low PTR thresholds (reorientation within a few increments)."""
    ph = ms.hcp_phase(lib, with_twin=1, nrate=10.0, tau0_mode=(60.0, 120.0, 200.0, 25.0),
                      voce_mode=[[5.0, 100.0, 5.0], [10.0, 200.0, 10.0], [20.0, 400.0, 20.0], [5.0, 50.0, 5.0]])
    ph.twin_thr1 = 5.0e-8
    ph.twin_thr2 = 1.0e-13
    return ph


def twin_phase_ratio(lib):
    """As twin_phase, with a PTR threshold that depends on F_eff / F_acc (golden hcp8_twin_ratio)."""
    ph = twin_phase(lib)
    ph.twin_thr1 = 4.0e-8
    ph.twin_thr2 = 1.0e-12
    return ph

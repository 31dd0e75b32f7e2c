"""Fixed-mesh wavelength sweep on one GPU (config-5 style at reduced size): per-point time with reuse."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.sweep import wavelength_sweep, group_index
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
npts = int(sys.argv[2]) if len(sys.argv) > 2 else 8
m, eps = strip_problem(n, slab_h=0.11)
b0 = Basis(m, ElementTriP0(), intorder=4)
core = np.zeros(m.nelements, dtype=bool); core[m.subdomains["core"]] = True; core[m.subdomains["slab"]] = True
def eps_of(wl):
    # toy Sellmeier-like dispersion of the core (vary_wavelength.py:41-53 uses tabulated n(lambda))
    e = np.full(m.nelements, 1.444**2)
    e[core] = (3.4777 - 0.08 * (wl - 1.55)) ** 2
    return e
wls = np.linspace(1.5, 1.6, npts)
t0 = time.time(); ne1 = wavelength_sweep(b0, eps_of, wls[:1], num_modes=1, order=2); t1 = time.time()
neff = wavelength_sweep(b0, eps_of, wls, num_modes=1, order=2); t2 = time.time()
print(f"ne={m.nelements}: first point {t1-t0:.2f}s; {npts} points {t2-t1:.2f}s = {(t2-t1)/npts:.3f} s/pt = {npts/(t2-t1):.2f} pts/s")
print("n_eff", np.round(neff[:, 0].real, 6)); print("n_g", np.round(group_index(wls, neff[:, 0].real), 4))
from femwell_b200.maxwell.waveguide import compute_modes
for wl in wls[:3]:
    t0 = time.time(); e = eps_of(wl); t1 = time.time()
    md = compute_modes(b0, e, wl, num_modes=1, order=2); t2 = time.time()
    tm = md.stats["timings"]
    print(f"eps {t1-t0:.3f}s compute_modes {t2-t1:.3f}s device {tm['total_ms']:.0f} ms [sym {tm['symbolic_ms']:.0f} lu_sym {tm['lu_symbolic_ms']:.0f} fac {tm['lu_factor_ms']:.0f} arn {tm['arnoldi_ms']:.0f} post {tm['postprocess_ms']:.0f} upl {tm['upload_ms']:.0f} dl {tm['download_ms']:.0f}] n_op {md.stats['n_op']}")

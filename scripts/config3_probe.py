"""BASELINE config 3 style run: bent rib waveguide, R = 5 um, DG-P1 complex eps with PML ramp, ~1M triangles."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from femwell_b200.basis import Basis
from femwell_b200.element import ElementDG, ElementTriP1
from femwell_b200.mesh import waveguide_mesh
from femwell_b200.maxwell.waveguide import compute_modes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 707
m = waveguide_mesh(n, slab_h=0.11, bbox=(-1.25, 4.25, -1.0, 1.22), extra_x=[2.25])
b0 = Basis(m, ElementDG(ElementTriP1()), intorder=4)
base = np.full(m.nelements, 1.0)
for name, nn in {"core": 3.48, "slab": 3.48, "box": 1.48, "clad": 1.0}.items():
    base[m.subdomains[name]] = nn**2
eps = np.repeat(base, 3).astype(complex)
xv = m.p[0, m.t].T.reshape(-1)
eps += -10j * np.maximum(0, xv - 2.25) ** 2
t0 = time.time(); straight = compute_modes(b0, eps, 1.55, num_modes=1, order=2); t1 = time.time()
print(f"straight: ne={m.nelements} N={straight[0].basis.N} n_eff={straight[0].n_eff} wall {t1-t0:.2f}s")
print("   ", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in straight.stats.items() if k != 'timings'})
t0 = time.time(); bent = compute_modes(b0, eps, 1.55, num_modes=1, order=2, radius=5.0, n_guess=straight[0].n_eff); t1 = time.time()
print(f"bent R=5: n_eff={bent[0].n_eff} wall {t1-t0:.2f}s overlap with straight {abs(straight[0].calculate_overlap(bent[0])):.6f}")
print("   ", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in bent.stats['timings'].items()})
print("   ", {k: (round(v, 1) if isinstance(v, float) else v) for k, v in bent.stats.items() if k != 'timings'})

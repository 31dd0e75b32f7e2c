import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes
for n in [int(a) for a in sys.argv[1:]]:
    t0 = time.time(); m, eps = strip_problem(n, slab_h=0.11); t1 = time.time()
    b0 = Basis(m, ElementTriP0(), intorder=4)
    for rep in range(2):
        t2 = time.time()
        modes = compute_modes(b0, eps, 1.55, num_modes=4, order=2)
        t3 = time.time()
        st = modes.stats
        print(f"n={n} ne={m.nelements} N={modes[0].basis.N} mesh {t1-t0:.2f}s compute_modes {t3-t2:.3f}s n_eff {np.round(modes.n_effs.real,6)}", flush=True)
        print("   ", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items() if k != 'timings'}, flush=True)
        print("   ", {k: (round(v, 2) if isinstance(v, float) else v) for k, v in st['timings'].items()}, flush=True)

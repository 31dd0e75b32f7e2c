import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes
n = int(sys.argv[1])
m, eps = strip_problem(n, slab_h=0.11)
b0 = Basis(m, ElementTriP0(), intorder=4)
for leaf in [int(a) for a in sys.argv[2:]]:
    for rep in range(2):
        modes = compute_modes(b0, eps, 1.55, num_modes=4, order=2, solver_options={"leaf_elements": leaf})
    st = modes.stats; tm = st["timings"]
    print(f"leaf={leaf} levels={st['lu_levels']} fronts={st['lu_fronts']} entries={st['lu_factor_entries']/1e9:.2f}G sym={st['t_symbolic_ms']:.0f} fac={st['t_factor_ms']:.0f} solve/op={st['t_solve_ms_total']/st['n_op']:.2f} n_op={st['n_op']} arn={st['t_arnoldi_ms']:.0f} post={tm['postprocess_ms']:.0f} total={tm['total_ms']:.0f} res={st['max_residual']:.1e}", flush=True)

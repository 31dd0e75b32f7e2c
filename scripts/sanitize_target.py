"""Small end-to-end solve exercised under compute-sanitizer (scripts/sanitize.sh)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import strip_problem
from femwell_b200.basis import Basis
from femwell_b200.element import ElementTriP0
from femwell_b200.maxwell.waveguide import compute_modes

for order, n, metallic in ((2, 12, False), (1, 16, True)):
    m, eps = strip_problem(n)
    b0 = Basis(m, ElementTriP0(), intorder=4)
    modes = compute_modes(b0, eps, 1.55, num_modes=2, order=order, metallic_boundaries=metallic)
    md = modes[0]
    print(order, modes.n_effs, md.calculate_power(), md.calculate_overlap(modes[1]), md.te_fraction,
          md.eval_error_estimator().sum(), md.calculate_intensity()[1].sum())

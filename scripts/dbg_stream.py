import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import panopaea_b200 as P
from panopaea_b200 import pcg
ctx = P.Context(0)
h, w = int(sys.argv[1]) if len(sys.argv) > 1 else 24, int(sys.argv[2]) if len(sys.argv) > 2 else 20
ctx.set_option("cg_kernel", int(os.environ.get("CGK", "0")))
g = P.Grid2d((h, w), ctx)
rng = np.random.default_rng(0)
b = rng.normal(size=(h, w)); b -= b.mean()
x, r, aux, s, B = (g.new_simplex_2() for _ in range(5))
B.upload(b)
print(pcg.solve_grid_laplacian(x, B, 20, 1e-3, r, aux, s, 0.05, (0, 0, 0, 0)))
print(np.abs(x.to_host()).max())

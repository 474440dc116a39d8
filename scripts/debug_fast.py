import sys; sys.path.insert(0, ".")
import numpy as np
from tests import scenarios as S
sc = S.periodic_box(9000, "M4", "cd10", jitter=0.0)
o = S.make_oracle(sc); m = S.make_cuda(sc, fp_mode="fast")
for k in range(2):
    o.evolve_once(); m.evolve_once()
    v = np.abs(o.get(0, "vxyz")).max(); h = o.get(0, "hpart").min()
    print("step", k, "vmax", v, "hmin", h, "v/h", v / h)
    for nm in ("divv", "curlv", "dtdivv", "axyz", "duint", "hpart", "step.omega"):
        a, b = m.get(0, nm), o.get(0, nm)
        d = np.abs(a - b)
        i = np.unravel_index(np.argmax(d), d.shape)
        print(f"  {nm:10s} max|d|={d.max():.3e} at {i} ref there={b[i]:.3e} max|ref|={np.abs(b).max():.3e} mean|ref|={np.abs(b).mean():.3e}")

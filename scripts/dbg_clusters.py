import sys, ctypes
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
torch.cuda.init(); torch.zeros(1, device="cuda")
from hulc_b200 import _lib
f = _lib.lib().cdll.hulc_debug_max_clusters
for split in (0, 1):
    for c in (8, 4, 2):
        n = ctypes.c_int(-1)
        rc = f(c, split, ctypes.byref(n))
        print("split", split, "cluster", c, "rc", rc, "max active clusters", n.value, "=> CTAs", n.value * c)

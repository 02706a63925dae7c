import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gwin_b200 import _lib
names = {0: "DFMA, shared multiplier/addend (reuse-friendly)", 1: "DFMA, 3 private source registers",
         2: "complex multiply chains (2 DMUL + 2 DFMA)"}
for v in (0, 1, 2):
    r = _lib.fp64_probe(0, v)
    print("variant %d  %-48s %.4e lane-ops/s = %.1f%% of 148x64x1.965e9" % (v, names[v], r, 100 * r / 1.861e13))

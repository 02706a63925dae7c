import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gwin_b200 import _lib
names = {0: "DFMA, shared multiplier/addend (reuse-friendly)", 1: "DFMA, 3 private source registers",
         2: "complex multiply chains (2 DMUL + 2 DFMA)", 3: "FP64 tensor-core MMA m8n8k4 (16 flop per lane-op)",
         4: "MMA m8n8k4 interleaved 1:1 with DFMA"}
for v in (0, 1, 2, 3, 4):
    r = _lib.fp64_probe(0, v)
    flop = {3: 16.0, 4: 18.0}.get(v, 2.0)
    print("variant %d  %-52s %.4e lane-ops/s = %.1f%% of 148x64x1.965e9   %.1f TFLOP/s"
          % (v, names[v], r, 100 * r / 1.861e13, r * flop / 1e12))

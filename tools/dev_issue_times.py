import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import util
from gwin_b200 import _lib
cfg = util.c3()
model = cfg.cuda_model()
plan = model.plan
plan.configure(mchirp_min=1.0)
rng = np.random.default_rng(1)
W = 8192
P = torch.tensor(np.ascontiguousarray(util.draw_bns(rng, W, cfg).T), device="cuda")
out = torch.empty(W, dtype=torch.float64, device="cuda")
lib = _lib.load()
st = torch.cuda.current_stream().cuda_stream
def call():
    _lib.check(lib.gwin_loglr_batch_strided(plan._h, 0, W, P.data_ptr(), 1, W, None, out.data_ptr(), None, None, st))
call(); torch.cuda.synchronize()
ts = []
t0 = time.perf_counter()
for i in range(8):
    a = time.perf_counter(); call(); ts.append(time.perf_counter() - a)
t_issue = time.perf_counter() - t0
torch.cuda.synchronize()
t_all = time.perf_counter() - t0
print("8 back-to-back device-API calls: issue %.2f ms, total %.2f ms; per-call host times [ms]:" % (t_issue * 1e3, t_all * 1e3), np.round(np.array(ts) * 1e3, 3))

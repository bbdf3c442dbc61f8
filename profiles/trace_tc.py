#!/usr/bin/env python
"""In-kernel timeline of k_tc_bilinear (needs librae_trace.so, built with -DRAE_TRACE)."""
import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from relation_autoencoder_b200 import _lib as L
L.LIB_PATH = os.path.join(ROOT, "librae_trace.so")
from relation_autoencoder_b200 import synthetic as SY
from relation_autoencoder_b200.engine import Engine
import bench

wlname = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = dict(SY.WORKLOADS[wlname]); B = wl["B"]
data, params, neg1, neg2 = bench._make_inputs(wl, 8 * B)
eng = Engine(wl["model"], wl["K"], wl["d"], wl["S"], B, wl["F"], wl["N"], wl["N_train"])
eng.set_params_numpy(params); eng.bind_split("train", data.indptr, data.indices, data.args1, data.args2)
eng.bind_epoch_negatives(neg1, neg2)
for b in range(4): eng.train_device(b, want_cost=False)
torch.cuda.synchronize()
buf = torch.zeros(256 * 64, dtype=torch.int64, device="cuda")
lib = C.CDLL(L.LIB_PATH)
lib.rae_debug_set_trace.argtypes = [C.c_void_p]
assert lib.rae_debug_set_trace(C.c_void_p(buf.data_ptr())) == 0
eng.train_device(4, want_cost=False)
torch.cuda.synchronize()
t = buf.cpu().numpy().reshape(256, 64)
names = {0: "entry", 1: "setup done", 2: "producer first issue", 3: "producer all issued", 4: "mma: a_full passed", 5: "q in tmem",
         6: "epi loop done", 7: "wp stored", 62: "final sync", 63: "dealloc"}
for i in range(6): names[8 + 2 * i] = "mma it%d t_empty ok" % i; names[9 + 2 * i] = "mma it%d b_full ok" % i
for i in range(8): names[24 + 2 * i] = "epi it%d wait t_full" % i; names[25 + 2 * i] = "epi it%d t_full ok" % i
names.update({32: "DQ entry", 33: "DQ setup done", 34: "DQ gen loop done", 35: "DQ final sync"})
for i in range(3): names[8 + 2 * i] = ("mma it%d t_empty ok | DQ mma it%d a_full ok" % (i, i)); names[9 + 2 * i] = "mma it%d b_full ok | DQ mma it%d b_full ok" % (i, i)
for i in range(4): names[36 + 2 * i] = "DQ gen it%d computed" % i; names[37 + 2 * i] = "DQ gen it%d published" % i
names.update({52: "DC entry", 53: "DC setup done", 55: "DC gen loop done", 54: "DC final sync"})
for i in range(3): names[56 + 2 * i] = "DC mma it%d a_full ok" % i; names[57 + 2 * i] = "DC mma it%d b_full ok" % i
for i in range(4): names[44 + 2 * i] = "DC gen it%d computed" % i; names[45 + 2 * i] = "DC gen it%d published" % i
for i in range(4): names[24 + 2 * i] = "epi it%d wait t_full" % i; names[25 + 2 * i] = "epi it%d t_full ok" % i
for i in range(0, 8, 2): names[24 + i] = "epi it%d wait t_full" % i; names[25 + i] = "epi it%d t_full ok" % i
names.update({16: "epi it4 ld0 done", 17: "epi it4 ld1 done", 18: "epi it4 chunk done", 20: "epi it6 ld0 done", 21: "epi it6 ld1 done", 22: "epi it6 chunk done"})
for i in range(4): names[8 + 2 * i] = "mma it%d t_empty ok" % i; names[9 + 2 * i] = "mma it%d b_full ok" % i
for k in list(names):
    if 32 <= k < 62: del names[k]
names.update({32: "DQ entry", 33: "DQ setup done", 34: "DQ gen loop done", 35: "DQ final sync"})
for i in range(4):
    names[36 + 3 * i] = "DQ gen it%d computed" % (40 + i); names[37 + 3 * i] = "DQ gen it%d a_empty ok" % (40 + i); names[38 + 3 * i] = "DQ gen it%d published" % (40 + i)
dqm = {}
for i in range(3):
    dqm[8 + 2 * i] = "DQ mma it%d a_full ok" % (40 + i); dqm[9 + 2 * i] = "DQ mma it%d b_full ok" % (40 + i); dqm[14 + i] = "DQ mma it%d issued" % (40 + i)
groups = {"FWD": [k for k in names if k < 32 or k >= 62], "DQ": [k for k in names if 32 <= k < 48]}
if os.environ.get("TRACE_DQ"):
    names.update(dqm)
    groups = {"DQ": [k for k in names if 32 <= k < 48] + list(dqm)}
for cta in (1, 127):
    r = t[cta]
    for gname, slots in groups.items():
        base = {"FWD": 0, "DQ": 32, "DC": 52}[gname]
        if r[base] == 0: continue
        print("CTA", cta, gname)
        for slot in sorted(slots, key=lambda s: r[s]):
            if r[slot] and abs(int(r[slot]) - int(r[base])) < 2000000:
                print("   %8d cyc  %7.2f us  [%2d] %s" % (r[slot] - r[base], (r[slot] - r[base]) / 1965.0, slot, names[slot]))
for cta in ():
    r = t[cta]
    if r[0] == 0: continue
    print("CTA", cta)
    for slot in sorted(names, key=lambda s: (r[s] if r[s] else 1 << 62)):
        if r[slot]:
            print("   %8d cyc  %7.2f us  %s" % (r[slot] - r[0], (r[slot] - r[0]) / 1965.0, names[slot]))
eng.close()

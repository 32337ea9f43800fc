"""Event trace of one tensor-memory RIC launch (CTA 0): python tools/tm_trace.py <step> [precision] [first] [count]
step = 1 + launch index inside stage 1 (21 = upconv1, 22 = conv_11, 6 = resnets.0.conv_0)."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import drawingspinup_b200 as dsu  # noqa: E402
from drawingspinup_b200 import capi, synth  # noqa: E402
from drawingspinup_b200.pipeline import DEFAULT_ARGS  # noqa: E402

step = int(sys.argv[1]) if len(sys.argv) > 1 else 21
prec = sys.argv[2] if len(sys.argv) > 2 else "fp16"
first = int(sys.argv[3]) if len(sys.argv) > 3 else 300
count = int(sys.argv[4]) if len(sys.argv) > 4 else 90
dev = torch.device("cuda:0")
m = dsu.GeneratorJ_RIC(precision=prec, **DEFAULT_ARGS)
m.load_state_dict(synth.to_torch_state_dict(synth.make_state_dict(1, out_gain=0.25)))
m = m.to(dev).eval()
c, p, _ = synth.make_frames(16, 512, 512, seed=3)
cd, pd = torch.from_numpy(c).to(dev), torch.from_numpy(p).to(dev)
with torch.no_grad():
    for _ in range(2):
        m.forward_frames(cd, pd, None)
    torch.cuda.synchronize()
    m.set_knob("tm_trace", step)
    m.forward_frames(cd, pd, None)
    torch.cuda.synchronize()
n = 32 + 5 * 1024
buf = (C.c_uint64 * n)()
capi.lib().dsu_debug_watchdog(m._handle, buf, n)
names = {1: "P start", 2: "P done ", 3: "I start", 4: "I commit", 5: "E start", 6: "E done ", 7: "W load ", 8: "E rows "}
evs = []
for role in range(5):
    for i in range(1024):
        v = buf[32 + role * 1024 + i]
        if v:
            evs.append((v & 0xFFFFFFFFFF, v >> 56, (v >> 40) & 0xFFFF))
evs.sort()
if not evs:
    print("no events recorded")
    sys.exit(1)
t0 = evs[0][0]
print("events: %d, span %.1f kclk" % (len(evs), (evs[-1][0] - t0) / 1e3))
sel = [e for e in evs if e[1] in (1, 2, 3, 4, 7) and first <= e[2] < first + count // 5] + [e for e in evs if e[1] in (5, 6, 8)]
sel.sort()
lo = min(e[0] for e in sel if e[1] == 1)
for t, ev, a in sel:
    if t < lo - 2000 or t > lo + 60000:
        continue
    print("%9.2f kclk  %-9s %d" % ((t - t0) / 1e3, names.get(ev, str(ev)), a))
# per-stage statistics over the steady state
import collections
ts = collections.defaultdict(dict)
for t, ev, a in evs:
    ts[a][ev] = t
js = sorted(j for j in ts if all(k in ts[j] for k in (1, 2, 3, 4)) and j + 1 in ts and 1 in ts[j + 1] and j > 50)
if js:
    P = sum(ts[j][2] - ts[j][1] for j in js) / len(js)
    gap = sum(ts[j][3] - ts[j][2] for j in js) / len(js)
    iss = sum(ts[j][4] - ts[j][3] for j in js) / len(js)
    per = (ts[js[-1]][1] - ts[js[0]][1]) / max(1, js[-1] - js[0])
    print("steady state over %d stages: produce %.0f clk, a_full->issuer %.0f clk, issue+commit %.0f clk, period %.0f clk/stage" % (len(js), P, gap, iss, per))

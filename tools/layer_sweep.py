"""Per-layer timing sweep over engine tuning knobs (env DSU_HALO_NS / DSU_HALO_SB), B=16, 512x512.
Development aid (run under gpurun): python tools/layer_sweep.py"""
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    import drawingspinup_b200 as dsu
    from drawingspinup_b200 import synth
    from drawingspinup_b200.pipeline import DEFAULT_ARGS
    stage = int(os.environ.get("SWEEP_STAGE", "2"))
    prec = os.environ.get("SWEEP_PREC", "fp16")
    cls = dsu.GeneratorJ if stage == 2 else dsu.GeneratorJ_RIC
    m = cls(precision=prec, **DEFAULT_ARGS)
    m.load_state_dict(synth.to_torch_state_dict(synth.make_state_dict(stage, out_gain=0.25)))
    m = m.to("cuda:0").eval()
    B = 16
    c, p, e = synth.make_frames(B, 512, 512, seed=1)
    with torch.no_grad():
        m.forward_frames(torch.from_numpy(c).cuda(), torch.from_numpy(p).cuda(),
                         torch.from_numpy(e).cuda() if stage == 2 else None)
    rows = m.profile_layers(B, 512, 512, reps=3)
    agg = {}
    for name, ms, fl in rows:
        key = "res" if name.startswith("resnets") else name
        a = agg.setdefault(key, [0.0, 0.0])
        a[0] += ms
        a[1] += fl
    tot = sum(v[0] for v in agg.values())
    print(json.dumps({"total_ms": round(tot, 2), **{k: (round(v[0], 3), round(v[1] / max(v[0], 1e-9) / 1e9, 1)) for k, v in agg.items()}}))
else:
    combos = [(2, "fp16", "", "", ""), (2, "fp16", "2", "2", ""), (2, "fp16", "2", "1", ""), (2, "fp16", "1", "2", ""), (2, "fp16", "1", "4", ""),
              (2, "fp16", "", "", "1"), (2, "fp16", "", "", "3"), (1, "fp16", "", "", ""), (2, "fp16x3", "", "", ""), (1, "fp16x3", "", "", "")]
    for stage, prec, ns, ks, tps in combos:
        env = dict(os.environ, **({'DSU_HALO_NS': ns, 'DSU_HALO_KS': ks} if ns else {}), **({'DSU_HALO_TPS': tps} if tps else {}),
                   SWEEP_STAGE=str(stage), SWEEP_PREC=prec)
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=env, capture_output=True, text=True)
        line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-300:]
        print("stage%d %-6s NS=%s KS=%s TPS=%s %s" % (stage, prec, ns, ks, tps, line), flush=True)

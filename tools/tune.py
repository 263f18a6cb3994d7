"""A/B sweep of the library's TB_* tuning knobs on one GPU: one ProverService, one synthetic batch, every knob set timed with the
built-in CUDA-event profiler.  usage: python tools/tune.py [ptx_per_step]"""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from taiga_b200 import ptx
import bench

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16
srs = bench.load_srs()
svc = ptx.ProverService(0, srs, c_workers=1, v_workers=1)
base = svc.synthesize_ptx(2, wseed=5)
rep = (P + 1) // 2
wit = {k: (np.concatenate([v] * rep)[: (2 * P if k.startswith("c_") else 4 * P)] if k.endswith(("_adv", "_inst")) else v) for k, v in base.items()}
cd, vd = torch.from_numpy(wit["c_adv"]).cuda(), torch.from_numpy(wit["v_adv"]).cuda()
seed = bytes(range(32))
ref = None
CONFIGS = [
    {}, {"TB_Q_ROWS": "2"}, {"TB_Q_ROWS": "2", "TB_Q_THREADS": "128"}, {"TB_Q_THREADS": "64"}, {"TB_Q_PARTS": "16"}, {"TB_Q_PARTS": "16", "TB_Q_ROWS": "2"}, {"TB_Q_PARTS": "4"},
    {"TB_MSM_BA_MIN_TERMS": str(1 << 30)}, {"TB_MSM_BA_ROUNDS": "9"}, {"TB_MSM_BA_ROUNDS": "10"}, {"TB_MSM_BA_ROUNDS": "13"}, {"TB_MSM_BA_CHUNK": "128"}, {"TB_MSM_BA_CHUNK": "512"},
    {"TB_NTT_TILE_LOG": "11"}, {},
]
extra = os.environ.get("TUNE_EXTRA")
if extra:
    CONFIGS = [json.loads(x) for x in extra.split(";")]
for cfg in CONFIGS:
    for k, v in cfg.items():
        os.environ[k] = v
    try:
        out = svc.build_ptx_batch(wit, seed, cd, vd)
        torch.cuda.synchronize()
        t0 = time.time()
        for _ in range(2):
            out = svc.build_ptx_batch(wit, seed, cd, vd)
        torch.cuda.synchronize()
        dt = (time.time() - t0) / 2
        svc.prof_enable(True); svc.serial = True
        svc.build_ptx_batch(wit, seed, cd, vd)
        pr = svc.prof_read(); svc.prof_enable(False); svc.serial = False
        same = None if ref is None else (out == ref)
        if ref is None:
            ref = out
        print("%-60s %7.2f ptx/s  same=%s  %s" % (json.dumps(cfg), P / dt, same, " ".join("%s=%.0f" % (k[:9], v[0]) for k, v in sorted(pr.items(), key=lambda kv: -kv[1][0])[:6])), flush=True)
    except Exception as ex:
        print("%-60s FAILED %r" % (json.dumps(cfg), ex), flush=True)
    for k in cfg:
        os.environ.pop(k)

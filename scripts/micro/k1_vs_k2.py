"""Crossover between the warp-per-pair kernel (K1) and the CTA-per-pair kernel (K2) for small
batches: time of the whole prefix_error_rates call, device-resident, per (P, R)."""
import os, sys, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    for p in (ROOT, os.path.join(ROOT, "pydrobert-pytorch_b200")):
        sys.path.insert(0, p)
    import numpy as np, torch, time
    import b200lev.functional as F
    rng = np.random.default_rng(0)
    res = {}
    for P in (32, 128, 512, 1024, 2048):
        for R in (64, 101, 200, 400, 800, 1600):
            tok = lambda n: torch.from_numpy(rng.integers(1, 1000, size=(R, n), dtype=np.int64)).cuda()
            ref, hyp = tok(P), tok(P)
            costs = dict(ins_cost=1.0, del_cost=2.0, sub_cost=1.0)  # non-uniform: wavefront kernels
            f = (lambda: F.edit_distance(ref, hyp, warn=False, **costs)) if os.environ.get("K12_FINAL") else (lambda: F.prefix_edit_distances(ref, hyp, warn=False, **costs))
            for _ in range(5): f()
            torch.cuda.synchronize(); t0 = time.perf_counter()
            n = 30
            for _ in range(n): f()
            torch.cuda.synchronize()
            res[f"{P}x{R}"] = (time.perf_counter() - t0) / n * 1e6
    print(json.dumps(res))
else:
    out = {}
    for k in ("0", "1"):
        env = dict(os.environ, B200LEV_CTA_KERNEL=k)
        r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
        out[k] = json.loads(r.stdout.strip().splitlines()[-1])
    print("PxR        K1(us)   K2(us)")
    for key in out["0"]:
        print(f"{key:10s} {out['0'][key]:8.1f} {out['1'][key]:8.1f}  {'K2' if out['1'][key] < out['0'][key] else 'K1'}")

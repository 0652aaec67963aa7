"""Timing experiment (needs a build with MTN_SCAN_ABLATIONS=1): per-CTA start/end times of one scan launch."""
import ctypes, os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from avse_challenge_b200 import CONFIGS, ops, _lib
hp = CONFIGS["S"]; di, R = hp.d_inner, hp.dt_rank; nd = ops.n_dbl_for(R); B, L = 32, 3999; M = B * L
g = torch.Generator(device="cuda").manual_seed(0)
u = (torch.randn(2, M, 2 * di, device="cuda", generator=g) * 0.5).to(torch.bfloat16)
dbl = torch.randn(M, 2 * nd, device="cuda", generator=g) * 0.5
xz = torch.randn(M, 2 * di, device="cuda", generator=g)
w_dt = torch.randn(2, di, R, device="cuda", generator=g) * R ** -0.5
dt_bias = torch.randn(2, di, device="cuda", generator=g) * 0.5 - 3.0
A2 = -torch.exp(torch.randn(2, di, 16, device="cuda", generator=g) * 0.5 + 0.5) * ops.LOG2E
Dk = torch.randn(2, di, device="cuda", generator=g)
y = torch.empty_like(u); hout = torch.zeros(2, B, di, 16, device="cuda")
lib = _lib.load()
for mode in sys.argv[1:] or ["full", "S"]:
    for _ in range(3):
        if mode == "S":
            ops.scan(u, dbl, xz, di, w_dt, dt_bias, A2, Dk, B, L, di, R, h_out=hout, summary_only=True)
        else:
            os.environ["MTN_SCAN_VARIANT"] = "0" if mode == "full" else mode
            ops.scan(u, dbl, xz, di, w_dt, dt_bias, A2, Dk, B, L, di, R, y=y)
    torch.cuda.synchronize()
    n = 256
    buf = (ctypes.c_ulonglong * (3 * n))()
    lib.mtn_debug_scan_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
    assert lib.mtn_debug_scan_times(buf, n) == 0
    a = np.array(buf, dtype=np.uint64).reshape(n, 3).astype(np.int64)
    t0 = a[:, 1].min()
    dur = (a[:, 2] - a[:, 1]) / 1e3
    per_sm = collections.defaultdict(list)
    for i in range(n):
        per_sm[int(a[i, 0])].append(i)
    print(f"== {mode}: kernel span {(a[:,2].max()-t0)/1e3:.0f} us; CTA duration us: min {dur.min():.0f} median {np.median(dur):.0f} max {dur.max():.0f}")
    print("   CTAs per SM histogram:", collections.Counter(len(v) for v in per_sm.values()), "SMs used", len(per_sm))
    late = np.argsort(-(a[:, 2] - t0))[:8]
    for i in late:
        print(f"   cta {i:3d} (x={i%8} b={i//8}) sm {a[i,0]:3d} start {(a[i,1]-t0)/1e3:7.0f} end {(a[i,2]-t0)/1e3:7.0f} dur {dur[i]:6.0f}  co-resident {per_sm[int(a[i,0])]}")
    solo = [dur[v[0]] for v in per_sm.values() if len(v) == 1]
    duo = [dur[i] for v in per_sm.values() if len(v) == 2 for i in v]
    print(f"   mean duration alone on SM {np.mean(solo) if solo else float('nan'):.0f} us, paired {np.mean(duo) if duo else float('nan'):.0f} us; fwd mean {dur[[i for i in range(n) if i%8<4]].mean():.0f} bwd mean {dur[[i for i in range(n) if i%8>=4]].mean():.0f}")

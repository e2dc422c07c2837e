import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "graphical-gan_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import gpu_util as U
from gg import cabi
np.set_printoptions(precision=4, suppress=True, linewidth=200)
for (a_mn, b_mn, N, K) in [(1, 0, 64, 32), (0, 1, 64, 32), (1,1,64,32)]:
    rs = np.random.RandomState(1)
    A = rs.randn(128, K).astype(np.float32)
    Bm = rs.randn(K, N).astype(np.float32)
    A_st = np.ascontiguousarray(A.T) if a_mn else A
    B_st = Bm if b_mn else np.ascontiguousarray(Bm.T)
    D = torch.full((128, N), float("nan"), device="cuda")
    Ad, Bd = U.dev(A_st), U.dev(B_st)
    cabi.call("gg_probe_umma_tf32", cabi.ptr(Ad), cabi.ptr(Bd), cabi.ptr(D), N, K, a_mn, b_mn, 0, cabi.stream_ptr())
    torch.cuda.synchronize()
    got = D.cpu().numpy().astype(np.float64)
    exact = A.astype(np.float64) @ Bm.astype(np.float64)
    print("case", a_mn, b_mn, N, K, "norm got", np.linalg.norm(got), "norm exact", np.linalg.norm(exact), "nan", np.isnan(got).sum())
    print(" got[0,:8]  ", got[0, :8]); print(" exact[0,:8]", exact[0, :8])
    print(" got[33,:8] ", got[33, :8]); print(" exact[33,:8]", exact[33, :8])
    print(" corr", (got * exact).sum() / (np.linalg.norm(got) * np.linalg.norm(exact) + 1e-30))
    # hypotheses: single k contributes
    for k in range(K):
        h = np.outer(A[:, k], Bm[k, :])
        c = (got * h).sum() / (np.linalg.norm(got) * np.linalg.norm(h) + 1e-30)
        if abs(c) > 0.3: print("  k", k, "corr", c)
    # which rows/cols non-zero
    nzr = np.where(np.abs(got).sum(1) > 1e-6)[0]; nzc = np.where(np.abs(got).sum(0) > 1e-6)[0]
    print(" nonzero rows", nzr[:40], len(nzr), " cols", nzc[:40], len(nzc))

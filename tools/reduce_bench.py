"""Micro-benchmark of chk_group_build + chk_reduce_apply on synthetic slot lists (run on the GPU box)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complexhyperbolickge_b200 import ops  # noqa: E402


def run(tag, N, Bq, P, w, scalars, hot, reps=20, coef=False, world=1):
    g = torch.Generator(device="cuda").manual_seed(0)
    S = Bq + P
    ids = torch.randint(0, N, (S,), generator=g, device="cuda")
    if hot:                                                   # Zipf-like heads / positive tails: a few rows named by ~100 slots
        pop = (1.0 / torch.arange(1, N + 1, device="cuda", dtype=torch.float64))
        hot_ids = torch.multinomial(pop / pop.sum(), 2 * Bq, replacement=True, generator=g)
        ids[:Bq] = hot_ids[:Bq]
        ids[Bq:2 * Bq] = hot_ids[Bq:]
    a_rows = torch.randn(Bq, w, generator=g, device="cuda")
    b_rows = torch.randn(P, w, generator=g, device="cuda")
    sc = torch.randn(P, generator=g, device="cuda")
    sh = torch.randn(Bq, generator=g, device="cuda")
    param, ssum = torch.randn(N, w, device="cuda"), torch.rand(N, w, device="cuda")
    ps = [torch.randn(N, 1, device="cuda") for _ in range(2)]
    ss = [torch.rand(N, 1, device="cuda") for _ in range(2)]
    work = ops.group_workspace(N, S, "cuda")
    hyper = torch.tensor([0.01, 1e-10, 0, 0, 0, 0, 0, 0], dtype=torch.float64, device="cuda")
    if coef:                                                  # computed source: query rows + three scalars per pair (per rank: Bq / world queries)
        q_rows = torch.randn(Bq, w, generator=g, device="cuda") * 0.05
        cf = torch.randn(P, 4, generator=g, device="cuda")
        cols = [dict(param=param, state0=ssum, dense=None, src=[(a_rows, 0, Bq, 0), (q_rows, Bq, S, 0)], pair=(cf, P // Bq, 0))]
        tag += " [pair coef]"
    else:
        cols = [dict(param=param, state0=ssum, dense=None, src=[(a_rows, 0, Bq, 0), (b_rows, Bq, S, 0)])]
    if scalars:
        cols.append(dict(param=ps[0], state0=ss[0], dense=None, src=[(sh, 0, Bq, 0)]))
        cols.append(dict(param=ps[1], state0=ss[1], dense=None, src=[(sc, Bq, S, 0)]))
    groups = ops._red_groups([dict(ids=ids, n_keys=N, slots_per_rank=S, world=1, work=work, cols=cols)])
    loss_part = torch.zeros(8, device="cuda")
    loss = torch.zeros((), device="cuda")
    step = torch.ones((), dtype=torch.int32, device="cuda")
    tg, tr = [], []
    for i in range(reps + 3):
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        ops.group_build(ids, N, work)
        e[1].record()
        ops.reduce_apply(param, ops.CHK_OPT_ADAGRAD, groups, hyper, finish=(loss_part, loss, step))
        e[2].record()
        torch.cuda.synchronize()
        if i >= 3:
            tg.append(e[0].elapsed_time(e[1]) * 1e3)
            tr.append(e[1].elapsed_time(e[2]) * 1e3)
    hdr = work[:4].tolist()
    cnt = torch.bincount(ids, minlength=N)
    byt = (Bq * w * 4 + P * 16 if coef else S * w * 4) + int((cnt > 0).sum()) * w * 4 * 4
    print(f"{tag:46s} group {np.median(tg):7.1f} us  reduce {np.median(tr):7.1f} us  rows {int((cnt > 0).sum())}  max seg {int(cnt.max())}  "
          f"long(>32) {int((cnt > 32).sum())}  {byt / np.median(tr) / 1e3:7.1f} GB/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "coef":
        for c in (False, True):
            run("fb237 uniform, entity + bh + bt", 14541, 500, 125500, 66, True, False, coef=c)
            run("fb237 hot heads, entity + bh + bt", 14541, 500, 125500, 66, True, True, coef=c)
            run("big4m uniform, entity + bh + bt", 4_000_000, 500, 50500, 514, True, False, coef=c)
            run("big4m x2 ranks' slots (DP receive side)", 4_000_000, 1000, 101000, 514, True, False, reps=10, coef=c)
            run("big4m x8 ranks' slots (DP receive side)", 4_000_000, 4000, 404000, 514, True, False, reps=5, coef=c)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "bigone":
        run("big4m uniform, entity only", 4_000_000, 500, 50500, 514, False, False, reps=3)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "big":
        run("big4m uniform, entity only", 4_000_000, 500, 50500, 514, False, False)
        run("big4m uniform, entity + bh + bt", 4_000_000, 500, 50500, 514, True, False)
        run("big4m x8 ranks' slots (DP receive side)", 4_000_000, 4000, 404000, 514, True, False, reps=5)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        run("fb237 uniform, entity only", 14541, 500, 125500, 66, False, False, reps=4)
        sys.exit(0)
    run("fb237 uniform, entity only", 14541, 500, 125500, 66, False, False)
    run("fb237 uniform, entity + bh + bt", 14541, 500, 125500, 66, True, False)
    run("fb237 hot heads, entity only", 14541, 500, 125500, 66, False, True)
    run("fb237 hot heads, entity + bh + bt", 14541, 500, 125500, 66, True, True)
    run("big4m uniform, entity only", 4_000_000, 500, 50500, 514, False, False)
    run("big4m uniform, entity + bh + bt", 4_000_000, 500, 50500, 514, True, False)
    run("big4m hot heads, entity + bh + bt", 4_000_000, 500, 50500, 514, True, True)
    run("big4m x8 ranks' slots (DP receive side)", 4_000_000, 4000, 404000, 514, True, False, reps=5)

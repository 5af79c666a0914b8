"""Per-kernel roofline probes (run on the GPU box): K1 query transform fwd/bwd at a large query count,
K3 gather scoring fwd/bwd at the training shapes.  Algorithmic bytes as defined in DESIGN.md; CUDA-event
timing on torch's current stream (the stream the C ABI launches on), L2 flushed between repetitions."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from complexhyperbolickge_b200 import ops  # noqa: E402

PEAK = 6530.0
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    PEAK = json.load(open(pk)).get("hbm_gbs", PEAK)
FLUSH = None


def timeit(fn, reps=5):
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def k1(kind, name, rank, nq, n_ent, n_rel2, dtype):
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 2 * (rank - 1)
    es = dtype.itemsize
    ent = torch.randn(n_ent, 2 * rank, generator=g, device="cuda", dtype=dtype) * float(np.sqrt(0.4 / (2 * rank)))
    rel = torch.randn(n_rel2, 2 * n, generator=g, device="cuda", dtype=dtype) * 0.05
    rd = torch.rand(n_rel2, 2 * n if kind == ops.CHK_ATT else n, generator=g, device="cuda", dtype=dtype) * 2 - 1
    ctx = torch.randn(n_rel2, n, generator=g, device="cuda", dtype=dtype) if kind == ops.CHK_ATT else None
    c = torch.rand(n_rel2, 1, generator=g, device="cuda", dtype=dtype) + 0.5
    h = torch.randint(0, n_ent, (nq,), generator=g, device="cuda")
    r = torch.randint(0, n_rel2, (nq,), generator=g, device="cuda")
    ms = timeit(lambda: ops.query_fwd(kind, rank, True, ent, rel, rd, ctx, c, h, r, grouped=False))
    byt = nq * (2 * 2 * rank * es + 16)                       # entity row in, query row out, ids
    ms_s = float("nan")
    if dtype == torch.float32 and rank <= 33 and nq >= 4096:  # thread-per-query variant incl. its argsort by relation
        ms_s = timeit(lambda: ops.query_fwd(kind, rank, True, ent, rel, rd, ctx, c, h, r, grouped=True))
    gq = torch.randn(nq, 2 * rank, generator=g, device="cuda", dtype=dtype)
    msb = timeit(lambda: ops.query_bwd(kind, rank, True, ent, rel, rd, ctx, c, h, r, gq))
    wrel = 2 * n + (2 * n if kind == ops.CHK_ATT else n) + (n if kind == ops.CHK_ATT else 0) + 1
    bytb = nq * ((2 * 2 * rank + 2 * rank + wrel) * es + 16)  # entity row + grad_q in, grad rows out
    print(f"K1 {name:8s} r={rank:3d} {str(dtype)[6:]:7s} nq={nq}: fwd {ms:7.3f} ms {byt / ms / 1e6:7.0f} GB/s ({byt / ms / 1e6 / PEAK:5.1%})"
          f" [thread-per-query + argsort by relation {ms_s:7.3f} ms {byt / ms_s / 1e6 / PEAK:5.1%}]"
          f"   bwd {msb:7.3f} ms {bytb / msb / 1e6:7.0f} GB/s ({bytb / msb / 1e6 / PEAK:5.1%})", flush=True)


def k3(rank, B, nt, n_ent, dtype):
    g = torch.Generator(device="cuda").manual_seed(0)
    es = dtype.itemsize
    ent = torch.randn(n_ent, 2 * rank, generator=g, device="cuda", dtype=dtype) * float(np.sqrt(0.4 / (2 * rank)))
    q = torch.randn(B, 2 * rank, generator=g, device="cuda", dtype=dtype) * float(np.sqrt(0.4 / (2 * rank)))
    bh = torch.randn(B, generator=g, device="cuda", dtype=dtype) * 0.1
    bt = torch.randn(n_ent, generator=g, device="cuda", dtype=dtype) * 0.1
    tails = torch.randint(0, n_ent, (B, nt), generator=g, device="cuda")
    ms = timeit(lambda: ops.score_gather_fwd(rank, B, nt, q, 1, 0, ent, tails, 0, bh, 1, 0, bt))
    byt = B * nt * (2 * rank * es + 8 + 2 * es) + B * 2 * rank * es
    gs = torch.randn(B, nt, generator=g, device="cuda", dtype=dtype)
    msb = timeit(lambda: ops.score_gather_bwd(rank, B, nt, q, 1, 0, ent, tails, 0, gs))
    bytb = B * nt * (2 * 2 * rank * es + 8 + es) + 2 * B * 2 * rank * es
    print(f"K3 r={rank:3d} {str(dtype)[6:]:7s} B={B} nt={nt} N={n_ent}: fwd {ms:7.3f} ms {byt / ms / 1e6:7.0f} GB/s ({byt / ms / 1e6 / PEAK:5.1%})"
          f"   bwd {msb:7.3f} ms {bytb / msb / 1e6:7.0f} GB/s ({bytb / msb / 1e6 / PEAK:5.1%})", flush=True)


if __name__ == "__main__":
    print("HBM peak used:", PEAK, "GB/s")
    for kind, name in ((ops.CHK_ROT, "FFTRotH"), (ops.CHK_REF, "FFTRefH"), (ops.CHK_ATT, "FFTAttH")):
        k1(kind, name, 33, 1 << 20, 1 << 20, 474, torch.float32)
    k1(ops.CHK_ROT, "FFTRotH", 65, 1 << 20, 1 << 20, 22, torch.float64)
    k1(ops.CHK_ROT, "FFTRotH", 257, 1 << 19, 1 << 20, 2000, torch.float32)
    k1(ops.CHK_ROT, "FFTRotH", 33, 500, 40943, 22, torch.float32)
    k3(33, 500, 101, 40943, torch.float32)
    k3(33, 500, 251, 14541, torch.float32)
    k3(33, 8192, 101, 123182, torch.float32)
    k3(257, 500, 101, 4_000_000, torch.float32)
    k3(257, 4096, 101, 4_000_000, torch.float32)
    k3(65, 500, 101, 40943, torch.float64)

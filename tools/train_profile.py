"""Training-step timeline helper (run on the GPU box, plain or under ncu): a few eager fused steps of a BASELINE shape so every
kernel of the chain is a separate launch, then the CUDA-graph step time for comparison.
    python tools/train_profile.py fb237|wn18rr|yago310|big4m [steps]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from complexhyperbolickge_b200 import synthetic  # noqa: E402
from complexhyperbolickge_b200.optim import N3  # noqa: E402
from complexhyperbolickge_b200.train import FusedKGOptimizer  # noqa: E402

CFG = {"fb237": ("FFTRefH", 33, "Adagrad", 0.02, 250, False), "wn18rr": ("FFTRotH", 33, "Adam", 3e-4, 100, True),
       "yago310": ("FFTAttH", 33, "Adagrad", 0.02, 100, False), "big4m": ("FFTRotH", 257, "Adagrad", 0.02, 100, False)}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "fb237"
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    name, rank, opt_name, lr, neg, dn = CFG[wl]
    dev = torch.device("cuda", 0)
    graph = synthetic.make_graph(wl, seed=0, n_train=200_000 if wl == "big4m" else None)
    model = bench.make_model(name, rank, "float", graph, dev)
    mk = (lambda ps: torch.optim.Adagrad(ps, lr=lr)) if opt_name == "Adagrad" else (lambda ps: torch.optim.Adam(ps, lr=lr))
    ex = synthetic.train_examples(graph)
    batches = bench.cycle_batches(ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))], 64, 500).to(dev)
    pc = False if os.environ.get("CHK_PAIR_COEF") == "0" else None      # A/B: stored tail-gradient rows instead of pair coefficients
    for use_graph in (False, True):
        opt = FusedKGOptimizer(model, N3(0.0), mk(model.parameters()), 500, 1, neg, dn, verbose=False, use_cuda_graph=use_graph, pair_coef=pc)
        n = steps if not use_graph else 60
        for i in range(3):
            opt.fused_step(batches[i])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(n):
            opt.fused_step(batches[(3 + i) % 64])
        e1.record()
        torch.cuda.synchronize()
        print(f"{wl} pair_coef={pc is None} graph={use_graph}: {e0.elapsed_time(e1) / n * 1e3:.1f} us/step (device), {(time.perf_counter() - t0) / n * 1e6:.1f} us/step (host)", flush=True)
        if os.environ.get("CHK_PROFILE_EAGER_ONLY"):
            break


if __name__ == "__main__":
    main()

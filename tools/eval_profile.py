"""Evaluation-step helper for ncu: a few 500-query filtered-ranking batches of a BASELINE shape through model.get_ranking.
    python tools/eval_profile.py big4m|wn18rr|fb237|yago310 [batches]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from complexhyperbolickge_b200 import synthetic  # noqa: E402

CFG = {"big4m": ("FFTRotH", 257, "float"), "wn18rr": ("FFTRotH", 33, "float"), "fb237": ("FFTRefH", 33, "float"),
       "yago310": ("FFTAttH", 33, "float"), "wn18rr_d": ("FFTRotH", 65, "double")}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else "wn18rr"
    nb = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    name, rank, dtype = CFG[wl]
    dev = torch.device("cuda", 0)
    graph = synthetic.make_graph(wl.split("_")[0], seed=0, n_train=200_000 if wl == "big4m" else None)
    model = bench.make_model(name, rank, dtype, graph, dev)
    model.eval()
    q = torch.from_numpy(graph["test"][:500 * nb])
    for _ in range(2):
        r = model.get_ranking(q, graph["filters"]["rhs"], batch_size=500)
    torch.cuda.synchronize()
    print(wl, "mean rank", r.mean().item())


if __name__ == "__main__":
    main()

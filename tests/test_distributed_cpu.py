"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU paths — sparse row-gradient
exchange of data-parallel training and the sharded-ranking count reduction.  No kernels run here (the product has
no CPU path); the per-shard scores of the ranking test come from the oracle, which tests may use."""
import os
import socket
from argparse import Namespace

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(fn, world, *args):
    port = _free_port()
    mp.spawn(_entry, args=(world, port, fn, args), nprocs=world, join=True)


def _entry(rank, world, port, fn, args):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        fn(rank, world, *args)
    finally:
        dist.destroy_process_group()


# ----------------------------------------------------------------------------------------------- DP training
def _w_exchange(rank, world):
    from complexhyperbolickge_b200 import parallel
    N, w = 97, 6
    contribs, rows_all = [], []
    for k in range(world):                      # every rank can rebuild everyone's contribution (seeded)
        g = torch.Generator().manual_seed(100 + k)
        rows = torch.randint(0, N, (23,), generator=g)
        dense = torch.zeros(N, w, dtype=torch.float64)
        dense.index_add_(0, rows, torch.randn(23, w, generator=g, dtype=torch.float64))
        contribs.append(dense)
        rows_all.append(rows)
    mine = contribs[rank].clone()
    out = parallel.exchange_sparse_rows(mine, rows_all[rank], None)
    want = sum(contribs) / world
    assert torch.allclose(out, want, rtol=0, atol=1e-15), (out - want).abs().max()
    # bit-identical on every rank (deterministic rank-order accumulation)
    gathered = [torch.empty_like(out) for _ in range(world)]
    dist.all_gather(gathered, out)
    assert all(torch.equal(gathered[0], t) for t in gathered)
    # empty contribution on one rank
    e = torch.zeros(N, w, dtype=torch.float64)
    if rank == 0:
        e[5] = 1.0
    out = parallel.exchange_sparse_rows(e, torch.tensor([5] if rank == 0 else [], dtype=torch.int64), None)
    assert out[5, 0].item() == 1.0 / world and out.abs().sum().item() == w / world


def test_sparse_gradient_exchange_world2():
    _run(_w_exchange, 2)


def _w_reduce_gradients(rank, world):
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import parallel
    args = Namespace(sizes=(50, 6, 50), rank=9, dropout=0, gamma=0, dtype="double", bias="learn", init_size=1e-3,
                     multi_c=True)
    torch.manual_seed(0)
    model = chk.FFTAttH(args)                   # CPU module: parameters only, no kernel is called
    full = {}
    for k in range(world):
        g = torch.Generator().manual_seed(7 + k)
        rows = torch.randint(0, 50, (11,), generator=g)
        grads = {}
        for name, p in model.named_parameters():
            d = torch.zeros_like(p)
            if name.split(".")[0] in parallel.SPARSE_TABLES:
                d.index_add_(0, rows, torch.randn((11,) + tuple(p.shape[1:]), generator=g, dtype=p.dtype))
            else:
                d.copy_(torch.randn(p.shape, generator=g, dtype=p.dtype))
            grads[name] = d
        full[k] = (rows, grads)
    rows, grads = full[rank]
    for name, p in model.named_parameters():
        p.grad = grads[name].clone()
    parallel.reduce_gradients(model, {"entity": rows, "bh": rows, "bt": rows}, None)
    for name, p in model.named_parameters():
        want = sum(full[k][1][name] for k in range(world)) / world
        assert torch.allclose(p.grad, want, rtol=0, atol=1e-14), name


def test_reduce_gradients_world2():
    _run(_w_reduce_gradients, 2)


# ----------------------------------------------------------------------------------------------- sharded ranking
def _w_sharded_ranking(rank, world):
    """Each rank counts over its entity shard (oracle scores stand in for the kernels), counts are summed with
    all_reduce: ranks must equal the single-process oracle ranking exactly, for any world size."""
    from conftest import golden_files, load_case, oracle_params, filters_from_arrays
    from complexhyperbolickge_b200.filters import FilterIndex
    from complexhyperbolickge_b200.ranking import shard_bounds
    from oracle import chk_oracle as O
    case = load_case(golden_files("rank_FFTRotH_double_trained")[0])
    p = oracle_params(case)
    filters = filters_from_arrays(case)["rhs"]
    qs = torch.from_numpy(case["test"][:40])
    want = O.get_ranking(p, qs, filters, batch_size=16)
    n_ent = case["n_ent"]
    lo, hi = shard_bounds(n_ent, world, rank)
    fi = FilterIndex.from_dict(filters, case["n_rel2"])
    indptr, idx = fi.batch_csr(qs.numpy())
    q, _ = O.query_fwd(p, qs[:, 0], qs[:, 1])
    bh_vals = p.bh[qs[:, 0]].unsqueeze(1)
    scores = O.score_all(p, q.unsqueeze(1), bh_vals).squeeze(-1)         # (b, N) oracle scores
    target = O.score_pairs(p, q.unsqueeze(1), bh_vals, qs[:, 2:3]).reshape(-1)
    counts = torch.zeros(len(qs), dtype=torch.int64)
    for i in range(len(qs)):
        s = scores[i, lo:hi]
        c = int((s >= target[i]).sum())
        f = idx[indptr[i]:indptr[i + 1]]
        f = f[(f >= lo) & (f < hi)]
        c -= int((scores[i, torch.from_numpy(f)] >= target[i]).sum()) if len(f) else 0
        counts[i] = c
    dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    assert torch.equal((counts + 1).float(), want), ((counts + 1).float() - want).abs().max()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ranking_counts(world):
    _run(_w_sharded_ranking, world)

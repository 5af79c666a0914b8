"""GPU, two ranks: data-parallel fused training (parallel.FusedDataParallelKGOptimizer) over a NON-divisible example count
(ragged last batch, odd global batch) equals the single-GPU fused epoch on the same shuffle and negatives, on both gradient
paths (sparse row exchange with the one-kernel receive side; dense all_reduce), replicas bit-identical; and the sharded
ranking equals the single-GPU ranking.  With >= 2 GPUs the ranks use NCCL (CUDA graph with the collectives inside); on a
one-GPU box both ranks share the GPU over gloo (host-synchronised collectives — no kernel waits on another rank — eager)."""
import os
import socket
from argparse import Namespace

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _hash_negs(batch, neg, n_ent):
    """Negatives as a function of the triple alone, so a rank's rows get the ids the single-GPU run gives the same rows."""
    j = torch.arange(1, neg + 1, device=batch.device).unsqueeze(0)
    v = (batch[:, 0:1] * 7919 + batch[:, 1:2] * 104729 + batch[:, 2:3] * 31 + j * 611953) % (n_ent - 1)
    return torch.where(v < batch[:, 2:3], v, v + 1)


def _mk_model(dtype, n_ent, name="FFTRotH"):
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    args = Namespace(sizes=(n_ent, 12, n_ent), rank=33, dropout=0, gamma=0, dtype=dtype, bias="learn", init_size=1e-3, multi_c=True)
    m = getattr(chk, name)(args).cuda()
    synthetic.trained_like_(m, 0)
    return m


def _examples(n_ent, n):
    g = torch.Generator().manual_seed(9)
    return torch.stack([torch.randint(0, n_ent, (n,), generator=g) % 50, torch.randint(0, 12, (n,), generator=g),
                        torch.randint(0, n_ent, (n,), generator=g)], 1)


def _worker(rank, world, port, backend, mode, opt_name, q):
    sparse = mode != "dense"
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    ngpu = torch.cuda.device_count()
    torch.cuda.set_device(rank % ngpu)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from complexhyperbolickge_b200.optim import N3
        from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer
        n_ent, neg, Bg = 600, 15, 45                      # odd global batch: ranks get 23 / 22 rows; 230 examples: ragged tail of 5
        m = _mk_model("double", n_ent)
        mk = (lambda ps: torch.optim.Adagrad(ps, lr=0.05)) if opt_name == "Adagrad" else (lambda ps: torch.optim.Adam(ps, lr=1e-3))

        class Fed(FusedDataParallelKGOptimizer):
            def get_neg_samples(self, b):
                return _hash_negs(b, neg, n_ent)
        opt = Fed(m, N3(0.0), mk(m.parameters()), Bg, 1, neg, False, verbose=False, process_group=None, sparse_exchange=sparse,
                  owner_sharded=None if mode == "owner" else False, use_cuda_graph=(backend == "nccl"))
        assert opt.world == world and opt.local_batch_size == 23
        assert opt.owner_sharded == (mode == "owner" and backend == "nccl")      # NCCL: peer-memory tables, update sharded by owner
        ex = _examples(n_ent, 230)
        torch.manual_seed(21 + rank)                      # ranks shuffle differently: rank 0's permutation must win
        losses = [opt.epoch(ex), opt.epoch(ex)]
        params = {k: p.detach().cpu().numpy() for k, p in m.named_parameters()}   # by value: this process exits before the parent reads
        # replicas must hold identical bits
        for k, p in m.named_parameters():
            mine = p.detach().clone()
            other = mine.clone()
            dist.broadcast(other, src=0)
            assert torch.equal(mine, other), f"replicas diverged in {k}"
        # sharded ranking == unsharded ranking on the trained weights
        filters = {}
        for h, r, t in ex.tolist():
            filters.setdefault((h, r), []).append(t)
        m.process_group = dist.group.WORLD
        sharded = m.get_ranking(ex[:64], filters, batch_size=20)
        m.process_group = None
        m.release_eval_cache()
        single = m.get_ranking(ex[:64], filters, batch_size=20)
        assert torch.equal(sharded, single)
        if opt.owner_sharded:                             # a bare step() leaves this rank's copy stale: evaluation must refuse
            opt.step(ex[:Bg])
            try:
                m.get_ranking(ex[:8], filters, batch_size=8)
                raise AssertionError("get_ranking ran on stale owner-sharded replicas")
            except RuntimeError as e:
                assert "sync_replicas" in str(e)
            opt.sync_replicas()
            m.get_ranking(ex[:8], filters, batch_size=8)
        if rank == 0:
            q.put((losses, params))
        # the captured CUDA graph holds NCCL work: release it before the communicator is torn down (destroy_process_group
        # waits for it otherwise)
        del opt
        import gc
        gc.collect()
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["owner", "sparse", "dense"])
@pytest.mark.parametrize("opt_name", ["Adagrad", "Adam"])
def test_dp_epoch_equals_single_gpu_epoch(mode, opt_name):
    """mode: owner = sparse exchange with owner-sharded tables in symmetric memory (NCCL only; under gloo it is the replicated
    sparse exchange), sparse = every replica applies every update, dense = dense all_reduce of the gradients."""
    sparse = mode != "dense"
    from complexhyperbolickge_b200.optim import N3
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    if sparse and opt_name == "Adam":
        pytest.skip("the sparse row exchange is Adagrad only")
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, backend, mode, opt_name, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        got = q.get(timeout=150)
        for p in procs:
            p.join(timeout=60)
            assert p.exitcode == 0, "a rank did not exit cleanly"
    finally:
        for p in procs:                                   # never leave a rank behind (it would hold the GPU and the port)
            if p.is_alive():
                p.kill()
    losses_dp, params_dp = got
    # single GPU, same shuffle (rank 0's seed) and the same per-triple negatives
    n_ent, neg, Bg = 600, 15, 45
    m = _mk_model("double", n_ent)
    mk = (lambda ps: torch.optim.Adagrad(ps, lr=0.05)) if opt_name == "Adagrad" else (lambda ps: torch.optim.Adam(ps, lr=1e-3))

    class Fed(FusedKGOptimizer):
        def get_neg_samples(self, b):
            return _hash_negs(b, neg, n_ent)
    opt = Fed(m, N3(0.0), mk(m.parameters()), Bg, 1, neg, False, verbose=False)
    ex = _examples(n_ent, 230)
    torch.manual_seed(21)
    losses = [opt.epoch(ex), opt.epoch(ex)]
    assert max(abs(a - b) for a, b in zip(losses, losses_dp)) <= 1e-10, (losses, losses_dp)
    for k, p in m.named_parameters():
        err = (p.detach().cpu() - torch.from_numpy(params_dp[k])).abs().max().item() / max(p.abs().max().item(), 1e-30)
        assert err <= 1e-9, (k, err)

"""Multi-GPU check (torchrun, one rank per GPU, NCCL): (1) entity-table-sharded get_ranking equals the
single-GPU ranks exactly; (2) a data-parallel training step (sparse row-gradient exchange) matches the
single-GPU step on the same global batch and negatives.  Prints one line per check on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py
"""
import os
import sys
from argparse import Namespace

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import complexhyperbolickge_b200 as chk  # noqa: E402
from complexhyperbolickge_b200 import synthetic  # noqa: E402
from complexhyperbolickge_b200.optim import N3  # noqa: E402
from complexhyperbolickge_b200.parallel import DataParallelKGOptimizer  # noqa: E402


class FixedNegs(DataParallelKGOptimizer):
    negs_global = None

    def get_neg_samples(self, input_batch):
        self._negs = self.negs_global[self.rank_id::self.world].to(input_batch.device)
        return self._negs


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    for name, r, dtype, algo in (("FFTRotH", 33, "float", "mma"), ("FFTAttH", 33, "double", "fma"), ("FFTRefH", 65, "float", "mma")):
        n_ent, n_rel2, nq = 30_011, 12, 400
        args = Namespace(sizes=(n_ent, n_rel2, n_ent), rank=r, dropout=0, gamma=0, dtype=dtype, bias="learn",
                         init_size=1e-3, multi_c=True)
        model = getattr(chk, name)(args).to(dev)
        synthetic.trained_like_(model, 0)                     # same seed -> identical replicas
        g = torch.Generator().manual_seed(1)
        qs = torch.stack([torch.randint(0, n_ent, (nq,), generator=g), torch.randint(0, n_rel2, (nq,), generator=g),
                          torch.randint(0, n_ent, (nq,), generator=g)], 1)
        filters = {(int(h), int(rr)): [int(t), int((3 * t + 1) % n_ent)] for h, rr, t in qs.numpy()}
        model.rank_algo = algo
        model.process_group = None
        single = model.get_ranking(qs, filters, batch_size=128)
        model.process_group = dist.group.WORLD
        sharded = model.get_ranking(qs, filters, batch_size=128)
        ok = torch.equal(single, sharded)
        if rank == 0:
            print(f"[ranking {name} r={r} {dtype} {algo} x{world}] sharded == single: {ok}  mean rank {single.mean().item():.2f}", flush=True)
        assert ok
        # ---- DP training step vs single-GPU step
        model.process_group = None
        B, neg = 64 * world, 50
        batch = torch.stack([torch.randint(0, n_ent, (B,), generator=g), torch.randint(0, n_rel2, (B,), generator=g),
                             torch.randint(0, n_ent, (B,), generator=g)], 1)
        negs = torch.randint(0, n_ent, (B, neg), generator=g)
        ref = getattr(chk, name)(args).to(dev)
        ref.load_state_dict(model.state_dict())
        o1 = FixedNegs(ref, N3(0.0), torch.optim.Adagrad(ref.parameters(), lr=0.05), B, 1, neg, False, verbose=False,
                       process_group=None)
        o1.world, o1.rank_id = 1, 0
        o1.negs_global = negs
        l1 = o1.step(batch)
        o2 = FixedNegs(model, N3(0.0), torch.optim.Adagrad(model.parameters(), lr=0.05), B, 1, neg, False, verbose=False,
                       process_group=dist.group.WORLD)
        o2.negs_global = negs
        l2 = o2.step(batch)
        lsum = l2.clone().double()
        dist.all_reduce(lsum)
        tol = 1e-12 if dtype == "double" else 2e-5
        worst = 0.0
        for (k, a), (_, b) in zip(ref.named_parameters(), model.named_parameters()):
            scale = max(a.detach().abs().max().item(), 1e-30)
            worst = max(worst, (a.detach() - b.detach()).abs().max().item() / scale)
        if rank == 0:
            print(f"[dp step {name} r={r} {dtype} x{world}] loss single {l1.item():.9f} mean-of-ranks {lsum.item() / world:.9f}  "
                  f"max |param diff| / max|param| after one Adagrad step = {worst:.2e}", flush=True)
        assert abs(l1.item() - lsum.item() / world) < 1e-4 and worst < (1e-9 if dtype == "double" else 1e-4)
        # ---- fused data-parallel step (CUDA graph + dense/sparse exchange + row-sparse Adagrad) vs single-GPU fused step
        from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer
        from complexhyperbolickge_b200.train import FusedKGOptimizer

        class FixedF(FusedKGOptimizer):
            def get_neg_samples(self, input_batch):
                return self._negs_static

        class FixedFDP(FusedDataParallelKGOptimizer):
            def get_neg_samples(self, input_batch):
                return self._negs_static

        # dense all_reduce of the small tables; sparse exchange with every replica applying every update; sparse exchange with
        # owner-sharded tables in symmetric memory (the big-table path under NCCL)
        for sparse, owner, tag in ((None, False, "dense"), (True, False, "sparse rows"), (True, None, "owner-sharded")):
            ref.load_state_dict(model.state_dict())
            f1 = FixedF(ref, N3(0.0), torch.optim.Adagrad(ref.parameters(), lr=0.05), B, 1, neg, False, verbose=False)
            f2 = FixedFDP(model, N3(0.0), torch.optim.Adagrad(model.parameters(), lr=0.05), B, 1, neg, False, verbose=False,
                          process_group=dist.group.WORLD, sparse_exchange=sparse, owner_sharded=owner)
            assert f2.owner_sharded == (tag == "owner-sharded")
            f1._negs_static = torch.zeros(B, neg, dtype=torch.int64, device=dev)
            f2._negs_static = torch.zeros(B // world, neg, dtype=torch.int64, device=dev)
            for it in range(4):
                batch = torch.stack([torch.randint(0, n_ent, (B,), generator=g), torch.randint(0, n_rel2, (B,), generator=g),
                                     torch.randint(0, n_ent, (B,), generator=g)], 1)
                batch[: B // 4, 0] = batch[0, 0]                   # a hot head and a hot tail: rows named by several ranks and
                batch[: B // 4, 2] = batch[1, 2]                   # several times per rank (claim / duplicate handling)
                batch = batch.to(dev)
                negs = torch.randint(0, n_ent, (B, neg), generator=g).to(dev)
                f1._negs_static.copy_(negs)
                f2._negs_static.copy_(negs[rank::world])
                f1.fused_step(batch)
                f2.step(batch)
            f2.sync_replicas()                                     # owner-sharded: every copy current again
            worst, replica = 0.0, 0.0
            for (k, a), (_, b) in zip(ref.named_parameters(), model.named_parameters()):
                worst = max(worst, (a.detach() - b.detach()).abs().max().item() / max(a.detach().abs().max().item(), 1e-30))
                lo, hi = b.detach().clone(), b.detach().clone()   # replicas must stay BIT-identical
                dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                replica = max(replica, (hi - lo).abs().max().item())
            lsum = f2._loss_sum.clone().double()
            dist.all_reduce(lsum)
            if rank == 0:
                print(f"[fused dp {name} r={r} {dtype} x{world} {tag}] 4 steps: loss single "
                      f"{f1._loss_sum.item() / 4:.9f} sum-over-ranks {lsum.item() / 4:.9f}  max |param diff| / max|param| = "
                      f"{worst:.2e}  max replica divergence = {replica:.1e}", flush=True)
            assert abs(f1._loss_sum.item() - lsum.item()) / 4 < (1e-10 if dtype == "double" else 1e-4)
            assert worst < (1e-9 if dtype == "double" else 5e-3)     # fp32: Adagrad normalises tiny early gradients
            assert replica == 0.0
            del f1, f2
        model.release_eval_cache()
    dist.barrier()
    if rank == 0:
        print("dp_check ok", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

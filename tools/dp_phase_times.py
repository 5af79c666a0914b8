"""Per-call device times of one data-parallel training step (torchrun, NCCL; eager, not graph-captured): every ops.* wrapper and
every torch.distributed collective of FusedDataParallelKGOptimizer.step is bracketed by CUDA events on its stream.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/dp_phase_times.py"""
import collections
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from complexhyperbolickge_b200 import ops, synthetic  # noqa: E402
from complexhyperbolickge_b200.optim import N3  # noqa: E402
from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer  # noqa: E402

REC = []


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn(*a, **k)
        e1.record()
        REC.append((name, e0, e1))
        return out
    setattr(mod, name, inner)


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    owner = None if os.environ.get("CHK_OWNER", "1") == "1" else False
    px = None if os.environ.get("CHK_PEER_EXCHANGE", "1") == "1" else False
    use_graph = os.environ.get("CHK_GRAPH", "0") == "1"
    graph = synthetic.make_graph("big4m", seed=0, n_train=200_000)
    model = bench.make_model("FFTRotH", 257, "float", graph, dev)
    opt = FusedDataParallelKGOptimizer(model, N3(0.0), torch.optim.Adagrad(model.parameters(), lr=0.02), 500 * world, 1, 100, False,
                                       verbose=False, process_group=dist.group.WORLD, use_cuda_graph=use_graph, owner_sharded=owner, peer_exchange=px,
                                       peer_dense=px)
    ex = synthetic.train_examples(graph)
    batches = bench.cycle_batches(ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))], 64, 500 * world).to(dev)
    for i in range(3):
        opt.step(batches[i])
    torch.cuda.synchronize()
    if not use_graph:
        for n in ("train_prep", "group_build", "query_fwd", "score_gather_train", "score_gather_train_peer", "peer_gather_rows", "query_bwd_into",
                  "reduce_apply", "dense_apply", "step_finish", "dp_all_gather", "dp_fused_apply"):
            wrap(ops, n)
    if not use_graph:
        for n in ("all_gather_into_tensor", "all_reduce"):
            wrap(dist, n)
    steps = 40 if use_graph else 8
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for i in range(steps):
        opt.step(batches[(3 + i) % 64])
    t1.record()
    torch.cuda.synchronize()
    agg = collections.OrderedDict()
    for name, e0, e1 in REC:
        agg.setdefault(name, []).append(e0.elapsed_time(e1) * 1e3)
    if rank == 0:
        print(f"big4m x{world} owner_sharded={opt.owner_sharded} peer_exchange={opt.peer_exchange} peer_dense={opt.peer_dense} "
              f"{'graph' if use_graph else 'eager'}: {t0.elapsed_time(t1) / steps * 1e3:.1f} us/step", flush=True)
        for name, v in agg.items():
            print(f"  {name:28s} calls/step {len(v) / steps:4.1f}  mean {sum(v) / len(v):8.1f} us  per step {sum(v) / steps:8.1f} us", flush=True)
    opt.check_peer_status()
    del opt
    import gc
    gc.collect()
    torch.cuda.synchronize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py — filtered full-ranking evaluation throughput of the chk_b200 hot path on B200.

    python bench.py --gpus 1 --steps 10 --warmup 3                # our arm (CUDA, through the C ABI)
    python bench.py --impl reference --steps 3 --warmup 1         # the reference's CPU algorithm (oracle port)
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # entity-table-sharded ranking over NCCL

Workload (BASELINE.json configs[4], the config the metric's "1/2/4/8 B200" refers to): FFTRotH rank=257,
fp32, synthetic 4,000,000-entity / 1,000-relation graph (2,000 relation rows with reciprocals), Zipf
heads/tails, filters built over all splits; one STEP = one evaluation batch of 500 filtered-ranking queries
(run.py eval_batch_size) against the whole entity table.  With N GPUs the table is row-sharded (strong
scaling: the step is fixed, each rank counts over N_ent/N rows, one int64 all_reduce per step).
The rank's table shard (8.2 GB / N) exceeds the 126 MB L2, so every step streams it from HBM.

One JSON line on stdout (rank 0).  `value` = queries/s with inputs resident in HBM; `e2e` = the same metric
through the public API model.get_ranking() with HOST query ids (host CSR lookup, H2D, D2H inside the timed
region); `roofline` = the dominant kernel (rank tile contraction) timed alone with CUDA events;
`cpu_baseline` = the oracle port on this box's host cores on a bounded sample; `train` = drop-in training
throughput (KGOptimizer contract + torch.optim) on BASELINE.json configs[1].
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "filtered_eval_queries_per_sec"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="big4m", choices=["big4m", "wn18rr", "fb237", "yago310"])
    ap.add_argument("--rank", type=int, default=None)
    ap.add_argument("--model", default=None)
    ap.add_argument("--dtype", default="float", choices=["float", "double"])
    ap.add_argument("--batch", type=int, default=500)
    ap.add_argument("--rank-algo", default=os.environ.get("CHK_RANK_ALGO", "auto"), choices=["auto", "fma", "mma"])
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


WORKLOAD_DEFAULTS = {"big4m": ("FFTRotH", 257), "wn18rr": ("FFTRotH", 33), "fb237": ("FFTRefH", 33),
                     "yago310": ("FFTAttH", 33)}


def config_dict(args, n_gpus, extra=None):
    from complexhyperbolickge_b200.synthetic import SHAPES
    model, rank = WORKLOAD_DEFAULTS[args.workload]
    model, rank = args.model or model, args.rank or rank
    n_ent, n_rel = SHAPES[args.workload][:2]
    cfg = {"workload": f"{model} rank={rank} {args.dtype} filtered full-ranking eval, synthetic {args.workload} "
                       f"({n_ent} entities, {n_rel} relations), eval batch {args.batch} queries/step",
           "baseline_config": "BASELINE.json configs[4]" if args.workload == "big4m" else args.workload,
           "entities": n_ent, "relation_rows": 2 * n_rel, "rank": rank, "eval_batch": args.batch,
           "parallelism": f"entity-table row-sharded x{n_gpus}, 1 int64 all_reduce/step" if n_gpus > 1 else "single GPU",
           "l2": "inputs larger than L2 (table shard streamed from HBM each step)"}
    if extra:
        cfg.update(extra)
    return cfg, model, rank


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "25", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append((time.time(), ln.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ts, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            inside = t0 - 0.05 <= ts <= t1 + 0.15
            try:
                if inside:
                    sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            if inside:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ watchdog
class Watchdog:
    """Per-phase deadlines.  A phase that exceeds its deadline (a hung collective, a rank that lost its GPU, a box that
    pages for minutes) must not hang the job: every rank dumps its Python stacks to stderr and exits 0; rank 0 first
    prints the JSON line with whatever has been measured so far plus an "error" key naming the phase."""

    def __init__(self, rank_id):
        self.rank_id, self.phase, self.deadline, self.partial = rank_id, "start", None, None
        self.emitted, self.stopped = False, False
        self.lock = threading.Lock()
        threading.Thread(target=self._run, daemon=True).start()

    def enter(self, phase, seconds):
        self.phase, self.deadline = phase, time.time() + seconds

    def stop(self):
        self.stopped = True

    def emit(self, line):
        with self.lock:
            if not self.emitted:
                self.emitted = True
                print(json.dumps(line), flush=True)

    def _run(self):
        import faulthandler
        while not self.stopped:
            time.sleep(1.0)
            if self.deadline is not None and time.time() > self.deadline and not self.stopped:
                sys.stderr.write(f"[bench watchdog] rank {self.rank_id}: phase '{self.phase}' exceeded its deadline\n")
                faulthandler.dump_traceback(file=sys.stderr, all_threads=True)
                sys.stderr.flush()
                if self.rank_id == 0 and not self.emitted:
                    line = dict(self.partial) if self.partial else {"metric": METRIC, "value": None, "unit": UNIT}
                    line["error"] = f"phase '{self.phase}' exceeded its deadline and was abandoned"
                    self.emit(line)
                os._exit(0)


# ------------------------------------------------------------------------------------------------ CPU arm
def oracle_eval_sample(model_name, rank, dtype, n_ent, n_rel2, n_queries, ent_slice, seed=0):
    """Time the oracle port's get_ranking on `n_queries` queries against `ent_slice` entity rows and
    extrapolate linearly in the table size (the reference's cost is linear in N: broadcast multiply-sum)."""
    from oracle import chk_oracle as O
    torch.set_num_threads(os.cpu_count())
    dt = torch.float32 if dtype == "float" else torch.float64
    g = torch.Generator().manual_seed(seed)
    n = 2 * (rank - 1)
    att = model_name == "FFTAttH"
    std = float(np.sqrt(0.4 / (2 * rank)))
    p = O.Params(O.KIND[model_name], rank, True, (torch.randn(ent_slice, 2 * rank, generator=g) * std).to(dt),
                 (torch.randn(n_rel2, 2 * n, generator=g) * 0.05).to(dt),
                 (torch.rand(n_rel2, 2 * n if att else n, generator=g) * 2 - 1).to(dt),
                 (torch.rand(n_rel2, 1, generator=g) * 1.5 + 0.5).to(dt), (torch.randn(ent_slice, 1, generator=g) * 0.1).to(dt),
                 (torch.randn(ent_slice, 1, generator=g) * 0.1).to(dt), torch.randn(n_rel2, n, generator=g).to(dt) if att else None)
    qs = torch.stack([torch.randint(0, ent_slice, (n_queries,), generator=g), torch.randint(0, n_rel2, (n_queries,), generator=g),
                      torch.randint(0, ent_slice, (n_queries,), generator=g)], 1)
    filters = {(int(h), int(r)): [int(t)] for h, r, t in qs.numpy()}
    t0 = time.perf_counter()
    O.get_ranking(p, qs, filters, batch_size=n_queries)
    dt_s = time.perf_counter() - t0
    return n_queries / (dt_s * (n_ent / ent_slice)), dt_s


def oracle_eval_budget(model_name, rank, dtype, n_ent, n_rel2, budget_s, max_samples=64):
    """Repeat bounded oracle samples until ~budget_s of CPU work is spent; returns (queries/s, cpu seconds, sample text)."""
    ent_slice = min(n_ent, 250_000 if rank > 65 else 160_000)
    nq = 2 if rank > 65 else 8
    inv, spent, n = 0.0, 0.0, 0
    while n < max_samples and (spent < budget_s or n < 2):
        v, t = oracle_eval_sample(model_name, rank, dtype, n_ent, n_rel2, nq, ent_slice, seed=n)
        inv += 1.0 / v
        spent += t
        n += 1
    sample = (f"{n} x ({nq} queries x {ent_slice} of {n_ent} entity rows), {spent:.1f} s of CPU work, extrapolated linearly in the "
              f"table size (the reference materialises a (b, N, r) complex temporary); oracle port of models/base.py:228-280 "
              f"on torch-CPU tensors, all host threads")
    return n / inv, spent, sample


def run_reference(args):
    """--impl reference: the reference's own CPU algorithm for the path (oracle port; the Python reference
    cannot travel to the GPU box), all host threads, same metric/config; each step a bounded sample."""
    rank_env = int(os.environ.get("RANK", "0"))
    if rank_env != 0:
        return
    cfg, model, rank = config_dict(args, args.gpus)
    from complexhyperbolickge_b200.synthetic import SHAPES
    n_ent, n_rel = SHAPES[args.workload][:2]
    vals, times, sample = [], [], ""
    for i in range(args.warmup + args.steps):
        v, t, sample = oracle_eval_budget(model, rank, args.dtype, n_ent, 2 * n_rel, budget_s=3.0, max_samples=8)
        if i >= args.warmup:
            vals.append(v)
            times.append(t)
    value = float(len(vals) / sum(1.0 / v for v in vals))       # total queries / total (extrapolated) time
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32" if args.dtype == "float" else "f64",
            "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "per step: " + sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def oracle_train_sample(steps=10):
    """CPU baseline of the training step (oracle port of neg_sampling_loss fwd + hand-derived bwd) on configs[1]."""
    from oracle import chk_oracle as O
    torch.set_num_threads(os.cpu_count())
    g = torch.Generator().manual_seed(0)
    n_ent, n_rel2, rank, B, neg = 14541, 474, 33, 500, 250
    n = 2 * (rank - 1)
    std = float(np.sqrt(0.4 / (2 * rank)))
    p = O.Params(O.REF, rank, True, torch.randn(n_ent, 2 * rank, generator=g) * std, torch.randn(n_rel2, 2 * n, generator=g) * 0.05,
                 torch.rand(n_rel2, n, generator=g) * 2 - 1, torch.rand(n_rel2, 1, generator=g) * 1.5 + 0.5,
                 torch.randn(n_ent, 1, generator=g) * 0.1, torch.randn(n_ent, 1, generator=g) * 0.1, None)
    ts = []
    for i in range(steps + 1):
        batch = torch.stack([torch.randint(0, n_ent, (B,), generator=g), torch.randint(0, n_rel2, (B,), generator=g),
                             torch.randint(0, n_ent, (B,), generator=g)], 1)
        negs = torch.randint(0, n_ent, (B, neg), generator=g)
        t0 = time.perf_counter()
        O.neg_sampling_loss(p, batch, negs)
        if i:
            ts.append(time.perf_counter() - t0)
    return {"value": B / float(np.mean(ts)), "unit": "triples/s", "cores": os.cpu_count(), "kind": "port",
            "sample": f"{steps} steps of B={B}, neg={neg} (forward + analytic backward, no optimizer), {sum(ts):.1f} s of CPU work"}


# ------------------------------------------------------------------------------------------------ train leg
def dp_train_leg(steps, warmup, device, pg, world):
    """Data-parallel training (parallel.DataParallelKGOptimizer): configs[1] shape, 500 triples per rank per step
    (weak scaling), sparse row-gradient all_gather + dense all_reduce of the relation tables over NCCL."""
    from argparse import Namespace
    import torch.distributed as dist
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    from complexhyperbolickge_b200.optim import N3
    from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer
    g = synthetic.make_graph("fb237", seed=0)
    args = Namespace(sizes=(g["n_ent"], g["n_rel2"], g["n_ent"]), rank=33, dropout=0, gamma=0, dtype="float",
                     bias="learn", init_size=1e-3, multi_c=True)
    model = chk.FFTRefH(args).to(device)
    synthetic.trained_like_(model, 0)
    B = 500 * world
    opt = FusedDataParallelKGOptimizer(model, N3(0.0), torch.optim.Adagrad(model.parameters(), lr=0.02), B, 1, 250, False,
                                       verbose=False, process_group=pg)
    steps *= 10
    ex = synthetic.train_examples(g)
    ex = ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))][: (steps + warmup) * B].to(device)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    loss = None
    for i in range(steps + warmup):
        if i == warmup:
            torch.cuda.synchronize()
            dist.barrier()
            ev0.record()
        opt.step(ex[i * B:(i + 1) * B])
    lv = opt._loss_sum.item() / (steps + warmup)
    ev1.record()
    torch.cuda.synchronize()
    t = torch.tensor([ev0.elapsed_time(ev1) / steps], device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item()
    return {"metric": "train_triples_per_sec", "value": B / (ms * 1e-3), "unit": "triples/s", "ms_per_step": ms,
            "scaling": "weak", "global_batch": B,
            "config": f"BASELINE.json configs[1] shape: FFTRefH rank=33 Adagrad neg=250, 500 triples per rank x{world}, "
                      "fused step per rank (CUDA graph) + all_gather of touched row ids + dense all_reduce of the (small) table gradients "
                      "[sparse row exchange for tables larger than the touched rows] + row-sparse Adagrad on the union",
            "mean_loss_rank0": lv}


def train_big_leg(model, graph, steps, warmup, device, pg, world):
    """Training on BASELINE.json configs[4] (the 4M-entity graph, FFTRotH rank 257): fused step per rank on 500 triples
    (neg = 100), row-sparse Adagrad on the 8.2 GB table.  N > 1: data parallel with replicated tables (weak scaling, 500
    triples per rank); the touched gradient rows travel (chk_claim_gather_rows + all_gather + rank-ordered scatter),
    never the dense 8.2 GB gradient."""
    import torch.distributed as dist
    from complexhyperbolickge_b200 import synthetic
    from complexhyperbolickge_b200.optim import N3
    from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    neg, B = 100, 500 * world
    model.train()
    adagrad = torch.optim.Adagrad(model.parameters(), lr=0.02)
    if world > 1:
        # beyond 2 ranks the step runs eagerly: it is exchange-bound (hundreds of MB of gradient rows per step), so the
        # ~50 launches cost nothing, and no multi-hundred-MB NCCL collective has to be captured into a CUDA graph
        opt = FusedDataParallelKGOptimizer(model, N3(0.0), adagrad, B, 1, neg, False, verbose=False, process_group=pg,
                                           use_cuda_graph=world <= 2)
    else:
        opt = FusedKGOptimizer(model, N3(0.0), adagrad, B, 1, neg, False, verbose=False)
    ex = synthetic.train_examples(graph)
    ex = ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))[: (steps + warmup) * B]]
    pinned = ex.pin_memory()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(steps + warmup):
        if i == warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            opt._loss_sum.zero_()
            ev0.record()
        b = pinned[i * B:(i + 1) * B].to(device, non_blocking=True)      # H2D of the step's batch inside the timed region
        if world > 1:
            opt.step(b)
        else:
            opt.fused_step(b)
    lv = opt._loss_sum.item() / steps
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    rank = model.rank
    bytes_per_triple = 2 * (2 + neg) * 2 * rank * 4
    out = {"metric": "train_triples_per_sec", "value": B / (ms * 1e-3), "unit": "triples/s", "ms_per_step": ms,
           "scaling": "weak", "global_batch": B, "mean_loss_rank0": lv,
           "config": f"BASELINE.json configs[4]: FFTRotH rank={rank} Adagrad, 500 triples per rank x{world}, neg={neg}, synthetic 4M-entity "
                     "graph; fused step (K1 + K3 + loss + adjoints + row-sparse Adagrad" + (", CUDA graph" if world <= 2 else ", eager launches")
                     + "), batch H2D from pinned memory each step"
                     + ("; replicated tables, sparse gradient-row exchange (claim-gather + all_gather + rank-ordered scatter)"
                        if world > 1 else ""),
           "algorithmic_bytes_per_triple": bytes_per_triple,
           "hbm_frac_per_gpu": B / world / (ms * 1e-3) * bytes_per_triple / 1e9 / 6530.0}
    if world > 1:
        m = 500 * (2 + neg)
        out["exchange_bytes_per_rank_per_step"] = world * m * (2 * rank * 4 + 2 * 4 + 8)
    del opt, adagrad
    for p in model.parameters():
        p.grad = None
    return out


def train_leg(steps, warmup, device):
    """Training throughput on BASELINE.json configs[1]: FFTRefH rank=33 Adagrad bs=500 neg=250 on the synthetic
    FB15k-237 shape.  Two numbers: the fused step (train.FusedKGOptimizer: one kernel chain + row-sparse Adagrad in a
    CUDA graph; same loss and update) and the unfused drop-in KGOptimizer contract loop (two model() calls, autograd,
    dense torch.optim).  Each step's batch comes from pinned host memory (H2D inside the timed region)."""
    from argparse import Namespace
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    from complexhyperbolickge_b200.optim import KGOptimizer, N3
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    g = synthetic.make_graph("fb237", seed=0)
    args = Namespace(sizes=(g["n_ent"], g["n_rel2"], g["n_ent"]), rank=33, dropout=0, gamma=0, dtype="float",
                     bias="learn", init_size=1e-3, multi_c=True)
    ex = synthetic.train_examples(g)
    ex = ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))]
    out = {}
    for mode, n_steps in (("fused", 10 * steps), ("drop_in", steps)):
        model = chk.FFTRefH(args).to(device)
        synthetic.trained_like_(model, 0)
        cls = FusedKGOptimizer if mode == "fused" else KGOptimizer
        opt = cls(model, N3(0.0), torch.optim.Adagrad(model.parameters(), lr=0.02), 500, 1, 250, False, verbose=False)
        pinned = ex[: (n_steps + warmup) * 500].pin_memory()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loss = None
        for i in range(n_steps + warmup):
            if i == warmup:
                torch.cuda.synchronize()
                if mode == "fused":
                    opt._loss_sum.zero_()
                ev0.record()
            b = pinned[i * 500:(i + 1) * 500].to(device, non_blocking=True)
            if mode == "fused":
                opt.fused_step(b)
            else:
                loss = opt.calculate_loss(b)
                loss.backward()
                opt.optimizer.step()
                opt.optimizer.zero_grad()
        lv = opt._loss_sum.item() / n_steps if mode == "fused" else loss.item()
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / n_steps
        out[mode] = {"value": 500.0 / (ms * 1e-3), "ms_per_step": ms, "loss": lv}
    bytes_per_triple = 2 * (2 + 250) * 66 * 4
    return {"metric": "train_triples_per_sec", "value": out["fused"]["value"], "unit": "triples/s",
            "ms_per_step": out["fused"]["ms_per_step"], "mean_loss": out["fused"]["loss"],
            "config": "BASELINE.json configs[1]: FFTRefH rank=33 Adagrad bs=500 neg=250, synthetic FB15k-237 shape; fused step "
                      "(K1 + K3 + loss + adjoints + row-sparse Adagrad, CUDA graph), batch H2D from pinned memory each step",
            "algorithmic_bytes_per_triple": bytes_per_triple,
            "hbm_frac": out["fused"]["value"] * bytes_per_triple / 1e9 / 6530.0,
            "drop_in_contract_loop": {"value": out["drop_in"]["value"], "ms_per_step": out["drop_in"]["ms_per_step"],
                                      "last_loss": out["drop_in"]["loss"],
                                      "what": "unfused KGOptimizer loop: 2 model() calls + autograd + dense torch.optim.Adagrad"}}


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank_id = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device: complexhyperbolickge_b200 has no CPU fallback")
    wd = Watchdog(rank_id)
    wd.enter("setup (NCCL init, synthetic graph, model, evaluation state)", 300)
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        if world >= 8:
            # The one 8-GPU run of round 1 hit the harness limit without printing a line (cause unknown).  The path's
            # collectives are tiny (8 B per query) or plain all_gathers, so the in-switch NVLS algorithms and CUDA-graph
            # buffer registration buy nothing here: take the plain ring/tree paths at 8 ranks.  Overridable from outside.
            os.environ.setdefault("NCCL_NVLS_ENABLE", "0")
            os.environ.setdefault("NCCL_GRAPH_REGISTER", "0")
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    from argparse import Namespace
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import ops, ranking, synthetic

    cfg, model_name, rank = config_dict(args, world)
    graph = synthetic.make_graph(args.workload, seed=0)
    n_ent, n_rel2 = graph["n_ent"], graph["n_rel2"]
    margs = Namespace(sizes=(n_ent, n_rel2, n_ent), rank=rank, dropout=0, gamma=0, dtype=args.dtype, bias="learn",
                      init_size=1e-3, multi_c=True)
    with torch.device(device):          # build the 8.2 GB table ON the GPU: no host copy (8 ranks x 12 GB of host RAM), no CPU init
        model = getattr(chk, model_name)(margs)
    model = model.to(device)
    synthetic.trained_like_(model, 0)
    model.eval()
    algo = args.rank_algo
    if algo == "auto":
        algo = "mma" if ops.mma_available() else "fma"
    model.rank_algo = algo
    model.process_group = pg
    cfg["rank_algo"] = algo
    b = args.batch
    K, W = args.steps, args.warmup
    findex = graph["filters"]["rhs"]
    test = graph["test"]
    reps = (b * (K + W) + len(test) - 1) // len(test)
    qall = np.concatenate([test] * reps)[: b * (K + W)]
    hist = np.diff(findex.indptr)
    cfg["filter_len"] = {"mean": float(hist.mean()), "p99": float(np.percentile(hist, 99)), "max": int(hist.max())}

    # ---- resident inputs for `value`
    steps_in = []
    for i in range(K + W):
        qb = qall[i * b:(i + 1) * b]
        indptr, idx = findex.batch_csr(qb)
        steps_in.append((torch.from_numpy(qb).to(device), torch.from_numpy(indptr).to(device),
                         torch.from_numpy(idx).to(device), int(idx.size)))
    with torch.no_grad():
        state = ranking.eval_state(model)
        ws = ops.rank_mma_workspace(rank, b, device) if state.algo == ops.CHK_RANK_MMA else None
        counts = torch.zeros(b, dtype=torch.int64, device=device)

        def step(i):
            counts.zero_()
            qd, ip, ix, tot = steps_in[i]
            ranking.rank_batch(model, state, qd, ip, ix, tot, counts, ws)
            if world > 1:
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)

        wd.enter("evaluation steps (resident inputs) + roofline probe", 180)
        for i in range(W):
            step(i)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        sampler = ClockSampler(local)
        if rank_id == 0:
            sampler.start()
        time.sleep(0.2)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        t_wall0 = time.time()
        launches0 = ops.launch_count
        ev0.record()
        for i in range(W, W + K):
            step(i)
        ev1.record()
        gpu_launches = ops.launch_count - launches0
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t_wall1 = time.time()
        ms_total = ev0.elapsed_time(ev1)
        if world > 1:
            t = torch.tensor([ms_total], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = t.item()
        clocks = sampler.stop(t_wall0, t_wall1) if rank_id == 0 else None
        ranks_check = (counts + 1).float().cpu()

        # ---- roofline: the dominant kernel alone (rank tile contraction), CUDA events on torch's stream
        qd, ip, ix, tot = steps_in[W]
        q, _ = ops.query_fwd(model.KIND, rank, True, model.entity.weight.detach(), model.rel.weight.detach(),
                             model.rel_diag.weight.detach(), None if model._ctx_weight() is None else model._ctx_weight().detach(),
                             model.c.weight.detach(), qd[:, 0].contiguous(), qd[:, 1].contiguous())
        qn = ops.row_hnorm(rank, q)
        bhv = model.bh.weight.detach().view(-1)[qd[:, 0]].contiguous()
        rows = model.entity.weight.detach()[qd[:, 2]].contiguous()
        tgt = ops.target_scores(rank, q, qn, bhv, rows, ops.row_hnorm(rank, rows),
                                model.bt.weight.detach().view(-1)[qd[:, 2]].contiguous())
        empty_ip = torch.zeros(b + 1, dtype=torch.int64, device=device)
        nrep = 3
        torch.cuda.synchronize()
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ops.rank_counts(state.algo, rank, q, qn, bhv, tgt, state.entity, state.hn, state.bt, state.lo, empty_ip, ix, 0,
                        counts, state.shadow, ws)
        call_ms, mma_ms = 0.0, 0.0
        mma = state.algo == ops.CHK_RANK_MMA
        if mma:                                   # the library records m0/m1 immediately around rank_mma_kernel
            m0.record(); m1.record()
            ops.rank_mma_profile_events(m0, m1)
        for _ in range(nrep):
            k0.record()
            ops.rank_counts(state.algo, rank, q, qn, bhv, tgt, state.entity, state.hn, state.bt, state.lo, empty_ip, ix,
                            0, counts, state.shadow, ws)
            k1.record()
            torch.cuda.synchronize()
            call_ms += k0.elapsed_time(k1) / nrep
            if mma:
                mma_ms += m0.elapsed_time(m1) / nrep
        if mma:
            ops.rank_mma_profile_events(None, None)
        kern_ms = mma_ms if mma else call_ms
        recheck = ops.rank_mma_status(ws) if ws is not None else (0, False)
        shard_rows = state.hi - state.lo
        flops = 8.0 * rank * b * shard_rows

    # ---- e2e through the public API with host buffers
    e2e = None
    wd.enter("end-to-end get_ranking", 120)
    if not args.no_e2e:
        # ONE public-API call over the K timed batches (what compute_metrics does): per batch the host builds the filter
        # CSR, copies ids + CSR host->device from pinned memory and the batch's ranks come back device->host, all
        # inside the timed region; the batches are pipelined (host prepares batch i+1 while the GPU counts batch i).
        qhost = torch.from_numpy(qall).pin_memory()
        model.get_ranking(qhost[:2 * b], findex, batch_size=b)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r_host = model.get_ranking(qhost[W * b:(W + K) * b], findex, batch_size=b)
        e1.record()
        torch.cuda.synchronize()
        io = model.last_eval_io
        e_ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([e_ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = t.item()
        same = bool(torch.equal(r_host[-b:], ranks_check))      # e2e ranks of the last batch == the resident-input ranks
        if not same:
            sys.stderr.write("[bench] e2e ranks differ from the resident-input ranks\n")
        e2e = {"value": b * K / (e_ms * 1e-3), "unit": UNIT, "ranks_equal_resident_path": same, "h2d_bytes_per_step": io["h2d_bytes"] // io["batches"],
               "d2h_bytes_per_step": io["d2h_bytes"] // io["batches"], "ms_per_step": e_ms / K,
               "api": f"model.get_ranking(host LongTensor[{K}*{b},3], FilterIndex, batch_size={b}): one call, {K} pipelined "
                      "batches, per batch 1 H2D copy (ids + filter CSR, pinned) and 1 D2H copy (ranks)"}

    # ---- the JSON line so far (rank 0); the training legs below only add keys to it
    line = None
    if rank_id == 0:
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        peak_tf = peaks.get("bf16_tflops", 1590.0)
        achieved = flops / (kern_ms * 1e-3) / 1e12
        traffic = None                  # dram bytes per launch of the dominant kernel, from the committed ncu --set full capture
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1 and args.workload == "big4m" and mma:
            traffic = json.load(open(tp)).get("rank_mma_kernel_big4m_dram_bytes_per_launch")
        issued = (8 * (rank - 1) * 3 + 8) if mma else 8 * rank
        roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                    "traffic": traffic,
                    "kernel": "rank_mma_kernel<0,0> (tcgen05 bf16x3 contraction + fused epilogue), timed alone with CUDA events recorded "
                              "around its launch inside chk_rank_counts" if mma else "rank_tile_kernel<float,1> (fp32 FMA)",
                    "kernel_ms": kern_ms, "call_ms": call_ms,
                    "call_what": "whole chk_rank_counts call: operand prep + rank_mma_kernel + exact re-check" if mma else "chk_rank_counts",
                    "call_frac": flops / (call_ms * 1e-3) / 1e12 / peak_tf,
                    "algorithmic_flop_per_pair": 8 * rank, "issued_flop_per_pair": issued,
                    "issued_frac": achieved / peak_tf * issued / (8 * rank),
                    "issued_vs_sustained_peak": (achieved * issued / (8 * rank)) / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                    "pairs_per_launch": b * shard_rows, "recheck_pairs_per_launch": recheck[0], "recheck_overflow": recheck[1],
                    "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1.59 PFLOP/s")}
        line = {"metric": METRIC, "value": b * K / (ms_total * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if args.dtype == "float" else "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
                "e2e": e2e, "gpu_launches": gpu_launches, "roofline": roofline,
                "mean_rank_check": float(ranks_check.mean())}
        wd.partial = line

    # ---- training legs (every rank takes part); each has its own deadline, a stuck leg is reported, not waited for
    del state, ws
    model.release_eval_cache()
    torch.cuda.empty_cache()
    def leg(key, what, seconds, fn, *a):
        """A training leg adds a key to the line; if it fails the headline (evaluation) numbers are still reported."""
        import traceback
        wd.enter(what, seconds)
        try:
            out = fn(*a)
        except Exception as e:                       # noqa: BLE001 — reported in the line, stack on stderr
            traceback.print_exc(file=sys.stderr)
            out = {"error": f"{type(e).__name__}: {e}"[:400]}
        if line is not None:
            line[key] = out

    if not args.no_train:
        if world > 1:
            leg("train", "data-parallel training leg (configs[1] shape)", 120, dp_train_leg, 20, 3, device, pg, world)
        if args.workload == "big4m":
            leg("train_big4m", "training leg on the 4M-entity table (configs[4])", 180, train_big_leg, model, graph, 50, 5,
                device, pg, world)
    del model
    torch.cuda.empty_cache()
    if rank_id != 0:
        wd.enter("process-group teardown", 30)
        dist.destroy_process_group()
        wd.stop()
        return
    if world == 1 and not args.no_train:
        leg("train", "single-GPU training leg (configs[1])", 240, train_leg, 20, 3, device)
    if world == 1 and not args.no_cpu_baseline:
        wd.enter("CPU baseline (oracle port on the host cores)", 300)
        v, spent, sample = oracle_eval_budget(model_name, rank, args.dtype, n_ent, n_rel2, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        if not args.no_train and "error" not in line.get("train", {"error": 1}):
            line["train"]["cpu_baseline"] = oracle_train_sample()
    wd.emit(line)
    if world > 1:
        wd.enter("process-group teardown", 30)
        dist.destroy_process_group()
    wd.stop()


if __name__ == "__main__":
    a = parse()
    # stdout carries exactly ONE JSON line: anything libraries print on fd 1 (e.g. NCCL's version banner) goes to stderr
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(_json_fd, "w")
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

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
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[0..3] legs and the K1/K3 kernel rooflines")
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
    pages for minutes) must not hang the job: every rank dumps its Python stacks to stderr and exits with code 3; rank 0
    first prints the JSON line with whatever has been measured so far plus an "error" key naming the phase."""

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
                os._exit(3)                       # the partial line is printed, but a hung phase is a failure


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


# ------------------------------------------------------------------------------------------------ shared helpers
def load_peaks():
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(pk)) if os.path.exists(pk) else {}


def make_model(model_name, rank, dtype, graph, device, multi_c=True):
    from argparse import Namespace
    import complexhyperbolickge_b200 as chk
    from complexhyperbolickge_b200 import synthetic
    margs = Namespace(sizes=(graph["n_ent"], graph["n_rel2"], graph["n_ent"]), rank=rank, dropout=0, gamma=0, dtype=dtype,
                      bias="learn", init_size=1e-3, multi_c=multi_c)
    with torch.device(device):          # build the table ON the GPU: no host copy (8 ranks x 12 GB of host RAM at 4M entities), no CPU init
        model = getattr(chk, model_name)(margs)
    model = model.to(device)
    synthetic.trained_like_(model, 0)
    return model


def cycle_batches(ex, n_batches, B):
    """n_batches batches of B rows from the example set, wrapping around (never slicing past the end)."""
    idx = (torch.arange(n_batches * B) % ex.shape[0]).view(n_batches, B)
    return ex[idx]


# ------------------------------------------------------------------------------------------------ evaluation
def measure_eval(model, graph, b, K, W, device, pg, world, rank_id, local, want_e2e, peaks, wd=None):
    """Filtered full-ranking throughput of `model` on `graph`: resident-input steps (`value`), the dominant kernel alone
    (roofline) and the public-API call with host buffers (`e2e`).  Returns the pieces of a JSON line (rank 0) or None."""
    import torch.distributed as dist
    from complexhyperbolickge_b200 import ops, ranking
    rank = model.rank
    findex = graph["filters"]["rhs"]
    test = graph["test"]
    reps = (b * (K + W) + len(test) - 1) // len(test)
    qall = np.concatenate([test] * reps)[: b * (K + W)]
    steps_in = [torch.from_numpy(qall[i * b:(i + 1) * b]).to(device) for i in range(K + W)]     # resident inputs: query ids in HBM
    findex.device_arrays(device)                                                                # the filter index lives on the device
    with torch.no_grad():
        state = ranking.eval_state(model)
        mma = state.algo == ops.CHK_RANK_MMA
        ws = ops.rank_mma_workspace(rank, b, device) if mma else None
        counts = torch.zeros(b, dtype=torch.int64, device=device)
        target_buf = torch.zeros(b, dtype=model.entity.weight.dtype, device=device)
        flags = torch.zeros(1, dtype=torch.int32, device=device)
        scratch = ops.eval_scratch(rank, b, model.entity.weight.dtype, device)

        def step(i):
            ranking.rank_batch_fused(model, state, findex, steps_in[i], counts, target_buf, flags, scratch, ws)
            if world > 1:
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)

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
        qd = steps_in[W]
        ix = torch.zeros(1, dtype=torch.int64, device=device)
        ctxw = model._ctx_weight()
        q, _ = ops.query_fwd(model.KIND, rank, bool(model.multi_c), model.entity.weight.detach(), model.rel.weight.detach(),
                             model.rel_diag.weight.detach(), None if ctxw is None else ctxw.detach(),
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
        algo_name = "mma" if mma else "fma"
    del state, ws

    # ---- e2e through the public API with host buffers
    e2e = None
    if want_e2e:
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
                      "batches, per batch 1 H2D copy (query ids, pinned), 1 chk_eval_batch call (filter index searched on the device) and 1 D2H copy (ranks)",
               "note": "the per-pass evaluation state (Hermitian norms, bf16 shadow: one pass over the table) is cached on the model "
                       "and was built by the warm call before the timed one; it is rebuilt only after a parameter update"}
    if rank_id != 0:
        return None
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    achieved = flops / (kern_ms * 1e-3) / 1e12
    issued = (8 * (rank - 1) * 3 + 8) if mma else 8 * rank
    roofline = {"bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                "traffic": None,
                "kernel": "rank_mma_kernel<0,0> (tcgen05 bf16x3 contraction + fused epilogue), timed alone with CUDA events recorded "
                          "around its launch inside chk_rank_counts" if mma else "rank_tile_kernel (exact FMA tier)",
                "kernel_ms": kern_ms, "call_ms": call_ms,
                "call_what": "whole chk_rank_counts call: operand prep + rank_mma_kernel + exact re-check" if mma else "chk_rank_counts",
                "call_frac": flops / (call_ms * 1e-3) / 1e12 / peak_tf,
                "algorithmic_flop_per_pair": 8 * rank, "issued_flop_per_pair": issued,
                "issued_frac": achieved / peak_tf * issued / (8 * rank),
                "issued_vs_sustained_peak": (achieved * issued / (8 * rank)) / peaks["bf16_tflops_sustained"] if peaks.get("bf16_tflops_sustained") else None,
                "pairs_per_launch": b * shard_rows, "pairs_per_s": b * shard_rows / (kern_ms * 1e-3),
                "recheck_pairs_per_launch": recheck[0], "recheck_overflow": recheck[1],
                "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst; kernel timed alone)" if peaks else "fallback 1.59 PFLOP/s")}
    hist = np.diff(findex.indptr)
    return {"value": b * K / (ms_total * 1e-3), "ms_per_step": ms_total / K, "clocks": clocks, "e2e": e2e, "gpu_launches": gpu_launches,
            "roofline": roofline, "mean_rank_check": float(ranks_check.mean()), "rank_algo": algo_name,
            "filter_len": {"mean": float(hist.mean()), "p99": float(np.percentile(hist, 99)), "max": int(hist.max())}}


# ------------------------------------------------------------------------------------------------ training
HBM_GBS_FALLBACK = 6650.0


def train_bytes_per_triple(rank, neg, itemsize, double_neg=False):
    """gathered tail rows + their gradient rows + the query row and its gradient = 2 (2 + neg) 2r s bytes per triple
    (with double_neg every negative also reads a head row and writes its gradient row: 2 (1 + neg) 2r s more)."""
    b = 2 * (2 + neg) * 2 * rank * itemsize
    if double_neg:
        b += 2 * (1 + neg) * 2 * rank * itemsize
    return b


def measure_train(model, graph, opt_name, lr, B, neg, double_neg, steps, warmup, device, pg, world, peaks, contract_steps=0):
    """Training throughput through the public optimizer API: FusedKGOptimizer.fused_step (world 1) or
    FusedDataParallelKGOptimizer.step (global batch B*world, weak scaling).  Every step's batch comes from pinned host memory
    (H2D inside the timed region); the mean loss is read back once after the timed steps."""
    import torch.distributed as dist
    from complexhyperbolickge_b200 import ops, synthetic
    from complexhyperbolickge_b200.optim import KGOptimizer, N3
    from complexhyperbolickge_b200.parallel import FusedDataParallelKGOptimizer
    from complexhyperbolickge_b200.train import FusedKGOptimizer
    model.train()
    mk = {"Adagrad": lambda ps: torch.optim.Adagrad(ps, lr=lr), "Adam": lambda ps: torch.optim.Adam(ps, lr=lr)}[opt_name]
    Bg = B * world
    torch_opt = mk(model.parameters())
    if world > 1:
        opt = FusedDataParallelKGOptimizer(model, N3(0.0), torch_opt, Bg, 1, neg, double_neg, verbose=False, process_group=pg,
                                           use_cuda_graph=os.environ.get("CHK_DP_GRAPH", "1") != "0")
    else:
        opt = FusedKGOptimizer(model, N3(0.0), torch_opt, Bg, 1, neg, double_neg, verbose=False)
    ex = synthetic.train_examples(graph)
    ex = ex[torch.randperm(ex.shape[0], generator=torch.Generator().manual_seed(0))]
    pinned = cycle_batches(ex, steps + warmup, Bg).pin_memory()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0, first = 0, ops.launch_count
    kernels_per_step = None
    for i in range(steps + warmup):
        if i == 1:                                                    # step 0 ran every wrapper once (twice when it also captured the graph)
            kernels_per_step = (ops.launch_count - first) / (2 if opt.use_cuda_graph else 1)
        if i == warmup:
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            opt._loss_sum.zero_()
            launches0 = ops.launch_count
            ev0.record()
        if world > 1:
            opt.step(pinned[i])                                       # rows rank::world, H2D inside
        else:
            opt.fused_step(pinned[i])                                 # pinned host batch: ONE async H2D copy into the step's id buffer
    launches = ops.launch_count - launches0
    lv = opt._loss_sum.double()
    if world > 1:
        dist.all_reduce(lv, op=dist.ReduceOp.SUM, group=pg)
    lv = lv.item() / steps                                            # D2H of the result inside the timed region
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / steps
    if world > 1:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    es = model.entity.weight.element_size()
    bpt = train_bytes_per_triple(model.rank, neg, es, double_neg)
    hbm = peaks.get("hbm_gbs", HBM_GBS_FALLBACK)
    out = {"metric": "train_triples_per_sec", "value": Bg / (ms * 1e-3), "unit": "triples/s", "ms_per_step": ms, "scaling": "weak",
           "global_batch": Bg, "mean_loss": lv, "optimizer": opt_name, "neg": neg, "double_neg": bool(double_neg),
           "kernels_per_step": kernels_per_step, "eager_launches_in_timed_region": launches,
           "h2d_bytes_per_step": (Bg // world) * 24, "d2h_bytes_total": 8,
           "algorithmic_bytes_per_triple": bpt,
           "hbm_frac_per_gpu": (Bg / world) / (ms * 1e-3) * bpt / 1e9 / hbm,
           "path": ("FusedDataParallelKGOptimizer.step: fused chain per rank on rows rank::world (CUDA graph incl. NCCL), "
                    + ("sparse exchange: all_gather" + (" by peer reads of symmetric buffers behind a flag barrier (chk_dp_all_gather, no NCCL call)"
                                                        if getattr(opt, "peer_exchange", False) else " (NCCL)")
                       + " of slot ids + per-rank contributions (head-gradient rows, query rows, 16 B of pair "
                       "coefficients per negative instead of its gradient row), one-kernel ordered reduce that rebuilds the rows + Adagrad in place"
                       + ("; tables owner-sharded in symmetric memory: K3 reads tail rows from their owner over NVLink, each rank updates only its own rows"
                          if getattr(opt, "owner_sharded", False) else "; every replica applies every update")
                       if getattr(opt, "sparse_entity", False) else "local segment-reduce into a flat dense gradient")
                    + ("; dense tables: gradient reduce-scatter + optimizer + parameter broadcast as ONE kernel over NVLink peer memory (chk_dp_fused_apply, no NCCL call)"
                       if getattr(opt, "peer_dense", False) else "; dense tables: ONE NCCL all_reduce + dense apply on every replica")
                    if world > 1 else
                    "FusedKGOptimizer.fused_step: prep(sampler) -> K1 -> K3 fwd+loss+bwd (pair coefficients) -> K1 adjoint -> ordered segment-reduce "
                    "rebuilding the tail-row gradients + optimizer (CUDA graph)")}
    if world > 1 and getattr(opt, "owner_sharded", False):
        # owner-sharded tables: the replicas are made current once per epoch (epoch() does it); timed here on its own and
        # amortised over the steps of one epoch of this synthetic graph
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        s0.record()
        opt.sync_replicas()
        s1.record()
        torch.cuda.synchronize()
        t = torch.tensor([s0.elapsed_time(s1)], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        steps_per_epoch = max(1, ex.shape[0] // Bg)
        out["owner_sharded"] = True
        out["replica_sync_ms_per_epoch"] = t.item()
        out["steps_per_epoch"] = steps_per_epoch
        out["value_incl_epoch_sync"] = Bg / ((ms + t.item() / steps_per_epoch) * 1e-3)
    if world > 1 and getattr(opt, "sparse_entity", False):
        pl = opt._plan(opt.local_batch_size)
        nbytes = pl.flat.numel() * es + pl.ent_ids.numel() * 8
        out["exchange_bytes_in_per_rank_per_step"] = (world - 1) * nbytes
        out["exchange_gbs_in_per_rank"] = (world - 1) * nbytes / (ms * 1e-3) / 1e9
        out["nvlink_ref_gbs"] = 770.0
    if contract_steps:
        del opt
        m2 = model
        c_opt = KGOptimizer(m2, N3(0.0), mk(m2.parameters()), B, 1, neg, double_neg, verbose=False)
        for p in m2.parameters():
            p.grad = None
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loss = None
        for i in range(contract_steps + 2):
            if i == 2:
                torch.cuda.synchronize()
                c0.record()
            bb = pinned[i % pinned.shape[0]][:B].to(device, non_blocking=True)
            loss = c_opt.calculate_loss(bb)
            loss.backward()
            c_opt.optimizer.step()
            c_opt.optimizer.zero_grad()
        last = loss.item()
        c1.record()
        torch.cuda.synchronize()
        cms = c0.elapsed_time(c1) / contract_steps
        out["drop_in_contract_loop"] = {"value": B / (cms * 1e-3), "ms_per_step": cms, "last_loss": last,
                                        "what": "unfused KGOptimizer loop: 2 model() calls + autograd + dense torch.optim"}
    for p in model.parameters():
        p.grad = None
    return out


# ------------------------------------------------------------------------------------------------ kernel rooflines
_FLUSH = None


def _time_kernel(fn, reps=5):
    """Median CUDA-event time of fn() on torch's current stream, L2 flushed (256 MB write) between repetitions."""
    global _FLUSH
    if _FLUSH is None:
        _FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        _FLUSH.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def kernel_rooflines(device, peaks):
    """K1 (query transform) and K3 (training pass) alone at sizes where they are not launch-latency bound, against the
    measured HBM copy bandwidth.  Algorithmic bytes as in DESIGN.md §4."""
    from complexhyperbolickge_b200 import ops
    hbm = peaks.get("hbm_gbs", HBM_GBS_FALLBACK)
    src = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s"
    g = torch.Generator(device=device).manual_seed(0)
    out = {}
    f32 = torch.float32
    # ---- K1 at 2^20 queries, rank 33 fp32 (entity table 1M rows = 264 MB > L2)
    rank, nq, n_ent, n_rel2 = 33, 1 << 20, 1_000_000, 474
    n = 2 * (rank - 1)
    ent = torch.randn(n_ent, 2 * rank, generator=g, device=device, dtype=f32) * float(np.sqrt(0.4 / (2 * rank)))
    rel = torch.randn(n_rel2, 2 * n, generator=g, device=device, dtype=f32) * 0.05
    rd = torch.rand(n_rel2, n, generator=g, device=device, dtype=f32) * 2 - 1
    c = torch.rand(n_rel2, 1, generator=g, device=device, dtype=f32) + 0.5
    h = torch.randint(0, n_ent, (nq,), generator=g, device=device)
    r = torch.randint(0, n_rel2, (nq,), generator=g, device=device)
    q_out = torch.empty(nq, 2 * rank, device=device, dtype=f32)
    c_out = torch.empty(nq, device=device, dtype=f32)
    byt = nq * (2 * 2 * rank * 4 + 16)
    ms_lane = _time_kernel(lambda: ops.query_fwd(ops.CHK_ROT, rank, True, ent, rel, rd, None, c, h, r, grouped=False, out=(q_out, c_out)))
    ms_tpq = _time_kernel(lambda: ops.query_fwd(ops.CHK_ROT, rank, True, ent, rel, rd, None, c, h, r, grouped=True, out=(q_out, c_out)))
    best = min(ms_lane, ms_tpq)
    out["roofline_k1"] = {"bound": "hbm", "achieved": byt / best / 1e6, "peak": hbm, "unit": "GB/s", "frac": byt / best / 1e6 / hbm, "traffic": None,
                          "kernel": "K1 query transform forward (FFTRotH rank 33 fp32, 2^20 queries, 1M-entity table), best of the lane-group kernel "
                                    "and the thread-per-query kernel incl. its counting sort by relation",
                          "ms_lane_group": ms_lane, "ms_thread_per_query": ms_tpq, "algorithmic_bytes_per_query": 2 * 2 * rank * 4 + 16,
                          "queries_per_launch": nq, "peak_source": src}
    gq = torch.randn(nq, 2 * rank, generator=g, device=device, dtype=f32)
    bufs = [torch.empty(nq, w, device=device, dtype=f32) for w in (2 * rank, 2 * n, n)] + [None, torch.empty(nq, device=device, dtype=f32)]
    bytb = nq * ((2 * 2 * rank + 2 * rank + 2 * n + n + 1) * 4 + 16)
    ms_b = _time_kernel(lambda: ops.query_bwd_into(ops.CHK_ROT, rank, True, ent, rel, rd, None, c, h, r, gq, *bufs))
    out["roofline_k1_bwd"] = {"bound": "hbm", "achieved": bytb / ms_b / 1e6, "peak": hbm, "unit": "GB/s", "frac": bytb / ms_b / 1e6 / hbm,
                              "traffic": None, "kernel": "K1 adjoint (lane-group kernel), same shape", "kernel_ms": ms_b,
                              "algorithmic_bytes_per_query": bytb // nq, "peak_source": src}
    del gq, bufs, q_out, c_out
    # ---- K3 training pass at B = 4096 (rank 33, neg 250) and B = 2048 (rank 257, neg 100)
    for key, rank, B, neg, n_ent in (("roofline_k3", 33, 4096, 250, 1_000_000), ("roofline_k3_r257", 257, 2048, 100, 1_000_000)):
        nt = neg + 1
        ent = torch.randn(n_ent, 2 * rank, generator=g, device=device, dtype=f32) * float(np.sqrt(0.4 / (2 * rank)))
        bh = torch.randn(n_ent, generator=g, device=device, dtype=f32) * 0.1
        bt = torch.randn(n_ent, generator=g, device=device, dtype=f32) * 0.1
        q = ent[torch.randint(0, n_ent, (B,), generator=g, device=device)].contiguous()
        heads = torch.randint(0, n_ent, (B,), generator=g, device=device)
        tails = torch.randint(0, n_ent, (B, nt), generator=g, device=device)
        hyper = torch.tensor([0.1, 1e-10, 1.0 / (B * nt), float(B), 0, 0, 1.0 / B, 0], dtype=torch.float64, device=device)
        lp, gs = torch.empty(B, device=device, dtype=f32), torch.empty(B, nt, device=device, dtype=f32)
        gq, coef, gbh = torch.empty(B, 2 * rank, device=device, dtype=f32), torch.empty(B * nt, 4, device=device, dtype=f32), torch.empty(B, device=device, dtype=f32)
        ms3 = _time_kernel(lambda: ops.score_gather_train(rank, B, nt, q, 1, 0, ent, tails, heads, 1, 0, bh, bt, hyper, lp, gs, gq, None, gbh,
                                                          pair_coef=coef))
        # per triple: (1+neg) gathered tail rows + per pair (tail id 8 B, bt 4 B, coefficients 16 B, d loss/d score 4 B) + the query row and its gradient
        by3 = B * (nt * (2 * rank * 4 + 32) + 2 * 2 * rank * 4)
        out[key] = {"bound": "hbm", "achieved": by3 / ms3 / 1e6, "peak": hbm, "unit": "GB/s", "frac": by3 / ms3 / 1e6 / hbm, "traffic": None,
                    "kernel": f"K3 training pass chk_score_gather_train (scores + loss + adjoint as pair coefficients, tail rows gathered once), "
                              f"rank {rank} fp32, B={B}, neg={neg}, 1M-entity table", "kernel_ms": ms3,
                    "algorithmic_bytes_per_triple": by3 // B, "triples_per_launch": B, "peak_source": src}
        del ent, bh, bt, q, tails, coef
    return out


# ------------------------------------------------------------------------------------------------ per-config legs
SMALL_CONFIGS = [
    # key, BASELINE.json index, workload, model, rank, dtype, train (optimizer, lr, neg, double_neg) or None
    ("configs[0]", 0, "wn18rr", "FFTRotH", 33, "float", ("Adam", 3e-4, 100, True)),
    ("configs[1]", 1, "fb237", "FFTRefH", 33, "float", ("Adagrad", 0.02, 250, False)),
    ("configs[2]", 2, "yago310", "FFTAttH", 33, "float", ("Adagrad", 0.02, 100, False)),
    ("configs[3]", 3, "wn18rr", "FFTRotH", 65, "double", None),
]


def small_config_legs(device, peaks, K, W, graphs, only=None):
    """BASELINE.json configs[0..3] on one GPU: filtered-eval queries/s (resident + e2e, own roofline entry) and, where the
    config names a training setup, triples/s through FusedKGOptimizer.  configs[3] (--dtype double) also states rank equality
    between the tensor-core tier (bf16x3 prefilter + fp64 re-check), the exact fp64 FMA tier and the CPU oracle."""
    from complexhyperbolickge_b200 import synthetic
    out = []
    for key, idx, workload, model_name, rank, dtype, train in SMALL_CONFIGS:
        if only is not None and idx not in only:
            continue
        if workload not in graphs:
            graphs[workload] = synthetic.make_graph(workload, seed=0)
        graph = graphs[workload]
        model = make_model(model_name, rank, dtype, graph, device)
        model.eval()
        entry = {"config": key, "workload": f"{model_name} rank={rank} {dtype}, synthetic {workload} ({graph['n_ent']} entities, "
                                            f"{graph['n_rel2'] // 2} relations), eval batch 500"}
        ev = measure_eval(model, graph, 500, K, W, device, None, 1, 0, 0, True, peaks)
        entry["eval"] = {"metric": METRIC, "unit": UNIT, **{k: ev[k] for k in ("value", "ms_per_step", "e2e", "roofline", "rank_algo", "gpu_launches", "mean_rank_check")}}
        if dtype == "double":
            entry["eval"]["fp64_check"] = fp64_rank_check(model, graph, model_name, rank)
        model.release_eval_cache()
        if train is not None:
            opt_name, lr, neg, dn = train
            entry["train"] = measure_train(model, graph, opt_name, lr, 500, neg, dn, 20 * K, 3, device, None, 1, peaks,
                                           contract_steps=10 if idx == 1 else 0)
        out.append(entry)
        del model
        torch.cuda.empty_cache()
    return out


def fp64_rank_check(model, graph, model_name, rank, n_q=24):
    """configs[3]: ranks of the tensor-core tier == ranks of the exact fp64 FMA tier on 500 queries, and == the CPU oracle
    (fp64 restatement of the reference) on n_q queries; the oracle call doubles as this config's CPU baseline."""
    from complexhyperbolickge_b200 import synthetic
    from oracle import chk_oracle as O
    test = torch.from_numpy(graph["test"][:500])
    findex = graph["filters"]["rhs"]
    keep = model.rank_algo
    model.rank_algo = "mma"
    r_mma = model.get_ranking(test, findex, batch_size=500)
    model.rank_algo = "fma"
    r_fma = model.get_ranking(test, findex, batch_size=500)
    model.rank_algo = keep
    model.release_eval_cache()
    p = O.Params.from_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()}, O.KIND[model_name], rank, True)
    qs = test[:n_q]
    fd = synthetic.filter_dict(findex)
    filters = {(int(h), int(r)): fd[(int(h), int(r))] for h, r, _ in qs.tolist()}
    torch.set_num_threads(os.cpu_count())
    t0 = time.perf_counter()
    r_or = O.get_ranking(p, qs, filters, batch_size=8)
    dt = time.perf_counter() - t0
    return {"ranks_mma_equal_exact_fp64_tier": bool(torch.equal(r_mma, r_fma)), "ranks_equal_cpu_oracle_fp64": bool(torch.equal(r_mma[:n_q], r_or)),
            "queries_checked": [500, n_q],
            "cpu_baseline": {"value": n_q / dt, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                             "sample": f"{n_q} queries against all {graph['n_ent']} entities, {dt:.1f} s of CPU work (oracle port, fp64)"}}


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
        dist.init_process_group("nccl", device_id=device)
        pg = dist.group.WORLD
    from complexhyperbolickge_b200 import ops, synthetic

    peaks = load_peaks()
    cfg, model_name, rank = config_dict(args, world)
    graphs = {args.workload: synthetic.make_graph(args.workload, seed=0)}
    graph = graphs[args.workload]
    n_ent, n_rel2 = graph["n_ent"], graph["n_rel2"]
    model = make_model(model_name, rank, args.dtype, graph, device)
    model.eval()
    algo = args.rank_algo
    if algo == "auto":
        algo = "mma" if ops.mma_available() else "fma"
    model.rank_algo = algo
    model.process_group = pg
    b, K, W = args.batch, args.steps, args.warmup

    wd.enter("evaluation steps (resident inputs) + roofline probe + end-to-end get_ranking", 300)
    ev = measure_eval(model, graph, b, K, W, device, pg, world, rank_id, local, not args.no_e2e, peaks, wd)

    # ---- the JSON line so far (rank 0); the legs below only add keys to it
    line = None
    if rank_id == 0:
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp) and world == 1 and args.workload == "big4m" and ev["rank_algo"] == "mma":
            ev["roofline"]["traffic"] = json.load(open(tp)).get("rank_mma_kernel_big4m_dram_bytes_per_launch")
        line = {"metric": METRIC, "value": ev["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ev["ms_per_step"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32" if args.dtype == "float" else "f64", "data": "synthetic", "config": cfg, "clocks": ev["clocks"],
                "e2e": ev["e2e"], "gpu_launches": ev["gpu_launches"], "roofline": ev["roofline"],
                "mean_rank_check": ev["mean_rank_check"], "rank_algo": ev["rank_algo"], "filter_len": ev["filter_len"],
                "train_examples_in_graph": int(graph["train"].shape[0]) * 2}
        wd.partial = line

    # ---- training legs (every rank takes part); each has its own deadline, a stuck leg is reported, not waited for
    model.release_eval_cache()
    model.process_group = None
    torch.cuda.empty_cache()

    def leg(key, what, seconds, fn, *a, **kw):
        """A leg adds a key to the line; if it fails the headline (evaluation) numbers are still reported."""
        import traceback
        wd.enter(what, seconds)
        try:
            out = fn(*a, **kw)
        except Exception as e:                       # noqa: BLE001 — reported in the line, stack on stderr
            traceback.print_exc(file=sys.stderr)
            out = {"error": f"{type(e).__name__}: {e}"[:400]}
        if line is not None:
            line[key] = out

    if not args.no_train:
        if args.workload == "big4m":
            leg("train_big4m", "training leg on the 4M-entity table (configs[4])", 240, measure_train, model, graph, "Adagrad", 0.02,
                500, 100, False, 50, 5, device, pg, world, peaks)
            if line is not None and "error" not in line["train_big4m"]:
                line["train_big4m"]["config"] = (f"BASELINE.json configs[4]: FFTRotH rank={rank} Adagrad, 500 triples per rank x{world}, neg=100, "
                                                 "synthetic 4M-entity graph, " + ("owner-sharded tables (symmetric memory)" if line["train_big4m"].get("owner_sharded") else "replicated tables"))
    del model
    torch.cuda.empty_cache()
    if not args.no_train:
        def cfg1_train():
            if "fb237" not in graphs:
                graphs["fb237"] = synthetic.make_graph("fb237", seed=0)
            m1 = make_model("FFTRefH", 33, "float", graphs["fb237"], device)
            out = measure_train(m1, graphs["fb237"], "Adagrad", 0.02, 500, 250, False, 200, 5, device, pg, world, peaks,
                                contract_steps=10 if world == 1 else 0)
            out["config"] = f"BASELINE.json configs[1]: FFTRefH rank=33 Adagrad neg=250, 500 triples per rank x{world}, synthetic FB15k-237 shape"
            return out
        leg("train", "training leg (configs[1] shape)", 180, cfg1_train)
    if world > 1:
        import gc
        wd.enter("process-group teardown", 60)
        gc.collect()                                  # captured CUDA graphs hold NCCL work: they must be gone before the communicator
        torch.cuda.synchronize()
        if rank_id != 0:
            dist.destroy_process_group()
            wd.stop()
            return
    if world == 1 and not args.no_configs:
        leg("configs", "BASELINE.json configs[0..3] (eval + train per config)", 420, small_config_legs, device, peaks, K, W, graphs)
        leg("kernel_rooflines", "K1 / K3 kernel rooflines", 120, kernel_rooflines, device, peaks)
        if "error" not in line["kernel_rooflines"]:
            line.update(line.pop("kernel_rooflines"))
    if world == 1 and not args.no_cpu_baseline:
        wd.enter("CPU baseline (oracle port on the host cores)", 300)
        v, spent, sample = oracle_eval_budget(model_name, rank, args.dtype, n_ent, n_rel2, budget_s=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample}
        if not args.no_train and "error" not in line.get("train", {"error": 1}):
            line["train"]["cpu_baseline"] = oracle_train_sample()
    wd.emit(line)
    if world > 1:
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

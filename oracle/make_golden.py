"""Generate tests/golden/*.npz by running the UNMODIFIED reference in this container.

    python oracle/make_golden.py            # needs /root/reference; writes tests/golden/

The fixtures pin the oracle (tests/test_oracle_golden.py, CPU) and the CUDA path
(tests/test_gpu_*.py, via the C-ABI).  Everything is seeded; regenerate only when the
fixture layout changes.  TEST INFRASTRUCTURE ONLY.
"""
import copy
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
MODELS = ["FFTRotH", "FFTRefH", "FFTAttH"]


def set_regime(model, regime, gen):
    """Three weight regimes of SURVEY §8c."""
    dt = model.entity.weight.dtype
    r = model.rank
    with torch.no_grad():
        def rn(t, std):
            t.copy_((torch.randn(t.shape, generator=gen, dtype=torch.float64) * std).to(dt))

        def ru(t, lo, hi):
            t.copy_((torch.rand(t.shape, generator=gen, dtype=torch.float64) * (hi - lo) + lo).to(dt))
        if regime == "init":
            rn(model.entity.weight, 1e-3)
            rn(model.rel.weight, 1e-3)
            ru(model.rel_diag.weight, -1, 1)
            model.c.weight.fill_(1.0)
            if hasattr(model, "context_vec"):
                rn(model.context_vec.weight, 1e-3)
            return
        std = {"trained": np.sqrt(0.4 / (2 * r)), "boundary": 0.3}[regime]
        rn(model.entity.weight, std)
        rn(model.rel.weight, {"trained": 0.05, "boundary": 2.0}[regime])
        ru(model.rel_diag.weight, -1, 1)
        ru(model.c.weight, 0.5, 2.0)
        rn(model.bh.weight, 0.1)
        rn(model.bt.weight, 0.1)
        if hasattr(model, "context_vec"):
            rn(model.context_vec.weight, 1.0)


def sd_np(model, prefix="p_"):
    return {prefix + k.replace(".weight", ""): v.detach().cpu().numpy().copy() for k, v in model.state_dict().items()}


def step_case(ref_models, name, dtype, multi_c, regime, rank, n_ent, n_rel2, B, neg, seed):
    gen = torch.Generator().manual_seed(seed)
    model = ref_shim.make_model(ref_models, name, n_ent, n_rel2, rank, dtype, multi_c)
    set_regime(model, regime, gen)
    batch = torch.stack([torch.randint(0, n_ent, (B,), generator=gen),
                         torch.randint(0, n_rel2, (B,), generator=gen),
                         torch.randint(0, n_ent, (B,), generator=gen)], 1)
    batch[1, 0] = batch[0, 0]          # duplicate head, duplicate relation: exercise the scatter-add
    batch[1, 1] = batch[0, 1]
    negs = torch.randint(0, n_ent - 1, (B, neg), generator=gen)
    negs = torch.where(negs < batch[:, 2:3], negs, negs + 1)
    out = sd_np(model)
    out.update(batch=batch.numpy(), neg=negs.numpy())
    # forward pieces
    (q, c), bh = model.get_queries(batch[:, :2].unsqueeze(1))
    out.update(q=q.detach().numpy(), c=c.detach().numpy())
    # eval-style scores against all candidates
    with torch.no_grad():
        lhs = model.get_queries(batch[:, :2])
        out["score_all"] = model.score(lhs, model.get_rhs(None)).numpy()
    # the reference loss (kg_optimizer.py:101-123) with injected negatives
    model.zero_grad()
    pos, _ = model(batch[:, :2].unsqueeze(1), batch[:, 2].unsqueeze(1))
    ngs, _ = model(batch[:, :2].unsqueeze(1), negs)
    loss = -torch.cat([F.logsigmoid(pos).view(-1), F.logsigmoid(-ngs).view(-1)]).mean()
    loss.backward()
    out.update(score_pos=pos.detach().numpy(), score_neg=ngs.detach().numpy(), loss=np.array(loss.item()))
    for k, p in model.named_parameters():
        out["g_" + k.replace(".weight", "")] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy().copy()
    out["meta"] = np.array([name, dtype, str(int(multi_c)), regime, str(rank), str(n_ent), str(n_rel2)])
    return out


def toy_graph(n_ent, n_rel, n_train, n_test, seed):
    rng = np.random.default_rng(seed)
    pop = 1.0 / np.arange(1, n_ent + 1)
    pop /= pop.sum()
    tot = n_train + 2 * n_test
    tr = np.stack([rng.choice(n_ent, tot * 2, p=pop), rng.integers(0, n_rel, tot * 2),
                   rng.choice(n_ent, tot * 2, p=pop)], 1)
    tr = np.unique(tr, axis=0)
    rng.shuffle(tr)
    tr = tr[:tot]
    train, valid, test = tr[:n_train], tr[n_train:n_train + n_test], tr[n_train + n_test:]
    lhs, rhs = {}, {}
    for h, r, t in tr:                                  # datasets/process.py:55-77
        rhs.setdefault((int(h), int(r)), set()).add(int(t))
        lhs.setdefault((int(t), int(r + n_rel)), set()).add(int(h))
    filters = {"lhs": {k: sorted(v) for k, v in lhs.items()}, "rhs": {k: sorted(v) for k, v in rhs.items()}}
    return train.astype(np.int64), valid.astype(np.int64), test.astype(np.int64), filters


def filters_to_arrays(filters):
    out = {}
    for side in ("lhs", "rhs"):
        keys = sorted(filters[side].keys())
        out[side + "_keys"] = np.array(keys, dtype=np.int64).reshape(-1, 2)
        lens = np.array([len(filters[side][k]) for k in keys], dtype=np.int64)
        out[side + "_indptr"] = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        out[side + "_vals"] = np.array([v for k in keys for v in filters[side][k]], dtype=np.int64)
    return out


def ranking_case(ref_models, name, dtype, regime, seed, multi_c=True, rank=9, n_ent=300, n_rel=8):
    gen = torch.Generator().manual_seed(seed)
    train, valid, test, filters = toy_graph(n_ent, n_rel, 1500, 120, seed)
    model = ref_shim.make_model(ref_models, name, n_ent, 2 * n_rel, rank, dtype, multi_c)
    set_regime(model, regime, gen)
    model.eval()
    out = sd_np(model)
    out.update(filters_to_arrays(filters))
    ex = torch.from_numpy(test)
    f = copy.deepcopy(filters)
    ranks_rhs = model.get_ranking(ex, f["rhs"], batch_size=50)
    q = torch.stack([ex[:, 2], ex[:, 1] + n_rel, ex[:, 0]], -1)
    ranks_lhs = model.get_ranking(q, f["lhs"], batch_size=50)
    mr, mrr, hits = model.compute_metrics(ex, copy.deepcopy(filters), batch_size=50)
    out.update(test=test, ranks_rhs=ranks_rhs.numpy(), ranks_lhs=ranks_lhs.numpy(),
               mr=np.array([mr["rhs"], mr["lhs"]]), mrr=np.array([mrr["rhs"], mrr["lhs"]]),
               hits=np.stack([hits["rhs"].numpy(), hits["lhs"].numpy()]))
    out["meta"] = np.array([name, dtype, str(int(multi_c)), regime, str(rank), str(n_ent), str(2 * n_rel)])
    return out


def loss_curve_case(ref_models, ref_reg, RefKGOptimizer, name, optim_name, lr, dtype="double", seed=0,
                    rank=9, n_ent=200, n_rel=6, B=64, neg=10, epochs=3):
    torch.manual_seed(seed)
    train, valid, test, filters = toy_graph(n_ent, n_rel, 600, 40, seed + 7)
    inv = train[:, [2, 1, 0]].copy()
    inv[:, 1] += n_rel
    examples = torch.from_numpy(np.vstack([train, inv]))            # datasets/kg_dataset.py:54-60
    model = ref_shim.make_model(ref_models, name, n_ent, 2 * n_rel, rank, dtype, True)
    out = sd_np(model, "p0_")
    opt = RefKGOptimizer(model, ref_reg.N3(0.0), getattr(torch.optim, optim_name)(model.parameters(), lr=lr),
                         batch_size=B, update_steps=1, neg_sample_size=neg, double_neg=False, verbose=False)
    opt.device = torch.device("cpu")
    neg_gen = torch.Generator().manual_seed(seed + 1)
    batches, negs, step_losses = [], [], []

    def get_neg(input_batch):                                       # same contract as kg_optimizer.py:92-99
        ns = torch.randint(0, n_ent - 1, (input_batch.shape[0], neg), generator=neg_gen)
        ns = torch.where(ns < input_batch[:, 2].unsqueeze(-1), ns, ns + 1)
        batches.append(input_batch.clone())
        negs.append(ns.clone())
        return ns
    opt.get_neg_samples = get_neg
    orig = opt.calculate_loss

    def calc(b):
        l = orig(b)
        step_losses.append(l.item())
        return l
    opt.calculate_loss = calc
    epoch_losses = []
    model.train()
    for ep in range(epochs):
        epoch_losses.append(opt.epoch(examples))
    nb = len(batches)
    lens = np.array([b.shape[0] for b in batches])
    out.update(sd_np(model, "pT_"))
    out.update(batch_cat=torch.cat(batches).numpy(), neg_cat=torch.cat(negs).numpy(), batch_lens=lens,
               step_losses=np.array(step_losses), epoch_losses=np.array(epoch_losses), lr=np.array(lr))
    out["meta"] = np.array([name, dtype, "1", optim_name, str(rank), str(n_ent), str(2 * n_rel)])
    assert nb == len(step_losses)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_models, ref_reg, RefKGOptimizer = ref_shim.load()
    torch.set_num_threads(4)
    seed = 100
    n = 0
    for name in MODELS:
        for dtype in ("double", "float"):
            for multi_c in (True, False):
                for regime in ("init", "trained", "boundary"):
                    seed += 1
                    case = step_case(ref_models, name, dtype, multi_c, regime, 9, 40, 6, 6, 4, seed)
                    np.savez_compressed(os.path.join(OUT, f"step_{name}_{dtype}_mc{int(multi_c)}_{regime}_r9.npz"), **case)
                    n += 1
        for rank, n_ent in ((33, 48), (65, 40)):
            seed += 1
            case = step_case(ref_models, name, "double", True, "trained", rank, n_ent, 6, 5, 3, seed)
            np.savez_compressed(os.path.join(OUT, f"step_{name}_double_mc1_trained_r{rank}.npz"), **case)
            n += 1
        seed += 1
        case = step_case(ref_models, name, "float", True, "trained", 33, 48, 6, 5, 3, seed)
        np.savez_compressed(os.path.join(OUT, f"step_{name}_float_mc1_trained_r33.npz"), **case)
        n += 1
    for name in MODELS:
        for dtype in ("double", "float"):
            for regime in ("trained", "init"):
                seed += 1
                case = ranking_case(ref_models, name, dtype, regime, seed)
                np.savez_compressed(os.path.join(OUT, f"rank_{name}_{dtype}_{regime}.npz"), **case)
                n += 1
    for name, on, lr in (("FFTRotH", "Adam", 3e-4), ("FFTRefH", "Adagrad", 0.02), ("FFTAttH", "Adagrad", 0.03)):
        case = loss_curve_case(ref_models, ref_reg, RefKGOptimizer, name, on, lr)
        np.savez_compressed(os.path.join(OUT, f"curve_{name}_{on}.npz"), **case)
        print(name, on, "epoch losses", case["epoch_losses"])
        n += 1
    print("wrote", n, "fixtures to", OUT)


def main_high_rank():
    """Supplement (does not touch the fixtures main() wrote): per-step forward values + autograd gradients of the
    reference at the ranks of BASELINE.json configs[3..4] (129, 257), fp64, trained-like regime, tiny tables."""
    os.makedirs(OUT, exist_ok=True)
    ref_models, _ref_reg, _RefKGOptimizer = ref_shim.load()
    torch.set_num_threads(4)
    seed, n = 900, 0
    for name in MODELS:
        for rank in (129, 257):
            seed += 1
            case = step_case(ref_models, name, "double", True, "trained", rank, 24, 4, 4, 3, seed)
            np.savez_compressed(os.path.join(OUT, f"step_{name}_double_mc1_trained_r{rank}.npz"), **case)
            n += 1
    print("wrote", n, "high-rank fixtures to", OUT)


if __name__ == "__main__":
    if "--high-rank" in sys.argv:
        main_high_rank()
    else:
        main()

"""CPU: host-side logic (filter CSR, shard bounds, model construction / state_dict layout) and the C-ABI
surface (library loads, every symbol declared in include/chk_b200.h is exported).  No compute calls."""
import ctypes
import os
import re
from argparse import Namespace

import numpy as np
import pytest
import torch

from conftest import ROOT, filters_from_arrays, golden_files, load_case


def test_capi_exports_every_declared_symbol():
    from complexhyperbolickge_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "chk_b200.h")).read()
    declared = set(re.findall(r"\b(chk_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build_native()
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/chk_b200.h but not exported"
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.lib().chk_abi_version() == 4


def test_no_cpu_fallback():
    from complexhyperbolickge_b200 import ops
    with pytest.raises(RuntimeError):
        ops.row_hnorm(9, torch.zeros(4, 18))
    import complexhyperbolickge_b200 as chk
    m = chk.FFTRotH(Namespace(sizes=(10, 2, 10), rank=9, dropout=0, gamma=0, dtype="float", bias="learn",
                              init_size=1e-3, multi_c=True))
    with pytest.raises(RuntimeError):
        m.get_queries(torch.zeros((2, 2), dtype=torch.int64))
    with pytest.raises(RuntimeError):
        m.get_ranking(torch.zeros((1, 3), dtype=torch.int64), {(0, 0): [0]}, 4)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "complexhyperbolickge_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle|import_module\([\"']oracle|oracle/", txt, re.M), f


@pytest.mark.parametrize("name", ["FFTRotH", "FFTRefH", "FFTAttH"])
@pytest.mark.parametrize("multi_c", [True, False])
def test_state_dict_layout_matches_reference(name, multi_c):
    """SURVEY §5 checkpoint row: keys and shapes must be interchangeable with the reference's model.pt."""
    import complexhyperbolickge_b200 as chk
    N, R2, r = 30, 6, 9
    m = getattr(chk, name)(Namespace(sizes=(N, R2, N), rank=r, dropout=0, gamma=0, dtype="double", bias="learn",
                                     init_size=1e-3, multi_c=multi_c))
    n = 2 * (r - 1)
    want = {"entity.weight": (N, 2 * r), "rel.weight": (R2, 2 * n), "rel_diag.weight": (R2, 2 * n if name == "FFTAttH" else n),
            "c.weight": (R2 if multi_c else 1, 1), "bh.weight": (N, 1), "bt.weight": (N, 1)}
    if name == "FFTAttH":
        want["context_vec.weight"] = (R2, n)
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    assert all(v.dtype == torch.float64 for v in m.state_dict().values())
    # golden fixture (reference state_dict) loads strictly
    case = load_case([p for p in golden_files("step_" + name) if f"mc{int(multi_c)}" in p and "double" in p][0])
    m2 = getattr(chk, name)(Namespace(sizes=(case["n_ent"], case["n_rel2"], case["n_ent"]), rank=case["rank"], dropout=0,
                                      gamma=0, dtype="double", bias="learn", init_size=1e-3, multi_c=multi_c))
    m2.load_state_dict({k[2:] + ".weight": torch.from_numpy(v) for k, v in case.items() if k.startswith("p_")})


def test_filter_index_matches_reference_semantics():
    from complexhyperbolickge_b200.filters import FilterIndex
    case = load_case(golden_files("rank_FFTRotH_double_trained")[0])
    filters = filters_from_arrays(case)
    fi = FilterIndex.from_dict(filters["rhs"], case["n_rel2"])
    fa = FilterIndex.from_arrays(case["rhs_keys"], case["rhs_indptr"], case["rhs_vals"], case["n_rel2"])
    qs = case["test"][:57]
    for f in (fi, fa):
        indptr, idx = f.batch_csr(qs)
        assert indptr.shape == (58,) and indptr[-1] == idx.size
        for i, (h, r, t) in enumerate(qs):
            want = sorted(set(filters["rhs"][(int(h), int(r))]) | {int(t)})      # base.py:266-267, deduplicated
            assert idx[indptr[i]:indptr[i + 1]].tolist() == want
    with pytest.raises(KeyError):
        fi.batch_csr(np.array([[10 ** 6, 0, 1]]))
    # ragged: empty filter dict, strict off -> just the targets
    e = FilterIndex.from_dict({}, 4)
    indptr, idx = e.batch_csr(np.array([[1, 2, 3], [4, 1, 0]]), strict=False)
    assert indptr.tolist() == [0, 1, 2] and idx.tolist() == [3, 0]


def test_shard_bounds_partition():
    from complexhyperbolickge_b200.ranking import shard_bounds
    for n in (1, 127, 128, 40943, 4_000_000):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert all(lo % 128 == 0 or lo == n for lo, _ in spans)


# ---- dataset reader: the reference's pickle format, golden made by the reference itself ----------
def test_dataset_reader_matches_reference_golden(tmp_path):
    import pickle as pkl
    from complexhyperbolickge_b200.datasets import KGDataset, build_filters, write_dataset
    from complexhyperbolickge_b200.filters import FilterIndex
    gold_dir = os.path.join(os.path.dirname(__file__), "golden")
    d = os.path.join(gold_dir, "dataset_toy")
    exp = np.load(os.path.join(gold_dir, "dataset_toy_expected.npz"))
    ds = KGDataset(d, False)
    assert tuple(exp["shape"]) == ds.get_shape()
    assert np.array_equal(ds.get_examples("train").numpy(), exp["train_examples"])        # reciprocal augmentation
    assert np.array_equal(ds.get_examples("test").numpy(), exp["test_examples"])
    assert np.array_equal(ds.get_examples("train", rel_idx=2).numpy(), exp["rel2_examples"])
    # build_filters == the reference's get_filters (stored in to_skip.pickle by the golden script)
    allx = np.concatenate([ds.data["train"], ds.data["valid"], ds.data["test"]], 0)
    lhs, rhs = build_filters(allx, ds.n_predicates // 2)
    assert lhs == ds.to_skip["lhs"] and rhs == ds.to_skip["rhs"]
    # write_dataset round trip: byte-compatible content
    write_dataset(str(tmp_path / "ds"), ds.data["train"], ds.data["valid"], ds.data["test"], ds.n_predicates // 2)
    ds2 = KGDataset(str(tmp_path / "ds"), False)
    assert ds2.to_skip == ds.to_skip and np.array_equal(ds2.data["valid"], ds.data["valid"])
    # CSR form: every evaluated query finds exactly the stored list (plus its own tail)
    fi = ds.filter_indices()
    assert isinstance(fi["rhs"], FilterIndex)
    test = ds.get_examples("test").numpy()
    indptr, idx = fi["rhs"].batch_csr(test)
    for i, (h, r, t) in enumerate(test):
        assert idx[indptr[i]:indptr[i + 1]].tolist() == sorted(set(ds.to_skip["rhs"][(int(h), int(r))]) | {int(t)})
    lhs_q = np.stack([test[:, 2], test[:, 1] + ds.n_predicates // 2, test[:, 0]], 1)
    indptr, idx = fi["lhs"].batch_csr(lhs_q)
    for i, (h, r, t) in enumerate(lhs_q):
        assert idx[indptr[i]:indptr[i + 1]].tolist() == sorted(set(ds.to_skip["lhs"][(int(h), int(r))]) | {int(t)})
    with pytest.raises(KeyError):
        fi["rhs"].batch_csr(np.array([[59, 9, 0]]))


# ---- bench.py plumbing that needs no GPU ------------------------------------------------------------------------------
def _load_bench():
    import importlib.util
    spec = importlib.util.spec_from_file_location("chk_bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_bench_watchdog_emits_partial_line_and_exits_nonzero():
    """A phase that exceeds its deadline must not hang the job: stacks to stderr, the line measured so far + "error" on
    stdout (rank 0), and a NON-zero exit code (3): a hung phase is a failure even though the partial line is kept."""
    import subprocess
    import sys
    code = ("import sys, time, importlib.util\n"
            f"spec = importlib.util.spec_from_file_location('b', r'{os.path.join(ROOT, 'bench.py')}')\n"
            "m = importlib.util.module_from_spec(spec); spec.loader.exec_module(m)\n"
            "wd = m.Watchdog(int(sys.argv[1])); wd.partial = {'metric': m.METRIC, 'value': 123.0}\n"
            "wd.enter('fake stuck phase', 1); time.sleep(20); print('NOT REACHED')\n")
    for rank, expect_line in ((0, True), (3, False)):
        r = subprocess.run([sys.executable, "-c", code, str(rank)], capture_output=True, text=True, timeout=60)
        assert r.returncode == 3 and "NOT REACHED" not in r.stdout
        assert "fake stuck phase" in r.stderr
        if expect_line:
            import json
            line = json.loads(r.stdout.strip())
            assert line["value"] == 123.0 and "fake stuck phase" in line["error"]
        else:
            assert r.stdout.strip() == ""


def test_bench_reference_arm_prints_one_json_line():
    """`bench.py --impl reference` (the oracle port on the host cores) on a small workload: one JSON line with the
    contract's keys; non-zero ranks print nothing."""
    import json
    import subprocess
    import sys
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "wn18rr", "--steps", "1", "--warmup", "0"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="0"))
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "filtered_eval_queries_per_sec" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, RANK="1", WORLD_SIZE="2"))
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_bench_cycles_examples_instead_of_slicing_past_the_end():
    """r1 N=8 hang: the DP leg sliced 812,000 rows out of 544,230 examples, so late steps got short / empty batches and the
    ranks issued collectives of different sizes.  Batches now wrap around: every batch has exactly B rows."""
    bench = _load_bench()
    ex = torch.arange(30).view(10, 3)
    out = bench.cycle_batches(ex, 7, 4)
    assert out.shape == (7, 4, 3)
    assert torch.equal(out.view(-1, 3)[:10], ex) and torch.equal(out.view(-1, 3)[10:20], ex)


def test_header_is_c_and_struct_layouts_match_ctypes(tmp_path):
    """include/chk_b200.h compiles as plain C (the boundary is a C ABI) and the structs it declares have the size / field
    offsets of their ctypes mirrors in _lib.py (an ABI drift between the header and the binding would corrupt arguments)."""
    import ctypes
    import shutil
    import subprocess
    from complexhyperbolickge_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "chk_b200.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(chk_red_col), sizeof(chk_red_group), '
                   'sizeof(chk_eval_args), sizeof(chk_dense_tab), sizeof(chk_table_desc), offsetof(chk_red_col, pair_coef), '
                   'offsetof(chk_red_group, cols), offsetof(chk_eval_args, scratch)); return 0; }\n')
    exe = tmp_path / "abi"
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", inc, str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    want = [ctypes.sizeof(_lib.RedCol), ctypes.sizeof(_lib.RedGroup), ctypes.sizeof(_lib.EvalArgs), ctypes.sizeof(_lib.DenseTab),
            ctypes.sizeof(_lib.TableDesc), _lib.RedCol.pair_coef.offset, _lib.RedGroup.cols.offset, _lib.EvalArgs.scratch.offset]
    assert got == want

import glob
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def golden_files(prefix):
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_case(path):
    z = np.load(path, allow_pickle=False)
    case = {k: z[k] for k in z.files}
    meta = [str(x) for x in case["meta"]]
    case["name"], case["dtype"], case["multi_c"] = meta[0], meta[1], bool(int(meta[2]))
    case["regime"], case["rank"], case["n_ent"], case["n_rel2"] = meta[3], int(meta[4]), int(meta[5]), int(meta[6])
    return case


def oracle_params(case, prefix="p_"):
    from oracle import chk_oracle as O
    t = lambda k: torch.from_numpy(case[prefix + k].copy()) if (prefix + k) in case else None
    return O.Params(O.KIND[case["name"]], case["rank"], case["multi_c"], t("entity"), t("rel"), t("rel_diag"),
                    t("c"), t("bh"), t("bt"), t("context_vec"))


def filters_from_arrays(case):
    out = {}
    for side in ("lhs", "rhs"):
        keys, indptr, vals = case[side + "_keys"], case[side + "_indptr"], case[side + "_vals"]
        out[side] = {(int(k[0]), int(k[1])): [int(v) for v in vals[indptr[i]:indptr[i + 1]]]
                     for i, k in enumerate(keys)}
    return out


# ---- observed parity margins: every GPU parity test records its worst observed error; written once per session to
# gpurun_out/parity_observed.json (the directory that travels back from the GPU box) and committed under profiles/.
_PARITY = {}


def record_parity(key, **vals):
    slot = _PARITY.setdefault(key, {})
    for k, v in vals.items():
        v = float(v)
        slot[k] = max(slot.get(k, v), v)


def pytest_sessionfinish(session, exitstatus):
    if not _PARITY:
        return
    import json
    out = os.path.join(ROOT, "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_observed.json"), "w") as f:
        json.dump(_PARITY, f, indent=1, sort_keys=True)

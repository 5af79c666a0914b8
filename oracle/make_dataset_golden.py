"""Golden fixture for the dataset reader (complexhyperbolickge_b200/datasets.py): a tiny dataset directory in the
reference's pickle format, with its filters built by the REFERENCE's own datasets/process.py:get_filters and its
training examples produced by the reference's own datasets/kg_dataset.py:KGDataset.get_examples, run unmodified in
the build container.  TEST INFRASTRUCTURE ONLY (like make_golden.py).

    python oracle/make_dataset_golden.py        # writes tests/golden/dataset_toy/*.pickle + dataset_toy_expected.npz
"""
import os
import pickle as pkl
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def main():
    ref_shim.load()                                   # mocks torch_geometric & co, puts /root/reference on sys.path
    import datasets.process as ref_process            # reference, unmodified
    from datasets.kg_dataset import KGDataset as RefKGDataset
    rng = np.random.default_rng(7)
    n_ent, n_rel = 60, 5
    tri = np.unique(np.stack([rng.integers(0, n_ent, 700), rng.integers(0, n_rel, 700), rng.integers(0, 12, 700)], 1), axis=0)
    tri[0] = [n_ent - 1, n_rel - 1, 3]                # make sure max ids appear in train (shape is inferred from train)
    rng.shuffle(tri[1:])
    train, valid, test = tri[:400], tri[400:480], tri[480:560]
    d = os.path.join(OUT, "dataset_toy")
    os.makedirs(d, exist_ok=True)
    for name, arr in (("train", train), ("valid", valid), ("test", test)):
        with open(os.path.join(d, name + ".pickle"), "wb") as f:
            pkl.dump(arr.astype("int64"), f)
    lhs, rhs = ref_process.get_filters(np.concatenate([train, valid, test], 0), n_rel)
    lhs = {(int(k[0]), int(k[1])): [int(x) for x in v] for k, v in lhs.items()}
    rhs = {(int(k[0]), int(k[1])): [int(x) for x in v] for k, v in rhs.items()}
    with open(os.path.join(d, "to_skip.pickle"), "wb") as f:
        pkl.dump({"lhs": lhs, "rhs": rhs}, f)
    ref = RefKGDataset(d, False)
    np.savez(os.path.join(OUT, "dataset_toy_expected.npz"), train_examples=ref.get_examples("train").numpy(),
             test_examples=ref.get_examples("test").numpy(), rel2_examples=ref.get_examples("train", rel_idx=2).numpy(),
             shape=np.array(ref.get_shape()))
    print("wrote", d)


if __name__ == "__main__":
    main()

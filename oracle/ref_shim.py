"""Import the UNMODIFIED reference (/root/reference) in the build container.  TEST INFRASTRUCTURE ONLY.

Used by ``oracle/make_golden.py`` (and nothing else) to run the reference itself and dump golden
vectors.  /root/reference does not exist on the GPU box, so nothing at run time may import this.

Recipe: mock the GNN-only dependencies that are not installed
(torch_scatter, torch_geometric*), never write bytecode into the read-only tree, and set the
instance attribute ``model.lift = True`` (HEAD's default ``lift=False`` crashes).
"""
import sys
from argparse import Namespace
from unittest.mock import MagicMock

REF_ROOT = "/root/reference"


def load():
    sys.dont_write_bytecode = True
    for m in ["torch_scatter", "torch_geometric", "torch_geometric.data", "torch_geometric.loader",
              "torch_geometric.utils", "torch_geometric.utils.map", "torch_geometric.utils.num_nodes",
              "torch_geometric.utils.mask", "torch_geometric.typing"]:
        sys.modules.setdefault(m, MagicMock())
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import models as ref_models                                   # noqa
    import optimizers.regularizers as ref_reg                     # noqa
    from optimizers.kg_optimizer import KGOptimizer as RefKGOptimizer  # noqa
    return ref_models, ref_reg, RefKGOptimizer


def make_model(ref_models, name, n_ent, n_rel2, rank, dtype="double", multi_c=True, bias="learn",
               init_size=1e-3):
    args = Namespace(sizes=(n_ent, n_rel2, n_ent), rank=rank, dropout=0, gamma=0, dtype=dtype,
                     bias=bias, init_size=init_size, multi_c=multi_c)
    model = getattr(ref_models, name)(args)
    model.lift = True
    return model

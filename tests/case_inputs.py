"""Builds the inputs of a golden case WITHOUT the reference (usable on the GPU box)."""
import json

import numpy as np
import torch

from tests import golden_util as GU
from gen_fvgn_steady_b200.mesh import synthetic as S


def case_meshes(name):
    """-> (meshes: list of numpy dicts, uvps: list of float32 arrays, npz)."""
    case = GU.CASES[name]
    z = GU.load_case(name)
    if isinstance(case["mesh"], str):
        meshes, uvps = GU.example_case_graphs(z)
        return meshes, uvps, z
    meshes, uvps = [], []
    for i, spec in enumerate(case["mesh"]):
        m, uvp = S.make_case(**spec)
        assert bytes.fromhex(GU.index_hash(m)) == z[f"index_hash.{i}"].tobytes(), "synthetic mesh drifted from golden"
        meshes.append(m)
        uvps.append(GU.perturbed_field(uvp, spec["seed"]))
    return meshes, uvps, z


def case_state_dict(z, dtype=torch.float32):
    shapes = {k: tuple(json.loads(s)) for k, s in zip(z["state_keys"].tolist(), z["state_shapes"].tolist())}
    return GU.golden_state_dict(shapes, dtype=dtype)


def product_graphs(meshes, uvps, device="cpu"):
    """The five batched graph objects (Graph_loader.py:503-784 + datapreprocessing :130-152)."""
    from gen_fvgn_steady_b200.mesh.batching import graphs_from_meshes
    return graphs_from_meshes(meshes, uvps, device)

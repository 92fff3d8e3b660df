"""The device-side synthetic mesh pipeline (mesh/synthetic_torch.py) against the numpy pipeline (mesh/synthetic.py, which is
pinned bit-exactly to the reference's converter by tests/golden index hashes): indices identical, floats to ~1 ulp."""
import numpy as np
import pytest
import torch

from gen_fvgn_steady_b200.mesh import synthetic as S
from gen_fvgn_steady_b200.mesh import synthetic_torch as ST

INT_KEYS = ("node|node_type", "face|face_node", "cells_node", "cells_index", "cells_face", "face|face_type",
            "face|neighbour_cell", "face_node_x", "support_edge")
F_KEYS = ("node|pos", "cell|centroid", "face|face_center_pos", "face|face_area", "unit_norm_v", "cell|cells_area",
          "A_node_to_node", "single_B_node_to_node", "extra_B_node_to_node", "theta_PDE", "dt_graph", "sigma", "uvp_dim",
          "target|uvp")


@pytest.mark.parametrize("kind,bc,nx,ny", [("quad", "cavity", 9, 7), ("tri", "channel", 6, 8), ("mixed", "cavity", 8, 8),
                                           ("mixed", "channel", 11, 5)])
def test_torch_pipeline_matches_numpy(kind, bc, nx, ny):
    a, ua = S.make_case(0, kind=kind, bc=bc, seed=2, nx=nx, ny=ny)
    b, ub = ST.make_case(0, kind=kind, bc=bc, seed=2, nx=nx, ny=ny, device="cpu")
    for k in INT_KEYS:
        assert np.array_equal(np.asarray(a[k]), b[k].numpy()), k
    for k in F_KEYS:
        x, y = np.asarray(a[k], dtype=np.float64), b[k].double().numpy()
        assert x.shape == y.shape, k
        assert np.allclose(x, y, rtol=2e-6 if a[k].dtype == np.float32 else 1e-12, atol=1e-12), (k, np.abs(x - y).max())
    assert np.array_equal(ua, ub.numpy())


RAW_KEYS = ("node|pos", "node|surf_mask", "node|node_type", "face|face_node", "cells_node", "cells_index", "cells_face")
EXTRACT_INT = ("cells_node", "cells_index", "cells_face", "face|face_type", "face|neighbour_cell", "face_node_x")
EXTRACT_F = ("cell|centroid", "face|face_center_pos", "face|face_area", "unit_norm_v", "cell|cells_area")


def _compare_extract(raw_np):
    a = S.extract_mesh_state(raw_np)
    b = ST.extract_mesh_state({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in raw_np.items()})
    for k in EXTRACT_INT:
        assert np.array_equal(np.asarray(a[k]), b[k].numpy()), k
    for k in EXTRACT_F:
        x, y = np.asarray(a[k], dtype=np.float64), b[k].double().numpy()
        assert x.shape == y.shape and np.allclose(x, y, rtol=1e-12, atol=1e-12), (k, np.abs(x - y).max())


def test_torch_extract_keeps_the_order_of_interleaved_cell_types():
    """Cells of different vertex counts in arbitrary order (what a Tecplot polygon export looks like): like the reference's
    sort_vertices_ccw (parse_to_h5.py:55-110) the converters regroup the per-slot arrays by vertex count, ascending, keeping
    the cells' order inside a group and their ids; the device-side converter returns exactly the numpy converter's arrays."""
    raw = S.make_grid_mesh(7, kind="mixed", bc="channel", seed=4)
    ci, cn, cf = raw["cells_index"], raw["cells_node"], raw["cells_face"]
    C = int(ci.max()) + 1
    starts = np.flatnonzero(np.r_[True, ci[1:] != ci[:-1]])
    ends = np.r_[starts[1:], ci.size]
    perm = np.random.default_rng(0).permutation(C)       # triangles and quads now alternate at random
    raw = dict(raw)
    raw["cells_node"] = np.concatenate([cn[starts[c]:ends[c]] for c in perm])
    raw["cells_face"] = np.concatenate([cf[starts[c]:ends[c]] for c in perm])
    raw["cells_index"] = np.concatenate([np.full(ends[c] - starts[c], i, dtype=ci.dtype) for i, c in enumerate(perm)])
    counts = np.bincount(raw["cells_index"])
    assert (counts[1:] != counts[:-1]).sum() > 10         # really interleaved
    _compare_extract({k: raw[k] for k in RAW_KEYS})


def test_torch_extract_on_the_polygon_example_mesh():
    """BASELINE configs[2] cylinder_flow_poly (polygon cells with 3 .. 8 vertices, parsed by the reference's Tecplot path; stored
    in the golden): re-extracting its connectivity gives the numpy converter's arrays."""
    from tests import golden_util as GU
    z = GU.load_case("cylinder_poly_v1")
    m = GU.mesh_from_npz(z)
    raw = {k: np.asarray(m[k]) for k in RAW_KEYS if k in m}
    if "node|surf_mask" not in raw:
        raw["node|surf_mask"] = np.zeros(raw["node|pos"].shape[0], dtype=bool)
    assert len(np.unique(np.bincount(raw["cells_index"]))) >= 4   # several polygon sizes
    _compare_extract(raw)

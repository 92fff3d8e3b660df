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

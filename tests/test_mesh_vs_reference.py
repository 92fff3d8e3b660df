"""Container-only (needs /root/reference): gen_fvgn_steady_b200.mesh.synthetic restates the reference's mesh pipeline --
Extract_mesh/parse_to_h5.extract_mesh_state (:257-496) and Load_mesh.CFDdatasetBase.transform_mesh (Load_mesh.py:523-565)
-- and must reproduce the UNMODIFIED reference on the same raw mesh: index arrays bit-exact, geometry to fp64 round-off."""
import io
import os
from contextlib import redirect_stdout

import numpy as np
import pytest
import torch

from oracle import ref_shims

pytestmark = pytest.mark.skipif(not ref_shims.reference_available(), reason="reference sources not present (GPU box)")

INDEX_KEYS = ("face|face_node", "cells_node", "cells_index", "cells_face", "face|face_type", "face|neighbour_cell", "face_node_x")
GEOM_KEYS = ("cell|centroid", "face|face_center_pos", "face|face_area", "unit_norm_v", "cell|cells_area")


def _reference_extract(raw, tmp_path):
    ref_shims.install()
    from Extract_mesh import parse_to_h5
    parse_to_h5.write_point_cloud_to_vtk = lambda *a, **k: None
    ds = {k: (torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v) for k, v in raw.items()}
    with redirect_stdout(io.StringIO()):
        return parse_to_h5.extract_mesh_state(ds, path={"file_dir": str(tmp_path), "case_name": "synthetic"})


@pytest.mark.parametrize("kind,bc,n", [("quad", "cavity", 7), ("tri", "channel", 6), ("mixed", "channel", 8), ("mixed", "cavity", 5)])
def test_extract_mesh_state_equals_reference(kind, bc, n, tmp_path):
    from gen_fvgn_steady_b200.mesh import synthetic as S
    raw = S.make_grid_mesh(n, kind=kind, bc=bc, seed=3)
    ref = _reference_extract(raw, tmp_path)
    mine = S.extract_mesh_state(raw)
    for k in INDEX_KEYS:
        assert np.array_equal(np.asarray(mine[k]).reshape(-1), ref[k].numpy().reshape(-1)), k
    for k in GEOM_KEYS:
        a, b = np.asarray(mine[k], dtype=np.float64).reshape(-1), ref[k].double().numpy().reshape(-1)
        assert a.shape == b.shape and np.abs(a - b).max() <= 1e-12 * max(np.abs(b).max(), 1.0), k


@pytest.mark.parametrize("kind,bc", [("quad", "cavity"), ("mixed", "channel")])
def test_transform_mesh_equals_reference(kind, bc, tmp_path):
    """Stencil (k-hop face_node_x, support_edge), WLSQ moment matrices, theta_PDE / sigma / uvp_dim / dt, Dirichlet targets and
    the initial field for one fixed (U, rho, mu, aoa, dt) choice -- the reference draws from a BC.json grid with one entry."""
    from oracle import ref_harness as H
    from gen_fvgn_steady_b200.mesh import synthetic as S
    raw = S.make_grid_mesh(6, kind=kind, bc=bc, seed=4)
    physics = dict(mean_u=1.5, mu=0.02, dt=0.4, aoa=0.0)
    mine, uvp_mine = S.transform_mesh(S.extract_mesh_state(raw), physics)
    ph = dict(S.DEFAULT_PHYSICS)
    ph.update(physics)
    bc_json = {
        "inflow": None, "wall": None, "outflow": None, "pressure_point": None, "surf": None, "periodic": None,
        "stencil|BC_extra_points": 4, "stencil|khops": ph["khops"],
        "theta_PDE": {"unsteady": ph["unsteady"], "continuity": ph["continuity"], "convection": ph["convection"], "grad_p": ph["grad_p"],
                      "inlet": [ph["mean_u"]] * 3, "rho": [ph["rho"]] * 3, "mu": [ph["mu"]] * 3, "source": [ph["source"]] * 3,
                      "aoa": [ph["aoa"]] * 3, "dt": ph["dt"], "L": ph["L"], "Re_max": 1e9, "Re_min": 0},
        "sigma": list(ph["sigma"]), "inlet_type": ph["inlet_type"], "init_field_type": ph["init_field_type"],
    }
    ref_mesh = _reference_extract(raw, tmp_path)
    params = ref_shims.ref_params()
    H.seed_all(0)
    ref, uvp_ref = H.attach_bc_and_transform(ref_mesh, bc_json, "synthetic", params)
    for k in ("face_node_x", "support_edge"):
        assert np.array_equal(np.asarray(mine[k]), ref[k].numpy()), k
    for k in ("A_node_to_node", "single_B_node_to_node", "extra_B_node_to_node", "theta_PDE", "sigma", "uvp_dim", "dt_graph", "target|uvp"):
        a, b = np.asarray(mine[k], dtype=np.float64).reshape(-1), ref[k].double().numpy().reshape(-1)
        assert a.shape == b.shape, (k, a.shape, b.shape)
        assert np.abs(a - b).max() <= 2e-6 * max(np.abs(b).max(), 1.0), (k, np.abs(a - b).max())
    assert np.abs(uvp_mine - uvp_ref.numpy()).max() <= 1e-6

"""Pins oracle/fvgn_oracle.py against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from oracle import fvgn_oracle as O
from tests import golden_util as GU
from tests.case_inputs import case_meshes, case_state_dict


def _run(name, dtype):
    case = GU.CASES[name]
    meshes, uvps, z = case_meshes(name)
    g = O.graphs_from_meshes(meshes, uvps, dtype)
    sd = {k: (v.clone().requires_grad_(True) if not k.startswith("node_norm.") else v)
          for k, v in case_state_dict(z, dtype).items()}
    res = O.nnmodel_forward(sd, g, net=case["net"], dataset_size=case["dataset_size"], return_aux=True,
                            conserved_form=case.get("conserved_form", True), integrator=case.get("integrator", "imex"))
    loss = O.script_loss(res)
    loss.backward()
    return res, loss, sd, z


@pytest.mark.parametrize("name", list(GU.CASES))
@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-10), ("f32", torch.float32, 2e-4)])
def test_oracle_matches_reference(name, tag, dtype, tol):
    res, loss, sd, z = _run(name, dtype)
    for k in ("loss_cont", "loss_mom_x", "loss_mom_y", "loss_press", "uvp_node", "uvp_cell", "decoder_out", "grad_phi"):
        ref = z[f"{tag}.{k}"]
        assert tuple(res[k].shape) == ref.shape, k
        assert GU.rel_err(res[k].detach(), ref) <= tol, (k, GU.rel_err(res[k].detach(), ref))
    assert abs(float(loss) - float(z[f"{tag}.loss"])) <= tol * max(1.0, abs(float(z[f"{tag}.loss"])))
    keys = z["param_keys"].tolist()
    norms = z[f"{tag}.grad_norm"]
    gtol = tol * 50 if tag == "f32" else 1e-8
    worst = 0.0
    for i, k in enumerate(keys):
        g = sd[k].grad if sd[k].grad is not None else torch.zeros_like(sd[k])
        samp = g.reshape(-1)[:: GU.GRAD_SAMPLE_STRIDE]
        ref = z[f"{tag}.grad_sample.{i}"]
        scale = max(float(norms[i]) / np.sqrt(max(g.numel(), 1)) * np.sqrt(len(ref)), 1e-30)
        err = float((samp.double() - torch.from_numpy(ref).double()).norm()) / scale
        worst = max(worst, err)
        assert abs(float(g.double().norm()) - norms[i]) <= gtol * max(norms[i], 1e-12) + 1e-12, (k, float(g.norm()), norms[i])
    assert worst <= gtol, worst


def test_fp32_reference_vs_fp64_reference_gap_documented():
    """How far the reference's own fp32 run is from its fp64 run: the floor for any fp32 parity claim."""
    z = GU.load_case("synth_ns_batch2_v2")
    for k in ("loss_mom_x", "decoder_out", "grad_phi"):
        assert GU.rel_err(z[f"f32.{k}"], z[f"f64.{k}"]) < 1e-3

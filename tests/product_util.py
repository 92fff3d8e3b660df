"""Shared by the emulated (CPU, container) and the real (-m gpu) parity tests: runs the PRODUCT
(gen_fvgn_steady_b200.FVMmodel.importer.NNmodel) on a golden case and compares with the vectors the
unmodified reference produced (tests/golden/*.npz)."""
import ctypes
import os

import numpy as np
import torch

from tests import golden_util as GU
from tests.case_inputs import case_meshes, case_state_dict, product_graphs


def use_emulated_kernels():
    """TEST ONLY: drive the product's host logic through the CPU SIMT emulation of its kernels."""
    from tests.emu import build_emu
    from gen_fvgn_steady_b200 import _lib
    _lib._set_library_for_tests(ctypes.CDLL(build_emu.build()), allow_host_tensors=True)


def use_real_kernels():
    from gen_fvgn_steady_b200 import _lib
    _lib._set_library_for_tests(None, allow_host_tensors=False)


def script_loss(out, p):
    lb = p.loss_press * out[3] + p.loss_cont * out[0] + p.loss_mom * out[1] + p.loss_mom * out[2]
    return torch.mean(torch.log(lb))


def run_product(name, device, precision="fp32"):
    from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel
    from gen_fvgn_steady_b200.utils.get_param import params as default_params
    case = GU.CASES[name]
    meshes, uvps, z = case_meshes(name)
    graphs = product_graphs(meshes, uvps, device)
    p = default_params(net=case["net"], dataset_size=case["dataset_size"], precision=precision,
                       conserved_form=case.get("conserved_form", True), integrator=case.get("integrator", "imex"))
    model = NNmodel(p)
    sd = case_state_dict(z)
    missing = model.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to(device)
    out = model(*graphs, is_training=True)
    loss = script_loss(out, p)
    loss.backward()
    return model, out, loss, z


def compare_with_golden(model, out, loss, z, tag, tol, gtol, report=None):
    """Asserts the product's outputs and parameter gradients against the reference's run `tag` ('f32' | 'f64')."""
    names = ("loss_cont", "loss_mom_x", "loss_mom_y", "loss_press", "uvp_node", "uvp_cell")
    errs = {}
    for k, v in zip(names, out):
        ref = z[f"{tag}.{k}"]
        assert tuple(v.shape) == ref.shape, (k, v.shape, ref.shape)
        errs[k] = GU.rel_err(v.detach().cpu(), ref)
    errs["decoder_out"] = GU.rel_err(model._last["decoder_out"].detach().cpu(), z[f"{tag}.decoder_out"])
    errs["grad_phi"] = GU.rel_err(model._last["grad_phi"].detach().cpu(), z[f"{tag}.grad_phi"])
    errs["loss"] = abs(float(loss) - float(z[f"{tag}.loss"])) / max(1.0, abs(float(z[f"{tag}.loss"])))
    keys = z["param_keys"].tolist()
    norms = z[f"{tag}.grad_norm"]
    named = dict(model.named_parameters())
    worst, worst_key = 0.0, None
    for i, k in enumerate(keys):
        g = named[k].grad
        g = torch.zeros_like(named[k]) if g is None else g
        g = g.detach().cpu()
        samp = g.reshape(-1)[:: GU.GRAD_SAMPLE_STRIDE]
        ref = z[f"{tag}.grad_sample.{i}"]
        scale = max(float(norms[i]) / np.sqrt(max(g.numel(), 1)) * np.sqrt(len(ref)), 1e-30)
        e1 = float((samp.double() - torch.from_numpy(ref).double()).norm()) / scale
        e2 = abs(float(g.double().norm()) - norms[i]) / max(norms[i], 1e-12) if norms[i] > 1e-12 else float(g.double().norm())
        if max(e1, e2) > worst:
            worst, worst_key = max(e1, e2), k
    errs["param_grad_worst"] = worst
    if report is not None:
        report.update(errs)
        report["param_grad_worst_key"] = worst_key
    bad = {k: v for k, v in errs.items() if v > (gtol if k == "param_grad_worst" else tol)}
    assert not bad, (bad, worst_key)
    return errs

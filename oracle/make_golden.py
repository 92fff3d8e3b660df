"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by executing the UNMODIFIED reference
(/root/reference/src, via oracle/ref_shims.py) on CPU in the build container.

    python -m oracle.make_golden            # regenerate every case in tests/golden_util.CASES
    python -m oracle.make_golden NAME ...   # only these cases

For each case it stores: the inputs that cannot be regenerated elsewhere (the example mesh after the
reference's own parser + extract_mesh_state + transform_mesh; synthetic meshes are regenerated from
the seeded generator and only their index hash is stored), and the reference's outputs in fp32 and in
an fp64 re-run: 4 loss terms, script loss, uvp_node, uvp_cell, decoder output, WLSQ gradient, and for
every parameter the gradient norm + a strided sample.  Weights are tests/golden_util.golden_state_dict
loaded through the reference's own load_state_dict (which also pins the state_dict key names).
"""
import copy
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as H  # noqa: E402
from oracle import ref_shims  # noqa: E402
from tests import golden_util as GU  # noqa: E402
from gen_fvgn_steady_b200.mesh import synthetic as S  # noqa: E402

REF_MESH_ROOT = "/root/reference/mesh_example"


def _to_t(mesh):
    out = {}
    for k, v in mesh.items():
        out[k] = torch.from_numpy(np.ascontiguousarray(v)) if isinstance(v, np.ndarray) else v
    return out


def build_case_inputs(case, params):
    """-> (list of mesh dicts with tensor values, list of init uvp tensors, extra npz payload)."""
    payload = {}
    if isinstance(case["mesh"], str):
        rel = case["mesh"].split(":", 1)[1]
        meshes, uvps = [], []
        for gi, seed in enumerate(case.get("batch_seeds", [0])):
            # one run of the reference's parser + transform_mesh per graph: the sampler (random.choice over the BC.json
            # grid, Load_mesh.py:133-211) draws this graph's inlet velocity / viscosity from its seed
            mesh, uvp = H.load_example(os.path.join(REF_MESH_ROOT, rel), params, seed=seed)
            uvp = torch.from_numpy(GU.perturbed_field(uvp.numpy(), seed))
            if gi == 0:
                for k in GU.MESH_KEYS_F64 + GU.MESH_KEYS_F32:
                    payload["mesh." + k] = mesh[k].numpy()
                for k in GU.MESH_KEYS_I:
                    payload["mesh." + k] = mesh[k].numpy().astype(np.int32)
                payload["uvp0"] = uvp.numpy()
            else:
                for k in ("theta_PDE", "dt_graph", "sigma", "uvp_dim", "target|uvp"):
                    payload[f"g{gi}.{k}"] = mesh[k].numpy()
                payload[f"g{gi}.uvp0"] = uvp.numpy()
            meshes.append(mesh)
            uvps.append(uvp)
        return meshes, uvps, payload
    meshes, uvps = [], []
    for i, spec in enumerate(case["mesh"]):
        m, uvp = S.make_case(**spec)
        payload[f"index_hash.{i}"] = np.frombuffer(bytes.fromhex(GU.index_hash(m)), dtype=np.uint8)
        uvp = GU.perturbed_field(uvp, spec["seed"])
        meshes.append(_to_t(m))
        uvps.append(torch.from_numpy(uvp))
    return meshes, uvps, payload


def run_reference(case, meshes, uvps, dtype):
    ref_shims.install()
    import FVMmodel.FVdiscretization.FVscheme as FVscheme
    torch.set_default_dtype(dtype)
    try:
        params = ref_shims.ref_params(net=case["net"], dataset_size=case["dataset_size"],
                                      conserved_form=case.get("conserved_form", True), integrator=case.get("integrator", "imex"))
        H.seed_all(0)
        model = H.make_ref_model(params, dtype)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        missing = model.load_state_dict(GU.golden_state_dict(shapes, dtype=dtype), strict=True)
        assert not missing.missing_keys and not missing.unexpected_keys
        model = model.to(dtype)
        graphs = H.build_ref_graphs(meshes, uvps, dtype)
        rec = {}
        orig_wlsq = FVscheme.node_based_WLSQ

        def wlsq_hook(**kw):
            out = orig_wlsq(**kw)
            rec["grad_phi"] = out[:, :, 0:2].detach().clone()
            return out
        FVscheme.node_based_WLSQ = wlsq_hook
        hk = model.simulator.register_forward_hook(lambda m, i, o: rec.__setitem__("decoder_out", o.detach().clone()))
        try:
            out = H.ref_train_step(model, graphs, params)
        finally:
            FVscheme.node_based_WLSQ = orig_wlsq
            hk.remove()
        res = {k: v.detach().numpy() for k, v in out.items()}
        res["grad_phi"] = rec["grad_phi"].numpy()
        res["decoder_out"] = rec["decoder_out"].numpy()
        gnorm, gsample = {}, {}
        for k, p in model.named_parameters():
            g = p.grad if p.grad is not None else torch.zeros_like(p)
            gnorm[k] = float(g.double().norm())
            gsample[k] = g.reshape(-1)[:: GU.GRAD_SAMPLE_STRIDE].numpy().copy()
        nb = {k: v.numpy().copy() for k, v in model.state_dict().items() if k.startswith("node_norm.")}
        return res, gnorm, gsample, shapes, nb
    finally:
        torch.set_default_dtype(torch.float32)


def main():
    os.makedirs(GU.GOLDEN_DIR, exist_ok=True)
    mpath = os.path.join(GU.GOLDEN_DIR, "MANIFEST.json")
    only = sys.argv[1:]  # optional: regenerate just these cases, keep the rest of the manifest
    manifest = json.load(open(mpath)) if (only and os.path.exists(mpath)) else {}
    for name, case in GU.CASES.items():
        if only and name not in only:
            continue
        params = ref_shims.ref_params(net=case["net"], dataset_size=case["dataset_size"],
                                      conserved_form=case.get("conserved_form", True), integrator=case.get("integrator", "imex"))
        meshes, uvps, payload = build_case_inputs(case, params)
        keys = None
        for tag, dt in (("f32", torch.float32), ("f64", torch.float64)):
            res, gnorm, gsample, shapes, nb = run_reference(case, copy.deepcopy(meshes), uvps, dt)
            for k, v in res.items():
                payload[f"{tag}.{k}"] = v
            keys = sorted(gnorm)
            payload[f"{tag}.grad_norm"] = np.array([gnorm[k] for k in keys], dtype=np.float64)
            for i, k in enumerate(keys):
                payload[f"{tag}.grad_sample.{i}"] = gsample[k]
            for k, v in nb.items():
                payload[f"{tag}.{k}"] = v
        payload["param_keys"] = np.array(keys)
        payload["state_keys"] = np.array(sorted(shapes))
        payload["state_shapes"] = np.array([json.dumps(list(shapes[k])) for k in sorted(shapes)])
        path = os.path.join(GU.GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **payload)
        manifest[name] = dict(bytes=os.path.getsize(path),
                              losses_f64={k: payload[f"f64.{k}"].reshape(-1).tolist()
                                          for k in ("loss_cont", "loss_mom_x", "loss_mom_y", "loss_press", "loss")})
        print(name, manifest[name])
    json.dump(manifest, open(mpath, "w"), indent=1)


if __name__ == "__main__":
    main()

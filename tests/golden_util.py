"""Shared by oracle/make_golden.py (container, runs the real reference) and the tests (anywhere).

Deterministic, numpy-seeded model weights keyed by the reference's state_dict names, the golden
case definitions, and tolerant comparison helpers."""
import hashlib
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GRAD_SAMPLE_STRIDE = 61

# synthetic cases are regenerated from the seeded generator; the example mesh is stored in the npz
CASES = {
    "poisson_quad_tri_v2": dict(net="TransFVGN_v2", mesh="example:poisson/cavity_poisson_quad_tri/mesh_tri.mphtxt",
                                dataset_size=1),
    # BASELINE.json configs[0]: the reference's own CPU-runnable Navier-Stokes case (N = 10 404 nodes, C = 10 201 cells)
    "lid_cavity_101_v2": dict(net="TransFVGN_v2", dataset_size=1,
                              mesh="example:lid_driven_cavity/lid_driven_cavity_101x101-Re=100/mesh.mphtxt"),
    # BASELINE.json configs[2]: mixed tri/quad cylinder-flow example mesh (N = 11 923 nodes, C = 22 065 cells), pure GN net
    "cylinder_tri_quad_v1": dict(net="TransFVGN_v1", dataset_size=1,
                                 mesh="example:cylinder_flow_tri_quad/mesh.mphtxt"),
    "synth_ns_batch2_v2": dict(net="TransFVGN_v2", dataset_size=100, mesh=[
        dict(n=0, nx=10, ny=8, kind="mixed", bc="channel", seed=3,
             physics=dict(mean_u=1.5, mu=0.02, dt=0.4, aoa=0.0)),
        dict(n=0, nx=7, ny=9, kind="tri", bc="cavity", seed=4,
             physics=dict(mean_u=0.8, mu=0.01, dt=0.5, source=0.3)),
    ]),
    "synth_ns_batch2_v1": dict(net="TransFVGN_v1", dataset_size=100, mesh=[
        dict(n=0, nx=9, ny=9, kind="quad", bc="channel", seed=5,
             physics=dict(mean_u=1.2, mu=0.015, dt=0.3, inlet_type="parabolic", init_field_type="parabolic")),
        dict(n=0, nx=6, ny=11, kind="mixed", bc="cavity", seed=6, physics=dict(mean_u=1.0, mu=0.01)),
    ]),
    # the reference's second residual formulation (FVscheme.py:276-511, --conserved_form False)
    "synth_ns_batch2_v1_nc": dict(net="TransFVGN_v1", dataset_size=100, conserved_form=False, mesh=[
        dict(n=0, nx=9, ny=9, kind="quad", bc="channel", seed=5,
             physics=dict(mean_u=1.2, mu=0.015, dt=0.3, inlet_type="parabolic", init_field_type="parabolic")),
        dict(n=0, nx=6, ny=11, kind="mixed", bc="cavity", seed=6, physics=dict(mean_u=1.0, mu=0.01)),
    ]),
    # BASELINE.json configs[3]: parametric steady NS on the NACA0012 airfoil example mesh (N = 16 861 nodes, C = 30 684 mixed
    # cells, WLSQ moment matrices with condition numbers up to 1e9: the gradient-reconstruction stress case), a batch of two
    # graphs with inlet velocities drawn from the BC.json grid by the reference's own sampler (seeds 0 and 1)
    "airfoil_naca0012_b2_v2": dict(net="TransFVGN_v2", dataset_size=100, batch_seeds=[0, 1],
                                   mesh="example:airfoil_L=1/farfield_NACA0012_with_quad_bc/mesh_2.mphtxt"),
    # BASELINE.json configs[2]: the polygonal cylinder-flow example (Tecplot FEPolygon, cells of 3-9 vertices that are NOT
    # grouped by type, N = 27 778 nodes, C = 17 436 cells) parsed by the reference's parse_tecplot.py
    "cylinder_poly_v1": dict(net="TransFVGN_v1", dataset_size=100, mesh="example:cylinder_flow_poly/mesh.dat"),
    # the two other time integrators of importer.py:192-201 (--integrator explicit / implicit; imex is the default)
    "synth_ns_batch2_v1_explicit": dict(net="TransFVGN_v1", dataset_size=100, integrator="explicit", mesh=[
        dict(n=0, nx=9, ny=9, kind="quad", bc="channel", seed=5,
             physics=dict(mean_u=1.2, mu=0.015, dt=0.3, inlet_type="parabolic", init_field_type="parabolic")),
        dict(n=0, nx=6, ny=11, kind="mixed", bc="cavity", seed=6, physics=dict(mean_u=1.0, mu=0.01)),
    ]),
    "synth_ns_batch2_v1_implicit": dict(net="TransFVGN_v1", dataset_size=100, integrator="implicit", mesh=[
        dict(n=0, nx=9, ny=9, kind="quad", bc="channel", seed=5,
             physics=dict(mean_u=1.2, mu=0.015, dt=0.3, inlet_type="parabolic", init_field_type="parabolic")),
        dict(n=0, nx=6, ny=11, kind="mixed", bc="cavity", seed=6, physics=dict(mean_u=1.0, mu=0.01)),
    ]),
}


def golden_state_dict(shapes, seed=1234, dtype=torch.float32):
    """shapes: {key: shape} in any order.  Values depend only on (sorted key, shape, seed)."""
    sd = {}
    for i, k in enumerate(sorted(shapes)):
        shp = tuple(shapes[k])
        rng = np.random.default_rng([seed, i])
        if k.startswith("node_norm."):
            v = np.ones(shp) if k.endswith(("acc_count", "num_accumulations")) else np.zeros(shp)
        elif "temperature" in k:
            v = np.full(shp, 0.5)
        elif len(shp) == 2:
            v = rng.standard_normal(shp) / np.sqrt(shp[1])
        elif k.endswith(".weight"):            # LayerNorm gain
            v = 1.0 + 0.1 * rng.standard_normal(shp)
        else:                                   # biases
            v = 0.1 * rng.standard_normal(shp)
        sd[k] = torch.from_numpy(np.asarray(v, dtype=np.float64)).to(dtype)
    return sd


def perturbed_field(uvp, seed):
    """Make the initial (piecewise-constant) field non-trivial so every term of the loss is exercised."""
    rng = np.random.default_rng([77, seed])
    return (np.asarray(uvp, dtype=np.float64) + 0.1 * rng.standard_normal(uvp.shape)).astype(np.float32)


def index_hash(mesh):
    h = hashlib.sha256()
    for k in ("face|face_node", "cells_node", "cells_index", "cells_face", "face|face_type", "face|neighbour_cell",
              "face_node_x", "support_edge", "node|node_type"):
        h.update(np.ascontiguousarray(np.asarray(mesh[k]), dtype=np.int64).tobytes())
    return h.hexdigest()


def rel_err(a, b):
    a = torch.as_tensor(a, dtype=torch.float64).reshape(-1)
    b = torch.as_tensor(b, dtype=torch.float64).reshape(-1)
    den = float(b.norm())
    return float((a - b).norm()) / (den if den > 0 else 1.0)


def load_case(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


MESH_KEYS_F64 = ("node|pos", "cell|centroid", "face|face_center_pos", "face|face_area", "unit_norm_v", "cell|cells_area")
MESH_KEYS_I = ("node|node_type", "face|face_node", "cells_node", "cells_index", "cells_face", "face|face_type",
               "face|neighbour_cell", "face_node_x", "support_edge")
MESH_KEYS_F32 = ("A_node_to_node", "single_B_node_to_node", "extra_B_node_to_node", "theta_PDE", "dt_graph", "sigma",
                 "uvp_dim", "target|uvp")


def mesh_from_npz(z, prefix="mesh.", graph=0):
    """Mesh dictionary of graph `graph` of an example-mesh case: graphs > 0 share the geometry / connectivity of graph 0
    and carry their own physical parameters and targets under the prefix g{graph}."""
    m = {}
    for k in MESH_KEYS_F64 + MESH_KEYS_I + MESH_KEYS_F32:
        own = f"g{graph}.{k}"
        v = z[own] if (graph > 0 and own in z) else z[prefix + k]
        m[k] = v.astype(np.int64) if k in MESH_KEYS_I else v
    return m


def example_case_graphs(z):
    """-> (meshes, uvps) of an example-mesh case (1 graph, or len(batch_seeds) graphs on the same mesh)."""
    n = 1
    while f"g{n}.uvp0" in z:
        n += 1
    return [mesh_from_npz(z, graph=i) for i in range(n)], [z["uvp0"] if i == 0 else z[f"g{i}.uvp0"] for i in range(n)]

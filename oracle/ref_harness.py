"""TEST INFRASTRUCTURE ONLY -- drives the unmodified reference (via ref_shims) on CPU.

Used by oracle/make_golden.py and by container-only tests.  Never imported by the product.

Follows the reference's own data path without h5py / PyG:
  mesh file --Cosmol_manager.extract_mesh (parse_comsol.py:455-528)--> mesh dict
  + BC.json keys, theta_PDE_list (Load_mesh.py:574-622, re-stated as 12 lines because the
    original reads the dict from an .h5 file)
  --CFDdatasetBase.transform_mesh (Load_mesh.py:523-565)--> mesh dict + init uvp
  --five graph objects (Graph_loader.py:503-784) with the PyG offset rules
    (Graph_loader.py:405-480)--> NNmodel.forward (importer.py:156-240)
"""
import copy
import io
import json
import os
import random
from contextlib import redirect_stdout

import numpy as np
import torch

from . import ref_shims


def seed_all(seed=0):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)


def extract_comsol_mesh(mesh_path):
    """Run the reference's COMSOL parser + extract_mesh_state on one .mphtxt file."""
    ref_shims.install()
    from Extract_mesh import parse_comsol
    file_dir = os.path.dirname(mesh_path)
    case_name = os.path.basename(file_dir)
    path = {"simulator": "COMSOL", "mesh_only": True, "file_dir": file_dir,
            "case_name": case_name, "file_name": os.path.basename(mesh_path)}
    parse_comsol.Cosmol_manager.save_to_vtu = lambda self, *a, **k: None
    parse_comsol.write_vtp_file = lambda *a, **k: None
    with redirect_stdout(io.StringIO()):
        mgr = parse_comsol.Cosmol_manager(mesh_file=mesh_path, data_file=None, file_dir=file_dir,
                                          case_name=case_name, path=path)
        mesh = mgr.extract_mesh(mesh_only=True)
    return mesh, file_dir, case_name


def attach_bc_and_transform(mesh, bc, case_name, params):
    """Load_mesh.py:574-622 fed from a dict instead of an .h5 file, then transform_mesh."""
    ref_shims.install()
    from utils import get_param
    from Load_mesh.Load_mesh import CFDdatasetBase
    import Load_mesh.Load_mesh as LM
    LM.generate_boundary_zone = lambda **k: None
    mesh = dict(mesh)
    mesh["case_name"] = case_name
    for k, v in bc.items():
        mesh[k] = v
    mesh["theta_PDE_bak"] = mesh["theta_PDE"]
    th = mesh["theta_PDE_bak"]
    with redirect_stdout(io.StringIO()):
        mesh["theta_PDE_list"] = get_param.generate_combinations(
            U_range=th["inlet"], rho_range=th["rho"], mu_range=th["mu"], source_range=th["source"],
            aoa_range=th["aoa"], dt=th["dt"], L=th["L"], Re_max=th["Re_max"], Re_min=th["Re_min"])
        mesh_t, init_uvp = CFDdatasetBase.transform_mesh(mesh, params)
    return mesh_t, init_uvp


def extract_tecplot_mesh(mesh_path):
    """Run the reference's Tecplot FEPolygon parser (parse_tecplot.py:50-679) + extract_mesh_state on one .dat file."""
    ref_shims.install()
    from Extract_mesh import parse_tecplot
    file_dir = os.path.dirname(mesh_path)
    case_name = os.path.basename(file_dir)
    path = {"simulator": "Tecplot", "mesh_only": True, "file_dir": file_dir, "case_name": case_name,
            "file_name": os.path.basename(mesh_path)}
    parse_tecplot.TecplotMesh.save_to_vtu = lambda self, *a, **k: None
    with redirect_stdout(io.StringIO()):
        mgr = parse_tecplot.TecplotMesh(mesh_file=mesh_path, data_file=None, file_dir=file_dir, case_name=case_name, path=path)
        mesh = mgr.extract_mesh(mesh_only=True)
    return mesh, file_dir, case_name


def load_example(mesh_path, params, seed=0):
    seed_all(seed)
    if mesh_path.endswith(".dat"):
        mesh, file_dir, case_name = extract_tecplot_mesh(mesh_path)
    else:
        mesh, file_dir, case_name = extract_comsol_mesh(mesh_path)
    bc = json.load(open(os.path.join(file_dir, "BC.json")))
    return attach_bc_and_transform(mesh, bc, case_name, params)


def build_ref_graphs(meshes, uvps, dtype=torch.float32):
    """Five batched graph objects exactly as Graph_loader.py:503-784 + PyG collate would make."""
    Data = ref_shims.Data
    gn = dict(x=[], edge_index=[], face=[], pos=[], node_type=[], y=[], batch=[])
    gx = dict(face_node_x=[], support_edge=[], A=[], B=[], Bx=[])
    ge = dict(face_type=[], face_area=[], face=[], pos=[], batch=[])
    gc = dict(edge_index=[], unv=[], area=[], pos=[], face=[], batch=[])
    gi = dict(theta=[], sigma=[], uvp_dim=[], dt=[])
    n0 = e0 = c0 = 0
    for b, (m, uvp) in enumerate(zip(meshes, uvps)):
        N = m["node|pos"].shape[0]
        E = m["face|face_node"].shape[1]
        C = m["cell|centroid"].shape[0]
        gn["x"].append(uvp.to(dtype))
        gn["edge_index"].append(m["face|face_node"].long() + n0)
        gn["face"].append(m["cells_node"].long() + n0)
        gn["pos"].append(m["node|pos"].to(dtype))
        gn["node_type"].append(m["node|node_type"].long().view(-1))
        gn["y"].append(m["target|uvp"].to(dtype))
        gn["batch"].append(torch.full((N,), b, dtype=torch.long))
        gx["face_node_x"].append(m["face_node_x"].long() + n0)
        gx["support_edge"].append(m["support_edge"].long() + n0)
        gx["A"].append(m["A_node_to_node"].to(dtype))
        gx["B"].append(m["single_B_node_to_node"].to(dtype))
        gx["Bx"].append(m["extra_B_node_to_node"].to(dtype))
        ge["face_type"].append(m["face|face_type"].long().view(-1))
        ge["face_area"].append(m["face|face_area"].to(dtype))
        ge["face"].append(m["cells_face"].long() + e0)
        ge["pos"].append(m["face|face_center_pos"].to(dtype))
        ge["batch"].append(torch.full((E,), b, dtype=torch.long))
        gc["edge_index"].append(m["face|neighbour_cell"].long() + c0)
        gc["unv"].append(m["unit_norm_v"].to(dtype))
        gc["area"].append(m["cell|cells_area"].to(dtype).view(-1))
        gc["pos"].append(m["cell|centroid"].to(dtype))
        gc["face"].append(m["cells_index"].long() + c0)
        gc["batch"].append(torch.full((C,), b, dtype=torch.long))
        gi["theta"].append(m["theta_PDE"].to(dtype))
        gi["sigma"].append(m["sigma"].to(dtype))
        gi["uvp_dim"].append(m["uvp_dim"].to(dtype))
        gi["dt"].append(m["dt_graph"].to(dtype))
        n0, e0, c0 = n0 + N, e0 + E, c0 + C
    B = len(meshes)
    cat = torch.cat
    graph_node = Data(x=cat(gn["x"]), edge_index=cat(gn["edge_index"], 1), face=cat(gn["face"]),
                      pos=cat(gn["pos"]), node_type=cat(gn["node_type"]), y=cat(gn["y"]),
                      batch=cat(gn["batch"]), num_graphs=B)
    graph_node_x = Data(face_node_x=cat(gx["face_node_x"], 1), support_edge=cat(gx["support_edge"], 1),
                        A_node_to_node=cat(gx["A"]), single_B_node_to_node=cat(gx["B"]),
                        extra_B_node_to_node=cat(gx["Bx"]), num_nodes=n0, num_graphs=B)
    graph_edge = Data(face_type=cat(ge["face_type"]), face_area=cat(ge["face_area"]), face=cat(ge["face"]),
                      pos=cat(ge["pos"]), batch=cat(ge["batch"]), num_graphs=B)
    graph_cell = Data(x=torch.zeros((c0, 3), dtype=dtype), edge_index=cat(gc["edge_index"], 1),
                      cells_face_unv=cat(gc["unv"]), cells_area=cat(gc["area"]), pos=cat(gc["pos"]),
                      face=cat(gc["face"]), batch=cat(gc["batch"]), num_graphs=B)
    graph_Index = Data(x=torch.arange(B), theta_PDE=cat(gi["theta"]), sigma=cat(gi["sigma"]),
                       uvp_dim=cat(gi["uvp_dim"]), dt_graph=cat(gi["dt"]), num_graphs=B)
    # Data_Pool.datapreprocessing (Graph_loader.py:130-152)
    graph_node.x = cat((graph_node.x[:, 0:3], graph_Index.theta_PDE[graph_node.batch]), dim=1)
    graph_node.norm_uvp = True
    graph_node.norm_global = True
    return graph_node, graph_node_x, graph_edge, graph_cell, graph_Index


def ref_loader_graphs(meshes, uvps, ids=None):
    """The five batch objects made by the REFERENCE's own dataset classes (Graph_loader.py:503-784: GraphNodeDataset,
    GraphNode_X_Dataset, GraphEdgeDataset, GraphCellDataset, Graph_INDEX_Dataset .get) on a stand-in for Data_Pool's
    storage (meta_pool / uvp_node_pool / init_loss, Graph_loader.py:60-128), batched with the PyG collate rule driven by
    the reference's CustomGraphData.__inc__ / __cat_dim__ (ref_shims.collate), then Data_Pool.datapreprocessing (:130-152).
    meshes: dicts of torch tensors with the converter keys; ids: the sampled graph indices (default: all, in order)."""
    ref_shims.install()
    import Load_mesh.Graph_loader as GL
    pool, off = [], 0
    for i, m in enumerate(meshes):
        m = dict(m)
        n = m["node|pos"].shape[0]
        m["global_idx"] = torch.arange(off, off + n)
        m.setdefault("case_name", f"case{i}")
        # the converter's per-graph scalars are rows of the batched [B, k] tensors (Load_mesh.py:133-211)
        for k in ("theta_PDE", "sigma", "uvp_dim", "dt_graph"):
            m[k] = torch.as_tensor(m[k]).reshape(1, -1)
        m["face|face_area"] = torch.as_tensor(m["face|face_area"]).reshape(-1, 1)
        off += n
        pool.append(m)

    class _Pool:   # the attributes the dataset classes read from Data_Pool
        meta_pool = pool
        uvp_node_pool = torch.cat([torch.as_tensor(u, dtype=torch.float32) for u in uvps], 0)
        init_loss = torch.full((len(pool),), 1.0)
        params = None
    ids = list(range(len(pool))) if ids is None else list(ids)
    batches = []
    for cls in (GL.GraphNodeDataset, GL.GraphNode_X_Dataset, GL.GraphEdgeDataset, GL.GraphCellDataset, GL.Graph_INDEX_Dataset):
        ds = cls(_Pool)
        batches.append(ref_shims.collate([ds.get(i) for i in ids]))
    graphs = GL.Data_Pool.datapreprocessing(*batches)
    graphs[0].norm_uvp, graphs[0].norm_global = True, True
    return graphs


def make_ref_model(params, dtype=torch.float32):
    ref_shims.install()
    from FVMmodel.importer import NNmodel
    model = NNmodel(params)
    return model.to(dtype)


def ref_train_step(model, graphs, params):
    """One forward + script-level loss (pre_train_Adam.py:160-191) + backward."""
    gn, gx, ge, gc, gi = [copy.copy(g) for g in graphs]
    gn.x = gn.x.clone()
    gn.norm_uvp, gn.norm_global = True, True
    out = model(graph_node=gn, graph_node_x=gx, graph_edge=ge, graph_cell=gc, graph_Index=gi,
                is_training=True)
    loss_cont, loss_mx, loss_my, loss_press, uvp_node, uvp_cell = out
    loss_batch = (params.loss_press * loss_press + params.loss_cont * loss_cont
                  + params.loss_mom * loss_mx + params.loss_mom * loss_my)
    loss = torch.mean(torch.log(loss_batch))
    model.zero_grad(set_to_none=True)
    loss.backward()
    return dict(loss_cont=loss_cont, loss_mom_x=loss_mx, loss_mom_y=loss_my, loss_press=loss_press,
                uvp_node=uvp_node, uvp_cell=uvp_cell, loss=loss)

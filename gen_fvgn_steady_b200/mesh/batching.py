"""Batches mesh dictionaries (the converter/loader keys) into the five graph objects of
src/Load_mesh/Graph_loader.py:503-784 with the PyG offset rules of CustomGraphData.__inc__/__cat_dim__ (:405-480)
and Data_Pool.datapreprocessing (:130-152).  Host-side plumbing; tensors may live on any device."""
import numpy as np
import torch

from ..data import Data


def _T(v, device=None):
    t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
    return t if device is None else t.to(device)


def graphs_from_meshes(meshes, uvps, device="cpu"):
    f32 = torch.float32
    gn = dict(x=[], edge_index=[], face=[], pos=[], node_type=[], y=[], batch=[])
    gx = dict(face_node_x=[], support_edge=[], A=[], B=[], Bx=[])
    ge = dict(face_type=[], face_area=[], face=[], pos=[], batch=[])
    gc = dict(edge_index=[], unv=[], area=[], pos=[], face=[], batch=[])
    gi = dict(theta=[], sigma=[], uvp_dim=[], dt=[])
    n0 = e0 = c0 = 0
    for b, (m, uvp) in enumerate(zip(meshes, uvps)):
        T = lambda v: _T(v, device)
        N, E, C = T(m["node|pos"]).shape[0], T(m["face|face_node"]).shape[1], T(m["cell|centroid"]).shape[0]
        gn["x"].append(T(uvp).to(f32)); gn["edge_index"].append(T(m["face|face_node"]).long() + n0)
        gn["face"].append(T(m["cells_node"]).long() + n0); gn["pos"].append(T(m["node|pos"]).to(f32))
        gn["node_type"].append(T(m["node|node_type"]).long().view(-1)); gn["y"].append(T(m["target|uvp"]).to(f32))
        gn["batch"].append(torch.full((N,), b, dtype=torch.long, device=device))
        gx["face_node_x"].append(T(m["face_node_x"]).long() + n0); gx["support_edge"].append(T(m["support_edge"]).long() + n0)
        gx["A"].append(T(m["A_node_to_node"]).to(f32)); gx["B"].append(T(m["single_B_node_to_node"]).to(f32))
        gx["Bx"].append(T(m["extra_B_node_to_node"]).to(f32))
        ge["face_type"].append(T(m["face|face_type"]).long().view(-1)); ge["face_area"].append(T(m["face|face_area"]).to(f32).view(-1, 1))
        ge["face"].append(T(m["cells_face"]).long() + e0); ge["pos"].append(T(m["face|face_center_pos"]).to(f32))
        ge["batch"].append(torch.full((E,), b, dtype=torch.long, device=device))
        gc["edge_index"].append(T(m["face|neighbour_cell"]).long() + c0); gc["unv"].append(T(m["unit_norm_v"]).to(f32))
        gc["area"].append(T(m["cell|cells_area"]).to(f32).view(-1)); gc["pos"].append(T(m["cell|centroid"]).to(f32))
        gc["face"].append(T(m["cells_index"]).long() + c0); gc["batch"].append(torch.full((C,), b, dtype=torch.long, device=device))
        gi["theta"].append(T(m["theta_PDE"]).to(f32).view(1, -1)); gi["sigma"].append(T(m["sigma"]).to(f32).view(1, -1))
        gi["uvp_dim"].append(T(m["uvp_dim"]).to(f32).view(1, -1)); gi["dt"].append(T(m["dt_graph"]).to(f32).view(1, -1))
        n0, e0, c0 = n0 + N, e0 + E, c0 + C
    B = len(meshes)
    cat = torch.cat
    graph_Index = Data(x=torch.arange(B, device=device), theta_PDE=cat(gi["theta"]), sigma=cat(gi["sigma"]),
                       uvp_dim=cat(gi["uvp_dim"]), dt_graph=cat(gi["dt"]), num_graphs=B)
    graph_node = Data(x=cat(gn["x"]), edge_index=cat(gn["edge_index"], 1), face=cat(gn["face"]), pos=cat(gn["pos"]),
                      node_type=cat(gn["node_type"]), y=cat(gn["y"]), batch=cat(gn["batch"]), num_graphs=B)
    graph_node.x = cat((graph_node.x[:, 0:3], graph_Index.theta_PDE[graph_node.batch]), dim=1)  # datapreprocessing :148-150
    graph_node.norm_uvp, graph_node.norm_global = True, True
    graph_node_x = Data(face_node_x=cat(gx["face_node_x"], 1), support_edge=cat(gx["support_edge"], 1),
                        A_node_to_node=cat(gx["A"]), single_B_node_to_node=cat(gx["B"]), extra_B_node_to_node=cat(gx["Bx"]),
                        num_nodes=n0, num_graphs=B)
    graph_edge = Data(face_type=cat(ge["face_type"]), face_area=cat(ge["face_area"]), face=cat(ge["face"]), pos=cat(ge["pos"]),
                      batch=cat(ge["batch"]), num_graphs=B)
    graph_cell = Data(x=torch.zeros((c0, 3), device=device), edge_index=cat(gc["edge_index"], 1), cells_face_unv=cat(gc["unv"]),
                      cells_area=cat(gc["area"]), pos=cat(gc["pos"]), face=cat(gc["face"]), batch=cat(gc["batch"]), num_graphs=B)
    return graph_node, graph_node_x, graph_edge, graph_cell, graph_Index

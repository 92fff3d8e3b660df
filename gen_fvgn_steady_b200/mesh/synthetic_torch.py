"""Device-side (torch, any device) version of mesh/synthetic.py for multi-million-cell synthetic meshes
(SURVEY.md section 8(d) config 5, section 8(f) row f2): raw jittered grid -> FV connectivity/geometry
(parse_to_h5.extract_mesh_state :257-496) -> stencil + WLSQ moments + initial field (Load_mesh.transform_mesh :523-565).

Index arrays are bit-identical to the numpy pipeline (tests/test_synthetic_torch.py), which itself is bit-identical to
the reference's converter on the same raw mesh.  Float fields agree to ~1 ulp (segment sums use a fixed sequential order).
Setup-time plumbing: plain torch ops, no custom kernels.
"""
import math

import numpy as np
import torch

from . import synthetic as S

NORMAL, INFLOW, OUTFLOW, WALL_BOUNDARY, PRESS_POINT, IN_WALL = 0, 1, 2, 3, 4, 5


def _unique_cols(a, nmax, return_inverse=False):
    key = a[0] * nmax + a[1]
    if return_inverse:
        uk, inv = torch.unique(key, return_inverse=True)
        return torch.stack([uk // nmax, uk % nmax]), inv.reshape(-1)
    uk = torch.unique(key)
    return torch.stack([uk // nmax, uk % nmax])


def _rowsum(v, ct):
    """Sequential (left-to-right) sum over groups of ct consecutive rows: the order np.bincount / index_add_ on CPU use."""
    v = v.reshape((-1, ct) + tuple(v.shape[1:]))
    out = v[:, 0].clone()
    for j in range(1, ct):
        out = out + v[:, j]
    return out


def make_grid_mesh(n, kind="quad", jitter=0.2, seed=0, bc="cavity", lx=1.0, ly=1.0, nx=None, ny=None, device="cpu"):
    nx = nx or n
    ny = ny or n
    rng = np.random.default_rng(seed)
    hx, hy = lx / nx, ly / ny
    ii_np, jj_np = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")
    pos = np.stack([jj_np * hx, ii_np * hy], axis=-1).astype(np.float64)
    interior = (ii_np > 0) & (ii_np < ny) & (jj_np > 0) & (jj_np < nx)
    d = rng.uniform(-jitter, jitter, size=pos.shape) * np.array([hx, hy])
    pos = torch.from_numpy((pos + d * interior[..., None]).reshape(-1, 2)).to(device)
    del d, interior
    ii = torch.from_numpy(ii_np).to(device)
    jj = torch.from_numpy(jj_np).to(device)
    nid = ii * (nx + 1) + jj
    ci, cj = torch.meshgrid(torch.arange(ny, device=device), torch.arange(nx, device=device), indexing="ij")
    ci, cj = ci.reshape(-1), cj.reshape(-1)
    n00, n01, n11, n10 = nid[ci, cj], nid[ci, cj + 1], nid[ci + 1, cj + 1], nid[ci + 1, cj]
    quads = torch.stack([n00, n01, n11, n10], dim=1)
    if kind == "quad":
        tri_sel = torch.zeros(len(quads), dtype=torch.bool, device=device)
    elif kind == "tri":
        tri_sel = torch.ones(len(quads), dtype=torch.bool, device=device)
    elif kind == "mixed":
        tri_sel = ((ci + cj) % 2) == 0
    else:
        raise ValueError(kind)
    q = quads[tri_sel]
    flip = (((ci[tri_sel] + cj[tri_sel]) // 2) % 2 == 0)[:, None]
    t_a = torch.where(flip, q[:, [0, 1, 2]], q[:, [0, 1, 3]])
    t_b = torch.where(flip, q[:, [0, 2, 3]], q[:, [1, 2, 3]])
    tris = torch.stack([t_a, t_b], dim=1).reshape(-1, 3)
    quads = quads[~tri_sel]

    node_type = torch.full((pos.shape[0],), NORMAL, dtype=torch.int64, device=device)
    I, J = ii.reshape(-1), jj.reshape(-1)
    left, right, bottom, top = J == 0, J == nx, I == 0, I == ny
    if bc == "cavity":
        node_type[top] = INFLOW
        wall = left | right | bottom
        was_in = node_type == INFLOW
        node_type[wall] = WALL_BOUNDARY
        node_type[wall & was_in] = IN_WALL
        node_type[nid[0, nx // 2]] = PRESS_POINT
    elif bc == "channel":
        node_type[left] = INFLOW
        wall = top | bottom
        was_in = node_type == INFLOW
        node_type[wall] = WALL_BOUNDARY
        node_type[wall & was_in] = IN_WALL
        was_wall = node_type == WALL_BOUNDARY
        node_type[right] = OUTFLOW
        node_type[right & was_wall] = WALL_BOUNDARY
    else:
        raise ValueError(bc)

    cells_node, cells_index, edges = [], [], []
    count = 0
    for elems in (tris, quads):
        if len(elems) == 0:
            continue
        k = elems.shape[1]
        cells_node.append(elems.reshape(-1))
        cells_index.append(torch.arange(count, count + len(elems), device=device).repeat_interleave(k))
        count += len(elems)
        e = torch.stack([elems, torch.roll(elems, -1, dims=1)], dim=2).reshape(-1, 2)
        edges.append(torch.sort(e, dim=1).values.T)
    edges = torch.cat(edges, dim=1)
    face_node, cells_face = _unique_cols(edges, int(edges.max()) + 1, return_inverse=True)
    return {"node|pos": pos, "node|surf_mask": torch.zeros(pos.shape[0], dtype=torch.bool, device=device),
            "node|node_type": node_type, "face|face_node": face_node, "cells_node": torch.cat(cells_node),
            "cells_index": torch.cat(cells_index), "cells_face": cells_face}


def extract_mesh_state(raw):
    """parse_to_h5.extract_mesh_state (:257-496) on torch tensors.  Cells may have any vertex count in any order (the polygon
    meshes of the Tecplot path included: tests/test_synthetic_torch.py); the slots of a cell must be contiguous.  Like the
    reference's sort_vertices_ccw the per-slot arrays come back regrouped by vertex count (one pass per distinct count)."""
    m = dict(raw)
    pos, node_type, face_node = m["node|pos"], m["node|node_type"], m["face|face_node"]
    cells_node, cells_index, cells_face = m["cells_node"], m["cells_index"], m["cells_face"]
    dev = pos.device
    C = int(cells_index.max()) + 1
    counts = torch.bincount(cells_index, minlength=C)
    ctypes = torch.unique(counts).tolist()
    slot_ct = counts[cells_index]
    centroid = torch.empty((C, 2), dtype=pos.dtype, device=dev)
    for ct in ctypes:
        mask = slot_ct == ct
        ids = cells_index[mask].reshape(-1, ct)[:, 0]
        centroid[ids] = _rowsum(pos[cells_node[mask]], ct) / ct
    face_center = (pos[face_node[0]] + pos[face_node[1]]) / 2.0
    new_cn, new_cf, new_ci = [], [], []
    for ct in ctypes:
        mask = slot_ct == ct
        cn = cells_node[mask].reshape(-1, ct)
        cf = cells_face[mask].reshape(-1, ct)
        ci = cells_index[mask]
        cc = centroid[ci.reshape(-1, ct)[:, 0]]
        rel = pos[cn] - cc[:, None, :]
        order = torch.sort(torch.atan2(rel[:, :, 1], rel[:, :, 0]), dim=1, stable=True).indices
        cn = torch.gather(cn, 1, order)
        relf = face_center[cf] - cc[:, None, :]
        order = torch.sort(torch.atan2(relf[:, :, 1], relf[:, :, 0]), dim=1, stable=True).indices
        cf = torch.gather(cf, 1, order)
        new_cn.append(cn.reshape(-1)); new_cf.append(cf.reshape(-1)); new_ci.append(ci)
    cells_node, cells_face, cells_index = torch.cat(new_cn), torch.cat(new_cf), torch.cat(new_ci)

    lt, rt = node_type[face_node[0]], node_type[face_node[1]]
    anyb = lambda t: (t == INFLOW) | (t == WALL_BOUNDARY) | (t == OUTFLOW) | (t == PRESS_POINT) | (t == IN_WALL)
    nob_in = lambda t: (t == WALL_BOUNDARY) | (t == IN_WALL) | (t == OUTFLOW) | (t == PRESS_POINT)
    face_type = torch.full((face_node.shape[1],), NORMAL, dtype=torch.int64, device=dev)
    face_type[(anyb(lt) & (rt == INFLOW)) | (anyb(rt) & (lt == INFLOW))] = INFLOW
    face_type[(anyb(lt) & (rt == WALL_BOUNDARY)) | (nob_in(rt) & (lt == WALL_BOUNDARY))] = WALL_BOUNDARY
    face_type[(anyb(lt) & (rt == OUTFLOW)) | (nob_in(rt) & (lt == OUTFLOW))] = OUTFLOW

    diff = pos[face_node[0]] - pos[face_node[1]]
    face_area = torch.sqrt((diff ** 2).sum(1, keepdim=True))
    E = face_node.shape[1]
    snd = torch.full((E,), -1, dtype=torch.int64, device=dev).scatter_reduce(0, cells_face, cells_index, "amax", include_self=True)
    rcv = torch.full((E,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=dev).scatter_reduce(
        0, cells_face, cells_index, "amin", include_self=True)
    neighbour_cell = torch.stack([rcv, snd])

    unv = torch.stack([-diff[:, 1], diff[:, 0]], dim=1)
    unv = unv / torch.sqrt((unv ** 2).sum(1, keepdim=True))
    f2c = face_center[cells_face] - centroid[cells_index]
    cfu = unv[cells_face]
    outward = (f2c * cfu).sum(1, keepdim=True) > 0.0
    cfu = torch.where(outward, cfu, -1.0 * cfu)
    surface_vec = cfu * face_area[cells_face]
    slot_ct = counts[cells_index]
    cells_area = torch.empty((C,), dtype=pos.dtype, device=dev)
    for ct in ctypes:
        mask = slot_ct == ct
        ids = cells_index[mask].reshape(-1, ct)[:, 0]
        closed = _rowsum(surface_vec[mask], ct)
        if not torch.allclose(closed, torch.zeros_like(closed), rtol=1e-5, atol=1e-8):
            raise ValueError("wrong unv calculation: sum_f S_f != 0")
        cells_area[ids] = _rowsum((0.5 * face_center[cells_face[mask]] * surface_vec[mask]).sum(1), ct)

    N = pos.shape[0]
    fx = []
    for ct in ctypes:
        cn = cells_node[slot_ct == ct].reshape(-1, ct)
        pairs = [torch.stack([cn.reshape(-1), torch.roll(cn, s, dims=1).reshape(-1)]) for s in range(1, ct)]
        p = torch.cat(pairs, dim=1)
        p = torch.sort(p[:, p[0] != p[1]], dim=0).values
        fx.append(_unique_cols(p, N))
    face_node_x = _unique_cols(torch.cat(fx, dim=1), N)
    m.update({"cells_node": cells_node, "cells_face": cells_face, "cells_index": cells_index, "cell|centroid": centroid,
              "face|face_center_pos": face_center, "face|face_type": face_type, "face|face_area": face_area,
              "face|neighbour_cell": neighbour_cell, "unit_norm_v": cfu, "cell|cells_area": cells_area,
              "face_node_x": face_node_x})
    return m


def k_hop_pairs(face_node, num_nodes, k_hop):
    """unique{exactly-k-step walks, k = 1..k_hop} as sorted pairs without self loops (Load_mesh.py:475-482)."""
    if k_hop not in (1, 2):
        raise NotImplementedError("synthetic_torch supports stencil|khops in {1,2}")
    keys = [face_node[0] * num_nodes + face_node[1]]  # face_node is already (min,max), unique
    if k_hop == 2:
        two = torch.cat([face_node, face_node.flip(0)], dim=1)
        order = torch.sort(two[0], stable=True).indices
        mid, nb = two[0][order], two[1][order]
        deg = torch.bincount(mid, minlength=num_nodes)
        ptr = torch.zeros(num_nodes + 1, dtype=torch.int64, device=face_node.device)
        torch.cumsum(deg, 0, out=ptr[1:])
        rep = deg[mid]                                     # each (mid -> a) pairs with every (mid -> b)
        a = nb.repeat_interleave(rep)
        start = ptr[mid].repeat_interleave(rep)
        total = int(rep.sum())
        first = torch.cumsum(rep, 0) - rep
        local = torch.arange(total, device=face_node.device) - first.repeat_interleave(rep)
        b = nb[start + local]
        keep = a < b
        keys.append(a[keep] * num_nodes + b[keep])
    uk = torch.unique(torch.cat(keys))
    return torch.stack([uk // num_nodes, uk % num_nodes])


def _moments(order, d):
    if order == "1st":
        return d
    if order == "2nd":
        return torch.cat([d, 0.5 * d ** 2, d[:, 0:1] * d[:, 1:2]], dim=1)
    raise NotImplementedError(order)


def wlsq_moments(order, pos, face_node_x, support_edge):
    """compute_normal_matrix (FVgrad.py:183-232) in float64; A is summed entry by entry in the reference's scatter order."""
    dev = pos.device
    N = pos.shape[0]
    out_i = torch.cat([face_node_x[0], face_node_x[1], support_edge[0]])
    in_i = torch.cat([face_node_x[1], face_node_x[0], support_edge[1]])
    perm = torch.sort(in_i, stable=True).indices
    col = out_i[perm]
    deg = torch.bincount(in_i, minlength=N)
    ptr = torch.zeros(N + 1, dtype=torch.int64, device=dev)
    torch.cumsum(deg, 0, out=ptr[1:])
    mdim = 2 if order == "1st" else 5
    A = torch.zeros((N, mdim, mdim), dtype=torch.float64, device=dev)
    rows = torch.arange(N, device=dev)
    for j in range(int(deg.max())):
        valid = deg > j
        r = rows[valid]
        d = pos[col[ptr[r] + j]] - pos[r]
        mm = _moments(order, d)
        w = 1.0 / torch.sqrt((d ** 2).sum(1, keepdim=True))
        A[r] += (mm * w)[:, :, None] * mm[:, None, :]
    d1 = pos[face_node_x[0]] - pos[face_node_x[1]]
    B = (_moments(order, d1) / torch.sqrt((d1 ** 2).sum(1, keepdim=True)))[:, :, None]
    dx = pos[support_edge[0]] - pos[support_edge[1]]
    Bx = (_moments(order, dx) / torch.sqrt((dx ** 2).sum(1, keepdim=True)))[:, :, None]
    return A, B, Bx


def transform_mesh(mesh, physics=None):
    ph = dict(S.DEFAULT_PHYSICS)
    ph.update(physics or {})
    m = dict(mesh)
    pos = m["node|pos"]
    dev = pos.device
    N = pos.shape[0]
    U, rho, mu = ph["mean_u"], ph["rho"], ph["mu"]
    diffusion = (mu / U) if ph["convection"] == 0 else (mu / (rho * U))
    Re = np.float32(rho * U * ph["L"]) / np.float32(mu) if mu != 0 else np.float32(0)
    Uin = (U * math.cos(math.radians(ph["aoa"])), U * math.sin(math.radians(ph["aoa"])))
    f32 = lambda a: torch.from_numpy(np.asarray(a, dtype=np.float32)).to(dev)
    m["theta_PDE"] = f32([[ph["unsteady"], ph["continuity"], ph["convection"], ph["grad_p"] / rho, diffusion,
                           ph["source"] / U, np.float32(Uin[0]), np.float32(Uin[1]), Re]])
    m["dt_graph"] = f32([[ph["dt"] * U]])
    m["sigma"] = f32([ph["sigma"]])
    m["uvp_dim"] = f32([[U, U, U * U]])
    extra = k_hop_pairs(m["face|face_node"], N, ph["khops"])
    m["face_node_x"] = torch.cat([m["face_node_x"], extra], dim=1)
    m["support_edge"] = torch.tensor([[0, 1], [1, 0]], dtype=torch.int64, device=dev)
    A, B, Bx = wlsq_moments(ph["order"], pos, m["face_node_x"], m["support_edge"])
    m["A_node_to_node"] = A.to(torch.float32)
    m["single_B_node_to_node"] = B.to(torch.float32)
    m["extra_B_node_to_node"] = Bx.to(torch.float32)
    nt = m["node|node_type"]
    if ph["init_field_type"] != "uniform" or ph["inlet_type"] != "uniform":
        raise NotImplementedError("synthetic_torch: uniform init/inlet profiles only (use mesh/synthetic.py otherwise)")
    uvp = torch.zeros((N, 3), dtype=torch.float32, device=dev)
    uvp[:, 0] = float(U)
    wall = nt == WALL_BOUNDARY
    inlet = (nt == INFLOW) | (nt == IN_WALL) | (nt == PRESS_POINT)
    inwall = nt == IN_WALL
    uvp[inlet, 0] = float(U)
    uvp[inlet, 1] = 0.0
    uvp[wall, 0:2] = 0
    uvp[inwall] = uvp[inwall] / 2.0
    m["target|uvp"] = (uvp[:, 0:2] / np.float32(U)).to(torch.float32)
    m["order"] = ph["order"]
    return m, uvp


def make_case(n, kind="quad", bc="cavity", jitter=0.2, seed=0, physics=None, nx=None, ny=None, device="cpu"):
    raw = make_grid_mesh(n, kind=kind, jitter=jitter, seed=seed, bc=bc, nx=nx, ny=ny, device=device)
    return transform_mesh(extract_mesh_state(raw), physics)

"""Synthetic 2-D unstructured meshes + the mesh -> FV-connectivity pipeline, vectorised numpy.

Host-side, setup-time code.  It produces
exactly the dictionary the reference's converter + loader produce, so that the hot path can be
driven on the GPU box where /root/reference does not exist:

  make_grid_mesh        raw mesh (what parse_comsol.py:455-505 hands to extract_mesh_state)
  extract_mesh_state    restates Extract_mesh/parse_to_h5.py:257-496 (centroid, face centres, CCW
                        sort :55-110, face typing :307-371, face length, neighbour cells, outward
                        unit normals, cell areas, 1-ring stencil face_node_x :132-150,474-491)
  transform_mesh        restates Load_mesh/Load_mesh.py:133-272,420-565 (theta_PDE, k-hop stencil
                        without de-duplication against the 1-ring, support_edge=[[0,1],[1,0]],
                        WLSQ moments A/B in float64 -> float32, initial field, Dirichlet targets)

Index arrays are bit-exact against the reference on the same raw mesh (tests/test_mesh_vs_reference.py,
container-only) and against committed golden hashes (tests/golden/).
"""
import math

import numpy as np

NORMAL, INFLOW, OUTFLOW, WALL_BOUNDARY, PRESS_POINT, IN_WALL = 0, 1, 2, 3, 4, 5  # utilities.py:7-13


# ----------------------------------------------------------------------------- raw mesh
def make_grid_mesh(n, kind="quad", jitter=0.2, seed=0, bc="cavity", lx=1.0, ly=1.0, nx=None, ny=None):
    """n x n cells on [0,lx]x[0,ly]; interior nodes jittered by +-jitter*h (seeded).

    kind: "quad" | "tri" (every quad split along a diagonal) | "mixed" (checkerboard of both,
    tri block first then quad block, as parse_comsol.py:466-484 orders element types).
    bc: "cavity" (top INFLOW lid, other sides WALL, one PRESS_POINT) | "channel"
    (left INFLOW, right OUTFLOW, top/bottom WALL).
    """
    nx = nx or n
    ny = ny or n
    rng = np.random.default_rng(seed)
    hx, hy = lx / nx, ly / ny
    ii, jj = np.meshgrid(np.arange(ny + 1), np.arange(nx + 1), indexing="ij")  # row i = y, col j = x
    pos = np.stack([jj * hx, ii * hy], axis=-1).astype(np.float64)
    interior = (ii > 0) & (ii < ny) & (jj > 0) & (jj < nx)
    d = rng.uniform(-jitter, jitter, size=pos.shape) * np.array([hx, hy])
    pos = pos + d * interior[..., None]
    pos = pos.reshape(-1, 2)
    nid = (ii * (nx + 1) + jj)

    ci, cj = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    ci, cj = ci.reshape(-1), cj.reshape(-1)
    n00 = nid[ci, cj]
    n01 = nid[ci, cj + 1]
    n11 = nid[ci + 1, cj + 1]
    n10 = nid[ci + 1, cj]
    quads = np.stack([n00, n01, n11, n10], axis=1)
    if kind == "quad":
        tri_sel = np.zeros(len(quads), bool)
    elif kind == "tri":
        tri_sel = np.ones(len(quads), bool)
    elif kind == "mixed":
        tri_sel = ((ci + cj) % 2) == 0
    else:
        raise ValueError(kind)
    q = quads[tri_sel]
    flip = ((ci[tri_sel] + cj[tri_sel]) // 2) % 2 == 0
    t_a = np.where(flip[:, None], q[:, [0, 1, 2]], q[:, [0, 1, 3]])
    t_b = np.where(flip[:, None], q[:, [0, 2, 3]], q[:, [1, 2, 3]])
    tris = np.stack([t_a, t_b], axis=1).reshape(-1, 3)
    quads = quads[~tri_sel]

    # node types, in the order parse_comsol.set_node_type applies them (:360-409)
    node_type = np.full(pos.shape[0], NORMAL, dtype=np.int64)
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

    # parse_comsol.extract_mesh :455-505
    cells_node, cells_index, edges = [], [], []
    count = 0
    for elems in (tris, quads):
        if len(elems) == 0:
            continue
        k = elems.shape[1]
        cells_node.append(elems.reshape(-1))
        cells_index.append(np.repeat(np.arange(count, count + len(elems)), k))
        count += len(elems)
        e = np.stack([elems, np.roll(elems, -1, axis=1)], axis=2).reshape(-1, 2)  # (v_m, v_{m+1}), last wraps
        edges.append(np.sort(e, axis=1).T)
    edges = np.concatenate(edges, axis=1)
    face_node, cells_face = _unique_cols(edges, return_inverse=True)
    return {
        "node|pos": pos,
        "node|surf_mask": np.zeros(pos.shape[0], bool),
        "node|node_type": node_type,
        "face|face_node": face_node.astype(np.int64),
        "cells_node": np.concatenate(cells_node).astype(np.int64),
        "cells_index": np.concatenate(cells_index).astype(np.int64),
        "cells_face": cells_face.astype(np.int64),
    }


def _unique_cols(a, return_inverse=False, nmax=None):
    """Lexicographically sorted unique columns of a [2,M] non-negative int array (== np.unique(axis=1)
    == torch.unique(dim=1)), done on a fused 64-bit key so it scales to 10^8 columns."""
    a = np.asarray(a, dtype=np.int64)
    base = int(a.max()) + 1 if nmax is None else int(nmax)
    key = a[0] * base + a[1]
    if return_inverse:
        uk, inv = np.unique(key, return_inverse=True)
        return np.stack([uk // base, uk % base]), inv.reshape(-1)
    uk = np.unique(key)
    return np.stack([uk // base, uk % base])


# ----------------------------------------------------------------------------- extract_mesh_state
def _segment_sum(src, index, n):
    out = np.zeros((n,) + src.shape[1:], dtype=src.dtype)
    np.add.at(out, index, src)
    return out


def _bincount_sum(src, index, n):
    """Sequential-order segment sum (same order as torch index_add_ on CPU) for 1-D/2-D float64."""
    if src.ndim == 1:
        return np.bincount(index, weights=src, minlength=n)
    return np.stack([np.bincount(index, weights=src[:, c], minlength=n) for c in range(src.shape[1])], axis=1)


def extract_mesh_state(raw):
    """Restatement of parse_to_h5.extract_mesh_state (:257-496).  Inputs/outputs: numpy arrays with
    the reference's dict keys and dtypes (Appendix B of SURVEY.md)."""
    m = dict(raw)
    pos = m["node|pos"]
    node_type = m["node|node_type"]
    face_node = m["face|face_node"]
    cells_node, cells_index, cells_face = m["cells_node"], m["cells_index"], m["cells_face"]
    C = int(cells_index.max()) + 1
    counts = np.bincount(cells_index, minlength=C)
    centroid = _bincount_sum(pos[cells_node], cells_index, C) / counts[:, None]  # :275-281
    face_center = (pos[face_node[0]] + pos[face_node[1]]) / 2.0  # :285

    # sort_vertices_ccw (:55-110): per cell type (ascending vertex count), angle sort about the centroid
    new_cn, new_cf, new_ci = [], [], []
    ctype_slot = counts[cells_index]
    for ct in np.unique(counts):
        mask = ctype_slot == ct
        cn = cells_node[mask].reshape(-1, ct)
        cf = cells_face[mask].reshape(-1, ct)
        ci = cells_index[mask]
        cc = centroid[ci.reshape(-1, ct)[:, 0]]
        rel = pos[cn] - cc[:, None, :]
        cn = np.take_along_axis(cn, np.argsort(np.arctan2(rel[:, :, 1], rel[:, :, 0]), axis=1, kind="stable"), 1)
        relf = face_center[cf] - cc[:, None, :]
        cf = np.take_along_axis(cf, np.argsort(np.arctan2(relf[:, :, 1], relf[:, :, 0]), axis=1, kind="stable"), 1)
        new_cn.append(cn.reshape(-1)); new_cf.append(cf.reshape(-1)); new_ci.append(ci)
    cells_node, cells_face, cells_index = np.concatenate(new_cn), np.concatenate(new_cf), np.concatenate(new_ci)

    # face typing (:307-371): both end nodes on the boundary and one of them of that type;
    # applied in the order inflow -> wall -> outflow so later rules win.  The reference's wall/outflow
    # masks list IN_WALL twice and omit INFLOW on one side (:339-343,:359-363); restated literally.
    lt, rt = node_type[face_node[0]], node_type[face_node[1]]
    anyb = lambda t: (t == INFLOW) | (t == WALL_BOUNDARY) | (t == OUTFLOW) | (t == PRESS_POINT) | (t == IN_WALL)
    nob_in = lambda t: (t == WALL_BOUNDARY) | (t == IN_WALL) | (t == OUTFLOW) | (t == PRESS_POINT)
    face_type = np.full(face_node.shape[1], NORMAL, dtype=np.int64)
    face_type[(anyb(lt) & (rt == INFLOW)) | (anyb(rt) & (lt == INFLOW))] = INFLOW
    face_type[(anyb(lt) & (rt == WALL_BOUNDARY)) | (nob_in(rt) & (lt == WALL_BOUNDARY))] = WALL_BOUNDARY
    face_type[(anyb(lt) & (rt == OUTFLOW)) | (nob_in(rt) & (lt == OUTFLOW))] = OUTFLOW

    diff = pos[face_node[0]] - pos[face_node[1]]
    face_area = np.sqrt((diff ** 2).sum(1, keepdims=True))  # :377-380
    E = face_node.shape[1]
    snd = np.full(E, -1, dtype=np.int64); np.maximum.at(snd, cells_face, cells_index)
    rcv = np.full(E, np.iinfo(np.int64).max, dtype=np.int64); np.minimum.at(rcv, cells_face, cells_index)
    neighbour_cell = np.stack([rcv, snd])  # :385-401

    unv = np.stack([-diff[:, 1], diff[:, 0]], axis=1)
    unv = unv / np.sqrt((unv ** 2).sum(1, keepdims=True))  # :408-409
    f2c = face_center[cells_face] - centroid[cells_index]
    cfu = unv[cells_face]
    outward = (f2c * cfu).sum(1, keepdims=True) > 0.0
    cfu = np.where(outward, cfu, -1.0 * cfu)  # :414-423
    surface_vec = cfu * face_area[cells_face]
    closed = _bincount_sum(surface_vec, cells_index, C)
    if not np.allclose(closed, 0.0, rtol=1e-5, atol=1e-8):
        raise ValueError("wrong unv calculation: sum_f S_f != 0")  # :430-438
    cells_area = _bincount_sum((0.5 * face_center[cells_face] * surface_vec).sum(1), cells_index, C)  # :446-453

    # 1-ring node stencil: unique unordered node pairs sharing a cell (:132-150,:474-491)
    fx = []
    ctype_slot = counts[cells_index]
    for ct in np.unique(counts):
        cn = cells_node[ctype_slot == ct].reshape(-1, ct)
        pairs = []
        for s in range(1, ct):
            pairs.append(np.stack([cn.reshape(-1), np.roll(cn, s, axis=1).reshape(-1)]))
        p = np.concatenate(pairs, axis=1)
        p = np.sort(p[:, p[0] != p[1]], axis=0)
        fx.append(_unique_cols(p, nmax=pos.shape[0]))
    face_node_x = _unique_cols(np.concatenate(fx, axis=1), nmax=pos.shape[0])

    m.update({
        "cells_node": cells_node, "cells_face": cells_face, "cells_index": cells_index,
        "cell|centroid": centroid, "face|face_center_pos": face_center, "face|face_type": face_type,
        "face|face_area": face_area, "face|neighbour_cell": neighbour_cell, "unit_norm_v": cfu,
        "cell|cells_area": cells_area, "face_node_x": face_node_x,
    })
    return m


# ----------------------------------------------------------------------------- transform_mesh
def k_hop_pairs(face_node, num_nodes, k_hop):
    """unique{ exactly-k-step walks, k=1..k_hop } as sorted unordered pairs without self loops
    (Load_mesh.py:475-482 + parse_to_h5.build_k_hop_edge_index :228-254)."""
    import scipy.sparse as sp
    two = np.concatenate([face_node, face_node[::-1]], axis=1)
    adj = sp.csr_matrix((np.ones(two.shape[1], dtype=np.float32), (two[0], two[1])), shape=(num_nodes, num_nodes))
    adj.sum_duplicates()
    keys = []
    ak = adj
    for k in range(1, k_hop + 1):
        if k > 1:
            ak = (ak @ adj).tocsr()
        coo = ak.tocoo()
        r, c = coo.row.astype(np.int64), coo.col.astype(np.int64)
        keep = r < c  # sort(0) of both directions + unique == keep the (min,max) copy
        keys.append(r[keep] * num_nodes + c[keep])
    uk = np.unique(np.concatenate(keys))
    return np.stack([uk // num_nodes, uk % num_nodes])


def wlsq_moments(order, pos, face_node_x, support_edge):
    """FVgrad.compute_normal_matrix :183-232 + FVorder.moments_order :7-86 in float64.
    Returns A [N,m,m], one-way B [X,m,1], extra B [Xs,m,1] (float64)."""
    two = np.concatenate([face_node_x, face_node_x[::-1]], axis=1)
    comp = np.concatenate([two, support_edge], axis=1)
    out_i, in_i = comp[0], comp[1]
    d = pos[out_i] - pos[in_i]
    if order == "1st":
        disp = d
    elif order == "2nd":
        disp = np.concatenate([d, 0.5 * d ** 2, d[:, 0:1] * d[:, 1:2]], axis=1)
    else:
        raise NotImplementedError(f"order {order}: only '1st' and '2nd' are on the supported path")
    w = 1.0 / np.sqrt((d ** 2).sum(1, keepdims=True))
    left = (disp * w)[:, :, None] * disp[:, None, :]
    N, mdim = pos.shape[0], disp.shape[1]
    A = np.zeros((N, mdim * mdim))
    left = left.reshape(-1, mdim * mdim)
    for c in range(mdim * mdim):
        A[:, c] = np.bincount(in_i, weights=left[:, c], minlength=N)
    B = (w * disp)[:, :, None]
    X = face_node_x.shape[1]
    return A.reshape(N, mdim, mdim), B[:X], B[2 * X:]


def velocity_profile(pos, mean_u, aoa, kind):
    """Load_mesh/Set_BC.py:6-66 ('uniform', 'uniform_aoa', 'parabolic')."""
    uv = np.zeros_like(pos)
    p = np.zeros((pos.shape[0], 1))
    if pos.shape[0] == 0:
        return uv.astype(np.float32), p.astype(np.float32)
    if kind == "parabolic":
        y = pos[:, 1] - pos[:, 1].min()
        uv[:, 0] = 6 * mean_u * y * (((y.max() - y.min()) - y) / (y.max() - y.min()) ** 2)
    elif kind == "uniform":
        uv[:, 0] = float(mean_u)
    elif kind == "uniform_aoa":
        uv[:, 0] = mean_u * math.cos(math.radians(aoa))
        uv[:, 1] = mean_u * math.sin(math.radians(aoa))
    else:
        raise ValueError(kind)
    return uv.astype(np.float32), p.astype(np.float32)


DEFAULT_PHYSICS = dict(  # a Navier-Stokes case in the shape of mesh_example/*/BC.json
    unsteady=1.0, continuity=1.0, convection=1.0, grad_p=1.0,
    mean_u=1.0, rho=1.0, mu=0.01, source=0.0, aoa=0.0, dt=0.5, L=1.0,
    sigma=(1.0, 1.0, 1.0), inlet_type="uniform", init_field_type="uniform", khops=2, order="2nd",
)


def transform_mesh(mesh, physics=None, structured_quad=None):
    """Restatement of CFDdatasetBase.transform_mesh (Load_mesh.py:523-565) for one fixed
    (U, rho, mu, source, aoa, dt, L) choice instead of random.choice over the BC.json grid."""
    ph = dict(DEFAULT_PHYSICS)
    ph.update(physics or {})
    m = dict(mesh)
    pos = m["node|pos"]
    N = pos.shape[0]
    U, rho, mu = ph["mean_u"], ph["rho"], ph["mu"]
    # set_theta_PDE :133-211
    diffusion = (mu / U) if ph["convection"] == 0 else (mu / (rho * U))
    Re = np.float32(rho * U * ph["L"]) / np.float32(mu) if mu != 0 else np.float32(0)
    Uin = (U * math.cos(math.radians(ph["aoa"])), U * math.sin(math.radians(ph["aoa"])))
    m["theta_PDE"] = np.array([[ph["unsteady"], ph["continuity"], ph["convection"], ph["grad_p"] / rho, diffusion,
                                ph["source"] / U, np.float32(Uin[0]), np.float32(Uin[1]), Re]], dtype=np.float32)
    m["dt_graph"] = np.array([[ph["dt"] * U]], dtype=np.float32)
    m["sigma"] = np.array([ph["sigma"]], dtype=np.float32)
    m["uvp_dim"] = np.array([[U, U, U * U]], dtype=np.float32)
    # construct_stencil :420-521
    extra = k_hop_pairs(m["face|face_node"], N, ph["khops"])
    m["face_node_x"] = np.concatenate([m["face_node_x"], extra], axis=1)
    m["support_edge"] = np.array([[0, 1], [1, 0]], dtype=np.int64)
    # calc_WLSQ_A_B_normal_matrix :246-272
    A, B, Bx = wlsq_moments(ph["order"], pos, m["face_node_x"], m["support_edge"])
    m["A_node_to_node"] = A.astype(np.float32)
    m["single_B_node_to_node"] = B.astype(np.float32)
    m["extra_B_node_to_node"] = Bx.astype(np.float32)
    # init_env :79-131
    nt = m["node|node_type"]
    uv, p = velocity_profile(pos, U, ph["aoa"], ph["init_field_type"])
    uvp = np.concatenate([uv, p], axis=1).astype(np.float32)
    wall = nt == WALL_BOUNDARY
    inlet = (nt == INFLOW) | (nt == IN_WALL) | (nt == PRESS_POINT)
    inwall = nt == IN_WALL
    iv, _ = velocity_profile(pos[inlet], U, ph["aoa"], ph["inlet_type"])
    uvp[inlet, 0:2] = iv[:, 0:2]
    uvp[wall, 0:2] = 0
    uvp[inwall] = uvp[inwall] / 2.0
    m["target|uvp"] = (uvp[:, 0:2] / np.float32(U)).astype(np.float32)
    m["order"] = ph["order"]
    return m, uvp


def make_case(n, kind="quad", bc="cavity", jitter=0.2, seed=0, physics=None, nx=None, ny=None):
    """raw mesh -> extract_mesh_state -> transform_mesh.  Returns (mesh dict of numpy arrays, init uvp)."""
    raw = make_grid_mesh(n, kind=kind, jitter=jitter, seed=seed, bc=bc, nx=nx, ny=ny)
    return transform_mesh(extract_mesh_state(raw), physics)

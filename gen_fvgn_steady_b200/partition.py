"""Cell partition of ONE large mesh over R ranks with a per-GnBlock halo exchange (SURVEY.md section 8(e).2,
BASELINE.json north_star: "a METIS-style cell partition of one large synthetic mesh with a per-layer halo exchange").

The reference has no multi-GPU path; this module defines one whose result equals the single-GPU run of the same mesh:

* cells are split by recursive coordinate bisection of their centroids (balanced, contiguous parts);
* a node / face is OWNED by the lowest rank among its incident cells;
* rank r keeps its owned cells plus `halo_layers` (default 3) layers of neighbouring cells (cells sharing a node);
  deeper halos trade redundant compute for fewer exchanges (6 layers: every 2nd block; 3 G + 2 layers: none at all, the
  only collectives left are the [B,k]-sized statistics and the gradient all-reduce).  One GnBlock
  reads the node latents within 3 hops (blocks.py: x' <- a2 <- a1(neighbours) <- e'(their edges) <- agg(both ends) <-
  x(neighbours)), and the WLSQ stencil reaches 2 hops, so after every ghost row has been refreshed from its owner one
  whole block (and the FV loss of the owned cells) is computed locally and exactly on the owned rows; the outer halo
  rows come out wrong and are overwritten by the next exchange.  One exchange per block, redundant compute on the halo;
* the local sub-mesh keeps every attribute of the global mesh (types, geometry, WLSQ moments) and is renumbered with
  the owned rows first, then the ghost rows grouped by owner (ascending global id inside a group): every receive is a
  contiguous row range, every send an index gather.  Ghost nodes / halo cells are given the graph id B + b, so all
  per-graph sums of the hot path (z-score statistics, Normalizer, residual norms) see the owned rows only; those sums
  are all-reduced over the ranks (gen_fvgn_steady_b200.parallel).

Everything here is setup-time index work in torch (runs on the mesh's device)."""
import torch

HALO_LAYERS = 3


def _t(v, device=None):
    import numpy as np
    t = v if torch.is_tensor(v) else torch.from_numpy(np.ascontiguousarray(v))
    return t if device is None else t.to(device)


def rcb_partition(centroid, nparts):
    """Recursive coordinate bisection: cell -> part id in [0, nparts).  Splits the longer extent at the rank that
    gives floor/ceil-balanced halves; ties are broken by cell id, so the result is deterministic."""
    C = centroid.shape[0]
    part = torch.zeros(C, dtype=torch.int64, device=centroid.device)
    todo = [(torch.arange(C, device=centroid.device), 0, nparts)]
    while todo:
        idx, base, k = todo.pop()
        if k == 1:
            part[idx] = base
            continue
        pts = centroid[idx].double()
        ext = pts.max(0).values - pts.min(0).values
        ax = int(torch.argmax(ext))
        order = torch.sort(pts[:, ax], stable=True).indices
        kl = k // 2
        nl = (idx.numel() * kl) // k
        todo.append((idx[order[:nl]], base, kl))
        todo.append((idx[order[nl:]], base + kl, k - kl))
    return part


def _grow(cell_mask, cells_node, cells_index, N, layers):
    """cells within `layers` node-sharing layers of the masked cells -> (mask, mask one layer earlier)."""
    C = cell_mask.shape[0]
    m = cell_mask.clone()
    prev = m
    for _ in range(layers):
        prev = m
        node_mask = torch.zeros(N, dtype=torch.bool, device=m.device)
        node_mask[cells_node[m[cells_index]]] = True
        touched = torch.zeros(C, dtype=torch.int32, device=m.device)
        touched.index_add_(0, cells_index, node_mask[cells_node].to(torch.int32))
        m = m | (touched > 0)
    return m, prev


class HaloPlan:
    """Exchange lists of one rank.  rows: 'node' / 'edge' -> dict(n_owned, n_local, send={peer: LongTensor local ids},
    recv={peer: (start, count)}, gid=LongTensor global ids of the local rows)."""

    def __init__(self, rank, world):
        self.rank, self.world = rank, world
        self.rows = {}
        self.n_owned_cells = 0
        self.cell_gid = None
        self.num_graphs = 1
        self.layers = HALO_LAYERS

    def wants_exchange(self, block_index, n_blocks):
        """Is a ghost refresh needed after GnBlock `block_index` (0-based) of `n_blocks`?  One block invalidates 3 more
        halo layers; the decoder + WLSQ + FV stage needs the 2 innermost layers exact.  With `layers` halo layers the
        ghosts are refreshed after every k = layers // 3 blocks and after the last one; 3 n_blocks + 2 layers never."""
        if self.layers >= 3 * n_blocks + 2:
            return False
        k = max(self.layers // 3, 1)
        return (block_index + 1) % k == 0 or block_index == n_blocks - 1

    def to(self, device):
        for r in self.rows.values():
            r["send"] = {q: v.to(device) for q, v in r["send"].items()}
            r["gid"] = r["gid"].to(device)
        if self.cell_gid is not None:
            self.cell_gid = self.cell_gid.to(device)
        return self

    def peers(self):
        ps = set()
        for r in self.rows.values():
            ps |= set(r["send"].keys()) | set(r["recv"].keys())
        return sorted(ps)

    def exchanged_rows(self, kind):
        r = self.rows[kind]
        return sum(int(v.numel()) for v in r["send"].values()), sum(c for _, c in r["recv"].values())


def _local_order(local_mask, owner, rank):
    """local rows: owned first (ascending global id), then ghosts grouped by owner rank (ascending global id)."""
    gid = torch.nonzero(local_mask).reshape(-1)
    own = owner[gid]
    key = torch.where(own == rank, torch.full_like(own, -1), own)
    order = torch.sort(key, stable=True).indices  # gid is ascending, stable sort keeps it inside a group
    gid, key = gid[order], key[order]
    n_owned = int((key == -1).sum())
    recv = {}
    if gid.numel() > n_owned:
        ghosts = key[n_owned:]
        peers, counts = torch.unique_consecutive(ghosts, return_counts=True)
        start = n_owned
        for q, c in zip(peers.tolist(), counts.tolist()):
            recv[int(q)] = (start, int(c))
            start += int(c)
    return gid, n_owned, recv


def build(mesh, uvp, world, rank, halo_layers=HALO_LAYERS, device=None):
    """-> (local mesh dict, local initial field, HaloPlan) for `rank` of `world`.

    mesh: the converter / loader dictionary of the GLOBAL mesh (SURVEY.md Appendix B keys, numpy or torch)."""
    if halo_layers < 3:
        raise ValueError("halo_layers must be >= 3 (one GnBlock reads latents 3 hops away)")
    dev = device
    cn = _t(mesh["cells_node"], dev).reshape(-1).long()
    cf = _t(mesh["cells_face"], dev).reshape(-1).long()
    ci = _t(mesh["cells_index"], dev).reshape(-1).long()
    fn = _t(mesh["face|face_node"], dev).long()
    pos = _t(mesh["node|pos"], dev)
    cen = _t(mesh["cell|centroid"], dev)
    dev = cn.device
    N, E, C = pos.shape[0], fn.shape[1], cen.shape[0]
    part = rcb_partition(cen, world)
    big = torch.full((1,), world, dtype=torch.int64, device=dev)
    node_owner = big.expand(N).clone().scatter_reduce_(0, cn, part[ci], reduce="amin", include_self=True)
    face_owner = big.expand(E).clone().scatter_reduce_(0, cf, part[ci], reduce="amin", include_self=True)

    def local_sets(r):
        cmask, cprev = _grow(part == r, cn, ci, N, halo_layers)
        slot = cmask[ci]
        nmask = torch.zeros(N, dtype=torch.bool, device=dev)
        nmask[cn[slot]] = True
        fmask = torch.zeros(E, dtype=torch.bool, device=dev)
        fmask[cf[slot]] = True
        inner = torch.zeros(N, dtype=torch.bool, device=dev)  # nodes of all but the outermost cell layer
        inner[cn[cprev[ci]]] = True
        return cmask, nmask, fmask, nmask & ~inner

    cmask, nmask, fmask, outer_nodes = local_sets(rank)
    halo = HaloPlan(rank, world)
    halo.layers = halo_layers
    node_gid, n_own_nodes, node_recv = _local_order(nmask, node_owner, rank)
    face_gid, n_own_faces, face_recv = _local_order(fmask, face_owner, rank)
    cell_gid, n_own_cells, _ = _local_order(cmask, part, rank)
    g2l_node = torch.full((N,), -1, dtype=torch.int64, device=dev)
    g2l_node[node_gid] = torch.arange(node_gid.numel(), device=dev)
    g2l_face = torch.full((E,), -1, dtype=torch.int64, device=dev)
    g2l_face[face_gid] = torch.arange(face_gid.numel(), device=dev)
    g2l_cell = torch.full((C,), -1, dtype=torch.int64, device=dev)
    g2l_cell[cell_gid] = torch.arange(cell_gid.numel(), device=dev)

    # what the peers hold as ghosts of rows owned here (every rank derives the same sets from the same global mesh)
    node_send, face_send = {}, {}
    for q in range(world):
        if q == rank:
            continue
        _, nm_q, fm_q, _ = local_sets(q)
        ids = torch.nonzero(nm_q & (node_owner == rank)).reshape(-1)
        if ids.numel():
            node_send[q] = g2l_node[ids]
        ids = torch.nonzero(fm_q & (face_owner == rank)).reshape(-1)
        if ids.numel():
            face_send[q] = g2l_face[ids]
    halo.rows["node"] = dict(n_owned=n_own_nodes, n_local=int(node_gid.numel()), send=node_send, recv=node_recv, gid=node_gid)
    halo.rows["edge"] = dict(n_owned=n_own_faces, n_local=int(face_gid.numel()), send=face_send, recv=face_recv, gid=face_gid)
    halo.n_owned_cells, halo.cell_gid = n_own_cells, cell_gid

    # ---- local sub-mesh: same keys, renumbered
    m = {}
    for k in ("node|pos", "node|node_type", "target|uvp"):
        m[k] = _t(mesh[k], dev)[node_gid]
    luvp = _t(uvp, dev)[node_gid]
    m["face|face_node"] = g2l_node[fn[:, face_gid]]
    for k in ("face|face_type", "face|face_area", "face|face_center_pos"):
        m[k] = _t(mesh[k], dev)[face_gid]
    nbc = _t(mesh["face|neighbour_cell"], dev).long()[:, face_gid]
    lnb = g2l_cell[nbc]
    lnb = torch.where(lnb < 0, lnb.flip(0), lnb)  # a neighbour outside the sub-mesh: the face becomes one-sided
    m["face|neighbour_cell"] = torch.where(lnb < 0, torch.zeros_like(lnb), lnb)
    # cell slots: all slots of the local cells, grouped by local cell id in local order
    slot_ids = torch.nonzero(cmask[ci]).reshape(-1)
    lcell = g2l_cell[ci[slot_ids]]
    sorder = torch.sort(lcell, stable=True).indices
    slot_ids, lcell = slot_ids[sorder], lcell[sorder]
    m["cells_node"] = g2l_node[cn[slot_ids]]
    m["cells_face"] = g2l_face[cf[slot_ids]]
    m["cells_index"] = lcell
    m["unit_norm_v"] = _t(mesh["unit_norm_v"], dev).reshape(-1, 2)[slot_ids]
    m["cell|cells_area"] = _t(mesh["cell|cells_area"], dev).reshape(-1)[cell_gid]
    m["cell|centroid"] = cen[cell_gid]
    # WLSQ stencil: pairs with both ends local (direction kept), moments of the global mesh
    fx = _t(mesh["face_node_x"], dev).long()
    keep = torch.nonzero(nmask[fx[0]] & nmask[fx[1]]).reshape(-1)
    m["face_node_x"] = g2l_node[fx[:, keep]]
    m["A_node_to_node"] = _t(mesh["A_node_to_node"], dev)[node_gid]
    m["single_B_node_to_node"] = _t(mesh["single_B_node_to_node"], dev)[keep]
    se = _t(mesh["support_edge"], dev).long()
    if bool(nmask[se].all()):
        m["support_edge"] = g2l_node[se]
        m["extra_B_node_to_node"] = _t(mesh["extra_B_node_to_node"], dev)
    else:
        # the global support pair is not here: attach the (required) pair to two nodes of the OUTERMOST halo layer, whose gradients
        # nobody reads; same weight / moment definition (FVorder.py:7-86: w = 1/|d|, m = [dx, dy, dx^2/2, dy^2/2, dx dy])
        cand = g2l_node[torch.nonzero(outer_nodes).reshape(-1)]
        if cand.numel() < 2:
            raise RuntimeError("partition: no outer halo layer to park the support edge on (halo_layers too small?)")
        a, b = int(cand[-1]), int(cand[-2])
        sel = torch.tensor([[a, b], [b, a]], dtype=torch.int64, device=dev)
        d = (m["node|pos"][sel[0]] - m["node|pos"][sel[1]]).double()
        nm = int(_t(mesh["extra_B_node_to_node"], dev).shape[1])
        mom = d if nm == 2 else torch.cat([d, 0.5 * d ** 2, d[:, 0:1] * d[:, 1:2]], 1)
        w = 1.0 / d.norm(dim=1, keepdim=True)
        m["support_edge"] = sel
        m["extra_B_node_to_node"] = (w * mom).reshape(2, nm, 1).to(_t(mesh["extra_B_node_to_node"], dev).dtype)
    for k in ("theta_PDE", "sigma", "uvp_dim", "dt_graph"):
        m[k] = _t(mesh[k], dev)
    for k in ("order",):
        if k in mesh:
            m[k] = mesh[k]
    return m, luvp, halo


def mark_partition(graphs, halo):
    """Give ghost nodes / halo cells the graph id B + b and duplicate the per-graph rows accordingly, so every
    per-graph reduction of the hot path runs over the owned rows only (the dummy graphs' results are discarded)."""
    gn, gx, ge, gc, gi = graphs
    B = int(gn.batch.max().item()) + 1
    n_own = halo.rows["node"]["n_owned"]
    gn.batch = gn.batch.clone()
    gn.batch[n_own:] += B
    gc.batch = gc.batch.clone()
    gc.batch[halo.n_owned_cells:] += B
    for k in ("theta_PDE", "sigma", "uvp_dim", "dt_graph"):
        v = getattr(gi, k)
        setattr(gi, k, torch.cat([v, v], 0))
    if hasattr(gi, "x"):
        gi.x = torch.cat([gi.x, gi.x], 0)
    halo.num_graphs = B
    for g in graphs:                      # the batch objects now describe 2 B graphs (B real + B ghost collectors)
        if getattr(g, "num_graphs", None) is not None:
            g.num_graphs = 2 * B
    gn._fvgn_halo = halo
    return graphs

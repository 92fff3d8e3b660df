"""GraphPlan: the int32 CSR structures built once per batch topology from the Load_mesh batch layout.

Every list is the *stable* grouping of the reference's scatter entry order by destination, so the
deterministic CSR reductions sum in exactly the order a sequential index_add_ would
(oracle/fvgn_oracle.py: csr_stable / node_incidence_plan / wlsq_entry_plan are the bit-exact targets).

Inputs are the five batched graphs of src/Load_mesh/Graph_loader.py:503-784:
  graph_node  : x, edge_index[2,E], face=cells_node[K], pos, node_type, y, batch
  graph_node_x: face_node_x[2,X], support_edge[2,2B], A_node_to_node, single_B_node_to_node, extra_B_node_to_node
  graph_edge  : face_type, face_area, face=cells_face[K], pos (face centres), batch
  graph_cell  : cells_face_unv[K,2], cells_area, pos (centroids), face=cells_index[K], batch
"""
import torch

from . import _lib

CHUNK_ROWS = 4096


def _i32(t):
    return t.to(torch.int32).contiguous()


def csr_stable(dest, n):
    """rowptr[n+1] (int32), perm (int64): entries grouped by destination row, original order kept inside a row.
    Device tensors: fvgn_csr_build (csrc/plan_build.cu: counting sort + per-row ordering, no library sort, no host sync)."""
    dest = dest.reshape(-1)
    if dest.is_cuda:
        if dest.dtype not in (torch.int32, torch.int64):
            dest = dest.to(torch.int64)
        dest = dest.contiguous()
        m = int(dest.shape[0])
        ptr = torch.empty(n + 1, dtype=torch.int32, device=dest.device)
        perm = torch.empty(max(m, 1), dtype=torch.int32, device=dest.device)
        ws = torch.empty(int(_lib.load().fvgn_csr_build_workspace_bytes(n)), dtype=torch.uint8, device=dest.device)
        _lib.call("fvgn_csr_build", _lib.ptr(dest), int(dest.dtype == torch.int64), m, n, _lib.iptr(ptr), _lib.iptr(perm), _lib.ptr(ws),
                  _lib.stream_ptr(dest.device))
        return ptr, perm[:m].to(torch.int64)
    # host tensors: only the CPU test harness (tests/test_plan.py, the emulator runs) builds plans on the host
    dest = dest.to(torch.int64)
    perm = torch.sort(dest, stable=True).indices
    counts = torch.bincount(dest, minlength=n)
    rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dest.device)
    torch.cumsum(counts, 0, out=rowptr[1:])
    return rowptr.to(torch.int32), perm


def _chunks(batch, nseg, rows_per_chunk=CHUNK_ROWS):
    """Row chunks of <= rows_per_chunk rows that never straddle a graph: (chunks[U,3] = (graph, row begin, row end),
    chunk_ptr[nseg+1], U).  Built with device ops only (no host round trip): U = N // rows_per_chunk + nseg is an upper
    bound of the chunk count; slots past chunk_ptr[nseg] are empty chunks (0, 0, 0) that no combine ever reads.
    `batch` must be sorted by graph (Load_mesh batches are)."""
    dev = batch.device
    n = int(batch.shape[0])
    U = max(n // rows_per_chunk + nseg, 1)
    counts = torch.bincount(batch.reshape(-1).to(torch.int64), minlength=nseg)[:nseg]
    row_ptr = torch.zeros(nseg + 1, dtype=torch.int64, device=dev)
    torch.cumsum(counts, 0, out=row_ptr[1:])
    cpg = (counts + rows_per_chunk - 1) // rows_per_chunk
    chunk_ptr = torch.zeros(nseg + 1, dtype=torch.int64, device=dev)
    torch.cumsum(cpg, 0, out=chunk_ptr[1:])
    ids = torch.arange(U, dtype=torch.int64, device=dev)
    seg = torch.searchsorted(chunk_ptr[1:].contiguous(), ids, right=True).clamp(max=nseg - 1)
    begin = row_ptr[seg] + (ids - chunk_ptr[seg]) * rows_per_chunk
    end = torch.minimum(begin + rows_per_chunk, row_ptr[seg + 1])
    live = ids < chunk_ptr[nseg]
    zero = torch.zeros_like(ids)
    chunks = torch.stack([torch.where(live, seg, zero), torch.where(live, begin, zero), torch.where(live, end, zero)], 1)
    return chunks.to(torch.int32).contiguous(), chunk_ptr.to(torch.int32), U


def num_graphs_of(graph, batch):
    """Number of graphs of a batch object without a device round trip when the loader says so (PyG Batch.num_graphs)."""
    ng = getattr(graph, "num_graphs", None)
    if ng is not None:
        return int(ng)
    return int(batch.max().item()) + 1 if batch.numel() > 0 else 1


def content_hash(tensors):
    """64-bit content hash of a list of device tensors (fvgn_hash_words; one 8-byte device -> host read)."""
    dev = tensors[0].device
    acc = torch.zeros(1, dtype=torch.int64, device=dev)
    st = _lib.stream_ptr(dev)
    for i, t in enumerate(tensors):
        t = t.contiguous()
        nbytes = t.numel() * t.element_size()
        if nbytes % 4 != 0:   # bool / int8 / 16-bit tensors of odd length: widen
            t = t.to(torch.int32)
            nbytes = t.numel() * 4
        _lib.call("fvgn_hash_words", _lib.ptr(t), nbytes // 4, 1000003 * (i + 1) + t.numel(), _lib.ptr(acc), st)
    return int(acc.item())


class GraphPlan:
    """Device-resident topology plan of one batch.  Build with GraphPlan.build(...) or GraphPlan.of(...) (cached)."""

    _by_content = {}   # content hash -> plan (most recent MAX_CONTENT_PLANS topologies; a plan of a 4 M-cell mesh holds ~2 GB)
    MAX_CONTENT_PLANS = 4

    @staticmethod
    def clear_cache():
        """Drop the content-keyed plans (plans attached to live batch objects stay with them)."""
        GraphPlan._by_content.clear()

    @staticmethod
    def _plan_inputs(graph_node, graph_node_x, graph_edge, graph_cell):
        ts = [graph_node.edge_index]
        if getattr(graph_node, "batch", None) is not None:
            ts.append(graph_node.batch)
        if graph_node_x is not None:
            ts += [graph_node.face, graph_node.pos, graph_node.node_type, graph_node.y, graph_node_x.face_node_x,
                   graph_node_x.support_edge, graph_node_x.A_node_to_node, graph_node_x.single_B_node_to_node,
                   graph_node_x.extra_B_node_to_node, graph_edge.face, graph_edge.pos, graph_edge.face_area, graph_edge.face_type,
                   graph_cell.face, graph_cell.pos, graph_cell.cells_area, graph_cell.cells_face_unv, graph_cell.batch]
        return ts

    @staticmethod
    def of(graph_node, graph_node_x=None, graph_edge=None, graph_cell=None, order="2nd"):
        """The plan of this batch.  Fast path: the batch object (or one sharing its index tensors) was seen before.  A
        loader that hands a NEW batch object every step (Graph_loader.py:830-1006) is recognised by CONTENT: a 64-bit
        hash of every tensor the plan is built from (connectivity, geometry, boundary targets, WLSQ matrices) costs one
        read pass + an 8-byte device -> host copy instead of the ~25 plan-building launches."""
        key = (graph_node.edge_index.data_ptr(), tuple(graph_node.edge_index.shape))
        key_fv = None if graph_node_x is None else (graph_node_x.face_node_x.data_ptr(), graph_cell.face.data_ptr(), order)
        plan = getattr(graph_node, "_fvgn_plan", None)
        if plan is not None and plan.key == key and (key_fv is None or plan.key_fv == key_fv):
            return plan
        halo = getattr(graph_node, "_fvgn_halo", None)
        h = None
        if graph_node.edge_index.is_cuda and halo is None:
            h = (content_hash(GraphPlan._plan_inputs(graph_node, graph_node_x, graph_edge, graph_cell)), order,
                 graph_node_x is not None, str(graph_node.edge_index.device))
            plan = GraphPlan._by_content.get(h)
            if plan is not None:
                graph_node._fvgn_plan = plan
                plan.key, plan.key_fv = key, key_fv
                return plan
        plan = GraphPlan.build(graph_node, graph_node_x, graph_edge, graph_cell, order)
        plan.key, plan.key_fv = key, key_fv
        plan.halo = halo  # cell-partition mode (partition.py)
        graph_node._fvgn_plan = plan
        if h is not None:
            if len(GraphPlan._by_content) >= GraphPlan.MAX_CONTENT_PLANS:
                GraphPlan._by_content.pop(next(iter(GraphPlan._by_content)))
            GraphPlan._by_content[h] = plan
        return plan

    @staticmethod
    def build(graph_node, graph_node_x=None, graph_edge=None, graph_cell=None, order="2nd"):
        p = GraphPlan()
        ei = graph_node.edge_index
        dev = ei.device
        p.device = dev
        N = int(graph_node.pos.shape[0]) if getattr(graph_node, "pos", None) is not None else int(graph_node.x.shape[0])
        E = int(ei.shape[1])
        p.N, p.E = N, E
        s, r = ei[0].to(torch.int64), ei[1].to(torch.int64)
        p.edge_s, p.edge_r = _i32(s), _i32(r)
        # node incidence, entry order = cat(senders, receivers)  (blocks.py:24-42,84-99)
        p.inc_ptr, perm = csr_stable(torch.cat([s, r]), N)
        edge = perm % max(E, 1)
        role = perm // max(E, 1)
        p.inc_code = _i32(edge * 2 + role)
        p.inc_nbr = _i32(torch.where(role == 0, r[edge], s[edge])) if E > 0 else _i32(edge)
        batch = getattr(graph_node, "batch", None)
        if batch is None:
            batch = torch.zeros(N, dtype=torch.int64, device=dev)
        p.batch_node = _i32(batch)
        p.B = num_graphs_of(graph_node, batch) if N > 0 else 1
        p.node_chunks, p.node_chunk_ptr, p.n_node_chunks = _chunks(batch, p.B)
        p.has_fv = False
        p.halo = None
        if graph_node_x is not None:
            p._build_fv(graph_node, graph_node_x, graph_edge, graph_cell, order)
        return p

    # ------------------------------------------------------------------ finite-volume part
    def _build_fv(self, graph_node, graph_node_x, graph_edge, graph_cell, order):
        dev, N, E = self.device, self.N, self.E
        f32 = torch.float32
        cells_node = graph_node.face.reshape(-1).to(torch.int64)
        cells_face = graph_edge.face.reshape(-1).to(torch.int64)
        cells_index = graph_cell.face.reshape(-1).to(torch.int64)
        C = int(graph_cell.pos.shape[0])
        K = int(cells_index.shape[0])
        self.C, self.K = C, K
        # slots sorted by cell (stable); identity for the reference's layouts (cells_index non-decreasing)
        self.cell_ptr, sperm = csr_stable(cells_index, C)
        self.slot_cell = _i32(cells_index[sperm])
        self.slot_node = _i32(cells_node[sperm])
        self.slot_face = _i32(cells_face[sperm])
        self.slot_unv = graph_cell.cells_face_unv.reshape(-1, 2)[sperm].to(f32).contiguous()
        self.face_slot_ptr, fperm = csr_stable(cells_face[sperm], E)
        self.face_slot = _i32(fperm)
        self.node_slot_ptr, nperm = csr_stable(cells_node[sperm], N)
        self.node_slot = _i32(nperm)
        self.batch_cell = _i32(graph_cell.batch)
        self.cell_chunks, self.cell_chunk_ptr, self.n_cell_chunks = _chunks(graph_cell.batch, self.B)
        # geometry (fp32 copies in the layout the kernels read)
        self.pos = graph_node.pos.to(f32).contiguous()
        # Dirichlet targets: the reference reads graph_node.y[:, 0:2] (importer.py:141-154); a loader may carry more columns
        if graph_node.y.dim() != 2 or graph_node.y.shape[1] < 2:
            raise ValueError(f"graph_node.y must be [N, >=2] (u, v targets), got {tuple(graph_node.y.shape)}")
        self.y = graph_node.y[:, 0:2].to(f32).contiguous()
        self.node_type = _i32(graph_node.node_type.reshape(-1))
        self.face_pos = graph_edge.pos.to(f32).contiguous()
        self.face_area = graph_edge.face_area.reshape(-1).to(f32).contiguous()
        self.face_type = _i32(graph_edge.face_type.reshape(-1))
        self.centroid = graph_cell.pos.to(f32).contiguous()
        self.cells_area = graph_cell.cells_area.reshape(-1).to(f32).contiguous()
        # WLSQ stencil: directed entries cat(fx, flip(fx), support_edge), (out -> in), grouped by 'in'  (FVgrad.py:264-273)
        fx = graph_node_x.face_node_x.to(torch.int64)
        se = graph_node_x.support_edge.to(torch.int64)
        out_i = torch.cat([fx[0], fx[1], se[0]])
        in_i = torch.cat([fx[1], fx[0], se[1]])
        self.w_ptr, wperm = csr_stable(in_i, N)
        self.w_col = _i32(out_i[wperm])
        A = graph_node_x.A_node_to_node.to(f32).contiguous()
        nm = int(A.shape[-1])
        B1 = graph_node_x.single_B_node_to_node.reshape(-1, nm).to(f32)
        Bx = graph_node_x.extra_B_node_to_node.reshape(-1, nm).to(f32)
        flip = B1.clone()
        flip[:, 0:2] *= -1  # FVgrad.py:301-306
        moments = torch.cat([B1, flip, Bx], 0)[wperm].contiguous()
        del flip
        self.w_nm = nm
        self.w_moments, self.w_A = moments, A
        self.w_row = _i32(in_i[wperm])
        self.w_tptr, self.w_tperm = csr_stable(self.w_col, N)
        self.w_trow = _i32(self.w_row[self.w_tperm].to(torch.int64))
        self._wq = {}
        self.wlsq_weights(2)
        self.has_fv = True

    def wlsq_weights(self, nq):
        """(q[nnz,nq], qsum[N,nq], qT[nnz,nq]) with the fp64 inverse moment matrix folded in (cached per nq)."""
        if nq not in self._wq:
            nnz = int(self.w_col.shape[0])
            q = torch.empty((nnz, nq), dtype=torch.float32, device=self.device)
            qsum = torch.empty((self.N, nq), dtype=torch.float32, device=self.device)
            _lib.call("fvgn_wlsq_weights", _lib.fptr(self.w_A), self.w_nm, _lib.iptr(self.w_ptr), _lib.fptr(self.w_moments),
                      nq, _lib.fptr(q), _lib.fptr(qsum), self.N, _lib.stream_ptr(self.device))
            self._wq[nq] = (q, qsum, q[self.w_tperm].contiguous())
        return self._wq[nq]

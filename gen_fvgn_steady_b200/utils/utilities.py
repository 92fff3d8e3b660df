"""NodeType enum and the cell <-> node pooling helpers of the reference (src/utils/utilities.py:7-61)."""
import enum


class NodeType(enum.IntEnum):
    NORMAL = 0
    INFLOW = 1
    OUTFLOW = 2
    WALL_BOUNDARY = 3
    PRESS_POINT = 4
    IN_WALL = 5


def _flat(cells_node, cells_index):
    if cells_node.shape != cells_index.shape:
        raise ValueError("wrong cells_node/cells_index dim")
    return cells_node.reshape(-1), cells_index.reshape(-1)


def calc_cell_centered_with_node_attr(node_attr, cells_node, cells_index, reduce="mean", map=True):
    """utilities.py:16-35: cell value = reduce over the cell's vertex slots of node_attr[cells_node] (map=True) or of the
    per-slot values themselves (map=False).  The scatter is the deterministic CSR segment sum (ops.segment_sum): fp32 sums
    in slot order, no atomics; differentiable.  Number of cells = max(cells_index) + 1 as torch_scatter infers it."""
    from .. import ops
    cells_node, cells_index = _flat(cells_node, cells_index)
    mapped = node_attr[cells_node.long()] if map else node_attr
    n = int(cells_index.max().item()) + 1 if cells_index.numel() > 0 else 0
    return ops.segment_sum(mapped, cells_index, n, reduce)


def calc_node_centered_with_cell_attr(cell_attr, cells_node, cells_index, reduce="mean", map=True):
    """utilities.py:38-61: node value = reduce over the node's cell-vertex slots of cell_attr[cells_index]."""
    from .. import ops
    cells_node, cells_index = _flat(cells_node, cells_index)
    mapped = cell_attr[cells_index.long()] if map else cell_attr
    n = int(cells_node.max().item()) + 1 if cells_node.numel() > 0 else 0
    return ops.segment_sum(mapped, cells_node, n, reduce)

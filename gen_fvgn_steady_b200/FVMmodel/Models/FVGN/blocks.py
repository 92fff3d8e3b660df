"""EdgeBlock / NodeBlock with the reference's constructor, attribute and state_dict layout
(src/FVMmodel/Models/FVGN/blocks.py:8-120); the forward passes are single fused CUDA ops."""
from torch import nn

from ....data import Data
from .... import ops
from ....plan import GraphPlan


def mlp_params(net):
    """Flat parameter list [w1,b1,w2,b2,w3,b3(,ln_g,ln_b)] of a build_mlp / build_mlp_from_num_layer module."""
    if isinstance(net[0], nn.Sequential):
        seq, ln = net[0], net[1]
        return [seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias, seq[4].weight, seq[4].bias, ln.weight, ln.bias]
    return [net[0].weight, net[0].bias, net[2].weight, net[2].bias, net[4].weight, net[4].bias]


def _precision(module):
    return getattr(module, "precision", None) or ops.default_precision()


class EdgeBlock(nn.Module):
    def __init__(self, input_size=None, custom_func=None):
        super().__init__()
        self.net = custom_func

    def forward(self, graph_node, graph_cell=None):
        plan = GraphPlan.of(graph_node)
        e_new = ops.apply(ops.EdgeBlockFn, graph_node.x, graph_node.edge_attr, plan, _precision(self), *mlp_params(self.net))
        return Data(x=graph_node.x, edge_attr=e_new, edge_index=graph_node.edge_index, face=getattr(graph_node, "face", None),
                    num_graphs=getattr(graph_node, "num_graphs", None), batch=getattr(graph_node, "batch", None),
                    _fvgn_plan=plan)


class NodeBlock(nn.Module):
    def __init__(self, input_size=None, custom_func=None):
        super().__init__()
        self.net = custom_func

    def forward(self, graph_node, graph_cell=None):
        plan = GraphPlan.of(graph_node)
        x_new = ops.apply(ops.NodeBlockFn, graph_node.x, graph_node.edge_attr, plan, _precision(self), *mlp_params(self.net))
        return Data(x=x_new, edge_attr=graph_node.edge_attr, edge_index=graph_node.edge_index,
                    face=getattr(graph_node, "face", None), num_graphs=getattr(graph_node, "num_graphs", None),
                    batch=getattr(graph_node, "batch", None), _fvgn_plan=plan)

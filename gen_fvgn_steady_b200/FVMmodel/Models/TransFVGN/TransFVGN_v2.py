"""TransFVGN_v2 Simulator (src/FVMmodel/Models/TransFVGN/TransFVGN_v2.py:12-104): Encoder -> 2 x (GnBlock x mp -> Transolver) -> Decoder."""
from torch import nn

from ..FVGN.EPD import Encoder, Decoder, GnBlock
from ..GraphTransolver.GraphTransolver import Transolver_block


class AttnProcessor(nn.Module):
    def __init__(self, message_passing_num=0, hidden_size=128, drop_out=False):
        super().__init__()
        if message_passing_num < 1:
            raise ValueError("message_passing_num must be greater than 0")
        self.GN_block_list = nn.ModuleList([GnBlock(hidden_size=hidden_size, drop_out=drop_out)
                                            for _ in range(message_passing_num)])
        self.TransBlock = Transolver_block(num_heads=8, hidden_dim=hidden_size, dropout=0, act="gelu", mlp_ratio=2, slice_num=32)

    def forward(self, latent_graph_node, graph_edge, first_block=0, total_blocks=None):
        from ....parallel import halo_refresh
        node_embedding = latent_graph_node.x
        latent = latent_graph_node
        total = total_blocks if total_blocks is not None else len(self.GN_block_list)
        from ....parallel import no_ghost_refresh
        whole = no_ghost_refresh(latent, total)
        for i, model in enumerate(self.GN_block_list):
            last = i == len(self.GN_block_list) - 1   # the Transolver block reads x in fp32
            latent = model(latent, keep_edge_latent=not (whole and first_block + i == total - 1),   # nothing reads the last e
                           latents_16bit=whole, x_fp32=last)
            latent = halo_refresh(latent, first_block + i, total)  # cell-partition mode only (no-op otherwise)
        latent.x = self.TransBlock(latent.x, latent.batch, halo=getattr(latent, "_fvgn_halo", None), num_graphs=getattr(latent, "num_graphs", None),
                                   embedding=node_embedding)
        latent._xh = self.TransBlock.last_shadow   # bf16 mode: (x, shadow) written by the block's last kernel
        return latent


class Simulator(nn.Module):
    def __init__(self, message_passing_num, edge_input_size, node_input_size, node_output_size, drop_out=False,
                 hidden_size=128, params=None):
        super().__init__()
        self.encoder = Encoder(node_input_size=node_input_size, edge_input_size=edge_input_size, hidden_size=hidden_size)
        self.processpr_list = nn.ModuleList([AttnProcessor(message_passing_num=message_passing_num, hidden_size=hidden_size,
                                                           drop_out=False) for _ in range(2)])
        self.decoder = Decoder(hidden_sze=hidden_size, node_output_size=node_output_size)

    def forward(self, graph_node=None, graph_edge=None, graph_cell=None):
        from ....parallel import no_ghost_refresh
        whole_ = no_ghost_refresh(graph_node, sum(len(m.GN_block_list) for m in self.processpr_list))
        latent, _ = self.encoder(graph_node, latents_16bit=whole_, x_fp32=True)   # x: the Transolver embedding of processor 1
        total = sum(len(m.GN_block_list) for m in self.processpr_list)
        first = 0
        for model in self.processpr_list:
            latent = model(latent, graph_edge, first, total)
            first += len(model.GN_block_list)
        return self.decoder(latent)

"""TransFVGN_v1 Simulator (src/FVMmodel/Models/TransFVGN/TransFVGN_v1.py:10-73): Encoder -> GnBlock x mp -> Transolver -> Decoder."""
from torch import nn

from ..FVGN.EPD import Encoder, Decoder, GnBlock
from ..GraphTransolver.GraphTransolver import Transolver_block


class Simulator(nn.Module):
    def __init__(self, message_passing_num, edge_input_size, node_input_size, node_output_size, drop_out=False,
                 hidden_size=128, params=None):
        super().__init__()
        self.encoder = Encoder(node_input_size=node_input_size, edge_input_size=edge_input_size, hidden_size=hidden_size)
        self.GN_block_list = nn.ModuleList([GnBlock(hidden_size=hidden_size, drop_out=drop_out)
                                            for _ in range(message_passing_num)])
        self.TransBlock = Transolver_block(num_heads=8, hidden_dim=hidden_size, dropout=0, act="gelu", mlp_ratio=2, slice_num=32)
        self.decoder = Decoder(hidden_sze=hidden_size, node_output_size=node_output_size)

    def forward(self, graph_node=None, graph_edge=None, graph_cell=None):
        from ....parallel import halo_refresh
        from ....parallel import no_ghost_refresh
        whole_ = no_ghost_refresh(graph_node, len(self.GN_block_list))
        latent, node_embedding = self.encoder(graph_node, latents_16bit=whole_, x_fp32=True)   # x: the Transolver embedding
        nblk = len(self.GN_block_list)
        whole = whole_
        for i, model in enumerate(self.GN_block_list):
            latent = model(latent, keep_edge_latent=not (whole and i == nblk - 1),   # nothing reads the last edge latent
                           latents_16bit=whole, x_fp32=i == nblk - 1)                        # the Transolver block reads x in fp32
            latent = halo_refresh(latent, i, nblk)  # cell-partition mode only (no-op otherwise)
        latent.x = self.TransBlock(latent.x, graph_node.batch, halo=getattr(latent, "_fvgn_halo", None), num_graphs=getattr(latent, "num_graphs", None),
                                   embedding=node_embedding)
        latent._xh = self.TransBlock.last_shadow   # bf16 mode: (x, shadow) written by the block's last kernel
        return self.decoder(latent)

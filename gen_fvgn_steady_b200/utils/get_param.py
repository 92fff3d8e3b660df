"""Default hyper-parameters of the reference hot path (src/utils/get_param.py:37-75), as a Namespace."""
import argparse


def params(**overrides):
    p = argparse.Namespace(
        net="TransFVGN_v2", batch_size=8, dataset_size=100, lr=5e-5, integrator="imex", norm_uvp=True, norm_global=True,
        ncn_smooth=True, conserved_form=True, order="2nd", loss_cont=6e4, loss_mom=5e4, loss_press=1.0,
        hidden_size=128, message_passing_num=3, node_phi_size=3, node_input_size=12, node_output_size=3,
        precision=None,  # "fp32" (SIMT parity) | "bf16" | "f16" (tcgen05 modes; f16 = TF32-grade significand); None -> $FVGN_PRECISION or fp32
    )
    for k, v in overrides.items():
        setattr(p, k, v)
    return p

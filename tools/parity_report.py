"""GPU: errors of the product against the reference's fp64 golden vectors for every golden case in every precision mode.

    python tools/parity_report.py [out.json] [mode ...]

Per (case, mode): relative error of the four loss terms, the script loss, uvp_node / uvp_cell, the decoder output, the WLSQ
gradient and the worst parameter gradient (stride-61 sample + norm, as tests/product_util.compare_with_golden measures them),
next to the reference's own fp32-vs-fp64 gap.  The asserted bars of tests/test_gpu_parity.py were chosen from this table.
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from tests import golden_util as GU  # noqa: E402
from tests import product_util as PU  # noqa: E402


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/parity_report.json"
    modes = sys.argv[2:] or ["fp32", "f16", "bf16"]
    PU.use_real_kernels()
    table = {}
    for name in GU.CASES:
        row = {}
        for mode in modes:
            rep = {}
            try:
                model, o, loss, z = PU.run_product(name, "cuda", mode)
                PU.compare_with_golden(model, o, loss, z, "f64", tol=1e30, gtol=1e30, report=rep)
                row[mode] = {k: (v if isinstance(v, str) or v is None else float(v)) for k, v in rep.items()}
            except Exception as e:  # noqa: BLE001
                row[mode] = {"error": repr(e)}
            torch.cuda.synchronize()
        z = GU.load_case(name)
        row["ref_fp32_gap"] = {k: GU.rel_err(z[f"f32.{k}"], z[f"f64.{k}"]) for k in
                               ("loss_cont", "loss_mom_x", "loss_mom_y", "loss_press", "uvp_node", "uvp_cell", "decoder_out", "grad_phi")}
        table[name] = row
        print(name, json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
    with open(out, "w") as f:
        json.dump(table, f, indent=1)


if __name__ == "__main__":
    main()

"""Debug: where does the captured step diverge from the eager step?"""
import copy
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import product_util as PU  # noqa: E402
from gen_fvgn_steady_b200.FVMmodel.importer import NNmodel  # noqa: E402
from gen_fvgn_steady_b200.graphed import GraphedTrainStep  # noqa: E402
from gen_fvgn_steady_b200.utils.get_param import params as default_params  # noqa: E402
from gen_fvgn_steady_b200.mesh import synthetic as S  # noqa: E402
from tests.case_inputs import product_graphs  # noqa: E402

PU.use_real_kernels()
dev = torch.device("cuda")
mesh, uvp = S.make_case(20, kind="mixed", bc="channel", seed=2)
prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
p = default_params(net="EPD", message_passing_num=2, dataset_size=1, precision=prec)
torch.manual_seed(0)
model_a = NNmodel(p).to(dev)
model_b = copy.deepcopy(model_a)
loss_fn = lambda out: PU.script_loss(out, p)

graphs = product_graphs([mesh], [uvp], dev)
x0 = graphs[0].x.clone()
opt_a = torch.optim.Adam(model_a.parameters(), lr=1e-3, fused=True, capturable=True)
snaps = []
for it in range(3):
    graphs[0].x, graphs[0].norm_uvp, graphs[0].norm_global = x0.clone(), True, True
    opt_a.zero_grad(set_to_none=True)
    loss = loss_fn(model_a(*graphs, is_training=True))
    loss.backward()
    grads = {k: v.grad.detach().clone() for k, v in model_a.named_parameters()}
    opt_a.step()
    snaps.append((float(loss), grads, {k: v.detach().clone() for k, v in model_a.named_parameters()}))
    print("eager", it, float(loss))

graphs = product_graphs([mesh], [uvp], dev)
opt_b = torch.optim.Adam(model_b.parameters(), lr=1e-3, fused=True, capturable=True)
gs = GraphedTrainStep(model_b, opt_b, graphs, loss_fn, warmup=3)
for k, v in model_b.named_parameters():
    st = opt_b.state[v]
    if float(st["step"]) != 0 or float(st["exp_avg"].abs().max()) != 0:
        print("state not reset", k, float(st["step"]))
        break
for it in range(3):
    l = float(gs.step())
    print("graph", it, l, "eager", snaps[it][0])
    worst = []
    for k, v in model_b.named_parameters():
        ge = snaps[it][1][k]
        gg = v.grad
        dg = float((gg - ge).abs().max()) / max(float(ge.abs().max()), 1e-30)
        dw = float((v.detach() - snaps[it][2][k]).abs().max())
        worst.append((dg, dw, k, float(opt_b.state[v]["step"])))
    worst.sort(reverse=True)
    for w in worst[:6]:
        print("   grad rel diff %.3e  weight abs diff %.3e  %s step=%g" % w)

# ---- lr = 0: every replay must be identical
print("=== lr=0 experiment")
model_a._last = None
model_c = copy.deepcopy(model_a)
graphs = product_graphs([mesh], [uvp], dev)
opt_c = torch.optim.Adam(model_c.parameters(), lr=0.0, fused=True, capturable=True)
gs = GraphedTrainStep(model_c, opt_c, graphs, loss_fn, warmup=3)
prev = None
for it in range(3):
    l = float(gs.step())
    cur = {"dec": model_c._last["decoder_out"].detach().clone(), "phi": model_c._last["phi"].detach().clone(),
           "out0": gs.out[0].detach().clone(), "uvp": gs.out[4].detach().clone()}
    cur.update({"g:" + k: v.grad.detach().clone() for k, v in model_c.named_parameters()})
    print("replay", it, l)
    if prev is not None:
        bad = [(k, float((cur[k] - prev[k]).abs().max())) for k in cur if not torch.equal(cur[k], prev[k])]
        print("   differing:", bad[:10])
    prev = cur
# eager forward on model_c (weights untouched by lr=0)
graphs2 = product_graphs([mesh], [uvp], dev)
l2 = float(loss_fn(model_c(*graphs2, is_training=True)))
print("eager forward after replays:", l2)

print("=== in-place weight perturbation under lr=0 graph")
names = [k for k, _ in model_c.named_parameters()]
groups = {"all": names, "weights(2D)": [k for k in names if dict(model_c.named_parameters())[k].dim() == 2],
          "biases/ln(1D)": [k for k in names if dict(model_c.named_parameters())[k].dim() == 1]}
for k in names:
    groups[k] = [k]
named = dict(model_c.named_parameters())
base = {k: v.detach().clone() for k, v in named.items()}
for gname, ks in groups.items():
    with torch.no_grad():
        for k, v in named.items():
            v.copy_(base[k])
        torch.manual_seed(1)
        for k in ks:
            named[k].add_(0.01 * torch.randn_like(named[k]))
    lg = float(gs.step())
    graphs2 = product_graphs([mesh], [uvp], dev)
    with torch.no_grad():
        le = float(loss_fn(model_c(*graphs2, is_training=True)))
    if lg != le or gname in ("all", "weights(2D)", "biases/ln(1D)"):
        print(f"{gname:60s} graph {lg:.9f} eager {le:.9f} {'MISMATCH' if lg != le else ''}")

"""A/B of a handle option on the one-node fitting step (BASELINE config 3):
python tools/fit_ab.py [B] [option] -> eager and CUDA-graph ms per step with option = 1 and 0, gradients compared."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import smplk
from smplk import synthetic
from smplk.body_models import fit_vertex_l2

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
opt = sys.argv[2] if len(sys.argv) > 2 else "bwd_overlap"
dev = torch.device("cuda:0")
model = synthetic.make_model("smplh", seed=0)
grads = {}
for val in (1, 0, 1, 0):
    dm = smplk.DeviceModel(model, device=0, options={opt: val})
    b, p, t = (torch.tensor(x, device=dev, requires_grad=True) for x in synthetic.make_inputs(model, B, seed=1))
    tgt = torch.randn(B, dm.V, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(3)) * 0.01
    with torch.no_grad():
        tgt += smplk.body_models.body_model_apply(dm, b + 0.05, p + 0.01, transl=t)[0]

    def step():
        for x in (b, p, t):
            x.grad = None
        fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()

    for _ in range(5):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        step()
    e1.record()
    torch.cuda.synchronize()
    eager = e0.elapsed_time(e1) / 50
    grads[val] = [x.grad.clone() for x in (b, p, t)]
    # the same step as a CUDA graph
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    for x in (b, p, t):
        x.grad = None
    with torch.cuda.graph(g):
        fit_vertex_l2(dm, b, p, tgt, transl=t).sum().backward()
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(50):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    graph = e0.elapsed_time(e1) / 50
    ggrads = [x.grad.clone() for x in (b, p, t)]
    same = all(torch.equal(a, c) for a, c in zip(grads[val], ggrads))
    print("%s=%d B=%d eager %.4f ms  graph %.4f ms  graph grads == eager grads: %s" % (opt, val, B, eager, graph, same), flush=True)
print("grads equal across option values:", all(torch.equal(a, c) for a, c in zip(grads[1], grads[0])))
print("grad digests (sum, sum |.|) betas / pose / transl:", [(float(g.double().sum()), float(g.double().abs().sum())) for g in grads[1]])

"""Prints shasta_adam_step_f32 (training.StreamAdam) next to torch.optim.Adam for tiny tensors (debug aid)."""
import torch
from shasta_b200 import training

for n in (1, 5):
    gen = torch.Generator(device="cpu").manual_seed(n)
    p0 = torch.randn(n, generator=gen)
    grads = [torch.randn(n, generator=gen) * (10.0 ** (k - 1)) for k in range(3)]
    kw = dict(lr=1e-2, weight_decay=1e-2, betas=(0.9, 0.999), eps=1e-8)
    pa = torch.nn.Parameter(p0.clone().cuda())
    pb = torch.nn.Parameter(p0.clone().cuda())
    oa = torch.optim.Adam([pa], **kw)
    ob = training.StreamAdam([pb], **kw)
    for g in grads:
        pa.grad = g.cuda()
        pb.grad = g.cuda()
        oa.step()
        ob.step()
        torch.cuda.synchronize()
        print(n, "p", pa.data.cpu().numpy(), pb.data.cpu().numpy(), "m", oa.state[pa]["exp_avg"].cpu().numpy(),
              ob.state[pb]["exp_avg"].cpu().numpy(), "v", oa.state[pa]["exp_avg_sq"].cpu().numpy(),
              ob.state[pb]["exp_avg_sq"].cpu().numpy(), "step", oa.state[pa]["step"], ob.state[pb]["step"])

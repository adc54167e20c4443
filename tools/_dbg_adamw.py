import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_sharding_gpu import _models
from prfl_b200.sharding import ShardedAdamW
cfg, sd, inp, make, kw = _models("t2v", 70)
a, b = make(), make()
opt_a = ShardedAdamW(a, lr=1e-3, weight_decay=0.01).attach_hooks()
opt_b = torch.optim.AdamW(b.parameters(), lr=1e-3, weight_decay=0.01)
for step in range(2):
    g = torch.Generator().manual_seed(300 + step)
    x = [torch.randn(inp["x"][0].shape, generator=g).cuda()]
    cot = torch.randn(16, *inp["x"][0].shape[1:], generator=g).cuda()
    (a(x=x, **kw)[0] * cot).sum().backward()
    shards = opt_a.reduce_gradients()
    (b(x=x, **kw)[0] * cot).sum().backward()
    names = dict(b.named_parameters())
    worst = []
    for ui, u in enumerate(opt_a.units):
        for n, (o, cnt, shp) in u.offsets.items():
            full = (u.sink.prefix + n) if u.kind == "resident" else n
            want = names[full].grad
            got = shards[ui][o:o + cnt].view(shp)
            if want is None:
                d = float(got.abs().max()); worst.append((d, full, "None-grad"))
            else:
                d = float((got - want.float()).abs().max() / (want.float().abs().max() + 1e-30)); worst.append((d, full, float(want.abs().max())))
    worst.sort(reverse=True)
    print(f"step {step} grad diffs (top 6):", worst[:6])
    na = opt_a.step(max_norm=1.0)
    for p in b.parameters():
        if p.grad is None:
            p.grad = torch.zeros_like(p)
    nb = torch.nn.utils.clip_grad_norm_(b.parameters(), 1.0)
    opt_b.step(); opt_b.zero_grad(set_to_none=True)
    print("norms", float(na), float(nb))
    full = opt_a.full_state_dict()
    diffs = sorted(((float((full[k].float() - v.float().cpu()).abs().max() / (v.float().abs().max().cpu() + 1e-12)), k) for k, v in b.state_dict().items()), reverse=True)
    print(f"step {step} weight diffs (top 8):", diffs[:8])
    k = diffs[0][1]
    da = (full[k].float() - b.state_dict()[k].float().cpu()).abs().flatten()
    i = int(da.argmax())
    print("worst element", k, i, float(full[k].flatten()[i]), float(b.state_dict()[k].flatten()[i].cpu()))

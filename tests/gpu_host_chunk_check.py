import sys, torch, time
sys.path.insert(0, '/root/repo')
import robustcap_b200 as rb
from robustcap_b200 import synthetic
assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
net = rb.Net(rb.ParametricModel(assets['smpl_file']))
net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
for B, T, ragged in ((256, 100, False), (300, 95, True), (1024, 300, False)):
    inp = synthetic.make_inputs(B, T, seed=11, conf='mixed')
    rb.Net.gravityc = inp['gravity'].clone()
    lengths = (torch.arange(B) * 7 % (T - 20) + 20).to(torch.int32) if ragged else None
    ft = torch.tensor([0., 0., 4.])
    pd, td = net.forward_offline(inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda(), first_tran=ft, lengths=lengths)
    pin = lambda x: x.contiguous().pin_memory()
    hj, ha, ho = pin(inp['j2dc']), pin(inp['accc']), pin(inp['oric'])
    hp = torch.empty(B, T, 24, 3, 3).pin_memory(); ht = torch.empty(B, T, 3).pin_memory()
    for _ in range(2):
        net.forward_offline(hj, ha, ho, first_tran=ft, lengths=lengths, out=(hp, ht))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        net.forward_offline(hj, ha, ho, first_tran=ft, lengths=lengths, out=(hp, ht))
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 3 * 1e3
    print('B=%d T=%d ragged=%s: host path == device path: %s %s | e2e %.1f ms = %.0f frames/s' % (B, T, ragged, bool(torch.equal(hp, pd.cpu())), bool(torch.equal(ht, td.cpu())), ms, B * T / ms * 1e3), flush=True)

"""GPU aid (not a test): A/B of the grouped-kernel variants (RC_PH_PAIR = 0: 128 x 128 tiles per CTA, 1: 256 x 128 CTA-pair tiles,
2: 256 x 256 CTA-pair tiles) — timing with CUDA events and the pose difference between the variants and against the fp32 SIMT
back end on the same batch.  The variant is a per-process switch, so every variant runs in a child process.
Usage: python tests/gpu_pair_ab.py [B] [T] [variants, e.g. 0,2] [conf]"""
import os
import subprocess
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def child(B, T, conf, out):
    import robustcap_b200 as rb
    from robustcap_b200 import synthetic
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    net = rb.Net(rb.ParametricModel(assets['smpl_file']))
    net.load_state_dict(synthetic.make_state_dict(0, 'contact'))
    inp = synthetic.make_inputs(B, T, seed=1000, conf=conf)
    rb.Net.gravityc = inp['gravity'].clone()
    j, a, o = inp['j2dc'].cuda(), inp['accc'].cuda(), inp['oric'].cuda()
    ft = torch.tensor([0., 0., 4.], device='cuda')
    net.set_gemm_mode(2)
    for _ in range(2):
        p, t = net.forward_offline(j, a, o, first_tran=ft)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        p, t = net.forward_offline(j, a, o, first_tran=ft)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print('RC_PH_PAIR=%s B=%d T=%d %s: %.2f ms per pass, %.1f us per frame, %.0f frames/s' % (
        os.environ.get('RC_PH_PAIR', '0'), B, T, conf, ms, ms / T * 1e3, B * T / ms * 1e3), flush=True)
    res = {'pose': p.cpu(), 'tran': t.cpu()}
    if os.environ.get('RC_AB_SIMT'):
        Ts = min(T, 10)
        net.set_gemm_mode(0)
        p0, t0 = net.forward_offline(j[:, :Ts].contiguous(), a[:, :Ts].contiguous(), o[:, :Ts].contiguous(), first_tran=ft)
        res['pose_simt'], res['tran_simt'] = p0.cpu(), t0.cpu()
    torch.save(res, out)


def angle(p, q):
    """Geodesic angle (rad) per joint, float64 atan2 form (acos of a float32 trace cannot resolve 1e-5 rad)."""
    a, b = p.double().reshape(-1, 3, 3), q.double().reshape(-1, 3, 3)
    d = a.transpose(1, 2) @ b
    s = 0.5 * torch.stack((d[:, 2, 1] - d[:, 1, 2], d[:, 0, 2] - d[:, 2, 0], d[:, 1, 0] - d[:, 0, 1]), 1).norm(dim=1)
    c = 0.5 * (d[:, 0, 0] + d[:, 1, 1] + d[:, 2, 2] - 1)
    return torch.atan2(s, c).view(p.shape[0], p.shape[1], 24)


if __name__ == '__main__':
    if len(sys.argv) > 1 and sys.argv[1] == '--child':
        child(int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5])
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    variants = (sys.argv[3] if len(sys.argv) > 3 else '0,2').split(',')
    conf = sys.argv[4] if len(sys.argv) > 4 else 'mixed'
    os.makedirs('gpurun_out', exist_ok=True)
    outs = {}
    for rep in range(2):
        for v in variants:
            out = '/tmp/pair_ab_%s.pt' % v
            env = dict(os.environ, RC_PH_PAIR=v)
            if rep == 0:
                env['RC_AB_SIMT'] = '1'
            r = subprocess.run([sys.executable, os.path.abspath(__file__), '--child', str(B), str(T), conf, out], env=env, timeout=900)
            if r.returncode != 0:
                print('variant %s FAILED rc=%d' % (v, r.returncode), flush=True)
                continue
            if rep == 0:
                outs[v] = torch.load(out)
    base = variants[0]
    for v, r in outs.items():
        Ts = r['pose_simt'].shape[1]
        an = angle(r['pose'][:, :Ts], r['pose_simt'])
        print('variant %s vs SIMT (%d frames): pose max %.2e rad, 99.9 %% %.2e, tran max %.2e m | finite %s' % (
            v, Ts, an.max().item(), an.flatten().kthvalue(int(an.numel() * 0.999)).values.item(),
            (r['tran'][:, :Ts] - r['tran_simt']).abs().max().item(), bool(torch.isfinite(r['pose']).all())), flush=True)
        if v != base and base in outs:
            an = angle(r['pose'], outs[base]['pose'])
            per_t = an.amax(dim=(0, 2))
            print('variant %s vs variant %s (all %d frames): pose max %.2e rad (frame %d), tran max %.2e m; max by frame decile: %s' % (
                v, base, r['pose'].shape[1], an.max().item(), int(per_t.argmax()), (r['tran'] - outs[base]['tran']).abs().max().item(),
                ' '.join('%.1e' % per_t[i * len(per_t) // 10:(i + 1) * len(per_t) // 10].max().item() for i in range(10))), flush=True)

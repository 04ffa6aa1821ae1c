"""GPU experiment (not a pytest): accuracy of one fused LSTM layer, fp32 SIMT vs tcgen05 split-fp16, against float64."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from robustcap_b200 import _lib, synthetic, Net, ParametricModel

def main():
    _lib.build()
    lib = _lib.load()
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    sd = synthetic.make_state_dict(0, 'default')
    net = Net(ParametricModel(assets['smpl_file']))
    net.load_state_dict(sd)
    net._ensure_native()
    B = 512
    st = net._state(B)
    g = torch.Generator().manual_seed(0)
    for ni, name in ((0, 'rnn2'), (3, 'rnn6'), (2, 'rnn4')):
        H = sd[name + '.rnn.weight_hh_l0'].shape[1]
        x = (torch.randn(B, H, generator=g) * 0.5).clamp(min=0).cuda()
        h = (torch.rand(B, H, generator=g) * 2 - 1).mul(0.6).cuda()
        c0 = (torch.randn(B, H, generator=g) * 0.5).cuda()
        wih, whh = sd[name + '.rnn.weight_ih_l0'].cuda().double(), sd[name + '.rnn.weight_hh_l0'].cuda().double()
        b = (sd[name + '.rnn.bias_ih_l0'] + sd[name + '.rnn.bias_hh_l0']).cuda().double()
        gates = x.double() @ wih.t() + h.double() @ whh.t() + b
        i, f, gg, o = gates.chunk(4, dim=1)
        cref = torch.sigmoid(f) * c0.double() + torch.sigmoid(i) * torch.tanh(gg)
        href = torch.sigmoid(o) * torch.tanh(cref)
        for mode, label in ((0, 'simt'), (1, 'tc')):
            c = c0.clone()
            hout = torch.empty(B, H, device='cuda')
            _lib.check(lib.rc_state_debug_lstm(st, ni, 0, mode, x.data_ptr(), h.data_ptr(), c.data_ptr(), hout.data_ptr(), _lib.stream()))
            torch.cuda.synchronize()
            eh = (hout.double() - href).abs()
            ec = (c.double() - cref).abs()
            print('%s H=%4d %-4s  h: max %.2e rms %.2e   c: max %.2e rms %.2e   mean signed c err %.2e'
                  % (name, H, label, eh.max(), eh.pow(2).mean().sqrt(), ec.max(), ec.pow(2).mean().sqrt(), (c.double() - cref).mean()))

if __name__ == '__main__':
    main()

r"""``Net`` — the RobustCap fusion network (reference ``net/sig_mp.py:23-274``) on the B200 library.

Same public surface as the reference class: ``Net()``, ``forward_online(j2dc, accc, oric, first_tran=None,
first_frame=False) -> (pose[24,3,3], tran[3])`` on the CPU, ``reset_states()``, ``load_state_dict`` with the
reference key set, and the class-level knobs (``gravityc``, ``conf_range``, ``use_flat_floor``, ``live`` ...).
New: ``forward_offline`` — defined as ``reset_states()`` followed by ``forward_online`` over the frames
— for one sequence ``[T, ...]`` or a batch ``[B, T, ...]`` of independent sequences.

Python only holds tensors and handles; every arithmetic step of the frame runs in CUDA kernels behind the C ABI
(``include/robustcap_b200.h``).  There is no CPU fallback.
"""
import ctypes

import torch

from . import _lib
from .model import ParametricModel
from .rnn import RNN, RNNWithInit

__all__ = ['Net', 'get_bbox_scale', 'sync_mp3d']


class Net(torch.nn.Module):
    # class-level knobs, same names and defaults as net/sig_mp.py:27-45
    hidden_size = 512
    conf_range = (0.7, 0.8)
    contact_threshold = 0.7
    smooth = 1
    use_flat_floor = True
    use_reproj_opt = False           # reference default; the closed-form re-projection tweak (:244-261) is dead code
    use_vision_updater = True
    use_imu_updater = True
    name = 'sig_mp'
    gravityc = torch.tensor([-0.0029, 0.9980, -0.0273])
    imu_num = 6
    height_threhold = 0.15
    distrance_threshold = 10
    tran_filter_num = 0.05
    live = False
    update_vision_freq = 30

    smpl_file = 'models/SMPL_male.pkl'   # config.paths.smpl_file
    body_model = None                    # shared ParametricModel (the reference keeps a module-level global)

    def __init__(self, body_model: ParametricModel = None):
        super().__init__()
        n = Net.imu_num
        self.rnn2 = RNNWithInit(input_size=n * 3 + n * 9, output_size=23 * 3, hidden_size=Net.hidden_size, num_rnn_layer=2, dropout=0.4)
        self.rnn3 = RNN(input_size=n * 3 + n * 9 + 23 * 3, output_size=3, hidden_size=Net.hidden_size, num_rnn_layer=2, dropout=0.4)
        self.rnn4 = RNN(input_size=n * 3 + n * 9 + 33 * 3, output_size=23 * 3, hidden_size=1024 + 256, num_rnn_layer=2, dropout=0.4)
        self.rnn6 = RNN(input_size=n * 3 + n * 9 + 33 * 3 + 23 * 3, output_size=3, hidden_size=1024, num_rnn_layer=2, dropout=0.4)
        self.rnn7 = RNN(input_size=n * 3 + n * 9 + 23 * 3, output_size=24 * 6, hidden_size=512, num_rnn_layer=2, dropout=0.1)
        self.rnn8 = RNN(input_size=n * 3 + n * 9 + 23 * 3, output_size=2, hidden_size=Net.hidden_size, num_rnn_layer=2, dropout=0.4)
        if self.live:                                     # sig_mp.py:91-93
            self.conf_range = (0.85, 0.9)
            self.tran_filter_num = 0.01
        if body_model is None:
            if Net.body_model is None:
                Net.body_model = ParametricModel(Net.smpl_file)
            body_model = Net.body_model
        self._body = body_model
        self._net = None            # rc_net handle
        self._states = {}           # batch size -> rc_state handle
        self._cfg_pushed = None
        self._cfg_raw = None
        self._grav_pushed = {}
        self._grav_seen = {}
        self._grav_keep = {}
        self._dirty = True
        self._use_graph = True
        self._fp = None
        # Any load into a sub-module (net.rnn2.load_state_dict(...), RNN(load_weight_file=...) as the reference's training / merge
        # scripts do, sig_mp.py:850-857) must reach the packed device copy too, not only Net.load_state_dict.
        mark = lambda module, incompatible: setattr(self, '_dirty', True)
        for m in self.modules():
            m.register_load_state_dict_post_hook(mark)

    # ---- native plumbing -------------------------------------------------------------------------------------------
    def _config(self):
        c = _lib.NetConfig()
        c.conf_lo, c.conf_hi = float(self.conf_range[0]), float(self.conf_range[1])
        c.tran_filter_num = float(self.tran_filter_num)
        c.contact_threshold = float(self.contact_threshold)
        c.height_threshold = float(self.height_threhold)
        c.distance_threshold = float(self.distrance_threshold)
        c.use_flat_floor = int(bool(self.use_flat_floor))
        c.live = int(bool(self.live))
        c.update_vision_freq = int(self.update_vision_freq)
        return c

    def _release(self):
        lib = _lib.load()
        for h in self._states.values():
            lib.rc_state_destroy(h)
        self._states = {}
        self._grav_pushed = {}
        self._grav_seen = {}
        self._grav_keep = {}
        if self._net is not None:
            lib.rc_net_destroy(self._net)
            self._net = None

    def __del__(self):
        try:
            self._release()
        except Exception:
            pass

    def _fingerprint(self):
        # storage address + in-place version counter of every parameter: catches `p.data = ...`, `p.copy_()`, optimiser steps
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def mark_dirty(self):
        """Force a re-upload of the weights on the next call (after editing parameters in a way torch cannot see)."""
        self._dirty = True

    def _ensure_native(self, check_params=False):
        lib = _lib.load()
        _lib.require_cuda()
        if check_params and not self._dirty and self._fingerprint() != self._fp:
            self._dirty = True
        if self._net is not None and not self._dirty:
            # the class-level knobs may be changed at any time (evaluate.py:254, 337, 392): compare the raw attribute values (cheap, this
            # runs on every streaming frame) and push a new config only when one moved
            raw = (self.conf_range, self.tran_filter_num, self.contact_threshold, self.height_threhold, self.distrance_threshold,
                   self.use_flat_floor, self.live, self.update_vision_freq)
            if raw != self._cfg_raw:
                cfg = self._config()
                key = bytes(cfg)
                if key != self._cfg_pushed:
                    _lib.check(lib.rc_net_set_config(self._net, ctypes.byref(cfg)))
                    self._cfg_pushed = key
                self._cfg_raw = raw
            return
        self._release()
        cfg = self._config()
        h = _lib.vp()
        _lib.check(lib.rc_net_create(ctypes.byref(h), self._body._native(), ctypes.byref(cfg)))
        self._net = h
        self._cfg_pushed = bytes(cfg)
        for key, val in self.state_dict().items():
            t = val.detach().to('cpu', torch.float32).contiguous()
            _lib.check(lib.rc_net_set_tensor(h, key.encode(), _lib.hptr(t), t.numel()))
        _lib.check(lib.rc_net_finalize(h))
        self._dirty = False
        self._fp = self._fingerprint()

    def _state(self, B):
        lib = _lib.load()
        if B not in self._states:
            h = _lib.vp()
            _lib.check(lib.rc_state_create(ctypes.byref(h), self._net, B))
            self._states[B] = h
        gc = self.gravityc                                  # callers overwrite it per sequence (evaluate.py:73, 180, 285, 337)
        seen = (id(gc), gc._version)
        if self._grav_seen.get(B) != seen:                  # same tensor object, not modified in place: nothing to do (per-frame path)
            g = gc.detach().to('cpu', torch.float32).reshape(3).contiguous()
            key = tuple(g.tolist())
            if self._grav_pushed.get(B) != key:
                _lib.check(lib.rc_state_set_gravity(self._states[B], _lib.hptr(g), _lib.stream()))
                self._grav_pushed[B] = key
            self._grav_seen[B] = seen
            self._grav_keep[B] = gc                         # keeps id() unique while cached
        return self._states[B]

    def load_state_dict(self, *args, **kwargs):
        out = super().load_state_dict(*args, **kwargs)
        self._dirty = True
        return out

    def set_gemm_mode(self, mode):
        """Batched (B > 8) back end: 3 = persistent SEQUENCE kernel (frames 1..T-1 of ``forward_offline`` in one launch: GEMM tiles
        and the per-frame row logic in one dependency queue), 2 = persistent grouped tcgen05 kernel, one launch per phase of the
        frame, 1 = tcgen05 one launch per layer, 0 = fp32 SIMT tiles.  All tcgen05 paths use split-fp16 operands (fp32-accurate)."""
        self._ensure_native()
        _lib.check(_lib.load().rc_net_set_gemm_mode(self._net, int(mode)))
        self._gemm_mode = int(mode)

    def set_seq_options(self, auto_max_streams=-1, warm_frames=0):
        """Policy of the persistent sequence kernel: batches of at most ``auto_max_streams`` streams use it in the default mode
        (0 = never), after ``warm_frames`` frames through the multi-launch path.  Negative / zero keep the current value."""
        self._ensure_native()
        _lib.check(_lib.load().rc_net_set_seq_options(self._net, int(auto_max_streams), int(warm_frames)))

    def weight_bytes(self):
        self._ensure_native()
        return int(_lib.load().rc_net_weight_bytes(self._net))

    # ---- reference API -------------------------------------------------------------------------------------------
    def reset_states(self):
        r"""Reset the hidden states and the tracker variables of the online stream (the B = 1 state ``forward_online`` uses;
        ``forward_offline`` resets its own batch state at the start of every call). sig_mp.py:95-104.
        Also the point where in-place parameter edits since the last call are detected (per-frame calls only see loads)."""
        if self._net is not None and not self._dirty and self._fingerprint() != self._fp:
            self._dirty = True
        if self._net is not None and not self._dirty and 1 in self._states:
            _lib.check(_lib.load().rc_state_reset(self._states[1], _lib.stream()))

    @staticmethod
    def cat(*x):
        return [torch.cat(_, dim=1) for _ in zip(*x)]

    @torch.no_grad()
    def forward_online(self, j2dc, accc, oric, first_tran=None, first_frame=False):
        r"""One frame. j2dc [33,3] (x, y on the z=1 plane, confidence), accc [6,3], oric [6,3,3] ->
        pose [24,3,3] (local, root = pelvis IMU) and tran [3], both on the CPU. sig_mp.py:113-274."""
        lib = _lib.load()
        self._ensure_native()
        _lib.require_cuda()
        st = self._state(1)
        on_dev = j2dc.is_cuda
        assert accc.is_cuda == on_dev and oric.is_cuda == on_dev, 'inputs must live on one device'
        prep = lambda x, n: x.detach().reshape(n).to(torch.float32).contiguous()
        dj, da, do = prep(j2dc, 99), prep(accc, 18), prep(oric, 54)
        dft = None
        if first_tran is not None:
            dft = prep(first_tran, 3)
            dft = dft.to(dj.device)
        pose = torch.empty(24, 3, 3)
        tran = torch.empty(3)
        _lib.check(lib.rc_forward_online(st, dj.data_ptr(), da.data_ptr(), do.data_ptr(), None if dft is None else dft.data_ptr(),
                                         int(bool(first_frame)), int(on_dev), pose.data_ptr(), tran.data_ptr(), _lib.stream()))
        return pose, tran

    @torch.no_grad()
    def forward_offline(self, j2dc, accc, oric, first_tran=None, first_frame=False, lengths=None, use_graph=None,
                        first_tran_mask=None, out=None, gravity=None, branch_log=None):
        r"""``reset_states()`` then ``forward_online`` over every frame (evaluate.py:75-85, 93), natively batched.

        j2dc [T,33,3] or [B,T,33,3]; accc [..,T,6,3]; oric [..,T,6,3,3].  ``first_tran`` ([3] or [B,3]) and
        ``first_frame`` (bool or bool[B]) apply to frame 0; ``lengths`` (int[B]) marks ragged batches (outputs beyond
        a sequence's length are zero); ``gravity`` ([B,3], device inputs only) gives every sequence its own camera-frame gravity
        (the reference sets ``net.gravityc`` per sequence, evaluate.py:73) instead of the class attribute.
        Returns pose [..,T,24,3,3], tran [..,T,3] on the device of ``j2dc``.
        CPU inputs go through the end-to-end host entry point (H2D copy, kernels, D2H copy)."""
        lib = _lib.load()
        self._ensure_native(check_params=True)
        dev = _lib.require_cuda()
        single = j2dc.dim() == 3
        if single:
            j2dc, accc, oric = j2dc.unsqueeze(0), accc.unsqueeze(0), oric.unsqueeze(0)
        B, T = j2dc.shape[0], j2dc.shape[1]
        st = self._state(B)
        use_graph = self._use_graph if use_graph is None else use_graph
        on_cpu = not j2dc.is_cuda
        where = dict(device='cpu' if on_cpu else dev, dtype=torch.float32)
        j = j2dc.detach().reshape(B, T, 99).to(**where).contiguous()
        a = accc.detach().reshape(B, T, 18).to(**where).contiguous()
        o = oric.detach().reshape(B, T, 54).to(**where).contiguous()
        ft = None
        if first_tran is not None:
            ft = first_tran.detach().reshape(-1, 3).to(**where).expand(B, 3).contiguous()
        ff = torch.as_tensor(first_frame).reshape(-1).to(torch.int32).expand(B) if not isinstance(first_frame, bool) \
            else torch.full((B,), int(first_frame), dtype=torch.int32)
        ftm = torch.full((B,), 2 if ft is not None else 0, dtype=torch.int32)
        if first_tran_mask is not None and ft is not None:     # per-sequence: which rows really pass first_tran
            ftm = torch.as_tensor(first_tran_mask).reshape(B).to(torch.int32).cpu() * 2
        flags = (ff.cpu() | ftm).to(torch.int32).contiguous()
        any_ff = int(bool((flags & 1).any()))
        use_flags = bool(flags.any())
        ln = None if lengths is None else torch.as_tensor(lengths).to('cpu', torch.int32).reshape(B).contiguous()
        if branch_log is not None:     # debug / parity aid: int32 [B, T] on the device, receives the RcBranch bits per frame
            assert not on_cpu and branch_log.is_cuda and branch_log.dtype == torch.int32 and branch_log.numel() == B * T
        _lib.check(lib.rc_state_set_branch_log(st, None if branch_log is None else branch_log.data_ptr()))
        if out is not None:            # caller-provided result buffers (e.g. pinned host memory), contiguous float32
            pose, tran = out[0].view(B, T, 24, 3, 3), out[1].view(B, T, 3)
            assert pose.is_cuda == (not on_cpu) and pose.dtype == torch.float32
        else:
            pose = torch.zeros(B, T, 24, 3, 3, **where)
            tran = torch.zeros(B, T, 3, **where)
        if on_cpu:
            assert gravity is None, 'per-sequence gravity needs device inputs'
            _lib.check(lib.rc_forward_sequence_host(st, T, _lib.hptr(j), _lib.hptr(a), _lib.hptr(o),
                                                    _lib.hptr(ln), _lib.hptr(ft),
                                                    _lib.hptr(flags) if use_flags else None, _lib.hptr(pose), _lib.hptr(tran),
                                                    int(use_graph), _lib.stream()))
        else:
            dflags = flags.to(dev) if use_flags else None
            dln = None if ln is None else ln.to(dev)
            dgr = None if gravity is None else gravity.detach().reshape(B, 3).to(dev, torch.float32).contiguous()
            _lib.check(lib.rc_forward_sequence(st, T, _lib.dptr(j), _lib.dptr(a), _lib.dptr(o), _lib.dptr(dln), _lib.dptr(dgr),
                                               _lib.dptr(ft), _lib.dptr(dflags), any_ff, _lib.dptr(pose), _lib.dptr(tran),
                                               int(use_graph), _lib.stream()))
            self._keepalive = (j, a, o, ft, dflags, dln, dgr)   # kernels are still in flight
        if single:
            return pose[0], tran[0]
        return pose, tran

    def debug_outputs(self, B=1):
        """Sub-net outputs of the last frame (tests): dict {2: [B,69], 3: [B,3], 4: [B,69], 6: [B,3], 7: [B,144], 8: [B,2]}."""
        lib = _lib.load()
        out = {}
        for k, w in ((2, 69), (3, 3), (4, 69), (6, 3), (7, 144), (8, 2)):
            t = torch.empty(B, w)
            _lib.check(lib.rc_state_debug_output(self._states[B], k, _lib.hptr(t), _lib.stream()))
            out[k] = t
        return out


def get_bbox_scale(uv):
    r"""max(bbox width, bbox height) over the key-point axis. sig_mp.py:277-284 (tensor plumbing)."""
    u_max, u_min = uv[..., 0].max(dim=-1).values, uv[..., 0].min(dim=-1).values
    v_max, v_min = uv[..., 1].max(dim=-1).values, uv[..., 1].min(dim=-1).values
    return torch.max(u_max - u_min, v_max - v_min)


def sync_mp3d(vert, joint):
    r"""33 MediaPipe points from SMPL vertices and joints. sig_mp.py:287-299 (index gather, bit-exact tables)."""
    from .constants import MP_MASK
    syn_3d = vert[torch.tensor(MP_MASK, device=vert.device)]
    syn_3d[11:17] = joint[16:22].clone()
    syn_3d[23:25] = joint[1:3].clone()
    syn_3d[25:27] = joint[4:6].clone()
    syn_3d[27:29] = joint[7:9].clone()
    return syn_3d

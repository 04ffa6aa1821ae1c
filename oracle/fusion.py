"""Oracle (test infrastructure): the six-LSTM fusion network, one frame at a time.  See ``oracle/__init__.py``.

Restates ``Net.forward_online`` / ``Net.reset_states`` (``net/sig_mp.py:23-274``) and the ``RNN`` / ``RNNWithInit``
building blocks it drives (``articulate/utils/torch/rnn.py:92-133, 174-219``) as explicit math on a flat weight
dictionary (the reference's ``state_dict`` key set).  ``forward_offline`` does not exist in the reference; it is
DEFINED as "reset, then forward_online frame by frame"  and that is what ``run`` does.
"""
import torch

from . import rotations as rot
from . import kinematics as kin

VEL_SCALE = 3            # config.py:97
MP_MASK = [332, 2809, 2800, 455, 6260, 3634, 3621, 583, 4071, 45, 3557, 1873, 4123, 1652, 5177, 2235, 5670, 2673,
           6133, 2319, 5782, 2746, 6191, 3138, 6528, 1176, 4662, 3381, 6727, 3387, 6787, 3226, 6624]  # config.py:99


class LstmStack:
    """``linear1 -> relu -> 2-layer LSTM (one time step) -> linear2`` (rnn.py:111-114 driven as in sig_mp.py:126-129).

    ``impl='manual'``: the cell written out (gate order i, f, g, o).
    ``impl='aten'``:   ``torch.nn.LSTM`` exactly as the reference calls it (float32 only) — used as the timed CPU
    baseline because it is the back end the reference really runs on.
    """

    def __init__(self, sd, prefix, dtype, impl='manual'):
        g = lambda k: sd[prefix + '.' + k].to(dtype)
        self.w1, self.b1 = g('linear1.weight'), g('linear1.bias')
        self.w2, self.b2 = g('linear2.weight'), g('linear2.bias')
        self.layers = [(g('rnn.weight_ih_l%d' % l), g('rnn.weight_hh_l%d' % l), g('rnn.bias_ih_l%d' % l),
                        g('rnn.bias_hh_l%d' % l)) for l in range(2)]
        self.H = self.w1.shape[0]
        self.dtype = dtype
        self.impl = impl
        if impl == 'aten':
            assert dtype == torch.float32
            self.mod = torch.nn.LSTM(self.H, self.H, 2)
            with torch.no_grad():
                for l in range(2):
                    for name, w in zip(('weight_ih', 'weight_hh', 'bias_ih', 'bias_hh'), self.layers[l]):
                        getattr(self.mod, '%s_l%d' % (name, l)).copy_(w)
            self.mod.eval()
        self.state = None   # [(h, c)] per layer, each [H]

    def zero_state(self):
        z = lambda: torch.zeros(self.H, dtype=self.dtype)
        return [(z(), z()), (z(), z())]

    def step(self, x):
        x = torch.relu(torch.addmv(self.b1, self.w1, x))
        if self.state is None:
            self.state = self.zero_state()
        if self.impl == 'aten':
            h = torch.stack([s[0] for s in self.state]).unsqueeze(1)
            c = torch.stack([s[1] for s in self.state]).unsqueeze(1)
            y, (h, c) = self.mod(x.view(1, 1, -1), (h, c))
            self.state = [(h[l, 0], c[l, 0]) for l in range(2)]
            x = y.view(-1)
        else:
            new = []
            for (wih, whh, bih, bhh), (h, c) in zip(self.layers, self.state):
                g = torch.addmv(bih, wih, x) + torch.addmv(bhh, whh, h)
                i, f, gg, o = g.chunk(4)
                c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
                h = torch.sigmoid(o) * torch.tanh(c)
                new.append((h, c))
                x = h
            self.state = new
        return torch.addmv(self.b2, self.w2, x)


class FusionOracle:
    """One stream of ``Net`` (net/sig_mp.py:23-274), non-training path."""

    def __init__(self, state_dict, body, dtype=torch.float32, live=False, lstm_impl='manual'):
        self.dtype = dtype
        self.body = body
        self.nets = {k: LstmStack(state_dict, 'rnn%d' % k, dtype, lstm_impl) for k in (2, 3, 4, 6, 7, 8)}
        self.init_net = [(state_dict['rnn2.init_net.%d.weight' % k].to(dtype),
                          state_dict['rnn2.init_net.%d.bias' % k].to(dtype)) for k in (0, 2, 4)]
        j = body.zero_pose()[0]
        self.bone = kin.joint_to_bone(j.unsqueeze(0), body.parent)[0]           # sig_mp.py:83-84
        self.parent = body.parent
        # class-level knobs (sig_mp.py:27-45, 91-93)
        self.live = live
        self.conf_range = (0.85, 0.9) if live else (0.7, 0.8)
        self.tran_filter_num = 0.01 if live else 0.05
        self.contact_threshold = 0.7
        self.height_threshold = 0.15
        self.distance_threshold = 10
        self.use_flat_floor = True
        self.update_vision_freq = 30
        self.gravity = torch.tensor([-0.0029, 0.9980, -0.0273], dtype=dtype)
        self.update_vision_count = 0
        self.j_temp = None
        self.reset()

    def reset(self):
        """sig_mp.py:95-104."""
        for n in self.nets.values():
            n.state = None
        self.last_pfoot = None
        self.last_tran = None
        self.floor_y = []
        self.first_reach = True

    # ---- helpers -------------------------------------------------------------------------------------------
    @staticmethod
    def _cat(*xs):
        return torch.cat([x.reshape(-1) for x in xs])

    def _normalise_keypoints(self, kp):
        """sig_mp.py:150-152 / 268-270: divide x, y by the bbox scale, make everything but keypoint 23 relative to it."""
        kp = kp.clone()
        kp[:, :2] = kp[:, :2] / kin.bbox_scale(kp)
        ref = kp[23:24, :2].clone()
        kp[24:, :2] = kp[24:, :2] - ref
        kp[:23, :2] = kp[:23, :2] - ref
        return kp

    def _foot_fk(self, poseg):
        """sig_mp.py:131-135: joint positions from global rotations and rest bone vectors."""
        pb = [torch.zeros(3, dtype=self.dtype)]
        for i in range(1, 24):
            pb.append(poseg[self.parent[i]] @ self.bone[i])
        return kin.bone_to_joint(torch.stack(pb).unsqueeze(0), self.parent)[0]

    # ---- one frame ---------------------------------------------------------------------------------------------
    @torch.no_grad()
    def step(self, j2dc, accc, oric, first_tran=None, first_frame=False, trace=None, branches=None):
        """``branches`` (optional list): receives one int per frame describing the data-dependent decisions of the translation /
        contact / floor logic, same bit layout as the CUDA path's debug log (``rc_rows.h`` RC_BR_*): 1 contact branch, 2 contact
        argmax, 4 snap to pc, 8 lerp to pc, 16 floor sample stored, 32 floor snap (other foot), 64 floor snap (near foot),
        128 rnn2 re-seeded by init_net."""
        dt = self.dtype
        br = 0
        j2dc, accc, oric = j2dc.to(dt).reshape(33, 3), accc.to(dt).reshape(6, 3), oric.to(dt).reshape(6, 3, 3)
        lo, hi = self.conf_range
        N = self.nets

        def run(k, x):
            y = N[k].step(x)
            if trace is not None:
                trace.append((k, y.clone()))
            return y

        c = j2dc[:, 2].mean().item()                                           # :138
        Rcr = oric[5]                                                          # :139
        accr = accc @ Rcr                                                      # :142
        orir = Rcr.t() @ oric                                                  # :143
        j3dr_i = run(2, self._cat(accr, orir))                                 # :144
        vr = run(3, self._cat(accr, orir, j3dr_i))                             # :145

        pc = None
        if c > lo or first_frame:                                              # :149-156
            j3dc = run(4, self._cat(accc, oric, self._normalise_keypoints(j2dc)))
            j3dr_v = j3dc.view(23, 3) @ Rcr
            if first_frame:
                pc = run(6, self._cat(accc, oric, j2dc, j3dc))
        if c >= hi:                                                            # :159-167
            j3dr = j3dr_v.reshape(-1)
            pc = run(6, self._cat(accc, oric, j2dc, j3dc))
        elif c > lo:
            k = (c - lo) / (hi - lo)
            j3dr = rot.lerp(j3dr_i.view(-1), j3dr_v.reshape(-1), k)
            pc = run(6, self._cat(accc, oric, j2dc, j3dc))
        else:
            j3dr = j3dr_i

        x7 = self._cat(accr, orir, j3dr)
        poseg6d = run(7, x7)                                                   # :169
        contact = torch.sigmoid(run(8, x7))                                    # :170

        poseg = rot.r6d_to_matrix(poseg6d).view(24, 3, 3)                      # :173
        pose = kin.ik_R(poseg.unsqueeze(0), self.parent)[0]                    # :174
        pose[0] = Rcr                                                          # :175

        if c >= hi and self.first_reach:                                       # :178-183
            self.first_reach = False
            br |= 128
            z = j3dr.reshape(-1)
            for n, (w, b) in enumerate(self.init_net):
                z = torch.addmv(b, w, z)
                if n < 2:
                    z = torch.relu(z)
            H = 512
            N[2].state = [(z[0:H].clone(), z[2 * H:3 * H].clone()), (z[H:2 * H].clone(), z[3 * H:4 * H].clone())]

        pfoot = self._foot_fk(poseg)[10:12] @ Rcr.t()                          # :186
        cmax = contact.max()
        thr = torch.tensor(self.contact_threshold, dtype=dt)                   # python scalar -> tensor dtype
        if bool(cmax < thr) or self.last_pfoot is None:                        # :187-190
            v = (Rcr @ vr.view(3, 1)).view(3) * VEL_SCALE / 60
        else:
            v = (self.last_pfoot - pfoot)[int(contact.argmax())]
            br |= 1 | (2 * int(contact.argmax()))
        tran = v if self.last_tran is None else self.last_tran + v             # :191-194

        if hi <= c:                                                            # :196-203
            k = min((c - lo) / (hi - lo), 1)
            if bool((pc - tran).norm() > self.distance_threshold) or self.tran_filter_num > 1:
                tran = pc
                br |= 4
            else:
                tran = rot.lerp(tran, pc, self.tran_filter_num * k)
                br |= 8
        tran = tran.reshape(3)

        g = self.gravity.to(dt)
        hthr = torch.tensor(self.height_threshold, dtype=dt)
        if (len(self.floor_y) < 11 and not first_frame and first_tran is None and bool(cmax > thr)
                and self.use_flat_floor and c >= hi):                          # :208-214
            p0 = torch.dot(pfoot[0] + tran, g) * g
            p1 = torch.dot(pfoot[1] + tran, g) * g
            self.floor_y.append(p1 if bool(p0.norm() < p1.norm()) else p0)
            br |= 16
        if self.use_flat_floor and len(self.floor_y) > 10 and bool(cmax > thr):  # :215-221
            p0 = torch.dot(pfoot[0] + tran, g) * g
            p1 = torch.dot(pfoot[1] + tran, g) * g
            mean = sum(self.floor_y[-6:]) / 6
            if bool(p0.norm() < p1.norm()) and bool((mean - p1).norm() < hthr):
                tran = tran + (mean - p1)
                br |= 32
            elif bool((mean - p0).norm() < hthr):
                tran = tran + (mean - p0)
                br |= 64
        if first_tran is not None:                                             # :222-225
            tran = first_tran.to(dt).reshape(3)
        elif first_frame:
            tran = pc.reshape(3)
        self.last_pfoot = pfoot                                                # :227

        # :228-242  mesh FK -> 33 synthetic MediaPipe points (every frame unless live)
        do_fk = (not self.live) or self.update_vision_count == 0
        if do_fk:
            _, joint, vert = self.body.forward_kinematics(pose.view(1, 24, 3, 3), tran=tran.view(1, 3), calc_mesh=True)
            j = kin.mediapipe_points(vert, joint, MP_MASK)[0]
            self._joint = joint[0]
            if self.live:
                self.j_temp = j
                self.update_vision_count = self.update_vision_freq
        else:
            j = self.j_temp
            self.update_vision_count -= 1

        # :263-271  vision updater: keep rnn6 / rnn4 state warm with re-projected synthetic keypoints
        if c <= lo and (self.update_vision_count == self.update_vision_freq or not self.live):
            kp = j / j[:, 2:]
            j3d = self._joint[1:] - self._joint[:1]
            run(6, self._cat(accc, oric, kp, j3d))
            run(4, self._cat(accc, oric, self._normalise_keypoints(kp)))

        self.last_tran = tran                                                  # :273
        if branches is not None:
            branches.append(br)
        return pose.view(24, 3, 3).clone(), tran.view(3).clone()

    def run(self, j2dc, accc, oric, first_tran=None, first_frame=False, gravity=None, reset=True, trace=None, branches=None):
        """forward_offline := reset + per-frame forward_online (evaluate.py:75-85, 93)."""
        if reset:
            self.reset()
        if gravity is not None:
            self.gravity = gravity.to(self.dtype)
        poses, trans = [], []
        for t in range(j2dc.shape[0]):
            kw = {'first_tran': first_tran, 'first_frame': first_frame} if t == 0 else {}
            p, tr = self.step(j2dc[t], accc[t], oric[t], trace=trace, branches=branches, **kw)
            poses.append(p)
            trans.append(tr)
        return torch.stack(poses), torch.stack(trans)

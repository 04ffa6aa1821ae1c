#!/usr/bin/env python
"""Benchmark of the RobustCap fusion + kinematics hot path on B200 (contract: see the task statement / DESIGN.md).

  python bench.py --gpus N --steps K --warmup W          one JSON line (rank 0)
  python bench.py --impl reference ...                   the reference's CPU implementation of the same path

Workload (BASELINE.json configs[2]): offline evaluation of `--seqs` (1024) independent sequences x `--frames` (300)
frames, synthetic 6-IMU + 33x3 key points with per-frame confidence U(0.6, 1.0) (all three branches of
net/sig_mp.py:149-167), random-init weights.  One "step" = one forward_offline pass over the batch.  Multi-GPU
(`--scaling strong`, default, the configuration BASELINE names): the 1024 sequences are sharded over the N ranks
(contiguous blocks, `robustcap_b200.distributed`), no data-path collective, one NCCL gather of [pose | tran] to rank 0
inside the timed region.  `--scaling weak` gives every rank `--seqs` sequences instead.

  value  frames/s, whole job, inputs already resident in HBM (device-timed with CUDA events, max over ranks)
  e2e    the same pass through the host-buffer C-ABI entry point (pinned host inputs -> H2D -> kernels -> D2H)
  roofline       dominant kernel = the persistent grouped tcgen05 GEMM (3 launches per frame run the whole LSTM stack), CUDA-event timed
  cpu_baseline   the oracle port of the reference loop on the host cores, bounded sample
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

import torch  # noqa: E402

METRIC = 'frames/sec (24-joint SMPL, 60 fps streams)'
FLOP_PER_FRAME = 2 * 60689920           # SURVEY.md §6 / BASELINE.md §3
WEIGHT_BYTES = 243.06e6


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--seqs', type=int, default=1024, help='sequences of the job (strong scaling) / per GPU (weak scaling)')
    ap.add_argument('--scaling', default='strong', choices=['strong', 'weak'])
    ap.add_argument('--no-occlusion', action='store_true', help='skip the occlusion-workload line (BASELINE configs[4])')
    ap.add_argument('--no-smplify', action='store_true', help='skip the fusion + FK + SMPLify line (BASELINE configs[3])')
    ap.add_argument('--frames', type=int, default=300)
    ap.add_argument('--conf', default='mixed')
    ap.add_argument('--cpu-frames', type=int, default=240, help='frames of the bounded CPU-baseline sample')
    ap.add_argument('--gemm-mode', type=int, default=2, help='2 = persistent grouped tcgen05 kernel (default), 1 = tcgen05 per layer, 0 = fp32 SIMT')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-stream-latency', action='store_true')
    return ap.parse_args()


def peaks():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return {'hbm_gbs': p['hbm_gbs'], 'tflops_burst': p['bf16_tflops'], 'tflops_sustained': p['bf16_tflops_sustained'],
                'source': 'measured'}
    return {'hbm_gbs': 6650.0, 'tflops_burst': 1590.0, 'tflops_sustained': 1400.0, 'source': 'fallback'}


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix='.csv')
            os.close(fd)
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits',
                                          '-lms', '100'], stdout=open(self.path, 'w'), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), f[5:9]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out = {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(mx), 'reasons': sorted(reasons), 'samples': len(sm)}
        return out


def dominant_traffic(gemm_mode=1):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full capture
    (profiles/r03_phase_traffic.json for the grouped kernel, written from the .ncu-rep by profiles/extract_phase.py); None when absent."""
    names = ('r03_phase_traffic.json', 'r01_phase_traffic.json') if gemm_mode == 2 else ('r01_tc_traffic.json',)   # newest capture first
    for name in names:
        path = os.path.join(REPO, 'profiles', name)
        if os.path.exists(path):
            return json.load(open(path)).get('dram_bytes_per_launch')
    return None


def build_net(seed=0):
    from robustcap_b200 import synthetic, Net, ParametricModel
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    body = ParametricModel(assets['smpl_file'])
    sd = synthetic.make_state_dict(seed, 'contact')     # contact/floor branches reachable (synthetic.make_state_dict)
    net = Net(body)
    net.load_state_dict(sd)
    return net, sd, assets


def make_cpu_reference(sd, assets, impl='aten'):
    """Oracle port of the reference loop with the thread count that maximises its throughput on this host: the work is
    GEMV-sized, and torch's default of one thread per core (128 on the B200 box) is far slower than a few threads."""
    from robustcap_b200 import synthetic
    from oracle.kinematics import BodyOracle
    from oracle.fusion import FusionOracle
    o = FusionOracle(sd, BodyOracle(assets['smpl_file']), lstm_impl=impl)
    ncpu = os.cpu_count() or 1
    inp = synthetic.make_inputs(1, 6, seed=98, conf='mixed')
    best, best_t = None, None
    for n in sorted({1, 2, 4, 8, 16, 32, ncpu}):
        if n > ncpu:
            continue
        torch.set_num_threads(n)
        per = cpu_reference_pass(o, inp, 4, warm=2)
        t = statistics.median(per)
        if best_t is None or t < best_t:
            best, best_t = n, t
    torch.set_num_threads(best)
    return o


def cpu_reference_pass(o, inp, frames, warm=0):
    """One bounded sample of the reference's per-frame loop (evaluate.py:75-85) restated by the oracle."""
    o.gravity = inp['gravity']
    o.reset()
    per = []
    for t in range(frames + warm):
        t0 = time.perf_counter()
        kw = {'first_tran': torch.tensor([0., 0., 4.])} if t == 0 else {}
        o.step(inp['j2dc'][0, t], inp['accc'][0, t], inp['oric'][0, t], **kw)
        if t >= warm:
            per.append(time.perf_counter() - t0)
    return per


def cpu_reference_rate(sd, assets, frames, conf, impl='aten', warm=10):
    from robustcap_b200 import synthetic
    o = make_cpu_reference(sd, assets, impl)
    inp = synthetic.make_inputs(1, frames + warm, seed=99, conf=conf)
    per = cpu_reference_pass(o, inp, frames, warm)
    return {'value': frames / sum(per), 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'kind': 'port',
            'sample': '1 sequence x %d frames (after %d warm-up frames), oracle/fusion.py with torch.nn.LSTM (%s) like the '
                      'reference, conf=%s' % (frames, warm, impl, conf),
            'ms_per_frame_p50': 1e3 * statistics.median(per)}


def workload_config(args, world, B_rank):
    """The `config` object of the JSON line — identical in both arms (the driver compares them)."""
    if args.scaling == 'strong':
        wl = ('offline_eval %d seq x %d frames sharded over %d GPU(s) (BASELINE configs[2]; configs[1] = the same net streaming at B=1, '
              'reported under "streaming")' % (args.seqs, args.frames, world))
    else:
        wl = 'offline_eval %d seq x %d frames per GPU, weak scaling (BASELINE configs[2] per GPU)' % (args.seqs, args.frames)
    return {'workload': wl, 'conf': args.conf, 'weights': 'random-init seed 0 (contact variant)'}


# ---- the reference arm ---------------------------------------------------------------------------------------------------------------
REF_ARCHIVE = os.path.join(REPO, 'baseline', '_ref', 'reference_src.tar.gz')


def load_real_reference():
    """The UNMODIFIED reference (its .py sources, archived by __graft_entry__.build() from /root/reference into the git-ignored
    baseline/_ref/, which travels to the GPU box), imported with stub modules for the viz / training deps and the seeded
    synthetic assets — the recipe of tests/golden/make_golden.py.  Returns its `Net` class, or None when the archive is absent."""
    if not os.path.exists(REF_ARCHIVE):
        return None
    import tarfile
    import types
    import warnings
    from robustcap_b200 import synthetic
    dst = tempfile.mkdtemp(prefix='robustcap_ref_')
    with tarfile.open(REF_ARCHIVE) as tf:
        tf.extractall(dst)
    for name in ('trimesh', 'pyrender', 'smplx', 'wandb'):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules['smplx'].SMPL = object
    thop = types.ModuleType('thop')
    thop.clever_format = lambda *a, **k: ''
    sys.modules['thop'] = thop
    # this arm times the reference's CPU implementation on the host cores: its modules pick `cuda` at import when one is visible
    os.environ['CUDA_VISIBLE_DEVICES'] = ''
    torch.cuda.is_available = lambda: False
    root = synthetic.default_asset_root()
    synthetic.write_assets(root, 0)
    os.chdir(root)                                  # the reference opens models/SMPL_male.pkl relative to the CWD at import
    sys.path.insert(0, os.path.join(dst, 'reference'))
    warnings.filterwarnings('ignore')
    from net.sig_mp import Net as RefNet            # noqa: E402  (the reference's own module)
    return RefNet


def reference_pass(net, inp, frames):
    """evaluate.py:75-85, 93 on one sequence: the reference's own per-frame loop."""
    with torch.no_grad():
        for t in range(frames):
            kw = {'first_tran': torch.tensor([0., 0., 4.])} if t == 0 else {}
            net.forward_online(inp['j2dc'][0, t], inp['accc'][0, t], inp['oric'][0, t], **kw)
        net.reset_states()


def best_threads(fn):
    """Thread count that maximises the throughput of a GEMV-sized CPU loop on this host (the default of one thread per core —
    128+ on the B200 box — is far slower than a few threads)."""
    ncpu = os.cpu_count() or 1
    best, best_t = 1, None
    for n in sorted({1, 2, 4, 8, 16, 32, ncpu}):
        if n > ncpu:
            continue
        torch.set_num_threads(n)
        fn()
        t0 = time.perf_counter()
        fn()
        t = time.perf_counter() - t0
        if best_t is None or t < best_t:
            best, best_t = n, t
    torch.set_num_threads(best)
    return best


def run_reference(args, rank, world):
    if rank != 0:
        return
    from robustcap_b200 import synthetic
    assets = synthetic.write_assets(synthetic.default_asset_root(), 0)
    sd = synthetic.make_state_dict(0, 'contact')
    frames = args.frames
    inp = synthetic.make_inputs(1, frames, seed=1000, conf=args.conf)
    RefNet = None
    try:
        RefNet = load_real_reference()
    except Exception as e:                           # fall back to the port, say why
        print('reference import failed (%s: %s); timing the oracle port instead' % (type(e).__name__, e), file=sys.stderr)
    if RefNet is not None:
        RefNet.gravityc = inp['gravity'].clone()
        net = RefNet()
        net.load_state_dict(sd)
        net.eval()
        probe = {k: (v[:, :8] if v.dim() > 1 else v) for k, v in inp.items()}
        best_threads(lambda: reference_pass(net, probe, 8))
        run = lambda: reference_pass(net, inp, frames)
        kind = 'reference'
        what = ('per step: ONE full sequence of %d frames through the unmodified reference (net/sig_mp.py Net.forward_online per frame + '
                'reset_states, evaluate.py:75-85,93) on the host cores, same weights / synthetic assets / inputs as rank 0 sequence 0' % frames)
    else:
        o = make_cpu_reference(sd, assets, 'aten')
        run = lambda: cpu_reference_pass(o, inp, frames)
        kind = 'port'
        what = ('per step: ONE full sequence of %d frames through the oracle port (oracle/fusion.py with torch.nn.LSTM, the ATen back end '
                'the reference runs on); baseline/_ref/reference_src.tar.gz was not found, so the reference itself could not be imported' % frames)
    times = []
    for i in range(max(args.warmup, 1) + args.steps):
        t0 = time.perf_counter()
        run()
        if i >= max(args.warmup, 1):
            times.append(time.perf_counter() - t0)
    ms = 1e3 * sum(times) / len(times)
    value = frames / (ms / 1e3)
    line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': max(args.warmup, 1),
            'ms_per_step': ms, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'impl': 'reference',
            'config': workload_config(args, world, None),
            'cpu_baseline': {'value': value, 'unit': 'frames/s', 'cores': torch.get_num_threads(), 'host_cpus': os.cpu_count(), 'kind': kind,
                             'sample': what + '; the reference processes sequences one after another (batch 1, stateful), so its frames/s does '
                                              'not depend on how many sequences the job has'},
            'e2e': {'value': value, 'unit': 'frames/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
            'gpu_launches': 0}
    print(json.dumps(line), flush=True)


def timed_passes(fn, n, barrier):
    """n calls of fn between two CUDA events (device time, ms per call)."""
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    barrier()
    return e0.elapsed_time(e1) / n


def main():
    args = parse()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if args.impl == 'reference':
        run_reference(args, rank, world)
        return
    assert torch.cuda.is_available(), 'bench.py needs a GPU (there is no CPU fallback for the product path)'
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    from robustcap_b200 import _lib, synthetic
    from robustcap_b200 import distributed as rdist
    if rank == 0:
        _lib.build()
        synthetic.write_assets(synthetic.default_asset_root(), 0)
    if dist is not None:
        dist.barrier()
    lib = _lib.load()
    net, sd, assets = build_net()
    net.set_gemm_mode(args.gemm_mode)
    T = args.frames
    if args.scaling == 'strong':
        lo, hi = rdist.shard_bounds(args.seqs, rank, world)
        total_seqs = args.seqs
    else:
        lo, hi = 0, args.seqs
        total_seqs = args.seqs * world
    B = hi - lo                                     # this rank's shard
    counts = [rdist.shard_bounds(args.seqs, r, world)[1] - rdist.shard_bounds(args.seqs, r, world)[0] for r in range(world)] \
        if args.scaling == 'strong' else [args.seqs] * world
    inp = synthetic.make_inputs(B, T, seed=1000 + rank, conf=args.conf)
    type(net).gravityc = inp['gravity'].clone()
    j, a, o = inp['j2dc'].to(dev), inp['accc'].to(dev), inp['oric'].to(dev)
    ft = torch.tensor([0., 0., 4.], device=dev)
    res = rdist.ShardedResult(B, T, dev)            # one flat [pose | tran] buffer per rank: ONE collective for both
    recv = None
    if world > 1 and rank == 0:
        recv = [torch.empty(c * T * 219, device=dev) for c in counts]

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step(graph=True, jj=j, aa=a, oo=o):
        # graph=True is the default behaviour of Net.forward_offline (multi-launch path: frames 1..T-1 replay one captured CUDA
        # graph; shards of <= 128 sequences: the persistent sequence kernel takes over after the warm-up frames)
        net.forward_offline(jj, aa, oo, first_tran=ft, use_graph=graph, out=(res.pose, res.tran))
        if dist is not None:
            if len(set(counts)) == 1:
                dist.gather(res.flat, recv if rank == 0 else None, dst=0)
            else:
                rdist.gather_flat(res, total_seqs, 0, None, recv)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    st = net._states[B]
    clocks = ClockSampler(local)
    clocks.start()
    launches0 = lib.rc_launch_count()
    ms_step = timed_passes(step, args.steps, barrier)
    launches = lib.rc_launch_count() - launches0
    clk = clocks.stop()
    # Roofline pass: the same K steps once more with plain stream launches, so that CUDA events can sit right around every launch
    # of the dominant kernel on its launch stream (kernels inside a graph replay cannot be bracketed).  Same kernels, same data.
    import ctypes
    step(False)
    barrier()
    _lib.check(lib.rc_profile_enable(st, 1))
    roof_ms = timed_passes(lambda: step(False), args.steps, barrier)
    tot_ms, nl, fpr = ctypes.c_double(), ctypes.c_int64(), ctypes.c_double()
    _lib.check(lib.rc_profile_collect(st, ctypes.byref(tot_ms), ctypes.byref(nl), ctypes.byref(fpr)))
    _lib.check(lib.rc_profile_enable(st, 0))

    def allmax(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_step = allmax(ms_step)
    frames_step = total_seqs * T
    value = frames_step / (ms_step / 1e3)

    # ---- occlusion workload (BASELINE configs[4]): key-point confidence 0 on 50 % of the frames -> IMU-only branch + vision updater ----
    occ = None
    if not args.no_occlusion:
        oi = synthetic.make_inputs(B, T, seed=2000 + rank, conf='occluded')
        oj, oa, oo = oi['j2dc'].to(dev), oi['accc'].to(dev), oi['oric'].to(dev)
        for _ in range(2):
            step(True, oj, oa, oo)
        occ_ms = allmax(timed_passes(lambda: step(True, oj, oa, oo), max(2, min(args.steps, 3)), barrier))
        occ = {'conf': 'occluded: confidence 0 on a seeded random 50 % of the frames, U(0.9, 1) elsewhere', 'ms_per_step': occ_ms,
               'value': frames_step / (occ_ms / 1e3), 'unit': 'frames/s'}
        del oj, oa, oo

    # ---- end-to-end through the host-buffer C-ABI entry point (pinned buffers) --------------------------------
    pin = lambda x: x.contiguous().pin_memory()
    hj, ha, ho = pin(inp['j2dc']), pin(inp['accc']), pin(inp['oric'])
    hp = torch.empty(B, T, 24, 3, 3).pin_memory()
    ht = torch.empty(B, T, 3).pin_memory()
    hft = torch.tensor([0., 0., 4.])
    for _ in range(2):
        net.forward_offline(hj, ha, ho, first_tran=hft, out=(hp, ht))
    barrier()
    t0 = time.perf_counter()
    reps = max(2, min(args.steps, 3))
    for _ in range(reps):
        net.forward_offline(hj, ha, ho, first_tran=hft, out=(hp, ht))
    barrier()
    e2e_ms = allmax(1e3 * (time.perf_counter() - t0) / reps)
    h2d = (hj.numel() + ha.numel() + ho.numel()) * 4 + 12
    d2h = (hp.numel() + ht.numel()) * 4
    h2d_all, d2h_all = allmax(h2d) * world, allmax(d2h) * world     # bytes of the whole job per step (equal shards)

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    pk = peaks()
    seq_kernel = args.gemm_mode == 3 or (args.gemm_mode == 2 and B <= 128)
    if args.gemm_mode >= 2:
        # dominant kernel: the persistent tcgen05 kernel(s); their launches of one pass run the whole LSTM stack of every stream-frame once
        dom_flop = fpr.value * B * T * args.steps
        if seq_kernel:
            kname = ('rc_seq_kernel<64> (persistent SEQUENCE kernel: frames 16..T-1 of the pass in ONE launch, one CTA per SM; 128 x 64 tcgen05 tiles of '
                     'every linear1 / LSTM / linear2 layer and the per-frame row jobs in one dependency queue) + rc_tc_phase_kernel for the 16 warm-up '
                     'frames; kind::f16 on split-fp16 operands, 3 MMAs per fp32-accurate product; CUDA events around every launch')
        else:
            kname = ('rc_tc_phase_kernel (persistent grouped tcgen05 GEMM, one CTA per SM, 128 x 128 tiles from a global queue: all linear1 / LSTM / '
                     'linear2 layers of a phase of the frame in one launch, 3 launches per frame; kind::f16 on split-fp16 operands, 3 MMAs per '
                     'fp32-accurate product; CUDA events around every launch)')
    else:
        dom_flop = fpr.value * 2 * B * T * args.steps
        kname = ('rc_tc_kernel<128,3,LSTM> (rnn4 fused LSTM layer [rows,2560]x[2560,5120]; tcgen05 kind::f16 on split-fp16 operands, '
                 '3 MMAs per fp32-accurate product; CUDA events on its launch stream while the other lane runs concurrently)')
    dom_tflops = dom_flop / (tot_ms.value * 1e-3) / 1e12 if tot_ms.value > 0 else 0.0
    per_gpu_ms = ms_step
    roofline = {'kernel': kname,
                'bound': 'tensor',
                'achieved': dom_tflops, 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                'frac': dom_tflops / pk['tflops_sustained'], 'peak_source': 'bf16 sustained, of ' + pk['source'],
                'traffic': dominant_traffic(args.gemm_mode), 'launches': int(nl.value), 'share_of_step': tot_ms.value / (roof_ms * args.steps),
                'timed_with': 'a second pass of the same K steps with stream launches (%.1f ms per step) — the value pass replays a CUDA graph' % roof_ms,
                'tensor_pipe_frac': 3 * dom_tflops / pk['tflops_sustained'],   # 3 fp16 MMAs are issued per algorithmic fp32 product
                'whole_path_tflops_per_gpu': FLOP_PER_FRAME * B * T / (per_gpu_ms * 1e-3) / 1e12,
                'weight_stream_gbs_per_gpu': WEIGHT_BYTES * T / (per_gpu_ms * 1e-3) / 1e9,
                'hbm_frac_weights': WEIGHT_BYTES * T / (per_gpu_ms * 1e-3) / 1e9 / pk['hbm_gbs'],
                'note': 'rank 0 of %d; per GPU a pass streams the 243 MB of weights once per frame: at %d sequences per GPU the frame is bound by %s'
                        % (world, B, 'the tensor pipe' if B >= 512 else 'the depth of its dependency chain (12 dependent GEMM layers + row logic), not by HBM or the tensor pipe')}
    cfg = workload_config(args, world, B)
    run_info = dict({'sequences_per_gpu': B,
                'l2': 'inputs (%.0f MB per GPU) + weights (243 MB) per step exceed the 126 MB L2' % ((hj.numel() + ha.numel() + ho.numel()) * 4 / 1e6),
                'final_gather': 'one nccl gather of the flat [pose | tran] buffer to rank 0 inside the timed region' if world > 1 else 'none',
                'launch': ('persistent sequence kernel after 16 warm-up frames (shards of <= 128 sequences)' if seq_kernel else
                           'CUDA-graph replay of the per-frame launch sequence (Net.forward_offline default)')})
    line = {'metric': METRIC, 'value': value, 'unit': 'frames/s', 'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32',
            'data': 'synthetic', 'config': cfg, 'run': run_info,
            'clocks': clk, 'gpu_launches': int(launches),
            'e2e': {'value': frames_step / (e2e_ms / 1e3), 'unit': 'frames/s', 'h2d_bytes_per_step': int(h2d_all), 'd2h_bytes_per_step': int(d2h_all),
                    'ms_per_step': e2e_ms},
            'roofline': roofline}
    if occ is not None:
        occ['weight_stream_gbs_per_gpu'] = WEIGHT_BYTES * T / (occ['ms_per_step'] * 1e-3) / 1e9
        occ['hbm_frac_weights'] = occ['weight_stream_gbs_per_gpu'] / pk['hbm_gbs']
        occ['whole_path_tflops_per_gpu'] = FLOP_PER_FRAME * B * T / (occ['ms_per_step'] * 1e-3) / 1e12
        occ['note'] = ('BASELINE configs[4]; every occluded frame runs the IMU-only branch AND the vision updater (rnn4 + rnn6 on re-projected key points, '
                       'net/sig_mp.py:263-271), so all six LSTM stacks still run for every stream-frame; HBM figure = 243 MB of weights per frame per GPU over the measured copy peak')
        line['occlusion'] = occ

    # ---- B=1 streaming latency (BASELINE configs[1]) ---------------------------------------------------------------
    if not args.no_stream_latency:
        s1 = synthetic.make_inputs(1, 300, seed=5, conf='high')
        net.reset_states()
        lat = []
        for t in range(300):
            t0 = time.perf_counter()
            net.forward_online(s1['j2dc'][0, t], s1['accc'][0, t], s1['oric'][0, t])
            lat.append(time.perf_counter() - t0)
        net.reset_states()
        dj, da, do = s1['j2dc'][0].to(dev), s1['accc'][0].to(dev), s1['oric'][0].to(dev)
        for _ in range(2):
            net.forward_offline(dj, da, do, use_graph=True)
        torch.cuda.synchronize()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        net.forward_offline(dj, da, do, use_graph=True)
        g1.record()
        torch.cuda.synchronize()
        dev_us = 1e3 * g0.elapsed_time(g1) / 300
        line['streaming'] = {'batch': 1, 'kernel': 'rc_stream2_kernel: one launch per frame, one CTA per SM, weights staged through a shared-memory ring by '
                                                   'cp.async.bulk (TMA), recurrent halves W_hh.h off the dependency chain, 8 split-phase grid barriers',
                             'latency_p50_us_forward_online': 1e6 * statistics.median(lat[20:]),
                             'device_us_per_frame': dev_us, 'weight_stream_gbs': WEIGHT_BYTES / (dev_us * 1e-6) / 1e9,
                             'hbm_frac': WEIGHT_BYTES / (dev_us * 1e-6) / 1e9 / pk['hbm_gbs'], 'peak_source': 'hbm copy, of ' + pk['source']}
    if not args.no_smplify:
        try:
            line['smplify'] = smplify_line(net, assets, dev, pk)
        except Exception as e:                       # never lose the headline line to the secondary workload
            line['smplify'] = {'error': '%s: %s' % (type(e).__name__, e)}
    if not args.no_cpu_baseline:
        line['cpu_baseline'] = cpu_reference_rate(sd, assets, args.cpu_frames, args.conf, 'aten')
        line['cpu_baseline']['host_cpus'] = os.cpu_count()
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def smplify_line(net, assets, dev, pk, seqs=64, frames=300, max_iter=5, cpu_frames=24):
    """BASELINE configs[3]: fusion RNN -> SMPL FK -> SMPLify with max_iter = 5 (the "5-iter inner loop"), `seqs` sequences x 300
    frames, everything device-resident: Net.forward_offline (tcgen05 path) -> key points of the result (FK + skinning of the 21
    vertices) -> smplify_runner_batch (one thread block per sequence runs torch.optim.LBFGS.step with the strong-Wolfe line search,
    closure = analytic objective + gradient).  Pixel key points = projection of the net's own output + N(0, 5^2) px."""
    from robustcap_b200 import synthetic, smplify, math as M
    import oracle.smplify as osm
    from oracle.kinematics import BodyOracle
    cwd = os.getcwd()
    os.chdir(os.path.dirname(os.path.dirname(assets['gmm_dir'])))       # the prior is opened relative to the CWD like the reference (prior.py:97)
    try:
        smplify.TemporalSMPLify.body_model = net._body
        inp = synthetic.make_inputs(seqs, frames, seed=4000, conf='high')
        type(net).gravityc = inp['gravity'].clone()
        j, a, o = inp['j2dc'].to(dev), inp['accc'].to(dev), inp['oric'].to(dev)
        ft = torch.tensor([0., 0., 4.], device=dev)
        cam_k = torch.tensor([[1000.0, 0, 960], [0, 1000, 540], [0, 0, 1]], device=dev)
        gen = torch.Generator().manual_seed(4001)
        noise = (5 * torch.randn(seqs, frames, 33, 2, generator=gen)).to(dev)
        conf = (0.5 + 0.5 * torch.rand(seqs, frames, 33, 1, generator=gen)).to(dev)

        def fused():
            pose, tran = net.forward_offline(j, a, o, first_tran=ft)
            _, kp3 = net._body.keypoints33(pose.reshape(-1, 24, 3, 3), tran.reshape(-1, 3))
            uv = (cam_k @ (kp3 / kp3[..., 2:]).unsqueeze(-1)).squeeze(-1)[..., :2].reshape(seqs, frames, 33, 2) + noise
            kp = torch.cat((uv, conf), dim=-1)
            return pose, tran, kp

        pose, tran, kp = fused()
        # one optimiser object for the whole run (the reference re-reads the GMM pickle and rebuilds everything per call, run.py:21)
        sm = smplify.TemporalSMPLify(cam_k=cam_k, imu_ori=o[0], step_size=1e-3, batch_size=frames, max_iter=max_iter, body_model=net._body)
        ign = sm.ign_mp_joints

        def run():
            # smplify_runner_batch without the per-call construction: start point (cv2.Rodrigues semantics), reference points, IMU
            # axis-angles, then ONE launch of the device-resident L-BFGS for all sequences
            pm = pose.reshape(seqs * frames, 24, 3, 3)
            aa = M.rotation_matrix_to_axis_angle(pm).reshape(seqs * frames, 72)
            _, ref3d = net._body.keypoints33(pm, tran.reshape(-1, 3))
            k2 = kp.reshape(seqs * frames, 33, 3).clone()
            k2[:, ign, 2] = 0.
            imu_aa = M.rotation_matrix_to_axis_angle(o.reshape(-1, 3, 3)).reshape(seqs * frames, 18)
            aa2, tr2, st = sm.optimise(aa, tran.reshape(-1, 3), k2[:, :, :2], k2[:, :, 2], ref3d, imu_aa)
            return M.axis_angle_to_rotation_matrix(aa2), tr2.reshape(seqs, frames, 3), st

        for _ in range(2):
            p2, t2, stats = run()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        reps = 3
        e0.record()
        for _ in range(reps):
            pose, tran, kp = fused()
        e1.record()
        for _ in range(reps):
            p2, t2, stats = run()
        e2.record()
        torch.cuda.synchronize()
        fus_ms, smp_ms = e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps
        evals = float(stats[:, 2].mean().item())
        nfr = seqs * frames
        moved_mm = 1e3 * (t2 - tran).abs().max().item()
        # bounded CPU sample: the oracle port of smplify_runner (autograd through the skinned points + torch.optim.LBFGS) on one short sequence
        body = BodyOracle(assets['smpl_file'])
        cf = min(cpu_frames, frames)
        t0 = time.perf_counter()
        osm.smplify_runner(body, os.path.join(assets['gmm_dir'], 'gmm_08.pkl'), pose[0, :cf].cpu(), tran[0, :cf].cpu(), kp[0, :cf].cpu().clone(),
                           o[0, :cf].cpu(), cf, cam_k.cpu(), lr=1e-3, loss_threshold=1e12, max_iter=max_iter)
        cpu_s = time.perf_counter() - t0
        alg_bytes = 1.2e3 * nfr * evals                     # ~1.2 KB and 0.15 MFLOP per frame and closure evaluation
        return {'workload': 'fusion -> FK -> SMPLify(max_iter=%d), %d seq x %d frames (BASELINE configs[3])' % (max_iter, seqs, frames),
                'value': nfr / ((fus_ms + smp_ms) * 1e-3), 'unit': 'frames/s (fused pipeline, device-timed)',
                'fusion_fk_ms': fus_ms, 'smplify_ms': smp_ms, 'smplify_frames_per_s': nfr / (smp_ms * 1e-3),
                'closure_evaluations': evals, 'iterations': float(stats[:, 3].mean().item()),
                'loss_first_mean': float(stats[:, 0].mean().item()), 'loss_final_mean': float(stats[:, 1].mean().item()),
                'max_translation_change_mm': moved_mm, 'gpu_launches_smplify': 1,
                'roofline': {'kernel': 'rc_smplify_lbfgs_kernel (one thread block per sequence: closure = GMM prior, warp per frame; forward / analytic '
                                       'backward, thread per frame; L-BFGS two-loop recursion + strong-Wolfe line search, reductions inside the block)',
                             'bound': 'hbm', 'achieved': alg_bytes / (smp_ms * 1e-3) / 1e9, 'peak': pk['hbm_gbs'], 'unit': 'GB/s',
                             'frac': alg_bytes / (smp_ms * 1e-3) / 1e9 / pk['hbm_gbs'], 'traffic': None,
                             'note': 'algorithmic 1.2 KB per frame and closure evaluation; %d resident blocks of one sequence each: the kernel is bound by '
                                     'the latency of the per-frame chain (24-joint FK, 33 points, backward), not by bandwidth' % seqs},
                'cpu_baseline': {'value': cf / cpu_s, 'unit': 'frames/s', 'kind': 'port', 'cores': torch.get_num_threads(),
                                 'sample': 'oracle/smplify.py smplify_runner (autograd + torch.optim.LBFGS, max_iter=%d) on 1 sequence x %d frames' % (max_iter, cf)}}
    finally:
        os.chdir(cwd)


if __name__ == '__main__':
    main()

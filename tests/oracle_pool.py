"""Test infrastructure: run the per-stream CPU oracle (oracle/fusion.py) on several sequences in parallel worker processes.

The oracle is a batch-1 Python loop like the reference (evaluate.py:75-85); 300-frame parity checks need a handful of sequences in
float32 AND float64, so the sequences are spread over a process pool (each worker rebuilds the seeded weights once)."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor
import multiprocessing as mp

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_W = {}


def _init(wseed, variant, smpl_file, threads):
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    import torch
    torch.set_num_threads(threads)
    from robustcap_b200 import synthetic
    from oracle.kinematics import BodyOracle
    from oracle.fusion import FusionOracle
    sd = synthetic.make_state_dict(wseed, variant)
    _W['o32'] = FusionOracle(sd, BodyOracle(smpl_file))
    _W['o64'] = FusionOracle(sd, BodyOracle(smpl_file, dtype=torch.float64), dtype=torch.float64)


def _run(job):
    import torch
    which, j, a, o, grav, kw = job
    oc = _W[which]
    br = []
    p, t = oc.run(torch.from_numpy(j), torch.from_numpy(a), torch.from_numpy(o), gravity=torch.from_numpy(grav), branches=br,
                  first_frame=kw.get('first_frame', False),
                  first_tran=None if kw.get('first_tran') is None else torch.tensor(kw['first_tran']))
    return p.numpy(), t.numpy(), br


def oracle_many(wseed, variant, smpl_file, seqs, gravity, workers=None, threads=2):
    """seqs: list of (j2dc[T,33,3], accc[T,6,3], oric[T,6,3,3], kwargs) CPU tensors.  Returns {'o32': [(pose, tran, branches)],
    'o64': [...]} in the same order (float32 / float64 oracle)."""
    import torch
    ncpu = os.cpu_count() or 2
    workers = workers or max(1, min(2 * len(seqs), ncpu // threads, 32))
    jobs = []
    for which in ('o32', 'o64'):
        for j, a, o, kw in seqs:
            jobs.append((which, j.numpy(), a.numpy(), o.numpy(), gravity.numpy(), kw))
    with ProcessPoolExecutor(workers, mp_context=mp.get_context('spawn'), initializer=_init,
                             initargs=(wseed, variant, smpl_file, threads)) as ex:
        res = list(ex.map(_run, jobs))
    n = len(seqs)
    conv = lambda r: (torch.from_numpy(r[0]), torch.from_numpy(r[1]), r[2])
    return {'o32': [conv(r) for r in res[:n]], 'o64': [conv(r) for r in res[n:]]}

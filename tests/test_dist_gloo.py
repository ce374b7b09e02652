"""world_size-2 gloo test of the multi-GPU host logic (candidate sharding + the single all-gather exchange step).
The per-rank scorer is the CPU oracle here (tests may use it); on a GPU box the same code path calls libgpfo."""
import os
import socket
import subprocess
import sys
import textwrap

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent("""
    import os, sys
    import numpy as np
    sys.path.insert(0, %r)
    import torch.distributed as td
    from gpflowopt_b200 import dist, synthetic
    from oracle import acq as oacq, gp as ogp
    td.init_process_group("gloo", rank=int(os.environ["RANK"]), world_size=int(os.environ["WORLD_SIZE"]))
    X, Y = synthetic.training_set(40, 3)
    h = synthetic.hypers(3)
    g = ogp.ScaledGPR(X, Y, ogp.MATERN52, h["lengthscales"], h["variance"], h["noise_variance"])
    ei = oacq.Leaf('ei', g, float(np.min(g.predict(X)[0])))
    # rank 0 owns the candidate matrix; the others start from garbage of the same shape
    pts = synthetic.candidates(1001, 3) if td.get_rank() == 0 else np.zeros((1001, 3))
    pts = dist.broadcast_points(pts)
    def local_argmax(rows):
        s = ei.value(rows)[:, 0]
        i = int(np.argmax(s))
        return float(s[i]), i
    v, i = dist.sharded_argmax(local_argmax, pts)
    full = ei.value(pts)[:, 0]
    assert i == int(np.argmax(full)) and v == full[i], (i, int(np.argmax(full)))
    # tie across shards: duplicate the winner into the other shard -> lowest global index must win on every rank
    pts2 = pts.copy()
    other = (i + 600) %% 1001
    pts2[other] = pts2[i]
    v2, i2 = dist.sharded_argmax(local_argmax, pts2)
    assert i2 == min(i, other) and v2 == v, (i2, i, other)
    # more ranks than rows, and an empty candidate set
    v3, i3 = dist.sharded_argmax(local_argmax, pts[:1])
    assert i3 == 0
    v4, i4 = dist.sharded_argmax(local_argmax, pts[:0])
    assert i4 == -1
    # scatter route used by MCOptimizer / CandidateOptimizer: every rank receives only its own row block of rank 0's matrix
    src = synthetic.candidates(1001, 3) if td.get_rank() == 0 else np.full((7, 2), np.nan)
    block, lo, n = dist.scatter_points(src)
    blo, bhi = dist.shard_bounds(1001, td.get_rank(), td.get_world_size())
    assert n == 1001 and lo == blo and block.shape == (bhi - blo, 3) and np.array_equal(block, pts[blo:bhi])
    for idx in (0, 500, 501, 1000):
        assert np.array_equal(dist.fetch_row(block, lo, idx, n, 3), pts[[idx]])
    from gpflowopt_b200.optim import CandidateOptimizer
    from gpflowopt_b200.domain import UnitCube
    class Acq(object):                                     # stands in for a gpflowopt_b200 acquisition (fused route)
        def argmax(self, rows):
            return local_argmax(np.asarray(rows))
    def inverse(x):
        return -ei.value(np.atleast_2d(x))
    inverse.acquisition = Acq()
    cands = pts if td.get_rank() == 0 else pts[::-1].copy()     # only rank 0's matrix counts
    res = CandidateOptimizer(UnitCube(3), cands).optimize(inverse)
    assert np.array_equal(res.x, pts[[i]]) and float(res.fun[0, 0]) == -v and res.nfev == 1001
    print("rank %%d ok %%d %%r" %% (td.get_rank(), i, v))
    td.destroy_process_group()
""") % ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_sharded_argmax_world_size_2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = _free_port()
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    # both ranks agree on the winner
    assert outs[0].split("ok")[1] == outs[1].split("ok")[1]


def test_single_process_pipeline_matches_one_shot():
    """dist.pipelined_argmax without a process group: large draws are produced chunk by chunk on a worker thread while the
    previous chunk is scored; rows, order, winner and tie-break are those of one produce(0, n) call; a producer error is
    re-raised on the caller's thread."""
    import numpy as np
    from gpflowopt_b200 import dist
    n, d = 300000, 8                                   # 2.4 M numbers: above the single-process pipelining threshold

    def produce(first, count, out=None):
        X = np.random.rand(count, d)
        if out is not None:
            out[...] = X
            return out
        return X

    def score(block):
        v = np.round(np.sin(block).sum(axis=1), 3)     # rounded: many exact ties across chunks
        i = int(np.argmax(v))
        return float(v[i]), i

    seen = []
    np.random.seed(5)
    v, i, x = dist.pipelined_argmax(n, d, produce, score, on_chunk_rows=seen.append)
    np.random.seed(5)
    rows = np.random.rand(n, d)
    ref = np.round(np.sin(rows).sum(axis=1), 3)
    assert seen == [37500] and i == int(np.argmax(ref)) and v == ref[i] and np.array_equal(x, rows[i:i + 1])
    assert np.array_equal(np.random.rand(3), np.random.RandomState(5).rand(n * d + 3)[-3:])     # the stream went on from the same place

    def bad(first, count, out=None):
        if first > 0:
            raise RuntimeError("boom")
        return produce(first, count, out)
    import pytest
    with pytest.raises(RuntimeError):
        dist.pipelined_argmax(n, d, bad, score)

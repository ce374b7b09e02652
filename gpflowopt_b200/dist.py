"""Candidate-row data parallelism: one process per GPU (torch.distributed), contiguous row shards, and ONE
exchange step per evaluation -- an all-gather of a 16-byte (value, global index) pair per rank followed by the
same lexicographic reduce (max value, then min index) on every rank.  NCCL has no MAXLOC, hence gather + reduce.

Every rank factorises the GPs redundantly (deterministic kernels -> bit-identical state), so there is no
broadcast on the setup side.  Without an initialised process group everything degenerates to a single shard.
"""
import numpy as np


def _pg():
    try:
        import torch.distributed as td
    except Exception:                                   # pragma: no cover
        return None
    return td if td.is_available() and td.is_initialized() else None


def world():
    td = _pg()
    return (td.get_rank(), td.get_world_size()) if td is not None else (0, 1)


def shard_bounds(n, rank, world_size):
    """Contiguous block [lo, hi) of rank ``rank``: blocks of ceil(n / world_size) rows, so that
    global index = lo + local index and the lowest-index tie-break stays trivial."""
    per = -(-n // world_size) if n > 0 else 0
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def _device_for_collectives(td):
    import torch
    if td.get_backend() == "nccl":
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def broadcast_points(points):
    """Makes all ranks score the same candidate matrix (rank 0's); identity without a process group."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return points
    import torch
    dev = _device_for_collectives(td)
    shape = torch.tensor(list(points.shape), dtype=torch.int64, device=dev)
    td.broadcast(shape, src=0)
    n, d = int(shape[0]), int(shape[1])
    buf = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev) if td.get_rank() == 0 \
        else torch.empty((n, d), dtype=torch.float64, device=dev)
    td.broadcast(buf, src=0)
    return buf.cpu().numpy()


def scatter_points(points):
    """Rank 0's candidate matrix split into the contiguous row blocks of ``shard_bounds``: every rank receives only its own
    block (N x D doubles leave rank 0 once instead of world_size times).  Returns (block, lo, n_total); under NCCL the block
    is a CUDA tensor that is scored in place through the device-pointer entry, otherwise a NumPy array."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return points, 0, points.shape[0]
    import torch
    dev = _device_for_collectives(td)
    rank, ws = td.get_rank(), td.get_world_size()
    shape = torch.tensor(list(points.shape), dtype=torch.int64, device=dev)
    td.broadcast(shape, src=0)
    n, d = int(shape[0]), int(shape[1])
    per = max(1, -(-n // ws))
    out = torch.empty((per, d), dtype=torch.float64, device=dev)
    chunks = None
    if rank == 0:
        full = torch.zeros((ws * per, d), dtype=torch.float64, device=dev)
        full[:n] = torch.from_numpy(np.ascontiguousarray(points, dtype=np.float64)).to(dev)
        chunks = list(full.split(per))
    td.scatter(out, scatter_list=chunks, src=0)
    lo, hi = shard_bounds(n, rank, ws)
    block = out[:hi - lo]
    return (block if block.is_cuda else block.numpy()), lo, n


def _produce_into(produce, first, count, out):
    """``produce(first, count, out=out)`` when the producer takes a destination buffer, ``produce(first, count)`` otherwise."""
    try:
        return produce(first, count, out=out)
    except TypeError:
        return produce(first, count)


_SINGLE_PIPELINE_MIN_ELEMS = 1 << 21       # below 2 M numbers the draw is a fraction of a millisecond: one chunk


def _pipelined_single(n_total, d, produce, score, rounds, on_chunk_rows):
    """One process, one GPU: the rows are produced in ``rounds`` chunks, chunk q + 1 on a worker thread (the native draw and the
    scoring call both release the GIL) while chunk q is scored -- the host side of MCOptimizer(domain, 10^6) (25 ms of a 180 ms
    inner optimisation at M = 2048) hides behind the device time.  Same rows in the same order as one ``produce(0, n_total)``
    call (successive draws continue np.random's stream); ties keep the lowest global index."""
    import threading
    C = -(-n_total // rounds)
    rounds = -(-n_total // C)
    if on_chunk_rows is not None:
        on_chunk_rows(C)                        # one kernel for every chunk: identical rows score identically
    bufs = [np.empty((C, d), dtype=np.float64), np.empty((C, d), dtype=np.float64)]
    box = {}

    def make(q):
        try:
            first = q * C
            count = min(C, n_total - first)
            rows = _produce_into(produce, first, count, bufs[q % 2][:count])
            box[q] = np.ascontiguousarray(rows, dtype=np.float64)
        except BaseException as exc:            # re-raised on the calling thread
            box[q] = exc

    best = [0.0, -1, np.empty((0, d))]
    make(0)
    for q in range(rounds):
        cur = box.pop(q)
        if isinstance(cur, BaseException):
            raise cur
        worker = None
        if q + 1 < rounds:
            worker = threading.Thread(target=make, args=(q + 1,))
            worker.start()
        try:
            v, i = score(cur)
        finally:
            if worker is not None:
                worker.join()
        if i >= 0 and v == v and (best[1] < 0 or v > best[0]):        # ascending chunks: ties keep the first
            best = [v, q * C + i, np.array(cur[i:i + 1], dtype=np.float64)]
    return best[0], best[1], best[2]


def pipelined_argmax(n_total, d, produce, score, rounds=8, on_chunk_rows=None):
    """The reference's single-stream candidate route under data parallelism (gpflowopt/design.py:55-70, 83-94;
    optim.py:152-164), pipelined: ONLY rank 0 produces rows -- in order, ``produce(first_row, count)`` returns rows
    [first_row, first_row + count) as a float64 ndarray, so successive ``np.random.rand`` draws form the reference's stream --
    and the rows travel in ``rounds`` scatter steps of one equal chunk per rank.  Rank 0 scores its own chunk on a worker
    thread, so producing + uploading round q+1 overlaps the device time of round q on every GPU.  Chunk j (global rows
    [j C, (j+1) C)) goes to rank j mod world_size; every rank folds its chunks lexicographically (max value, lowest GLOBAL
    index) and keeps a host copy of its best row.

    ``score(block) -> (best value, best local index)``; ``block`` is a CUDA tensor under NCCL, an ndarray otherwise.
    ``on_chunk_rows(C)`` is called once on every rank, before the first score, with the rows per chunk.
    Returns (value, global index, row 1 x d) on every rank; index -1 when no row has a finite score."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        if n_total * d >= _SINGLE_PIPELINE_MIN_ELEMS and rounds >= 2:
            return _pipelined_single(n_total, d, produce, score, rounds, on_chunk_rows)
        rows = produce(0, n_total)
        v, i = score(rows) if n_total > 0 else (0.0, -1)
        return v, i, (np.array(rows[i:i + 1], dtype=np.float64) if i >= 0 else np.empty((0, d)))
    import threading
    import torch
    dev = _device_for_collectives(td)
    rank, ws = td.get_rank(), td.get_world_size()
    meta = torch.tensor([int(n_total), int(d)], dtype=torch.int64, device=dev)
    td.broadcast(meta, src=0)
    n_total, d = int(meta[0]), int(meta[1])
    rounds = max(1, min(int(rounds), -(-n_total // ws))) if n_total > 0 else 0
    C = -(-n_total // (ws * rounds)) if n_total > 0 else 0          # rows per chunk
    if on_chunk_rows is not None:
        on_chunk_rows(C)
    best = [0.0, -1, np.empty((0, d))]

    def fold(block, first_row, count):
        if count <= 0:
            return
        v, i = score(block[:count])
        if i >= 0 and v == v and (best[1] < 0 or v > best[0]):        # chunks arrive in ascending global order: ties keep the first
            row = block[i:i + 1]
            best[:] = [v, first_row + i, np.array(row.cpu().numpy() if hasattr(row, 'cpu') else row, dtype=np.float64)]

    worker = None
    staging = None
    import os, time
    debug = os.environ.get("GPFO_DIST_DEBUG") and rank == 0
    for q in range(rounds):
        t_round = time.perf_counter()
        first = q * ws * C                                             # first global row of this round
        out = torch.empty((C, d), dtype=torch.float64, device=dev)
        chunks = None
        if rank == 0:
            count = max(0, min(n_total - first, ws * C))
            if dev.type == "cuda":                  # rows are produced straight into a pinned staging buffer (reused every round)
                if staging is None:
                    staging = torch.empty((ws * C, d), dtype=torch.float64, pin_memory=True)
                rows = _produce_into(produce, first, count, staging.numpy()[:count]) if count > 0 else None
                if rows is not None and rows is not staging.numpy() and not np.shares_memory(rows, staging.numpy()):
                    staging.numpy()[:count] = rows
                if count < ws * C:
                    staging.numpy()[count:] = 0.0
                t_prod = time.perf_counter()
                full = staging.to(dev, non_blocking=True)
            else:
                full = torch.zeros((ws * C, d), dtype=torch.float64, device=dev)
                if count > 0:
                    full[:count] = torch.from_numpy(np.ascontiguousarray(produce(first, count), dtype=np.float64))
            chunks = list(full.split(C))
        t_sc0 = time.perf_counter()
        td.scatter(out, scatter_list=chunks, src=0)
        if debug:
            torch.cuda.synchronize() if dev.type == "cuda" else None
            print("round %d: produce %.3f s, h2d issue %.3f s, scatter %.3f s" % (
                q, (t_prod if dev.type == "cuda" else t_sc0) - t_round, t_sc0 - (t_prod if dev.type == "cuda" else t_sc0),
                time.perf_counter() - t_sc0), flush=True)
        mine = first + rank * C
        count = max(0, min(n_total - mine, C))
        block = out if out.is_cuda else out.numpy()
        if rank == 0:                                                 # score while the next round is being produced
            if worker is not None:
                t_j = time.perf_counter()
                worker.join()
                if debug:
                    print("          join of the previous scoring thread %.3f s" % (time.perf_counter() - t_j), flush=True)
            worker = threading.Thread(target=fold, args=(block, mine, count))
            worker.start()
        else:
            fold(block, mine, count)
    if worker is not None:
        worker.join()
    v, i = reduce_pairs(*allgather_pair(best[0], best[1]))
    if i < 0:
        return v, i, np.empty((0, d))
    owner = int((i // C) % ws)
    row = torch.empty((1, d), dtype=torch.float64, device=dev)
    if rank == owner:
        row.copy_(torch.from_numpy(best[2]))
    td.broadcast(row, src=owner)
    return v, i, row.cpu().numpy()


def fetch_row(block, lo, index, n_total, d):
    """Row ``index`` of the scattered matrix on every rank (its owner broadcasts D doubles)."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return np.array(block[index - lo:index - lo + 1], dtype=np.float64)
    import torch
    dev = _device_for_collectives(td)
    ws = td.get_world_size()
    per = max(1, -(-n_total // ws))
    owner = int(index // per)
    row = torch.empty((1, d), dtype=torch.float64, device=dev)
    if td.get_rank() == owner:
        src = block[index - lo:index - lo + 1]
        row.copy_(src if hasattr(src, 'is_cuda') else torch.from_numpy(np.ascontiguousarray(src)))
    td.broadcast(row, src=owner)
    return row.cpu().numpy()


def broadcast_int(value):
    """Rank 0's integer on every rank (seed of the device candidate generator)."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return int(value)
    import torch
    t = torch.tensor([int(value)], dtype=torch.int64, device=_device_for_collectives(td))
    td.broadcast(t, src=0)
    return int(t[0])


def reduce_pairs(values, indices):
    """Lexicographic reduce over per-rank (value, global index) pairs: max value, ties -> lowest index;
    index < 0 marks an empty shard; NaN never wins."""
    best_v, best_i = 0.0, -1
    for v, i in zip(values, indices):
        v, i = float(v), int(i)
        if i < 0 or v != v:
            continue
        if best_i < 0 or v > best_v or (v == best_v and i < best_i):
            best_v, best_i = v, i
    return best_v, best_i


def allgather_pair(value, index):
    """The one exchange step: all-gather of (f64 value, i64 global index)."""
    td = _pg()
    if td is None or td.get_world_size() == 1:
        return [value], [index]
    import torch
    dev = _device_for_collectives(td)
    ws = td.get_world_size()
    # one 16-byte record per rank: the int64 index travels bit-exactly inside a float64 lane
    rec = torch.empty(2, dtype=torch.float64, device=dev)
    rec[0] = float(value)
    rec[1:2].view(torch.int64)[0] = int(index)
    out = torch.empty(2 * ws, dtype=torch.float64, device=dev)
    td.all_gather_into_tensor(out, rec) if hasattr(td, "all_gather_into_tensor") and td.get_backend() == "nccl" \
        else td.all_gather(list(out.view(ws, 2).unbind(0)), rec)
    out = out.view(ws, 2).cpu()
    vals = [float(out[r, 0]) for r in range(ws)]
    idxs = [int(out[r, 1:2].view(torch.int64)[0]) for r in range(ws)]
    return vals, idxs


def sharded_argmax(local_argmax, points):
    """``local_argmax(rows) -> (best value, best local index)``; returns the global (value, index)."""
    rank, ws = world()
    lo, hi = shard_bounds(points.shape[0], rank, ws)
    if hi > lo:
        v, i = local_argmax(points[lo:hi])
        i = i + lo if i >= 0 else -1
    else:
        v, i = 0.0, -1
    vals, idxs = allgather_pair(v, i)
    return reduce_pairs(vals, idxs)

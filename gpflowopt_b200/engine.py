"""Host-side engine: turns an acquisition tree into (GP states, postfix program, cell tables) for libgpfo and
keeps the device copy in sync with the Python objects.

It plays the role GPflow's AutoFlow plays in the reference (acquisition/acquisition.py:247-267): AutoFlow caches
a tf.Graph + Session per object and re-runs the whole graph (Cholesky included) on every call; here the device
state is re-uploaded / re-factorised only when a model's state version changed, and a call is one gpfo_eval.
"""
import os

import numpy as np

from . import _lib

_default_device = None


def set_device(index):
    """Selects the CUDA device new engines are created on (default: $LOCAL_RANK or 0)."""
    global _default_device
    _default_device = int(index)


def get_device():
    if _default_device is not None:
        return _default_device
    return int(os.environ.get("LOCAL_RANK", "0"))


class ProgramBuilder(object):
    """The object handed to ``Acquisition.build_acquisition`` in place of the reference's ``Xcand`` tensor.

    Leaves and aggregations append postfix ops; the return values are opaque node tokens.
    """

    def __init__(self, slot_of):
        self._slot_of = slot_of          # id(DataScaler) -> first moment slot
        self.ops = []
        self.params = []
        self.cells = None                # (pf_ext, lb, ub, version key)

    def _param(self, value):
        self.params.append(float(np.asarray(value, dtype=np.float64).ravel()[0]))
        return len(self.params) - 1

    def _leaf(self, op, model, value, column=0):
        self.ops.append((op, self._slot_of[id(model)] + column, 0, self._param(value)))
        return len(self.ops) - 1

    def ei(self, model, fmin):
        return self._leaf(_lib.OP_EI, model, fmin)

    def poi(self, model, fmin):
        return self._leaf(_lib.OP_POI, model, fmin)

    def pof(self, model, threshold):
        return self._leaf(_lib.OP_POF, model, threshold)

    def lcb(self, model, sigma):
        return self._leaf(_lib.OP_LCB, model, sigma)

    def mes(self, model, samples):
        samples = np.asarray(samples, dtype=np.float64).ravel()
        first = len(self.params)
        self.params.extend(float(v) for v in samples)
        self.ops.append((_lib.OP_MES, self._slot_of[id(model)], samples.size, first))
        return len(self.ops) - 1

    def mean(self, model, column=0):
        self.ops.append((_lib.OP_MEAN, self._slot_of[id(model)] + column, 0, 0))
        return len(self.ops) - 1

    def hvpoi(self, models, pf_ext, lb, ub, key):
        slots = []
        for m in models:
            s0 = self._slot_of[id(m)]
            slots.extend(range(s0, s0 + m.Y.shape[1]))
        if slots != list(range(slots[0], slots[0] + len(slots))):
            raise _lib.GpfoError(_lib.ERR_UNSUPPORTED, "HVPoI objectives must occupy consecutive moment slots")
        self.ops.append((_lib.OP_HVPOI, slots[0], len(slots), 0))
        self.cells = (pf_ext, lb, ub, key)
        return len(self.ops) - 1

    def reduce(self, kind, nodes):
        self.ops.append((_lib.OP_SUM if kind == 'sum' else _lib.OP_PROD, len(nodes), 0, 0))
        return len(self.ops) - 1

    def scale(self, node, factor):
        self.ops.append((_lib.OP_SCALE, 0, 0, self._param(factor)))
        return len(self.ops) - 1


class Engine(object):
    def __init__(self, device=None):
        self.ctx = _lib.Context(get_device() if device is None else device)
        self._models = []                # DataScaler objects, index = gp id
        self._versions = {}              # gp id -> (id(scaler), state version, slot0)
        self._program_key = None
        self._cells_key = None

    def __deepcopy__(self, memo):
        return None                      # device state is never copied; a copy gets (or adopts) its own engine

    def close(self):
        if self.ctx is not None:
            self.ctx.close()
            self.ctx = None

    # ---- models -----------------------------------------------------------------------------------------
    def sync_models(self, models):
        """Makes device GP i mirror ``models[i]``; re-factorises only what changed."""
        models = list(models)
        if len(models) > _lib.MAX_GPS:
            raise _lib.GpfoError(_lib.ERR_UNSUPPORTED, "more than %d GPs in one acquisition" % _lib.MAX_GPS)
        slot = 0
        slot_of = {}
        for i, m in enumerate(models):
            R = m.wrapped.Y.shape[1]
            key = (id(m), m.state_version, slot)
            if self._versions.get(i) != key:
                self.ctx.gp_set(i, slot0=slot, **m.device_state())
                self._versions[i] = key
                self._program_key = None
            slot_of[id(m)] = slot
            slot += R
            m._engine = self
        if slot > _lib.MAX_SLOTS:
            raise _lib.GpfoError(_lib.ERR_UNSUPPORTED, "more than %d model outputs in one acquisition" % _lib.MAX_SLOTS)
        if len(models) != len(self._models):
            self.ctx.gp_count_set(len(models))
            for i in list(self._versions):
                if i >= len(models):
                    del self._versions[i]
        self._models = models
        return slot_of

    def predict(self, scaler, X):
        models = self._models if any(m is scaler for m in self._models) else self._models + [scaler]
        self.sync_models(models)
        gp_id = [i for i, m in enumerate(self._models) if m is scaler][0]
        return self.ctx.predict(gp_id, X, scaler.wrapped.Y.shape[1])

    # ---- evaluation ---------------------------------------------------------------------------------------
    def prepare(self, acq):
        """Uploads everything ``acq`` (a node of the tree owning this engine) needs; returns nothing."""
        root = acq.highest_parent
        slot_of = self.sync_models(root._all_models())
        b = ProgramBuilder(slot_of)
        acq.build_acquisition(b)
        key = (tuple(b.ops), tuple(b.params))
        if key != self._program_key:
            self.ctx.program_set(np.array(b.ops, dtype=np.int32), np.array(b.params, dtype=np.float64))
            self._program_key = key
        if b.cells is not None and b.cells[3] != self._cells_key:
            self.ctx.cells_set(b.cells[0], b.cells[1], b.cells[2])
            self._cells_key = b.cells[3]

    def evaluate(self, acq, X, gradients=False):
        X = np.asarray(X)
        if X.ndim != 2 or X.dtype != np.float64:
            # AutoFlow placeholders are (float64, [None, None]) (acquisition/acquisition.py:248,260)
            X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        self.prepare(acq)
        vals, grads, _, _ = self.ctx.eval(X, want_values=True, want_grads=gradients)
        return (vals, grads) if gradients else vals

    def argmax(self, acq, X):
        """Best (value, index) over the candidate rows without bringing the scores back to the host.  ``X`` is a host
        array or a CUDA float64 tensor on this engine's device (torch is only the buffer owner: the rows are read in place)."""
        if hasattr(X, 'data_ptr') and getattr(X, 'is_cuda', False):
            import torch
            assert X.dtype == torch.float64 and X.dim() == 2
            X = X.contiguous()
            torch.cuda.current_stream(X.device).synchronize()      # the library launches on its own stream
            self.prepare(acq)
            return self.ctx.eval_device(X.data_ptr(), X.shape[0], X.shape[1])
        X = np.atleast_2d(np.asarray(X, dtype=np.float64))
        self.prepare(acq)
        _, _, bv, bi = self.ctx.eval(X, want_values=False, want_grads=False)
        return bv, bi

    def pin_variant_for(self, acq, n):
        """Pins the forward kernel the library would choose for batches of ``n`` rows; returns the variant to restore.
        Used when ONE candidate set is scored in several calls (pipelined multi-GPU route): with one kernel for all calls a
        row's score does not depend on the call it lands in, so duplicates tie exactly and the lowest index wins."""
        self.prepare(acq)
        previous = self.ctx.variant
        if previous == 0 and n > 0:
            self.ctx.set_variant(self.ctx.select_variant(n))
        return previous

    def argmax_random(self, acq, seed, first_row, n, lower, upper):
        """Best (value, LOCAL index) over rows first_row .. first_row+n-1 of the device-generated candidate stream."""
        self.prepare(acq)
        _, bv, bi = self.ctx.eval_random(seed, first_row, n, lower, upper)
        return bv, bi

    def evaluate_random(self, acq, seed, first_row, n, lower, upper):
        """Scores (n x 1) of rows first_row .. first_row+n-1 of the device-generated candidate stream."""
        self.prepare(acq)
        vals, _, _ = self.ctx.eval_random(seed, first_row, n, lower, upper, want_values=True)
        return vals

    def argmax_device(self, acq, x_ptr, N, D, values_ptr=None):
        self.prepare(acq)
        return self.ctx.eval_device(x_ptr, N, D, values_ptr)


def model_engine(scaler):
    """Engine serving ``predict_f`` of a DataScaler: the acquisition tree's engine if it has one."""
    if getattr(scaler, '_engine', None) is None or scaler._engine.ctx is None:
        scaler._engine = Engine()
    return scaler._engine

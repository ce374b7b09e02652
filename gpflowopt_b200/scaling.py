"""DataScaler: keeps the data inside a GP model scaled (mirror of gpflowopt/scaling.py).

Inputs go through ``input_transform`` (domain -> unit cube once ``Acquisition.enable_scaling`` ran, identity
before), outputs through ``output_transform`` (zero mean / unit variance when ``normalize_output`` is on).  On
the device both maps are folded into the contraction kernel (prologue / epilogue), so ``build_predict`` here
only describes them; ``predict_f`` runs through libgpfo.
"""
import numpy as np

from .domain import UnitCube
from .gpr import GPR, DataHolder
from .transforms import DataTransform, LinearTransform


class DataScaler(object):
    def __init__(self, model, domain=None, normalize_Y=False):
        assert isinstance(model, GPR), "the device path supports GP regression models only (UNSUPPORTED_MODEL)"
        self.wrapped = model
        n_inputs = model.X.shape[1]
        n_outputs = model.Y.shape[1]
        self._input_transform = (domain or UnitCube(n_inputs)) >> UnitCube(n_inputs)
        self._normalize_Y = normalize_Y
        self._output_transform = LinearTransform(np.ones(n_outputs), np.zeros(n_outputs))
        self._engine = None
        self.X = model.X.value
        self.Y = model.Y.value

    # ---- attribute forwarding to the wrapped model (role of gpflowopt/models.py ModelWrapper) -----------
    def __getattr__(self, item):
        if item in ('wrapped', '_input_transform', '_output_transform', '_normalize_Y', '_engine'):
            raise AttributeError(item)
        return getattr(self.wrapped, item)

    @property
    def state_version(self):
        return (self.wrapped.state_version, self._input_transform.version, self._output_transform.version)

    # ---- transforms ------------------------------------------------------------------------------------
    @property
    def input_transform(self):
        return self._input_transform

    @input_transform.setter
    def input_transform(self, t):
        assert isinstance(t, DataTransform)
        X = self.X.value                       # un-scale with the old transform
        self._input_transform.assign(t)
        self.X = X                             # re-scale with the new one

    @property
    def output_transform(self):
        return self._output_transform

    @output_transform.setter
    def output_transform(self, t):
        assert isinstance(t, DataTransform)
        Y = self.Y.value
        self._output_transform.assign(t)
        self.Y = Y

    @property
    def normalize_output(self):
        return self._normalize_Y

    @normalize_output.setter
    def normalize_output(self, flag):
        self._normalize_Y = flag
        if not flag:
            n = self.Y.value.shape[1]
            self.output_transform = LinearTransform(np.ones(n), np.zeros(n))
        else:
            self.Y = self.Y.value

    # ---- data (original units outside, scaled inside) ------------------------------------------------------
    @property
    def X(self):
        return DataHolder(self.input_transform.backward(self.wrapped.X.value))

    @X.setter
    def X(self, x):
        self.wrapped.X = self.input_transform.forward(x.value if isinstance(x, DataHolder) else x)

    @property
    def Y(self):
        return DataHolder(self.output_transform.backward(self.wrapped.Y.value))

    @Y.setter
    def Y(self, y):
        value = y.value if isinstance(y, DataHolder) else np.atleast_2d(y)
        if self.normalize_output:              # gpflowopt/scaling.py:179-180 (population std, ddof=0)
            self.output_transform.assign(~LinearTransform(value.std(axis=0), value.mean(axis=0)))
        self.wrapped.Y = self.output_transform.forward(value)

    # ---- prediction ----------------------------------------------------------------------------------------
    def device_state(self):
        """Everything gpfo_gp_set needs for this model."""
        in_scale, in_shift = self.input_transform.diagonal()
        out_scale, out_shift = self.output_transform.diagonal()
        k = self.wrapped.kern
        return dict(X=self.wrapped.X.value, Y=self.wrapped.Y.value, kind=k.kind, lengthscales=k.lengthscales,
                    variance=k.variance, noise_variance=self.wrapped.likelihood.variance, in_scale=in_scale,
                    in_shift=in_shift, out_scale=out_scale, out_shift=out_shift)

    def predict_f(self, Xnew):
        """Mean and variance of the latent function at Xnew (original units), N x R each."""
        from .engine import model_engine
        return model_engine(self).predict(self, np.atleast_2d(Xnew))

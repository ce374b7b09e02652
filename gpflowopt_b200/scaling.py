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


def _raw(value):
    return value.value if isinstance(value, DataHolder) else np.atleast_2d(value)


class _Transform(object):
    """Descriptor of one of the two data maps of a DataScaler.  Assigning a new map keeps the data the caller sees fixed:
    the wrapped model's copy is taken back to original units with the old map and re-scaled with the new one."""

    def __init__(self, field, data_name):
        self._field, self._data_name = field, data_name

    def __get__(self, scaler, owner=None):
        return self if scaler is None else getattr(scaler, self._field)

    def __set__(self, scaler, transform):
        assert isinstance(transform, DataTransform)
        original = getattr(scaler, self._data_name).value
        getattr(scaler, self._field).assign(transform)
        setattr(scaler, self._data_name, original)


class DataScaler(object):
    """A GPR seen through an input map (domain -> unit cube) and an output map (standardisation): callers read and write X, Y
    and predictions in original units, the wrapped model only ever holds scaled data (interface of gpflowopt/scaling.py:27-220;
    attribute forwarding as gpflowopt/models.py's ModelWrapper)."""

    _OWN = ('wrapped', '_in_map', '_out_map', '_standardise', '_engine')

    input_transform = _Transform('_in_map', 'X')
    output_transform = _Transform('_out_map', 'Y')

    def __init__(self, model, domain=None, normalize_Y=False):
        assert isinstance(model, GPR), "the device path supports GP regression models only (UNSUPPORTED_MODEL)"
        self.wrapped = model
        d_in, d_out = model.X.shape[1], model.Y.shape[1]
        cube = UnitCube(d_in)
        self._in_map = (cube if domain is None else domain) >> cube
        self._out_map = LinearTransform(np.ones(d_out), np.zeros(d_out))
        self._standardise = normalize_Y
        self._engine = None
        self.X, self.Y = model.X.value, model.Y.value          # scales what the model holds

    def __getattr__(self, name):                              # only reached for names this object does not define
        if name in DataScaler._OWN:
            raise AttributeError(name)
        return getattr(self.wrapped, name)

    @property
    def state_version(self):
        return (self.wrapped.state_version, self._in_map.version, self._out_map.version)

    # ---- output standardisation on / off ---------------------------------------------------------------------------
    @property
    def normalize_output(self):
        return self._standardise

    @normalize_output.setter
    def normalize_output(self, flag):
        self._standardise = flag
        if flag:
            self.Y = self.Y.value                             # recomputes mean / std and rescales
        else:
            d_out = self.Y.value.shape[1]
            self.output_transform = LinearTransform(np.ones(d_out), np.zeros(d_out))

    # ---- data: original units outside, scaled inside -----------------------------------------------------------------
    @property
    def X(self):
        return DataHolder(self._in_map.backward(self.wrapped.X.value))

    @X.setter
    def X(self, x):
        self.wrapped.X = self._in_map.forward(_raw(x))

    @property
    def Y(self):
        return DataHolder(self._out_map.backward(self.wrapped.Y.value))

    @Y.setter
    def Y(self, y):
        y = _raw(y)
        if self._standardise:                                 # population moments (ddof = 0), gpflowopt/scaling.py:179-180
            self._out_map.assign(~LinearTransform(y.std(axis=0), y.mean(axis=0)))
        self.wrapped.Y = self._out_map.forward(y)

    # ---- prediction ------------------------------------------------------------------------------------------------
    def device_state(self):
        """Everything gpfo_gp_set needs for this model."""
        in_scale, in_shift = self._in_map.diagonal()
        out_scale, out_shift = self._out_map.diagonal()
        k = self.wrapped.kern
        return dict(X=self.wrapped.X.value, Y=self.wrapped.Y.value, kind=k.kind, lengthscales=k.lengthscales,
                    variance=k.variance, noise_variance=self.wrapped.likelihood.variance, in_scale=in_scale,
                    in_shift=in_shift, out_scale=out_scale, out_shift=out_shift)

    def predict_f(self, Xnew):
        """Mean and variance of the latent function at Xnew (original units), N x R each."""
        from .engine import model_engine
        return model_engine(self).predict(self, np.atleast_2d(Xnew))

"""A minimal stand-in for the TensorFlow-1 graph/session pair the reference is written against.

The reference's device boundary is ``tf.Session.run(fetch, feed_dict)`` on symbolic tensors
(/root/reference/auto_pose/ae/codebook.py:63, ae_train.py:128).  Here a ``Tensor`` is a named node with an
evaluation closure that launches the CUDA kernels through the C ABI; ``Session.run`` evaluates fetches with the
feeds bound to ``Placeholder`` nodes and returns numpy arrays, exactly like TF.  ``Session.run_device`` returns the
raw torch CUDA tensors instead (no device->host copy) for callers that stay on the GPU.
"""
from contextlib import contextmanager

import numpy as np
import torch

_scope_stack = []


@contextmanager
def variable_scope(name):
    """tf.variable_scope: prefixes variable names (auto_pose/ae/ae_factory.py:131)."""
    _scope_stack.append(name)
    try:
        yield
    finally:
        _scope_stack.pop()


def scoped(name):
    return "/".join(_scope_stack + [name]) if _scope_stack else name


class Tensor:
    def __init__(self, name, shape=None, dtype=np.float32, fn=None):
        self.name = scoped(name)
        self._shape = tuple(shape) if shape is not None else None
        self.dtype = dtype
        self._fn = fn

    @property
    def shape(self):
        return self._shape

    def get_shape(self):
        return _Shape(self._shape)

    def __repr__(self):
        return "<aae Tensor %s shape=%s>" % (self.name, self._shape)


class _Shape(tuple):
    def as_list(self):
        return list(self)


class Placeholder(Tensor):
    def __init__(self, dtype=np.float32, shape=None, name="Placeholder"):
        super().__init__(name, shape, dtype, None)


def placeholder(dtype=np.float32, shape=None, name="Placeholder"):
    return Placeholder(dtype, shape, name)


class Variable(Tensor):
    """Holds a device tensor; evaluates to its current value."""

    def __init__(self, initial_value, dtype=np.float32, trainable=False, name="Variable"):
        arr = np.asarray(initial_value, dtype=dtype)
        super().__init__(name, arr.shape, dtype, None)
        self.trainable = trainable
        self._host = arr
        self._dev = None
        self.on_assign = None  # hook: called with the new numpy value (e.g. to rebuild a codebook handle)

    def assign(self, value):
        value = np.ascontiguousarray(np.asarray(value, dtype=self.dtype))
        if value.shape != self._shape:
            raise ValueError("assign to %s: shape %s != %s" % (self.name, value.shape, self._shape))
        self._host = value
        self._dev = None
        if self.on_assign is not None:
            self.on_assign(value)

    def value(self):
        return self._host


class RunContext:
    def __init__(self, session, feeds):
        self.session = session
        self.feeds = feeds
        self.memo = {}
        self.touched = []     # device modules whose kernels ran in this call (their range guard is checked by Session.run)

    def get(self, tensor):
        key = id(tensor)
        if key in self.memo:
            return self.memo[key]
        if key in self.feeds:
            v = self.feeds[key]
        elif isinstance(tensor, Variable):
            v = tensor.value()
        elif tensor._fn is not None:
            v = tensor._fn(self)
        else:
            raise ValueError("placeholder %s was not fed" % tensor.name)
        self.memo[key] = v
        return v


class Session:
    """Device + stream context.  ``config`` is accepted and ignored (tf.ConfigProto in the reference)."""

    def __init__(self, device=None, config=None):
        if not torch.cuda.is_available():
            raise RuntimeError("augmentedautoencoder_b200 needs a CUDA device: there is no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self._copy_stream = None

    @property
    def copy_stream(self):
        """Side stream for host<->device copies that overlap compute (Codebook.nearest_rotation_async)."""
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        return self._copy_stream

    @property
    def stream_ptr(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def _bind(self, feed_dict):
        feeds = {}
        for k, v in (feed_dict or {}).items():
            feeds[id(k)] = v
        return feeds

    def _evaluate(self, fetches, feed_dict):
        ctx = RunContext(self, self._bind(feed_dict))
        with torch.cuda.device(self.device):
            if isinstance(fetches, (list, tuple)):
                return [ctx.get(f) for f in fetches], ctx
            return ctx.get(fetches), ctx

    def run_device(self, fetches, feed_dict=None):
        """Asynchronous: torch CUDA tensors, no host synchronisation (and therefore no range-guard check: Encoder.check_range)."""
        return self._evaluate(fetches, feed_dict)[0]

    def run(self, fetches, feed_dict=None):
        out, ctx = self._evaluate(fetches, feed_dict)

        def host(v):
            if isinstance(v, torch.Tensor):
                return v.detach().cpu().numpy()
            return v

        res = [host(v) for v in out] if isinstance(fetches, (list, tuple)) else host(out)
        for m in ctx.touched:               # results are on the host now: a value outside the tensor-core range must not pass silently
            m.check_range(self.device)
        return res

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def to_device_input(x, device):
    """Feeds may be numpy (uint8 / float) or torch tensors already on the device.  Returns a contiguous CUDA tensor
    that is uint8 or float32; float64 numpy input (the reference's x/255.) is rounded to float32 as TF's feed does."""
    if isinstance(x, torch.Tensor):
        t = x
        if t.device != device:
            t = t.to(device, non_blocking=True)
    else:
        a = np.asarray(x)
        if a.dtype != np.uint8:
            a = a.astype(np.float32, copy=False)
        t = torch.from_numpy(np.ascontiguousarray(a)).to(device, non_blocking=True)
    if t.dtype != torch.uint8:
        t = t.to(torch.float32)
    return t.contiguous()

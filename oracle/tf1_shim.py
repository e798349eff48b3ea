"""A torch-backed, eager emulation of the TensorFlow-1.x API subset that the reference's BA layer uses.
TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Purpose: the reference (`/root/reference/bundlenet.py`, `legacy/ba.py`, `legacy/utils_python.py`) is TF-1.x graph-building
Python-2 code and TensorFlow is not installable here.  With this module registered as `tensorflow`, the reference's OWN
source files execute statement by statement on torch-CPU tensors (eagerly, float64), so that `tests/golden/gen_ref_golden.py`
can record what the reference code itself computes and `tests/test_oracle_pinned.py` can hold the oracle to it.

What is third-party restatement here (TensorFlow is not part of /root/reference; version unpinned, <= 1.15):
  * `tf.contrib.resampler.resampler`: bilinear, taps outside the map read 0, nothing sampled unless -1 < x < w and -1 < y < h;
  * `tf.matrix_solve` / `tf.linalg.solve`: LU with partial pivoting (torch.linalg.solve = LAPACK getrf/getrs);
  * `tf.qr(full_matrices=True)`: Householder QR (LAPACK geqrf); `tf.nn.selu` constants; `tf.nn.l2_normalize` epsilon 1e-12;
  * `tf.pad(mode='REFLECT')`, `tf.nn.conv1d` with a width-1 filter (= a dense layer), ordinary elementwise / shape ops.
The custom op library `utils.so` (`tf.load_op_library`) resolves to a literal replay of the op's cuBLAS call chain
(oracle/gemm_chain.py, itself checked on the GPU against the reference's compiled `utils.cu`, tests/test_gpu_reference_pin.py).

Everything the reference code itself states — formulas, operation order, stack axes (including the axis-0 stack of VMatrix,
bundlenet.py:45), slicing, damping, level schedule — is NOT restated: it runs from the reference's files.
"""
from __future__ import annotations

import contextlib
import sys
import types
from typing import Dict

import numpy as np
import torch

DEFAULT_DTYPE = torch.float64
VARIABLES: Dict[str, torch.Tensor] = {}          # name -> value (preset by the caller, else initialised on first use)
_RNG = torch.Generator().manual_seed(0)

SELU_ALPHA = 1.6732632423543772848170429916717
SELU_SCALE = 1.0507009873554804934193349852946


class TensorShape(tuple):
    def as_list(self):
        return list(self)


def _get_shape(self):
    return TensorShape(int(s) for s in self.shape)


torch.Tensor.get_shape = _get_shape              # the reference calls tensor.get_shape() everywhere


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    dt = like.dtype if isinstance(like, torch.Tensor) and like.dtype.is_floating_point else DEFAULT_DTYPE
    return torch.as_tensor(x, dtype=dt if isinstance(x, float) or (isinstance(x, (list, tuple)) and x and isinstance(x[0], float)) else None)


def _dtype(dt):
    if dt is None:
        return DEFAULT_DTYPE
    if dt in ("int32", "int64"):
        return torch.int64
    if dt is _tf.float32 or dt is _tf.float64:
        return DEFAULT_DTYPE
    if dt is _tf.int32 or dt is _tf.int64:
        return torch.int64
    return dt


def _axis_kw(kw):
    keep = kw.pop("keepdims", None)
    if keep is None:
        keep = kw.pop("keep_dims", False)
    return bool(keep)


# ------------------------------------------------------------------------------------------- module objects
_tf = types.ModuleType("tensorflow")
_tf.float32 = "tf.float32"; _tf.float64 = "tf.float64"; _tf.int32 = "tf.int32"; _tf.int64 = "tf.int64"
_tf.AUTO_REUSE = object()


@contextlib.contextmanager
def _scope(*a, **k):
    yield


_tf.name_scope = _scope
_tf.variable_scope = _scope
_tf.control_dependencies = _scope

# elementwise
_tf.sqrt = torch.sqrt; _tf.cos = torch.cos; _tf.sin = torch.sin; _tf.square = torch.square; _tf.abs = torch.abs
_tf.floor = torch.floor; _tf.identity = lambda x, name=None: x; _tf.stop_gradient = lambda x, name=None: x.detach()
_tf.multiply = lambda a, b, name=None: _t(a, b) * _t(b, a)
_tf.add = lambda a, b, name=None: _t(a, b) + _t(b, a)
_tf.div = lambda a, b, name=None: _t(a, b) / _t(b, a)
_tf.pow = lambda a, b, name=None: torch.pow(_t(a, b), _t(b, a))
_tf.maximum = lambda a, b, name=None: torch.maximum(_t(a, b), _t(b, a).to(_t(a, b).dtype))
_tf.add_n = lambda xs, name=None: sum(xs[1:], xs[0])
_tf.to_float = lambda x, name=None: _t(x).to(DEFAULT_DTYPE) if not isinstance(x, float) else torch.tensor(x, dtype=DEFAULT_DTYPE)
_tf.cast = lambda x, dtype=None, name=None: x.to(_dtype(dtype))
_tf.less = lambda a, b, name=None: torch.as_tensor(a) < torch.as_tensor(b)
_tf.equal = lambda a, b, name=None: torch.as_tensor(a) == torch.as_tensor(b)
_tf.logical_not = torch.logical_not
_tf.logical_and = torch.logical_and


def _clip_by_value(x, lo, hi, name=None):
    return torch.clamp(x, min=float(lo) if x.dtype.is_floating_point else int(lo), max=float(hi) if x.dtype.is_floating_point else int(hi))


_tf.clip_by_value = _clip_by_value


# shapes
def _shape_list(shape):
    return [int(s) for s in (shape.tolist() if isinstance(shape, torch.Tensor) else shape)]


_tf.zeros = lambda shape, dtype=None, name=None: torch.zeros(_shape_list(shape), dtype=_dtype(dtype))
_tf.ones = lambda shape, dtype=None, name=None: torch.ones(_shape_list(shape), dtype=_dtype(dtype))
_tf.range = lambda a, b=None, name=None: torch.arange(int(a), dtype=torch.int64) if b is None else torch.arange(int(a), int(b), dtype=torch.int64)
_tf.reshape = lambda x, shape, name=None: x.reshape(_shape_list(shape))
_tf.transpose = lambda x, perm=None, name=None: x.permute(*perm) if perm is not None else x.t()
_tf.expand_dims = lambda x, axis=None, name=None, dim=None: x.unsqueeze(axis if axis is not None else dim)
_tf.tile = lambda x, multiples, name=None: x.repeat(*_shape_list(multiples))
_tf.concat = lambda xs, axis, name=None: torch.cat(list(xs), dim=axis)
_tf.stack = lambda xs, axis=0, name=None: torch.stack([_t(x) for x in xs], dim=axis)
_tf.unstack = lambda x, num=None, axis=0, name=None: list(torch.unbind(x, dim=axis))
_tf.gather = lambda params, indices, name=None: params[indices]
_tf.meshgrid = lambda *a, **k: torch.meshgrid(*a, indexing=k.get("indexing", "xy"))


def _squeeze(x, axis=None, name=None, squeeze_dims=None):
    if axis is None:
        axis = squeeze_dims
    if axis is None:
        return x.squeeze()
    if isinstance(axis, (list, tuple)):
        for a in sorted([a % x.dim() for a in axis], reverse=True):
            x = x.squeeze(a)
        return x
    return x.squeeze(axis)


_tf.squeeze = _squeeze


def _split(value, num_or_size_splits, axis=0, num=None, name=None):
    if isinstance(num_or_size_splits, int):
        return list(torch.chunk(value, num_or_size_splits, dim=axis))
    return list(torch.split(value, list(num_or_size_splits), dim=axis))


_tf.split = _split


def _eye(num_rows, num_columns=None, batch_shape=None, dtype=None, name=None):
    e = torch.eye(num_rows, num_columns if num_columns is not None else num_rows, dtype=_dtype(dtype))
    if batch_shape:
        e = e.expand(*[int(b) for b in batch_shape], *e.shape).clone()
    return e


_tf.eye = _eye


def _pad(x, paddings, mode="CONSTANT", name=None):
    if mode.upper() != "REFLECT":
        raise NotImplementedError(mode)
    for d, (lo, hi) in enumerate(paddings):      # REFLECT: mirror without repeating the border element
        if lo == 0 and hi == 0:
            continue
        n = x.shape[d]
        idx = torch.tensor(list(range(lo, 0, -1)) + list(range(n)) + list(range(n - 2, n - 2 - hi, -1)), dtype=torch.int64)
        x = x.index_select(d, idx)
    return x


_tf.pad = _pad


# reductions
def _reduce(fn):
    def f(x, axis=None, name=None, **kw):
        keep = _axis_kw(kw)
        if isinstance(x, (list, tuple)):
            x = torch.stack([torch.as_tensor(v) for v in x])
        if axis is None:
            return fn(x)
        return fn(x, dim=axis, keepdim=keep)
    return f


_tf.reduce_sum = _reduce(torch.sum)
_tf.reduce_mean = _reduce(torch.mean)
_tf.reduce_any = _reduce(lambda x, **k: torch.any(x, **k))
_tf.reduce_all = _reduce(lambda x, **k: torch.all(x, **k))


def _norm(x, ord="euclidean", axis=None, name=None, **kw):
    keep = _axis_kw(kw)
    if axis is None:
        return torch.sqrt((x * x).sum())
    return torch.sqrt((x * x).sum(dim=axis, keepdim=keep))


_tf.norm = _norm


# linear algebra
def _matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return a @ b


_tf.matmul = _matmul
_tf.matrix_diag_part = lambda x, name=None: torch.diagonal(x, dim1=-2, dim2=-1)
_tf.matrix_diag = lambda x, name=None: torch.diag_embed(x)
_tf.matrix_solve = lambda a, b, adjoint=False, name=None: torch.linalg.solve(a, b)
_tf.matrix_inverse = lambda a, adjoint=False, name=None: torch.linalg.inv(a)
_tf.qr = lambda a, full_matrices=False, name=None: tuple(torch.linalg.qr(a, mode="complete" if full_matrices else "reduced"))
_tf.linalg = types.SimpleNamespace(solve=lambda a, b, adjoint=False, name=None: torch.linalg.solve(a, b))

# control flow (eager)
_tf.cond = lambda pred, true_fn, false_fn, name=None: true_fn() if bool(pred) else false_fn()


def _while_loop(cond, body, loop_vars, back_prop=True, parallel_iterations=10, name=None, **kw):
    v = list(loop_vars)
    while bool(cond(*v)):
        v = list(body(*v))
    return v


_tf.while_loop = _while_loop
_tf.Assert = lambda *a, **k: None


# variables
def _get_variable(name=None, shape=None, initializer=None, dtype=None, **kw):
    if name not in VARIABLES:
        VARIABLES[name] = initializer([int(s) for s in shape]) if initializer is not None else torch.zeros([int(s) for s in shape], dtype=DEFAULT_DTYPE)
    return VARIABLES[name]


_tf.get_variable = _get_variable
_tf.zeros_initializer = lambda: (lambda shape: torch.zeros(shape, dtype=DEFAULT_DTYPE))
_tf.keras = types.SimpleNamespace(initializers=types.SimpleNamespace(
    he_normal=lambda seed=None: (lambda shape: torch.randn(shape, generator=_RNG, dtype=DEFAULT_DTYPE) * (2.0 / shape[-2]) ** 0.5)))
_tf.placeholder = lambda dtype=None, shape=None, name=None: torch.zeros(_shape_list(shape), dtype=_dtype(dtype))


# nn
def _selu(x, name=None):
    return SELU_SCALE * torch.where(x > 0, x, SELU_ALPHA * (torch.exp(x) - 1.0))


def _conv1d(value, filters, stride, padding, name=None, **kw):
    if filters.shape[0] != 1 or stride != 1:
        raise NotImplementedError("only width-1, stride-1 conv1d (a dense layer) is used by the reference")
    return value @ filters[0]


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None, dim=None):
    axis = axis if axis is not None else dim
    return x * torch.rsqrt(torch.clamp((x * x).sum(dim=axis, keepdim=True), min=epsilon))


_tf.nn = types.SimpleNamespace(selu=_selu, tanh=lambda x, name=None: torch.tanh(x),
                               elu=lambda x, name=None: torch.where(x > 0, x, torch.exp(x) - 1.0),
                               conv1d=_conv1d, bias_add=lambda v, b, name=None: v + b, l2_normalize=_l2_normalize)


# tf.contrib.resampler.resampler (third party; see the module docstring)
def _resampler(data, warp, name=None):
    nb, h, w, C = data.shape
    x, y = warp[..., 0], warp[..., 1]
    fx, fy = torch.floor(x), torch.floor(y)
    cx, cy = fx + 1.0, fy + 1.0
    dx, dy = cx - x, cy - y
    inside = (x > -1.0) & (y > -1.0) & (x < float(w)) & (y < float(h))
    flat = data.reshape(nb, h * w, C)

    def point(px, py):
        ok = (px >= 0) & (py >= 0) & (px <= w - 1) & (py <= h - 1)
        idx = (py.clamp(0, h - 1).long() * w + px.clamp(0, w - 1).long()).unsqueeze(-1).expand(-1, -1, C)
        return torch.gather(flat, 1, idx) * ok.unsqueeze(-1).to(data.dtype)

    out = (point(fx, fy) * (dx * dy).unsqueeze(-1) + point(cx, cy) * ((1 - dx) * (1 - dy)).unsqueeze(-1)
           + point(fx, cy) * (dx * (1 - dy)).unsqueeze(-1) + point(cx, fy) * ((1 - dx) * dy).unsqueeze(-1))
    return out * inside.unsqueeze(-1).to(data.dtype)


_tf.contrib = types.SimpleNamespace(resampler=types.SimpleNamespace(resampler=_resampler))
def _cosine_distance(labels, predictions, axis=None, weights=1.0, scope=None, dim=None, **kw):
    """tf.losses.cosine_distance (TF 1.x): losses = 1 - sum(labels * predictions, axis, keepdims); default reduction SUM_BY_NONZERO_WEIGHTS = mean."""
    axis = axis if axis is not None else dim
    return (1.0 - (labels * predictions).sum(dim=axis, keepdim=True)).mean()


_tf.losses = types.SimpleNamespace(cosine_distance=_cosine_distance)
_tf.train = types.SimpleNamespace()


# the custom op library
class _OpLibrary:
    """`tf.load_op_library('./utils.so')`: EquationConstruction(+Grad) as the literal cuBLAS chain of utils.cu:331-414, 625-690."""

    @staticmethod
    def equation_construction(jacobian=None, gradient=None, difference=None, name=None):
        from . import gemm_chain
        J, G, d = (t.detach().cpu().double().numpy() for t in (jacobian, gradient, difference))
        lefts, rights = [], []
        for b in range(J.shape[0]):
            l, r = gemm_chain.equation_construction_chain(J[b], G[b], d[b])
            lefts.append(l); rights.append(r)
        return (torch.tensor(np.stack(lefts), dtype=jacobian.dtype), torch.tensor(np.stack(rights), dtype=jacobian.dtype))

    @staticmethod
    def equation_construction_grad(jacobian, gradient, difference, left_grad, right_grad, name=None):
        from . import gemm_chain
        J, G, d, lg, rg = (t.detach().cpu().double().numpy() for t in (jacobian, gradient, difference, left_grad, right_grad))
        outs = [gemm_chain.equation_construction_grad_chain(J[b], G[b], d[b], lg[b], rg[b]) for b in range(J.shape[0])]
        return tuple(torch.tensor(np.stack([o[i] for o in outs]), dtype=jacobian.dtype) for i in range(3))

    jacobian_construction = None                  # legacy/ba.py:13 binds it; nothing on the path calls it


_tf.load_op_library = lambda path: _OpLibrary()

# tensorflow.python.framework.{ops,dtypes}, tensorflow.python.ops.array_ops (imported by the reference, barely used)
_ops = types.ModuleType("tensorflow.python.framework.ops")
_ops.name_scope = _scope
_ops.RegisterGradient = lambda name: (lambda fn: fn)
_dtypes = types.ModuleType("tensorflow.python.framework.dtypes")
_array_ops = types.ModuleType("tensorflow.python.ops.array_ops")
_python = types.ModuleType("tensorflow.python"); _framework = types.ModuleType("tensorflow.python.framework")
_pyops = types.ModuleType("tensorflow.python.ops")
_python.framework = _framework; _python.ops = _pyops; _framework.ops = _ops; _framework.dtypes = _dtypes; _pyops.array_ops = _array_ops
_tf.python = _python

tf = _tf


def install() -> types.ModuleType:
    """Register the shim as `tensorflow` (refuses to shadow a real TensorFlow)."""
    if "tensorflow" in sys.modules and sys.modules["tensorflow"] is not _tf:
        raise RuntimeError("a real tensorflow is already imported; the shim is only for environments without it")
    sys.modules.update({"tensorflow": _tf, "tensorflow.python": _python, "tensorflow.python.framework": _framework,
                        "tensorflow.python.framework.ops": _ops, "tensorflow.python.framework.dtypes": _dtypes,
                        "tensorflow.python.ops": _pyops, "tensorflow.python.ops.array_ops": _array_ops})
    return _tf


# The only edits made to the reference's text before executing it: Python-2 constructs that do not parse / mean something else on
# Python 3.  Each is (old, new); the loader asserts that every `old` occurs, so that a changed reference is noticed.
PY2_FIXES = {
    "bundlenet.py": [('print "lambda_shape",lambda_prediction.get_shape()', 'pass  # (py2 print statement)'),
                     ("[nbatch/2:nbatch,:,:,:]", "[nbatch//2:nbatch,:,:,:]"), ("[0:nbatch/2,:,:,:]", "[0:nbatch//2,:,:,:]")],
    "legacy/ba.py": [],
    "legacy/utils_python.py": [],
}


def load_reference(ref_root: str, rel_path: str, module_name: str, extra_modules=None) -> types.ModuleType:
    """Execute a reference source file (read from `ref_root`, never copied into the repo) as a module under the shim."""
    import os
    install()
    src = open(os.path.join(ref_root, rel_path)).read()
    for old, new in PY2_FIXES.get(rel_path, []):
        if old not in src:
            raise RuntimeError(f"{rel_path}: expected text not found (reference changed?): {old[:40]!r}")
        src = src.replace(old, new)
    mod = types.ModuleType(module_name)
    mod.__file__ = os.path.join(ref_root, rel_path)
    for k, v in (extra_modules or {}).items():
        sys.modules[k] = v
    sys.modules[module_name] = mod
    exec(compile(src, mod.__file__, "exec"), mod.__dict__)
    return mod

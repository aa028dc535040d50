"""A minimal stand-in for `jax` so that the UNMODIFIED reference file sofacontrol/SSM/ssm.py can be imported and run
in the build container.  TEST INFRASTRUCTURE ONLY (oracle/__init__.py) -- never imported by the product package.

The reference uses exactly four pieces of JAX (ssm.py:6-8, 164-178, 198-235, 281-296):
  * `jax.numpy` as an array namespace (`dot`, `asarray`, `eye`, `linalg.inv`, `ndarray`)   -> numpy, float64
  * `jax.scipy.special` as a lambdify namespace (no function of it is ever called)        -> empty module
  * `jax.jit(fun, static_argnums=...)`                                                      -> identity
  * `jax.jacobian(fun, argnums)`                                                            -> EXACT forward-mode
    differentiation with dual numbers: every intermediate carries (value, tangent block); `+ - * / **` and `dot`
    propagate tangents by the chain rule, so the result is the analytic Jacobian evaluated in float64 (no finite
    differences, no truncation error) -- the same quantity XLA's autodiff returns, up to summation order.

Precision: JAX defaults to float32 unless `jax_enable_x64` is set, and the reference never sets it (SURVEY.md
section 7).  BASELINE.json asks for FP64 parity, so the shim computes in float64 -- it plays the role of
`jax.config.update("jax_enable_x64", True)`.

`install()` registers the shim as `jax`, `jax.numpy`, `jax.scipy`, `jax.scipy.special` in `sys.modules` (only if no
real jax is importable); `oracle.refimport.load_ssm()` calls it and then imports the reference module.
"""
import sys
import types

import numpy as np


class Dual:
    """value `v` (ndarray, any shape) + tangent `t` (shape v.shape + (nt,)): d value / d seed directions."""
    __array_priority__ = 1000          # ndarray <op> Dual defers to Dual.__r<op>__

    def __init__(self, v, t):
        self.v = np.asarray(v, dtype=np.float64)
        self.t = np.asarray(t, dtype=np.float64)

    # -- container protocol: `self.rom_phi(*x)` unpacks the state into scalars (ssm.py:168)
    @property
    def shape(self):
        return self.v.shape

    @property
    def ndim(self):
        return self.v.ndim

    def __len__(self):
        return self.v.shape[0]

    def __iter__(self):
        for i in range(self.v.shape[0]):
            yield Dual(self.v[i], self.t[i])

    def __getitem__(self, i):
        return Dual(self.v[i], self.t[i])

    @staticmethod
    def lift(x, nt):
        if isinstance(x, Dual):
            return x
        x = np.asarray(x, dtype=np.float64)
        return Dual(x, np.zeros(x.shape + (nt,)))

    def _nt(self):
        return self.t.shape[-1]

    # -- arithmetic
    def __neg__(self):
        return Dual(-self.v, -self.t)

    def __add__(self, o):
        o = Dual.lift(o, self._nt())
        return Dual(self.v + o.v, self.t + o.t)
    __radd__ = __add__

    def __sub__(self, o):
        o = Dual.lift(o, self._nt())
        return Dual(self.v - o.v, self.t - o.t)

    def __rsub__(self, o):
        return Dual.lift(o, self._nt()) - self

    def __mul__(self, o):
        o = Dual.lift(o, self._nt())
        return Dual(self.v * o.v, self.t * o.v[..., None] + o.t * self.v[..., None])
    __rmul__ = __mul__

    def __truediv__(self, o):
        o = Dual.lift(o, self._nt())
        q = self.v / o.v
        return Dual(q, (self.t - o.t * q[..., None]) / o.v[..., None])

    def __rtruediv__(self, o):
        return Dual.lift(o, self._nt()) / self

    def __pow__(self, p):
        if isinstance(p, Dual):
            raise TypeError("Dual ** Dual is not needed by ssm.py")
        p = float(p)
        return Dual(self.v ** p, (p * self.v ** (p - 1.0))[..., None] * self.t)


def _stack(items):
    nt = next(i._nt() for i in items if isinstance(i, Dual))
    items = [Dual.lift(i, nt) for i in items]
    return Dual(np.stack([i.v for i in items]), np.stack([i.t for i in items]))


def _asarray(a, dtype=None):
    if isinstance(a, Dual):
        return a
    if isinstance(a, (list, tuple)) and any(isinstance(i, Dual) for i in a):
        return _stack(list(a))
    return np.asarray(a, dtype=np.float64 if dtype is None else dtype)


def _dot(a, b):
    """jnp.dot for the shapes ssm.py uses: (r, k) . (k,) and (r, k) . (k, c), either side possibly Dual."""
    da, db = isinstance(a, Dual), isinstance(b, Dual)
    if not da and not db:
        return np.dot(a, b)
    if not da:
        a = np.asarray(a, dtype=np.float64)
        # tangent of a . b = a . t_b, contracted over b's first axis (b is (k,) or (k, c))
        return Dual(np.dot(a, b.v), np.tensordot(a, b.t, axes=([a.ndim - 1], [0])))
    if not db:
        b = np.asarray(b, dtype=np.float64)
        tv = np.moveaxis(np.tensordot(a.t, b, axes=([a.v.ndim - 1], [0])), a.v.ndim - 1, -1)
        return Dual(np.dot(a.v, b), tv)
    return _dot(a, b.v) + _dot(a.v, b)


def _jacobian(fun, argnums=0):
    """Exact forward-mode Jacobian of `fun` w.r.t. the positional argument(s) `argnums` (int or tuple)."""
    single = isinstance(argnums, int)
    nums = (argnums,) if single else tuple(argnums)

    def jac(*args):
        out = []
        for a in nums:
            x = np.asarray(args[a], dtype=np.float64)
            seeded = list(args)
            seeded[a] = Dual(x, np.eye(x.size).reshape(x.shape + (x.size,)))
            y = fun(*seeded)
            if not isinstance(y, Dual):                     # the function does not depend on this argument
                y = Dual.lift(y, x.size)
            out.append(np.ascontiguousarray(y.t.reshape(y.v.shape + x.shape)))     # row-major like a jax array
        return out[0] if single else tuple(out)
    return jac


def _jit(fun=None, **kwargs):
    if fun is None:
        return lambda f: f
    return fun


def build_modules():
    jax = types.ModuleType("jax")
    jnp = types.ModuleType("jax.numpy")
    jsp = types.ModuleType("jax.scipy")
    special = types.ModuleType("jax.scipy.special")
    linalg = types.ModuleType("jax.numpy.linalg")
    linalg.inv = np.linalg.inv
    jnp.linalg = linalg
    jnp.dot = _dot
    jnp.asarray = _asarray
    jnp.array = _asarray
    jnp.eye = np.eye
    jnp.zeros = np.zeros
    jnp.ndarray = np.ndarray
    jnp.float64 = np.float64
    jsp.special = special
    jax.numpy = jnp
    jax.scipy = jsp
    jax.jit = _jit
    jax.jacobian = _jacobian
    jax.jacfwd = _jacobian
    jax.__srcb200_shim__ = True
    return {"jax": jax, "jax.numpy": jnp, "jax.numpy.linalg": linalg, "jax.scipy": jsp, "jax.scipy.special": special}


def install():
    """Registers the shim unless a real jax is importable.  Returns True when the shim is the active `jax`."""
    if "jax" in sys.modules:
        return bool(getattr(sys.modules["jax"], "__srcb200_shim__", False))
    try:
        import jax  # noqa: F401
        return False
    except ImportError:
        pass
    sys.modules.update(build_modules())
    return True

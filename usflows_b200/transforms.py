"""Transform layers of the flow stack -- drop-in mirrors of the reference's `src/usflows/transforms.py`.

Same class names, constructor arguments, parameter names/shapes (state-dict compatible) and method surface
(`forward`, `backward`, `log_abs_det_jacobian`, `is_feasible`, `add_jitter`, `log_prior`, `sign`, `simplify`,
`matrix`/`bias`/`inverse_matrix`).  The arithmetic runs in the sm_100a kernels of libusflows_b200.so:
weight-side work (L@U, triangular inverses, Householder products, log-dets) is computed once per weight
version by the preparation kernels and cached; batch-side work goes through `engine.run_layers`.
There is no CPU path: calling a layer with a non-CUDA tensor raises.

Direction convention (as in the reference): `forward` = sampling direction (latent -> data),
`backward` = density direction (data -> latent).
"""
from __future__ import annotations

import math
from typing import Iterable, List, Optional

import torch
from torch import nn
from torch.nn import init

from . import ops


def _versions(params) -> tuple:
    return tuple((p.data_ptr(), p._version) for p in params)


def _mm64(a: torch.Tensor, b: torch.Tensor, tri: int = 0) -> torch.Tensor:
    out = torch.empty(a.shape[0], b.shape[1], dtype=torch.float64, device=a.device)
    ops.matmul_f64(a.contiguous(), b.contiguous(), out, tri)
    return out


class BaseTransform(nn.Module):
    """Contract of every layer (reference transforms.py:23-69)."""

    bijective = True

    def is_feasible(self) -> bool:
        return True

    def add_jitter(self, jitter: float = 1e-6) -> None:
        pass

    def jitter(self, jitter: float = 1e-6) -> None:  # the reference's abstract name (transforms.py:37-39)
        self.add_jitter(jitter)

    def forward(self, x: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        from . import engine
        return engine.run_layers([self], "forward", x)

    def backward(self, y: torch.Tensor, context: Optional[torch.Tensor] = None) -> torch.Tensor:
        from . import engine
        return engine.run_layers([self], "backward", y)

    def _call(self, x):
        return self.forward(x)

    def _inverse(self, y):
        return self.backward(y)

    def log_abs_det_jacobian(self, x, y, context=None):
        raise NotImplementedError

    def log_prior(self):
        """Uniform (pseudo-)prior (transforms.py:62-64)."""
        return 0.0

    def simplify(self):
        return self

    def sign(self):
        return 1

    # the pyro / torch `Transform` surface the reference's layers inherit (transforms.py:195-202, 247-251): every layer
    # maps real vectors (one event dim) to real vectors and keeps no value cache
    @property
    def domain(self):
        return torch.distributions.constraints.independent(torch.distributions.constraints.real, 1)

    @property
    def codomain(self):
        return torch.distributions.constraints.independent(torch.distributions.constraints.real, 1)

    def with_cache(self, cache_size: int = 1):
        return self

    # engine hooks -------------------------------------------------------------------------------
    def _ladj_device(self) -> Optional[torch.Tensor]:
        """Device tensor [2] = (forward log|det J|, #infeasible entries), or None when identically 0."""
        return None


# ------------------------------------------------------------------------------------------------
class ScaleTransform(BaseTransform):
    """y = scale * x  (transforms.py:73-171)."""

    def __init__(self, in_dims, prior_scale: float = 1.0):
        super().__init__()
        self.in_dims = in_dims
        self.prior_scale = prior_scale
        self.dim = math.prod(in_dims) if isinstance(in_dims, Iterable) else in_dims
        self.scale = nn.Parameter(torch.empty(in_dims))
        self.init_params()

    def init_params(self):
        bound = 1 / math.sqrt(self.dim) if self.dim > 0 else 0
        init.uniform_(self.scale, -bound, bound)            # transforms.py:99-103

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return self._ladj_device()[0]                          # sum log|scale| (:135-144)

    def _ladj_device(self):
        key = _versions([self.scale])
        if getattr(self, "_ladj_key", None) != key:
            ops.require_cuda(self.scale, "ScaleTransform.scale")
            out = torch.empty(2, dtype=torch.float32, device=self.scale.device)
            with ops.on_device(out):
                ops.vec_logabs(self.scale.detach().reshape(-1), out)
            self._ladj_cache, self._ladj_key = out, key
        return self._ladj_cache

    def sign(self) -> int:
        return 1 if int((self.scale < 0).sum()) % 2 == 0 else -1

    def is_feasible(self) -> bool:
        return bool((self.scale != 0).all())

    def add_jitter(self, jitter: float = 1e-6) -> None:
        # the reference's version references a non-existent attribute (transforms.py:154-157); fixed here
        with torch.no_grad():
            self.scale.add_(torch.randn_like(self.scale) * jitter)

    def log_prior(self):
        return 0


class Permute(BaseTransform):
    """y = x[..., permutation]  (transforms.py:174-251)."""

    volume_preserving = True

    def __init__(self, permutation: torch.Tensor, *, dim: int = -1, cache_size: int = 1):
        super().__init__()
        if dim >= 0:
            raise ValueError("'dim' keyword argument must be negative")
        self.permutation = permutation
        self.dim = dim

    @property
    def inv_permutation(self):
        result = torch.empty_like(self.permutation, dtype=torch.long)
        result[self.permutation] = torch.arange(self.permutation.size(0), dtype=torch.long,
                                                device=self.permutation.device)
        return result

    def log_abs_det_jacobian(self, x, y, context=None):
        return torch.zeros(x.size()[:-1], dtype=x.dtype, device=x.device)

    def to(self, device):
        self.permutation = self.permutation.to(device)
        return super().to(device)


class LeakyReLUTransform(BaseTransform):
    """y = leaky_relu(x, alpha)  (transforms.py:417-474).

    `log_abs_det_jacobian` is per row: log(alpha) * #{x_j < 0}.  The reference sums log(y/x) over the whole
    tensor including the batch axis (NaN at x = 0, transforms.py:474); on an unbatched vector -- the only
    case the reference tests -- both agree.
    """

    def __init__(self, alpha: float = 0.01):
        if alpha == 0:
            raise ValueError("alpha must be positive")
        super().__init__()
        self.alpha = alpha

    def log_abs_det_jacobian(self, x, y, context=None):
        from . import engine
        return engine.leaky_relu_ladj(x, self.alpha)


# ------------------------------------------------------------------------------------------------
class AffineTransform(BaseTransform):
    """y = A x + b with getters for A, b, A^-1 (transforms.py:697-750).  Subclasses fill `_prepare`."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim
        self.input_shape = dim

    def _prep_params(self) -> List[torch.Tensor]:
        return list(self.parameters())

    def _prepared(self) -> dict:
        """dict(matrix [d,d], inverse_matrix [d,d], bias [d], ladj [2]) for the current weight version."""
        key = _versions(self._prep_params())
        if getattr(self, "_prep_key", None) != key:
            for p in self._prep_params():
                ops.require_cuda(p, f"{type(self).__name__} parameter")
            with torch.no_grad(), ops.on_device(self._prep_params()[0]):
                self._prep_cache = self._prepare()
            self._prep_key = key
        return self._prep_cache

    def _prepare(self) -> dict:
        raise NotImplementedError

    def matrix(self) -> torch.Tensor:
        return self._prepared()["matrix"]

    def inverse_matrix(self) -> torch.Tensor:
        return self._prepared()["inverse_matrix"]

    def bias(self) -> torch.Tensor:
        return self._prepared()["bias"]

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return self._prepared()["ladj"][0]

    def _ladj_device(self):
        return self._prepared()["ladj"]

    def _to_plane_linear(self) -> "PlaneBijectiveLinearTransform":
        """The same map as a plain (W, b, W^-1) layer (transforms.py:732-747).  The prepared log|det| of this layer
        travels with it instead of being re-derived by a dense `slogdet` of W."""
        p = self._prepared()
        bias = p["bias"].clone()
        if isinstance(self, LUTransform):                  # its bias() IS the `bias_vector` parameter (:1295-1297)
            bias = nn.Parameter(bias)
        return PlaneBijectiveLinearTransform(self.dim, p["matrix"].clone(), bias, p["inverse_matrix"].clone(),
                                             ladj=p["ladj"][0].clone())

    def simplify(self):
        return self._to_plane_linear()                                         # transforms.py:749-750


class PlaneBijectiveLinearTransform(AffineTransform):
    """y = W x + b with W, b and W^-1 given as plain tensors (transforms.py:618-695): what `simplify()` turns every LU /
    Householder / sequential affine layer into for verification back ends.  Holds the reference's modules (`forth`, `back`:
    state-dict keys `forth.weight`, `forth.bias`, `back.weight`, `back.bias`); not meant to be trained, bijectivity is
    not enforced.  `log|det W|` is a construction-time constant as in the reference (`slogdet` there, `:653-654`); pass
    `ladj` when it is known (the LU layers know theirs exactly).  The reference requires `m_inv`; here a missing one is
    computed once at construction."""

    volume_preserving = False

    def __init__(self, dim: int, m: torch.Tensor, bias: torch.Tensor, m_inv: Optional[torch.Tensor] = None, *,
                 ladj: Optional[torch.Tensor] = None):
        super().__init__(dim)
        bias_is_parameter = isinstance(bias, nn.Parameter)
        m, bias = m.detach(), bias.detach()
        if tuple(m.shape) != (dim, dim) or tuple(bias.shape) != (dim,):
            raise ValueError("m must be [dim, dim] and bias [dim]")
        with torch.no_grad():
            m_inv = torch.linalg.inv(m.double()).to(m.dtype) if m_inv is None else m_inv.detach()
            if ladj is None:
                ladj = torch.linalg.slogdet(m.double())[1].to(m.dtype)
            back_bias = -(m_inv.double() @ bias.double()).to(m.dtype)          # :649
        # `self.bias_vector = bias` (:639): a parameter (and a state-dict key) exactly when the caller hands one in, which
        # the reference's LUTransform.bias() does and its SequentialAffineTransform.bias() does not
        self.bias_vector = nn.Parameter(bias) if bias_is_parameter else bias
        self.forth = nn.Linear(dim, dim, bias=True)
        self.forth.weight = nn.Parameter(m)
        self.forth.bias = nn.Parameter(bias)
        self.back = nn.Linear(dim, dim, bias=True)
        self.back.weight = nn.Parameter(m_inv)
        self.back.bias = nn.Parameter(back_bias)
        self.register_buffer("ladj", torch.as_tensor(ladj, dtype=m.dtype).detach().reshape(()).to(m.device),
                             persistent=False)

    @property
    def m_inv(self) -> torch.Tensor:
        return self.back.weight

    def _prep_params(self) -> List[torch.Tensor]:
        return [self.forth.weight, self.forth.bias, self.back.weight]

    def _prepare(self) -> dict:
        W, Winv = self.forth.weight.detach().contiguous(), self.back.weight.detach().contiguous()
        ladj = torch.stack([self.ladj.to(W.device, torch.float32), torch.zeros((), dtype=torch.float32, device=W.device)])
        return dict(matrix=W, inverse_matrix=Winv, bias=self.forth.bias.detach(), ladj=ladj,
                    matrix64=W.double(), inverse64=Winv.double())

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return self.ladj

    def matrix(self) -> torch.Tensor:
        return self.forth.weight

    def inverse_matrix(self) -> torch.Tensor:
        return self.back.weight

    def bias(self) -> torch.Tensor:
        return self.forth.bias

    def simplify(self):
        return self


class LUTransform(AffineTransform):
    """y = (L U) x + b with L unit lower, U upper triangular (transforms.py:1178-1379)."""

    volume_preserving = False

    def __init__(self, dim: int, prior_scale: float = 1.0):
        super().__init__(dim)
        self.L_raw = nn.Parameter(torch.empty(dim, dim))
        self.U_raw = nn.Parameter(torch.empty(dim, dim))
        self.bias_vector = nn.Parameter(torch.empty(dim))
        self.prior_scale = prior_scale
        self.init_params()
        self.L_mask = torch.tril(torch.ones(dim, dim), diagonal=-1)
        self.U_mask = torch.triu(torch.ones(dim, dim), diagonal=0)
        self.L_raw.register_hook(lambda grad: grad * self.L_mask.to(grad.device))   # transforms.py:1212-1213
        self.U_raw.register_hook(lambda grad: grad * self.U_mask.to(grad.device))

    def init_params(self):
        """Same init distributions as transforms.py:1215-1240."""
        init.kaiming_uniform_(self.L_raw, nonlinearity="relu")
        with torch.no_grad():
            self.L_raw.copy_(self.L_raw.tril(diagonal=-1).fill_diagonal_(1))
        init.kaiming_uniform_(self.U_raw, nonlinearity="relu")
        with torch.no_grad():
            self.U_raw.fill_diagonal_(0)
            d = self.dim
            sign = -torch.ones(d) + 2 * torch.bernoulli(0.5 * torch.ones(d))
            scale = self.prior_scale * torch.ones(d) * 1 / d if self.prior_scale is not None else torch.ones(d)
            self.U_raw += sign * torch.normal(torch.zeros(d), scale).exp().diag()
            self.U_raw.copy_(self.U_raw.triu())
        bound = 1 / math.sqrt(self.dim) if self.dim > 0 else 0
        init.uniform_(self.bias_vector, -bound, bound)

    def _prepare(self) -> dict:
        d, dev = self.dim, self.L_raw.device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
        T, X, tmp = new(2, d, d), new(2, d, d), new(2, d, d)
        L, U = T[0], new(d, d)
        ops.lu_assemble(self.L_raw.detach(), self.U_raw.detach(), L, T[1], transpose_u=True)   # L, U^T (:1271-1279)
        ops.transpose(T[1], U)
        # inverse(L), inverse(U) (:1291-1292): both lower-triangular inverses (L and U^T) in one batched recursive-doubling
        # pass (csrc/train.cuh; the per-matrix panel sweep took 0.5 ms at d = 784 and ~10 ms at d = 3072 per inverse)
        ops.tri_inverse_batched(T, X, tmp, 0b01)
        Linv, Uinv = X[0], new(d, d)
        ops.transpose(X[1], Uinv)
        ladj = new(2)
        ops.lu_logabsdet(self.U_raw.detach(), ladj)                            # sum log|diag U| (:1303-1320)
        # matrix = L @ U (:1281-1283), inverse = U^-1 @ L^-1 (:1293) as fp64 tensor-core products that skip the zero halves
        # of the factors: what the engine composes neighbouring affine layers from.  The fp32 `matrix` / `inverse_matrix`
        # are these rounded once (the fp32 CUDA-core products they used to be cost as much as everything else in the
        # preparation at d = 3072, and carried the rounding of a K-term fp32 sum).
        W64 = _mm64(L.double(), U.double(), ops.TRI_LOWER_UPPER)
        Winv64 = _mm64(Uinv.double(), Linv.double(), ops.TRI_UPPER_LOWER)
        W, Winv = W64.float(), Winv64.float()
        return dict(matrix=W, inverse_matrix=Winv, bias=self.bias_vector.detach(), ladj=ladj,
                    L=L, U=U, L_inv=Linv, U_inv=Uinv, matrix64=W64, inverse64=Winv64)

    def to_linear(self) -> "PlaneBijectiveLinearTransform":
        """transforms.py:1365-1369 as written there: `M_inv = L @ U`, `M = inverse(M_inv)`,
        `PlaneBijectiveLinearTransform(dim, M, bias_vector, M_inv)` -- the plain layer whose `forth` weight is (L U)^-1
        (`simplify()` / `_to_plane_linear` above is the one whose forward equals this layer's)."""
        p = self._prepared()
        return PlaneBijectiveLinearTransform(self.dim, p["inverse_matrix"].clone(), self.bias_vector,
                                             p["matrix"].clone(), ladj=-p["ladj"][0].clone())

    @property
    def L(self) -> torch.Tensor:
        return self._prepared()["L"]

    @property
    def U(self) -> torch.Tensor:
        return self._prepared()["U"]

    def sign(self):
        return self.U_raw.diag().prod().sign()

    def to(self, device):
        self.L_mask = self.L_mask.to(device)
        self.U_mask = self.U_mask.to(device)
        self.device = device
        return super().to(device)

    def is_feasible(self) -> bool:
        return bool((self.U_raw.diag() != 0).all())

    def add_jitter(self, jitter: float = 1e-6) -> None:
        with torch.no_grad():
            self.U_raw.diagonal().add_(torch.randn(self.dim, device=self.U_raw.device) * jitter)

    def log_prior(self):
        x = self.U_raw.diag().abs().log()
        return -(x * x).sum() / (2 * self.prior_scale ** 2) - x.sum()


class BlockLUTransform(LUTransform):
    """An LU layer over the leading axis of `in_dims`, applied to every block of the remaining axes (a 1x1 convolution for
    `[C, H, W]`): transforms.py:1488-1622.  The same map as `BlockAffineTransform(in_dims, LUTransform(in_dims[0]))` with
    the parameters `L_raw`, `U_raw`, `bias_vector` at the top level; log|det| = sum log|U_kk| x prod(in_dims[1:]).
    (`sign()` raises in the reference -- it reads a `block_transform` the class does not have; here it is the LU sign to
    the power of the block count.)"""

    def __init__(self, in_dims: Iterable[int], prior_scale: float = 1.0):
        self.in_dims = list(in_dims)
        self.block_size = self.in_dims[0]
        self.input_rank = len(self.in_dims) - 1
        self.n_blocks = math.prod(self.in_dims[1:])
        if self.input_rank not in (0, 2):
            raise NotImplementedError("usflows_b200: in_dims=[d] (flat) and [C, H, W] (1x1 convolution) are built")
        super().__init__(self.block_size, prior_scale)

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return super().log_abs_det_jacobian(x, y, context) * self.n_blocks

    def _ladj_device(self):
        ladj = self._prepared()["ladj"]
        return ladj * torch.tensor([float(self.n_blocks), 1.0], device=ladj.device)

    def sign(self):
        return super().sign() ** self.n_blocks

    def log_prior(self, correlated: bool = False):
        """Log-normal prior on |diag U| (transforms.py:1565-1591): precision of the (optionally negatively correlated)
        covariance `prior_scale^2 / d * I`."""
        d = self.block_size
        x = self.U_raw.diag().abs().log()
        if correlated:
            cov = (-1 / d * torch.ones(d, d, device=x.device) + (1 + 1 / d) * torch.eye(d, device=x.device)) * self.prior_scale ** 2
        else:
            cov = torch.eye(d, device=x.device) * (self.prior_scale ** 2 / d)
        return -(x * (torch.linalg.inv(cov) @ x)).sum() - x.sum()

    def simplify(self):
        """Plain-matrix form, as `BlockAffineTransform.simplify` (the reference inherits `AffineTransform.simplify`, which
        drops the block structure: transforms.py:749-750)."""
        return BlockAffineTransform(self.in_dims, LUTransform._to_plane_linear(self)).simplify() \
            if self.input_rank == 2 else BlockAffineTransform(self.in_dims, LUTransform._to_plane_linear(self))


class Rotation(AffineTransform):
    """Rotation by `angle` in the coordinate plane `plane` of R^dim (transforms.py:476-556); log|det| = 0.  `forward` is
    the reference's map.  `backward` is its INVERSE: the reference's backward overwrites y[plane[0]] and then uses the
    overwritten value for y[plane[1]] (transforms.py:521-524), which is not the inverse rotation (round-trip error of
    order sin(angle)) -- a deviation of the same kind as SURVEY Q1."""

    ladj = 0

    def __init__(self, dim: int, plane, angle: float):
        if dim < 2:
            raise ValueError("dim must be at least 2")
        if plane[0] == plane[1]:
            raise ValueError("plane must be a tuple of different indices")
        if dim <= max(plane):
            raise ValueError("plane indices must be smaller than dim")
        super().__init__(dim)
        self.plane = tuple(plane)
        self.angle = angle
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)      # follows .to(device): where the map lives

    def as_matrix(self) -> torch.Tensor:
        R = torch.eye(self.dim)
        i, j = self.plane
        c, s = math.cos(self.angle), math.sin(self.angle)
        R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
        return R

    def _matrix64(self) -> torch.Tensor:
        R = torch.eye(self.dim, dtype=torch.float64)
        i, j = self.plane
        c, s = math.cos(self.angle), math.sin(self.angle)
        R[i, i], R[i, j], R[j, i], R[j, j] = c, -s, s, c
        return R

    def _prep_params(self) -> List[torch.Tensor]:
        return [self._anchor]

    def _prepare(self) -> dict:
        dev = self._anchor.device
        R = self._matrix64().to(dev)
        return dict(matrix=R.float(), inverse_matrix=R.t().contiguous().float(),
                    bias=torch.zeros(self.dim, dtype=torch.float32, device=dev),
                    ladj=torch.zeros(2, dtype=torch.float32, device=dev), matrix64=R, inverse64=R.t().contiguous())

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return self.ladj

    def sign(self):
        return 1

    def simplify(self):
        return self


class CompositeRotation(Rotation):
    """Rotations applied one after the other, `y = R_n ... R_1 x` (transforms.py:558-616); `backward` is the inverse (see
    `Rotation`).  `as_matrix()` returns the reference's product `R_1 R_2 ... R_n` literally -- which is the matrix of the
    rotations applied in the OPPOSITE order (transforms.py:580-584), not of `forward`; `matrix()` is the map `forward`
    applies."""

    def __init__(self, rotations: List[Rotation]):
        rotations = list(rotations)
        AffineTransform.__init__(self, rotations[0].dim)
        self.rotations = rotations
        self.input_shape = rotations[0].dim
        self.register_buffer("_anchor", torch.zeros(1), persistent=False)

    def as_matrix(self) -> torch.Tensor:
        R = torch.eye(self.input_shape)
        for rot in self.rotations:
            R = torch.matmul(R, rot.as_matrix())
        return R

    def _matrix64(self) -> torch.Tensor:
        R = torch.eye(self.dim, dtype=torch.float64)
        for rot in self.rotations:                     # forward applies rot_1 first: x -> R_n ... R_1 x
            R = rot._matrix64() @ R
        return R


class HouseholderTransform(AffineTransform):
    """y = H x, H = w_0 prod_k (I - 2 v_k v_k^T / v_k.v_k)  (transforms.py:752-872); log|det| = 0."""

    ladj = 0

    def __init__(self, dim: int, nvs: int = 1, device="cpu"):
        super().__init__(dim)
        self.nvs = nvs
        indices = torch.randperm(dim)
        w = torch.zeros((dim, dim))
        w[torch.arange(dim), indices] = 1.0
        self.vk_householder = nn.Parameter(0.2 * torch.randn(nvs, dim))
        self.w_0 = nn.Parameter(w, requires_grad=False)
        self.to(device)

    def _construct_householder_permutation(self) -> torch.Tensor:
        """w_0 prod_k (I - 2 v_k v_k^T / v_k.v_k)  (transforms.py:795-809): the prepared matrix."""
        return self.matrix()

    def _prepare(self) -> dict:
        d, dev = self.dim, self.w_0.device
        W = self.w_0.detach().clone()
        work = torch.empty(d, dtype=torch.float32, device=dev)
        for k in range(self.nvs):
            ops.householder_right(W, self.vk_householder.detach()[k].contiguous(), work)  # :795-809
        Wt = torch.empty(d, d, dtype=torch.float32, device=dev)
        ops.transpose(W, Wt)                                                               # :864-868
        return dict(matrix=W, inverse_matrix=Wt, bias=torch.zeros(d, dtype=torch.float32, device=dev),
                    ladj=torch.zeros(2, dtype=torch.float32, device=dev), matrix64=W.double(), inverse64=Wt.double())


class SequentialAffineTransform(AffineTransform):
    """Composition of affine maps as one matrix + bias (transforms.py:1381-1486)."""

    def __init__(self, transforms: Iterable[AffineTransform]):
        transforms = list(transforms)
        dim = transforms[0].dim
        if any(t.dim != dim for t in transforms):
            raise ValueError("All transforms must have the same dimension")
        super().__init__(dim)
        self.transforms = nn.ModuleList(transforms)

    def _prepare(self) -> dict:
        parts = [t._prepared() for t in self.transforms]
        d, dev = self.dim, parts[0]["matrix"].device
        new = lambda *s: torch.empty(*s, dtype=torch.float32, device=dev)  # noqa: E731
        # matrix: I @ A_1 @ A_2 ... (:1457-1462); inverse: I @ A_n^-1 @ ... @ A_1^-1 (:1464-1469)
        M = parts[0]["matrix"]
        for p in parts[1:]:
            out = new(d, d)
            ops.matmul_f32(M, p["matrix"], out)
            M = out
        Minv = parts[-1]["inverse_matrix"]
        for p in parts[-2::-1]:
            out = new(d, d)
            ops.matmul_f32(Minv, p["inverse_matrix"], out)
            Minv = out
        # bias: b = 0; b = b @ A_k + b_k (:1471-1476)
        b = parts[0]["bias"]
        for p in parts[1:]:
            out = new(1, d)
            ops.matmul_f32(b.reshape(1, d), p["matrix"], out, bias=p["bias"])
            b = out.reshape(d)
        ladj = torch.stack([p["ladj"] for p in parts]).sum(0)                  # :1429-1446
        M64, Minv64 = parts[0]["matrix64"], parts[-1]["inverse64"]
        for p in parts[1:]:
            M64 = _mm64(M64, p["matrix64"])
        for p in parts[-2::-1]:
            Minv64 = _mm64(Minv64, p["inverse64"])
        return dict(matrix=M, inverse_matrix=Minv, bias=b, ladj=ladj, matrix64=M64, inverse64=Minv64)

    def is_feasible(self) -> bool:
        return all(t.is_feasible() for t in self.transforms)

    def add_jitter(self, jitter: float = 1e-6) -> None:
        for t in self.transforms:
            t.add_jitter(jitter)

    def sign(self):
        return math.prod([t.sign() for t in self.transforms])

    def to(self, device):
        for t in self.transforms:
            t.to(device)
        self.device = device
        return super().to(device)


class BlockAffineTransform(BaseTransform):
    """Applies an AffineTransform over `in_dims[0]` (transforms.py:874-1029): `F.linear` for flat `in_dims=[d]`, a 1x1
    convolution over the channels for `in_dims=[C, H, W]` (run as the same contraction over channels-last rows, see
    image_engine.py); log|det| = inner x prod(in_dims[1:])."""

    def __init__(self, in_dims: Iterable[int], block_transform: AffineTransform):
        super().__init__()
        self.in_dims = list(in_dims)
        if block_transform.dim != self.in_dims[0]:
            raise ValueError("block_transform dim must match input dim")
        self.block_size = self.in_dims[0]
        self.input_rank = len(self.in_dims) - 1
        self.n_blocks = math.prod(self.in_dims[1:])
        if self.input_rank not in (0, 2):
            raise NotImplementedError("usflows_b200: in_dims=[d] (flat) and [C, H, W] (1x1 convolution) are built")
        self.block_transform = block_transform

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return self.block_transform.log_abs_det_jacobian(x, y, context) * self.n_blocks

    def _ladj_device(self):
        return self.block_transform._ladj_device() * torch.tensor(
            [float(self.n_blocks), 1.0], device=self.block_transform._ladj_device().device)

    def sign(self):
        return self.block_transform.sign() ** self.n_blocks

    def is_feasible(self) -> bool:
        return self.block_transform.is_feasible()

    def add_jitter(self, jitter: float = 1e-6) -> None:
        self.block_transform.add_jitter(jitter)

    def to(self, device):
        self.block_transform.to(device)
        return super().to(device)

    def _to_block_plane_linear(self) -> "BlockAffineTransform":
        return BlockAffineTransform(self.in_dims, self.block_transform._to_plane_linear())     # transforms.py:993-1002

    def simplify(self):
        """Plain-matrix form (transforms.py:1004-1020): a `Bijective1x1Conv2d` for `[C, H, W]` events, else a block of a
        `PlaneBijectiveLinearTransform`."""
        if len(self.in_dims) == 3:
            p = self.block_transform._prepared()
            C = self.block_size
            return Bijective1x1Conv2d(p["matrix"].clone().view(C, C, 1, 1), p["bias"].clone().view(C),
                                      inv_weight=p["inverse_matrix"].clone().view(C, C, 1, 1), ladj=p["ladj"][0].clone(),
                                      n_blocks=self.n_blocks)
        return self._to_block_plane_linear()


class Bijective1x1Conv2d(BaseTransform):
    """y = W * x + b as a 1x1 convolution with given weights (transforms.py:1031-1176): the `[C, H, W]` counterpart of
    `PlaneBijectiveLinearTransform`.  Buffers `weight [C, C, 1, 1]`, `bias [C]`, `inv_weight`, and the reference's
    `forward_conv` / `inverse_conv` modules as parameter holders (same state-dict keys); runs as the C x C contraction
    over channels-last rows like `BlockAffineTransform`.  log|det J| = H * W * log|det W| (`:1128-1146`): `H * W` comes
    from the tensor when `log_abs_det_jacobian` is called with one, from `n_blocks` (set by `simplify()` or by the
    `Flow` that holds the layer, from its event shape) inside a flow.  `inv_weight` / `ladj` / `n_blocks` are
    extensions: the reference re-derives the first two with `torch.inverse` / `slogdet` (the default here as well)."""

    def __init__(self, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
                 inv_weight: Optional[torch.Tensor] = None, ladj: Optional[torch.Tensor] = None,
                 n_blocks: Optional[int] = None):
        super().__init__()
        weight = weight.detach()
        if weight.dim() != 4 or weight.shape[2] != 1 or weight.shape[3] != 1:
            raise ValueError("Weight must be 4D tensor with shape (C, C, 1, 1)")
        self.in_channels, self.out_channels = weight.shape[1], weight.shape[0]
        if self.in_channels != self.out_channels:
            raise ValueError("Input and output channels must be equal for bijective 1x1 conv")
        C = self.in_channels
        self.dim = self.block_size = C
        self.n_blocks = n_blocks
        with torch.no_grad():
            w2 = weight.reshape(C, C)
            if inv_weight is None:
                inv_weight = torch.linalg.inv(w2.double()).to(weight.dtype).view(C, C, 1, 1)       # :1092-1101
            if ladj is None:
                ladj = torch.linalg.slogdet(w2.double())[1].to(weight.dtype)                       # :1139-1140
        self.register_buffer("weight", weight)
        self.register_buffer("bias", None if bias is None else bias.detach())
        self.register_buffer("inv_weight", inv_weight.detach())
        self.forward_conv = nn.Conv2d(C, C, kernel_size=1, bias=False)
        self.forward_conv.weight = nn.Parameter(self.weight)
        if bias is not None:
            self.forward_conv.bias = nn.Parameter(self.bias)
        self.inverse_conv = nn.Conv2d(C, C, kernel_size=1, bias=False)
        self.inverse_conv.weight = nn.Parameter(self.inv_weight)
        self.register_buffer("ladj", torch.as_tensor(ladj, dtype=weight.dtype).detach().reshape(()).to(weight.device),
                             persistent=False)

    def _prepared(self) -> dict:
        """Same contract as `AffineTransform._prepared` (per-pixel C x C map; `ladj` is per block).  Read from the
        parameters the reference's forward / backward use (`forward_conv`, `inverse_conv`)."""
        C = self.in_channels
        fw, iw, fb = self.forward_conv.weight, self.inverse_conv.weight, self.forward_conv.bias
        key = _versions([fw, iw] + ([] if fb is None else [fb]))
        if getattr(self, "_prep_key", None) != key:
            ops.require_cuda(fw, "Bijective1x1Conv2d.forward_conv.weight")
            W, Winv = fw.detach().reshape(C, C).contiguous(), iw.detach().reshape(C, C).contiguous()
            b = torch.zeros(C, dtype=W.dtype, device=W.device) if fb is None else fb.detach()
            ladj = torch.stack([self.ladj.to(W.device, torch.float32), torch.zeros((), dtype=torch.float32, device=W.device)])
            self._prep_cache = dict(matrix=W, inverse_matrix=Winv, bias=b, ladj=ladj, matrix64=W.double(),
                                    inverse64=Winv.double())
            self._prep_key = key
        return self._prep_cache

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        if x is not None and x.dim() == 4:                       # per batch element, as the reference (:1142-1146)
            return self.ladj * (x.shape[2] * x.shape[3]) * torch.ones(x.shape[0], device=x.device)
        if self.n_blocks is None:
            raise RuntimeError("Bijective1x1Conv2d: the spatial size is unknown (pass x, or set n_blocks)")
        return self.ladj * self.n_blocks

    def _ladj_device(self):
        if self.n_blocks is None:
            raise RuntimeError("Bijective1x1Conv2d: the spatial size is unknown (set n_blocks = H * W)")
        ladj = self._prepared()["ladj"]
        return ladj * torch.tensor([float(self.n_blocks), 1.0], device=ladj.device)

    def sign(self):
        return torch.linalg.slogdet(self.forward_conv.weight.detach().reshape(self.dim, self.dim).double())[0]

    def is_feasible(self) -> bool:
        return bool(torch.isfinite(self.ladj)) and float(self.ladj) > math.log(1e-6)               # |det W| > 1e-6 (:1148-1151)

    def update_inverse_weight(self) -> None:
        """Inverse and log|det| re-derived from the forward weight (:1092-1101)."""
        C = self.in_channels
        with torch.no_grad():
            w2 = self.forward_conv.weight.detach().reshape(C, C).double()
            inv = torch.linalg.inv(w2).to(self.weight.dtype).view(C, C, 1, 1)
            self.inverse_conv.weight.copy_(inv)
            self.inv_weight.copy_(inv)
            self.ladj.copy_(torch.linalg.slogdet(w2)[1].to(self.weight.dtype))

    def add_jitter(self, jitter: float = 1e-6) -> None:
        """weight += jitter * I, inverse and log|det| re-derived (`jitter`, :1153-1168)."""
        with torch.no_grad():
            self.forward_conv.weight.reshape(self.in_channels, -1).diagonal().add_(jitter)
            self.weight.copy_(self.forward_conv.weight)
        self.update_inverse_weight()


class InverseTransform(BaseTransform):
    """Swaps forward/backward of the wrapped transform and negates its log-det (transforms.py:349-414)."""

    def __init__(self, transform: BaseTransform):
        super().__init__()
        self.transform = transform
        self.bijective = transform.bijective

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return -self.transform.log_abs_det_jacobian(x, y, context)

    def _ladj_device(self):
        inner = self.transform._ladj_device()
        if inner is None:
            return None
        return inner * torch.tensor([-1.0, 1.0], device=inner.device)

    def sign(self):
        return self.transform.sign()

    def is_feasible(self) -> bool:
        return self.transform.is_feasible()

    def add_jitter(self, jitter: float = 1e-6) -> None:
        self.transform.add_jitter(jitter)

    def simplify(self):
        return InverseTransform(self.transform.simplify())


class MaskedCoupling(BaseTransform):
    """y = x + (1 - mask) * conditioner(x * mask)  (additive coupling, transforms.py:254-347); log|det| = 0.

    The conditioner must be a `usflows_b200.nn.DenseNN` (Linear/ReLU stack): its GEMMs, the masking and the
    residual add/sub run fused in the contraction kernels' epilogues."""

    def __init__(self, mask: torch.Tensor, conditioner: nn.Module):
        super().__init__()
        self.mask = mask
        self.conditioner = conditioner
        self.input_shape = mask.shape

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        return 0.0

    def sign(self):
        return 1.0

    def to(self, device):
        self.mask = self.mask.to(device)
        return super().to(device)

    def _raw(self) -> dict:
        """Conditioner weights / biases and the flat mask, for the engine's planner (which either folds the mask
        into the first / last Linear or re-orders the features so both halves are contiguous)."""
        from .nn import ConvNet, ConvNet2D
        if isinstance(self.conditioner, ConvNet2D) or (isinstance(self.conditioner, ConvNet) and not self.conditioner.is_vector):
            # image-shaped events (image_engine.py): ConvNet2D, or ConvNet's convolutional branch (networks.py:308-377)
            ws = [m.weight.detach().reshape(m.weight.shape[0], -1) for m in self.conditioner.modules()
                  if isinstance(m, nn.Conv2d)]
            dev = ws[0].device
            return dict(net="convnet2d", module=self.conditioner, weights=ws,
                        mask=self.mask.to(dev).reshape(-1).to(torch.float32))
        if isinstance(self.conditioner, ConvNet):          # the reference's own MLP-style conditioner (networks.py:287-307)
            from . import engine
            desc = engine._convnet_desc(self.conditioner)
            dev = desc["first"][0].device
            return dict(net="convnet", desc=desc, weights=engine._convnet_weights(desc),
                        mask=self.mask.to(dev).reshape(-1).to(torch.float32))
        from .nn import mlp_layers
        lin = mlp_layers(self.conditioner)
        for l in lin:
            ops.require_cuda(l.weight, "conditioner parameter")
        dev = lin[0].weight.device
        return dict(weights=[l.weight.detach() for l in lin], biases=[l.bias.detach() for l in lin],
                    mask=self.mask.to(dev).reshape(-1).to(torch.float32))

    def _prepared(self) -> dict:
        """Mask folded into the first / last Linear: (x*m) W1^T = x (W1 diag(m))^T and
        (1-m) * (h W3^T + b3) = h (diag(1-m) W3)^T + (1-m) b3 -- exact, multiplying by 0/1."""
        from .nn import mlp_layers
        params = list(self.conditioner.parameters())
        key = _versions(params) + (self.mask.data_ptr(), bool(getattr(self.conditioner, "zero_context_default", False)))
        lin = mlp_layers(self.conditioner)
        if getattr(self, "_prep_key", None) != key:
            for p in params:
                ops.require_cuda(p, "conditioner parameter")
            dev = params[0].device
            m = self.mask.to(dev).reshape(-1).to(torch.float32).contiguous()
            inv = (1 - m).contiguous()
            with torch.no_grad(), ops.on_device(dev):
                ws = [l.weight.detach() for l in lin]
                bs = [l.bias.detach() for l in lin]
                w_first = torch.empty_like(ws[0])
                ops.scale_rows_cols(ws[0], w_first, colf=m)
                if len(lin) == 1:
                    tmp = torch.empty_like(w_first)
                    ops.scale_rows_cols(w_first, tmp, rowf=inv)
                    ws = [tmp]
                else:
                    w_last = torch.empty_like(ws[-1])
                    ops.scale_rows_cols(ws[-1], w_last, rowf=inv)
                    ws = [w_first] + ws[1:-1] + [w_last]
                b_last = torch.empty_like(bs[-1])
                ops.scale_rows_cols(bs[-1].reshape(1, -1), b_last.reshape(1, -1), colf=inv)
                bs = bs[:-1] + [b_last]
            self._prep_cache = dict(weights=ws, biases=bs)
            self._prep_key = key
        return self._prep_cache


class MaskedAffineCoupling(MaskedCoupling):
    """y = x * exp((1 - mask) * s) + (1 - mask) * t  with (s, t) = conditioner(x * mask), s clamped to
    [log_scale_min_clip, log_scale_max_clip]; forward log|det J| of a row = sum_j (1 - mask_j) s_j.

    EXTENSION: the reference's `MaskedCoupling` (transforms.py:254-347) is additive only; this scale-and-shift form is
    the "masked affine coupling" the task names, with the clamped log-scale of pyro's `AffineCoupling`.  The
    conditioner is a `usflows_b200.nn.DenseNN` with `param_dims=[d, d]` (log-scale block first, then shift).  Its
    log-det is data dependent, so a flow containing it carries a per-row log-det vector next to the model constant."""

    def __init__(self, mask: torch.Tensor, conditioner: nn.Module, log_scale_min_clip: float = -5.0,
                 log_scale_max_clip: float = 3.0):
        super().__init__(mask, conditioner)
        d = int(mask.numel())
        out = list(conditioner.layers)[-1].weight.shape[0]
        if out != 2 * d:
            raise ValueError(f"MaskedAffineCoupling: the conditioner must emit 2*d = {2 * d} values (param_dims=[d, d]), got {out}")
        self.log_scale_min_clip = float(log_scale_min_clip)
        self.log_scale_max_clip = float(log_scale_max_clip)

    def _raw(self) -> dict:
        raw = super()._raw()
        raw.update(affine=True, clip=(self.log_scale_min_clip, self.log_scale_max_clip))
        return raw

    def log_abs_det_jacobian(self, x=None, y=None, context=None):
        """Per-row forward log|det J| (needs the input x of `forward`, or equivalently its output y: the conditioner
        sees only the masked features, which the layer leaves unchanged)."""
        from . import engine
        ref = x if x is not None else y
        if ref is None:
            raise ValueError("MaskedAffineCoupling.log_abs_det_jacobian needs x or y (the log-det is data dependent)")
        m = self.mask.to(ref.device).reshape(-1).to(torch.float32)
        st = engine.run_mlp(self.conditioner, (ref.reshape(-1, m.numel()) * m).contiguous())
        s = st[:, :m.numel()].clamp(self.log_scale_min_clip, self.log_scale_max_clip)
        return ((1 - m) * s).sum(-1).reshape(ref.shape[:-1])

"""Multi-vectors: the reference's dict-of-vectors container and its device-resident slab form.

  * ``MultiVector``      host container {Multiindex -> Vector}; mirrors
                         spuq/application/egsz/multi_vector.py:29-166 (mapping API, vector-space
                         operations, ``active_indices`` = sorted keys, pickling hooks).
  * ``SlabMultiVector``  subclass whose storage is ONE column-major N x |Lambda| fp64 slab in HBM
                         (a torch tensor of shape (L, N): column mu is row mu of the tensor, i.e.
                         element (i, mu) at ``data_ptr + 8*(i + N*mu)``), columns in
                         ``active_indices()`` order -- the layout of the reference's Euclidean
                         flatten (multi_vector.py:388-403) and of ``FullTensor.from_list``
                         (spuq/linalg/tensor_vector.py:87-90).  All arithmetic runs in the CUDA
                         kernels of libspuq_sg (axpy / scal / dot); torch only owns the memory.

``SlabMultiVector`` *is a* ``MultiVector`` (the reference type-checks ``apply`` and ``pcg``
arguments with isinstance, multi_operator2.py:38, pcg.py:14), so it can be handed to any code that
expects the reference container: ``w[mu]`` gives a column view, ``w[mu] = vec`` copies into the
column, ``w.keys()``, ``w.active_indices()``, ``w.copy()``, ``+ - *``, ``inner(a, b)`` all work.
"""
import numpy as np
import torch

from . import _abi
from .lambda_tables import as_index_array, sort_permutation
from .linalg import CanonicalBasis, FlatVector, Scalar, Vector, check_basis, inner
from .multiindex import Multiindex, MultiindexSet

__all__ = ["MultiVector", "MultiVectorSharedBasis", "SlabMultiVector", "SlabColumn"]


class MultiVectorBasis:
    """multi_vector.py:515-531 (with the reference's attribute-name slip repaired)."""

    def __init__(self, multivec, single_basis=False):
        self.single_basis = single_basis
        active = multivec.active_indices()
        if single_basis:
            self.basis = multivec[active[0]].basis
        else:
            self.basis = {mu: multivec[mu].basis for mu in active}

    @property
    def dim(self):
        if self.single_basis:
            return self.basis.dim
        return {mu: b.dim for mu, b in self.basis.items()}

    def active_indices(self):
        return sorted(self.basis.keys()) if not self.single_basis else None


class MultiVector(Vector):
    """Host container; every operation loops over the per-mu vectors like the reference."""

    def __init__(self, on_modify=lambda: None, multivector=None):
        self.mi2vec = dict()
        self.on_modify = on_modify
        if multivector is not None:
            for mu, vec in multivector.items():
                self[mu] = vec

    @property
    def basis(self):
        return MultiVectorBasis(self)

    @property
    def dim(self):
        return {mu: self[mu].dim for mu in self.active_indices()}

    @property
    def max_order(self):
        return max(len(mu) for mu in self.keys())

    def __getitem__(self, mi):
        if not isinstance(mi, Multiindex):
            raise TypeError("MultiVector keys are Multiindex objects")
        return self.mi2vec[mi]

    def __setitem__(self, mi, val):
        if not isinstance(mi, Multiindex) or not isinstance(val, Vector):
            raise TypeError("MultiVector maps Multiindex -> Vector")
        self.on_modify()
        self.mi2vec[mi] = val

    def __len__(self):
        return len(self.mi2vec)

    def keys(self):
        return self.mi2vec.keys()

    def values(self):
        return self.mi2vec.values()

    def items(self):
        return self.mi2vec.items()

    iteritems = items

    def active_indices(self):
        return sorted(self.keys())

    def copy(self):
        mv = type(self)()
        for mi in self.keys():
            mv[mi] = self[mi].copy()
        return mv

    def set_defaults(self, multiindex_set, init_vector):
        if not isinstance(multiindex_set, MultiindexSet) or not isinstance(init_vector, Vector):
            raise TypeError("set_defaults(MultiindexSet, Vector)")
        self.on_modify()
        for mi in multiindex_set:
            self[Multiindex(mi)] = init_vector.copy()

    def set_zero(self):
        for mu in self.active_indices():
            self[mu].set_zero()

    def flatten(self):
        """Euclidean vector: per-mu coefficient arrays concatenated in active_indices() order."""
        return FlatVector(np.concatenate([np.asarray(self[mu].coeffs) for mu in self.active_indices()]))

    def __eq__(self, other):
        return type(self) == type(other) and self.mi2vec == other.mi2vec

    __hash__ = None

    def __neg__(self):
        new = self.copy()
        for mi in self.active_indices():
            new[mi] = -self[mi]
        return new

    def __iadd__(self, other):
        assert self.active_indices() == other.active_indices()
        self.on_modify()
        for mi in self.active_indices():
            self[mi] += other[mi]
        return self

    def __isub__(self, other):
        assert self.active_indices() == other.active_indices()
        self.on_modify()
        for mi in self.active_indices():
            self[mi] -= other[mi]
        return self

    def __imul__(self, other):
        assert isinstance(other, Scalar)
        self.on_modify()
        for mi in self.keys():
            self[mi] *= other
        return self

    def __inner__(self, other):
        assert isinstance(other, MultiVector)
        s = 0.0
        for mi in self.keys():
            s += inner(self[mi], other[mi])
        return s

    def __repr__(self):
        return "<%s keys=%s>" % (type(self).__name__, sorted(self.mi2vec.keys()))

    def __getstate__(self):
        odict = self.__dict__.copy()
        odict.pop("on_modify", None)
        return odict

    def __setstate__(self, d):
        self.__dict__.update(d)
        self.on_modify = lambda: None


class MultiVectorSharedBasis(MultiVector):
    """All vectors live on one spatial basis (multi_vector.py:434-512)."""

    def __init__(self, on_modify=lambda: None, multivector=None):
        self.single_basis = True
        super().__init__(on_modify, multivector)

    @property
    def basis(self):
        return MultiVectorBasis(self, single_basis=True)

    def __setitem__(self, mi, val):
        if len(self) > 0:
            assert val.basis == self[self.active_indices()[0]].basis
        super().__setitem__(mi, val)

    def refine(self, cell_ids=None):
        """Every vector prolongated onto the refined basis (multi_vector.py:474-479)."""
        _, prolongate, _ = self[self.active_indices()[0]].basis.refine(cell_ids)
        mv = type(self)()
        for mi in self.keys():
            mv[mi] = prolongate(self[mi])
        return mv


# ---- device helpers ------------------------------------------------------------------------------
_scratch = {}


def _reduce_buffers(dev):
    key = str(dev)
    buf = _scratch.get(key)
    if buf is None:
        nbytes = int(_abi.lib().spuq_sg_reduce_scratch_bytes())
        buf = (torch.zeros(nbytes // 8 + 1, dtype=torch.float64, device=dev),
               torch.zeros(2, dtype=torch.float64, device=dev))
        _scratch[key] = buf
    return buf


def _axpy(alpha, x, y):
    """y += alpha * x on contiguous fp64 device tensors (K3)."""
    assert x.numel() == y.numel() and x.is_contiguous() and y.is_contiguous()
    _abi.check(_abi.lib().spuq_sg_axpy(x.numel(), float(alpha), None, None, _abi.ptr(x), _abi.ptr(y),
                                        _abi.stream_ptr()), "axpy")


def _scal(alpha, x):
    assert x.is_contiguous()
    _abi.check(_abi.lib().spuq_sg_scal(x.numel(), float(alpha), _abi.ptr(x), _abi.stream_ptr()), "scal")


def _dot(x, y):
    assert x.numel() == y.numel() and x.is_contiguous() and y.is_contiguous()
    scratch, out = _reduce_buffers(x.device)
    _abi.check(_abi.lib().spuq_sg_dot(x.numel(), _abi.ptr(x), _abi.ptr(y), _abi.ptr(out), _abi.ptr(scratch),
                                       _abi.stream_ptr()), "dot")
    return float(out[0].item())


class SlabColumn(Vector):
    """View of one column of a slab; behaves like the reference's per-mu vector.  ``coeffs``
    returns a host copy (numpy); ``tensor`` is the live device view."""

    def __init__(self, tensor, basis):
        self.tensor = tensor
        self._basis = basis

    @property
    def basis(self):
        return self._basis

    @property
    def dim(self):
        return self.tensor.numel()

    @property
    def coeffs(self):
        return self.tensor.detach().cpu().numpy()

    @coeffs.setter
    def coeffs(self, value):
        self.tensor.copy_(torch.as_tensor(np.ascontiguousarray(value, dtype=np.float64)))

    def as_array(self):
        return self.coeffs

    def flatten(self):
        return FlatVector(self.coeffs, self._basis)

    def copy(self):
        return SlabColumn(self.tensor.clone(), self._basis)

    def set_zero(self):
        self.tensor.zero_()

    def _other_tensor(self, other):
        check_basis(self.basis, other.basis)
        if isinstance(other, SlabColumn):
            return other.tensor
        return torch.as_tensor(np.ascontiguousarray(other.coeffs, dtype=np.float64)).to(self.tensor.device)

    def __iadd__(self, other):
        _axpy(1.0, self._other_tensor(other), self.tensor)
        return self

    def __isub__(self, other):
        _axpy(-1.0, self._other_tensor(other), self.tensor)
        return self

    def __imul__(self, other):
        if not isinstance(other, Scalar):
            raise TypeError("a vector can only be scaled by a scalar")
        _scal(other, self.tensor)
        return self

    def __neg__(self):
        out = self.copy()
        _scal(-1.0, out.tensor)
        return out

    def __inner__(self, other):
        return _dot(self.tensor, self._other_tensor(other))

    def __rinner__(self, other):
        return self.__inner__(other)

    def __eq__(self, other):
        if isinstance(other, SlabColumn):
            return self._basis == other._basis and bool(torch.equal(self.tensor, other.tensor))
        if isinstance(other, FlatVector):
            return self._basis == other.basis and bool((self.coeffs == other.coeffs).all())
        return False

    __hash__ = None


class _SharedBasisView:
    """What ``w.basis`` returns for a slab: ``.basis`` is the single spatial basis (the reference's
    ``MultiVectorBasis(single_basis=True)``, multi_vector.py:447-450)."""

    single_basis = True

    def __init__(self, basis, keys):
        self.basis = basis
        self._keys = keys

    @property
    def dim(self):
        return self.basis.dim

    def active_indices(self):
        return list(self._keys)


class SlabMultiVector(MultiVector):
    """MultiVector stored as one column-major N x L fp64 slab on the device."""

    def __init__(self, indices=None, n=None, basis=None, slab=None, on_modify=lambda: None,
                 _sorted=False):
        self.on_modify = on_modify
        self.single_basis = True
        if indices is None:
            self._index_array = np.zeros((0, 0), dtype=np.int32)
            self._keys, self._pos = [], {}
            self._n = n
            self._basis = basis
            self.slab = None
            return
        arr = as_index_array(indices)
        if not _sorted:
            arr = np.ascontiguousarray(arr[sort_permutation(arr)])
        if slab is None:
            if n is None:
                n = basis.dim
            slab = torch.zeros((arr.shape[0], int(n)), dtype=torch.float64, device=_abi.device())
        assert slab.dtype == torch.float64 and slab.dim() == 2 and slab.is_contiguous()
        assert slab.shape[0] == arr.shape[0]
        self.slab = slab
        self._n = int(slab.shape[1])
        self._basis = basis if basis is not None else CanonicalBasis(self._n)
        assert self._basis.dim == self._n
        self._set_indices(arr)

    def _set_indices(self, arr):
        self._index_array = arr
        self._keys = [Multiindex(row) for row in arr]
        self._pos = {mu: k for k, mu in enumerate(self._keys)}

    # -- constructors ---------------------------------------------------------------------------
    @classmethod
    def from_multivector(cls, mv):
        """Upload any MultiVector whose vectors share one basis (host -> device copy)."""
        if isinstance(mv, SlabMultiVector):
            return mv.copy()
        active = mv.active_indices()
        first = mv[active[0]]
        for mu in active:
            check_basis(first.basis, mv[mu].basis)
        host = np.empty((len(active), first.dim), dtype=np.float64)
        for k, mu in enumerate(active):
            host[k] = np.asarray(mv[mu].coeffs, dtype=np.float64)
        out = cls(active, basis=first.basis, _sorted=False)
        out.slab.copy_(torch.from_numpy(host))
        return out

    @classmethod
    def from_array(cls, indices, array, basis=None):
        """``array``: N x L (columns in the order of ``indices``) numpy array or (L, N) tensor."""
        arr = as_index_array(indices)
        perm = sort_permutation(arr)
        if isinstance(array, torch.Tensor):
            t = array.to(_abi.device(), torch.float64)[torch.as_tensor(perm.astype(np.int64), device=array.device)]
            return cls(arr[perm], slab=t.contiguous(), basis=basis, _sorted=True)
        host = np.ascontiguousarray(np.asarray(array, dtype=np.float64).T[perm])
        out = cls(arr[perm], n=host.shape[1], basis=basis, _sorted=True)
        out.slab.copy_(torch.from_numpy(host))
        return out

    def to_multivector(self, cls=MultiVector):
        """Download into a host container of FlatVectors."""
        host = self.slab.detach().cpu().numpy()
        mv = cls()
        for k, mu in enumerate(self._keys):
            mv[mu] = FlatVector(host[k].copy(), self._basis)
        return mv

    def to_array(self):
        """N x L numpy array, columns in active_indices() order."""
        return self.slab.detach().cpu().numpy().T.copy()

    def like(self, slab=None):
        """New slab vector over the same index set (zero-filled unless ``slab`` is given)."""
        if slab is None:
            slab = torch.zeros_like(self.slab)
        out = SlabMultiVector.__new__(SlabMultiVector)
        out.on_modify = lambda: None
        out.single_basis = True
        out.slab, out._n, out._basis = slab, self._n, self._basis
        out._index_array, out._keys, out._pos = self._index_array, self._keys, self._pos
        return out

    # -- mapping API ----------------------------------------------------------------------------
    @property
    def index_array(self):
        return self._index_array

    @property
    def n(self):
        return self._n

    @property
    def basis(self):
        return _SharedBasisView(self._basis, self._keys)

    @property
    def spatial_basis(self):
        return self._basis

    @property
    def dim(self):
        return {mu: self._n for mu in self._keys}

    @property
    def max_order(self):
        return max(len(mu) for mu in self._keys)

    @property
    def mi2vec(self):
        return {mu: self[mu] for mu in self._keys}

    def __len__(self):
        return len(self._keys)

    def keys(self):
        return list(self._keys)

    def values(self):
        return [self[mu] for mu in self._keys]

    def items(self):
        return [(mu, self[mu]) for mu in self._keys]

    iteritems = items

    def active_indices(self):
        return list(self._keys)

    def __getitem__(self, mi):
        if not isinstance(mi, Multiindex):
            raise TypeError("MultiVector keys are Multiindex objects")
        return SlabColumn(self.slab[self._pos[mi]], self._basis)       # KeyError when absent

    def __setitem__(self, mi, val):
        if not isinstance(mi, Multiindex) or not isinstance(val, Vector):
            raise TypeError("MultiVector maps Multiindex -> Vector")
        self.on_modify()
        if self.slab is None:
            self._basis = self._basis or val.basis
            self._n = val.dim
            self.slab = torch.zeros((0, self._n), dtype=torch.float64, device=_abi.device())
        check_basis(self._basis, val.basis)
        if mi not in self._pos:
            self._insert(mi)
        col = self.slab[self._pos[mi]]
        if isinstance(val, SlabColumn):
            if val.tensor.data_ptr() != col.data_ptr():
                col.copy_(val.tensor)
        else:
            col.copy_(torch.from_numpy(np.ascontiguousarray(val.coeffs, dtype=np.float64)))

    def insert_zero_columns(self, new_indices):
        """Grow Lambda by several indices at once (``Marking.refine_y``, marking2.py:205-208): the new
        columns are zero, indices already present are left alone (the reference would overwrite them
        with zeros; a marking never proposes active indices), and the slab is re-laid-out in sorted
        order with ONE gather on the device instead of one per index."""
        fresh = []
        for mi in new_indices:
            if not isinstance(mi, Multiindex):
                raise TypeError("MultiVector keys are Multiindex objects")
            if mi not in self._pos and mi not in fresh:
                fresh.append(mi)
        if not fresh:
            return
        if self.slab is None:
            raise ValueError("insert_zero_columns needs a vector that already has a basis")
        self.on_modify()
        n_old = len(self._keys)
        width = max([self._index_array.shape[1] if self._index_array.size else 0, 1] + [len(mi) for mi in fresh])
        old = as_index_array(self._index_array, width) if n_old else np.zeros((0, width), np.int32)
        add = np.array([mi.padded(width) for mi in fresh], dtype=np.int32).reshape(len(fresh), width)
        arr = np.vstack([old, add])
        perm = sort_permutation(arr)
        new_slab = torch.zeros((arr.shape[0], self._n), dtype=torch.float64, device=self.slab.device)
        src = torch.as_tensor(perm.astype(np.int64), device=self.slab.device)
        keep = src < n_old
        if n_old:
            new_slab[keep] = self.slab[src[keep]]
        self.slab = new_slab
        self._set_indices(np.ascontiguousarray(arr[perm]))

    def refine(self, cell_ids=None):
        """``MultiVectorSharedBasis.refine`` (multi_vector.py:474-479) on the slab: the shared basis is refined
        once and its prolongation applied to ALL columns in one launch (a rectangular sparse matrix times
        the slab) instead of once per multi-index.  The basis' ``refine`` must return the reference's
        triple (new_basis, prolongate, restrict); a ``prolongate`` that can state its matrix (``as_csr`` /
        ``as_matrix``) runs on the device, any other callable column by column on the host."""
        new_basis, prolongate, _ = self.spatial_basis.refine(cell_ids)
        mat = None
        if hasattr(prolongate, "as_csr"):
            mat = prolongate.as_csr()
        elif hasattr(prolongate, "as_matrix"):
            import scipy.sparse as sps
            m = prolongate.as_matrix()
            mat = sps.csr_matrix(m if sps.issparse(m) else np.asarray(m, dtype=np.float64))
        if mat is not None:
            from .multi_operator import CsrSlabOperator
            op = CsrSlabOperator(mat, len(self))
            fine = torch.empty((len(self), mat.shape[0]), dtype=torch.float64, device=self.slab.device)
            op.apply_slab(self.slab, fine)
            op.plan.close()
            out = type(self)(self._index_array, n=mat.shape[0], basis=new_basis, slab=fine, _sorted=True)
            return out
        mv = MultiVectorSharedBasis()
        for mi in self.keys():
            mv[mi] = prolongate(self[mi])
        return type(self).from_multivector(mv)

    def _insert(self, mi):
        """Grow Lambda by one index (Lambda growth between solves, marking2.py:205-208): the new
        column is zero-initialised and the slab is re-laid-out in sorted order on the device."""
        width = max(self._index_array.shape[1] if self._index_array.size else 0, len(mi), 1)
        old = as_index_array(self._index_array, width) if len(self._keys) else np.zeros((0, width), np.int32)
        arr = np.vstack([old, np.asarray(mi.padded(width), dtype=np.int32)[None, :]])
        perm = sort_permutation(arr)
        new_slab = torch.zeros((arr.shape[0], self._n), dtype=torch.float64, device=self.slab.device)
        src = torch.as_tensor(perm.astype(np.int64), device=self.slab.device)
        keep = src < len(self._keys)
        if len(self._keys):
            new_slab[keep] = self.slab[src[keep]]
        self.slab = new_slab
        self._set_indices(np.ascontiguousarray(arr[perm]))

    def set_defaults(self, multiindex_set, init_vector):
        if not isinstance(multiindex_set, MultiindexSet) or not isinstance(init_vector, Vector):
            raise TypeError("set_defaults(MultiindexSet, Vector)")
        self.on_modify()
        arr = as_index_array(np.asarray(multiindex_set.arr))
        if len(self._keys):
            for row in arr:
                self[Multiindex(row)] = init_vector
            return
        arr = np.ascontiguousarray(arr[sort_permutation(arr)])
        self._basis = self._basis or init_vector.basis
        self._n = init_vector.dim
        col = torch.from_numpy(np.ascontiguousarray(init_vector.coeffs, dtype=np.float64)).to(_abi.device())
        self.slab = col.unsqueeze(0).repeat(arr.shape[0], 1).contiguous()
        self._set_indices(arr)

    def set_zero(self):
        self.slab.zero_()

    def copy(self):
        return self.like(self.slab.clone())

    def flatten(self):
        return FlatVector(self.slab.detach().cpu().numpy().reshape(-1).copy())

    # -- vector space -------------------------------------------------------------------------
    def _same_set(self, other):
        if isinstance(other, SlabMultiVector):
            ok = (self._index_array is other._index_array or
                  (self._index_array.shape == other._index_array.shape and
                   bool((self._index_array == other._index_array).all())))
            assert ok, "index sets differ"
            check_basis(self._basis, other._basis)
            return other.slab
        assert self.active_indices() == other.active_indices(), "index sets differ"
        return SlabMultiVector.from_multivector(other).slab

    def __iadd__(self, other):
        self.on_modify()
        _axpy(1.0, self._same_set(other), self.slab)
        return self

    def __isub__(self, other):
        self.on_modify()
        _axpy(-1.0, self._same_set(other), self.slab)
        return self

    def __imul__(self, other):
        assert isinstance(other, Scalar)
        self.on_modify()
        _scal(other, self.slab)
        return self

    def __neg__(self):
        out = self.copy()
        _scal(-1.0, out.slab)
        return out

    def __inner__(self, other):
        assert isinstance(other, MultiVector)
        return _dot(self.slab, self._same_set(other))

    def __rinner__(self, other):
        return self.__inner__(other)

    def __eq__(self, other):
        if not isinstance(other, SlabMultiVector):
            return False
        return (self._keys == other._keys and self._basis == other._basis
                and bool(torch.equal(self.slab, other.slab)))

    __hash__ = None

    def __repr__(self):
        return "<%s keys=%s>" % (type(self).__name__, self._keys)

    # -- pickling: index array + host copy of the slab (run_SFEM2.py:170-185 pickles w_history) -------
    def __getstate__(self):
        return {"index_array": self._index_array, "basis": self._basis,
                "slab": None if self.slab is None else self.slab.detach().cpu().numpy()}

    def __setstate__(self, d):
        self.on_modify = lambda: None
        self.single_basis = True
        self._basis = d["basis"]
        host = d["slab"]
        self._n = None if host is None else host.shape[1]
        self.slab = None if host is None else torch.from_numpy(host).to(_abi.device())
        self._set_indices(d["index_array"])

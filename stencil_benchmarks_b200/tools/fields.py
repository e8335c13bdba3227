"""Host field allocation with the reference's layout / padding / alignment rules.

Same contract as ``alloc_array`` / ``nbytes`` of the reference
(stencil_benchmarks/tools/array.py:53-208):

* ``layout`` is a permutation; the axis holding ``ndim-1`` has unit stride, the
  axis holding ``0`` the largest stride;
* with ``alignment > 0`` the extent of the unit-stride axis is padded so that all
  other strides are multiples of ``alignment`` bytes, and the element at
  ``index_to_align`` (the first interior point for stencil fields,
  stencils/base.py:103-104) sits on an ``alignment``-byte boundary;
* memory comes from an ``alloc(nbytes)`` callable returning a buffer object, so
  the B200 backend can hand out pinned (page-locked) host memory for fast
  H2D/D2H, where the reference uses its small-page / huge-page allocators.
"""

import ctypes
from typing import Any, Callable, Optional, Sequence

import numpy as np


def _default_alloc(nbytes: int):
    return bytearray(nbytes)


def _buffer_address(buffer) -> int:
    return ctypes.addressof(ctypes.c_char.from_buffer(buffer))


def padded_strides(shape: Sequence[int], itemsize: int, layout: Sequence[int], alignment: int):
    """Byte strides and total byte extent for a padded field."""
    ndim = len(shape)
    strides = [0] * ndim
    extent = itemsize
    for rank in reversed(range(ndim)):  # rank ndim-1 = fastest axis
        axis = list(layout).index(rank)
        strides[axis] = extent
        extent *= shape[axis]
        if rank == ndim - 1 and alignment:
            extent = -(-extent // alignment) * alignment
    return tuple(strides), extent


def alloc_array(
    shape: Sequence[int],
    dtype,
    layout: Sequence[int],
    alignment: int = 0,
    index_to_align: Optional[Sequence[int]] = None,
    alloc: Optional[Callable[[int], Any]] = None,
) -> np.ndarray:
    """Allocate an uninitialised, padded and aligned ndarray (see module docstring).

    >>> x = alloc_array((2, 3), "int32", (0, 1), alignment=64)
    >>> x.strides
    (64, 4)
    >>> x.ctypes.data % 64
    0
    >>> y = alloc_array((4, 5, 6), "float64", (2, 1, 0), 128, index_to_align=(1, 1, 1))
    >>> (y.ctypes.data + 8 + y.strides[1] + y.strides[2]) % 128, y.strides[1] % 128
    (0, 0)
    """
    shape = tuple(int(s) for s in shape)
    layout = tuple(int(v) for v in layout)
    dtype = np.dtype(dtype)
    alignment = int(alignment)
    ndim = len(shape)
    if sorted(layout) != list(range(ndim)):
        raise ValueError("invalid layout specification")
    if alignment < 0:
        raise ValueError("alignment must be non-negative")
    if index_to_align is None:
        index_to_align = (0,) * ndim
    if len(index_to_align) != ndim:
        raise ValueError("dimension mismatch")
    if alloc is None:
        alloc = _default_alloc

    strides, extent = padded_strides(shape, dtype.itemsize, layout, alignment)
    buffer = alloc(extent + alignment)
    offset = 0
    if alignment:
        anchor = _buffer_address(buffer) + sum(s * i for s, i in zip(strides, index_to_align))
        offset = -anchor % alignment
    return np.ndarray(shape=shape, dtype=dtype, buffer=buffer, offset=offset, strides=strides)


def nbytes(data: np.ndarray) -> int:
    """Bytes from the first to the last element of ``data``, padding included.

    This is what a flat 1-D copy (``cudaMemcpy``) of the field has to move.

    >>> nbytes(np.zeros((2, 3), dtype="int32"))
    24
    >>> nbytes(alloc_array((2, 3), "int32", (0, 1), alignment=64))
    76
    """
    return int(sum((n - 1) * s for n, s in zip(data.shape, data.strides)) + data.itemsize)

"""Minimal NIfTI-1 reader / writer (single-file `.nii` / `.nii.gz`), so the ACDC / DSB15 volumes written by the
reference's preprocessing scripts (src/acdc_preprocess.py:74-77: `nib.save(nib.Nifti1Image(video, np.eye(4)), ...)`)
can be read on boxes without nibabel.  `read` returns what `np.asarray(nib.load(path).dataobj)` returns: the array in
the header's shape (first index fastest on disk), scaled by scl_slope / scl_inter when the header asks for it.

Header fields used (NIfTI-1, 348 bytes): sizeof_hdr@0 (endianness probe), dim[8]@40, datatype@70, bitpix@72,
vox_offset@108, scl_slope@112, scl_inter@116, magic@344.
"""
import gzip
import struct

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}


class NiftiError(ValueError):
    pass


def _open(path, mode):
    path = str(path)
    return gzip.open(path, mode) if path.endswith('.gz') else open(path, mode)


def read(path):
    with _open(path, 'rb') as f:
        raw = f.read()
    if len(raw) < 348:
        raise NiftiError(f'{path}: shorter than a NIfTI-1 header')
    for end in ('<', '>'):
        if struct.unpack_from(end + 'i', raw, 0)[0] == 348:
            break
    else:
        raise NiftiError(f'{path}: sizeof_hdr is not 348 (not a NIfTI-1 file)')
    if raw[344:347] not in (b'n+1', b'ni1'):
        raise NiftiError(f'{path}: bad magic {raw[344:348]!r}')
    if raw[344:347] == b'ni1':
        raise NiftiError(f'{path}: header/image pairs (.hdr/.img) are not supported')
    dim = struct.unpack_from(end + '8h', raw, 40)
    ndim = dim[0]
    if not 1 <= ndim <= 7:
        raise NiftiError(f'{path}: dim[0] = {ndim}')
    shape = tuple(int(d) for d in dim[1:1 + ndim])
    code = struct.unpack_from(end + 'h', raw, 70)[0]
    if code not in _DTYPES:
        raise NiftiError(f'{path}: unsupported datatype code {code}')
    dtype = np.dtype(_DTYPES[code]).newbyteorder(end)
    vox_offset, slope, inter = struct.unpack_from(end + '3f', raw, 108)
    off = max(int(vox_offset), 352)
    n = int(np.prod(shape))
    if len(raw) < off + n * dtype.itemsize:
        raise NiftiError(f'{path}: truncated data ({len(raw) - off} bytes for {n} x {dtype})')
    data = np.frombuffer(raw, dtype=dtype, count=n, offset=off).reshape(shape, order='F')
    data = data.astype(dtype.newbyteorder('='), copy=False)
    if np.isfinite(slope) and slope != 0 and np.isfinite(inter) and not (slope == 1 and inter == 0):
        data = data * np.float64(slope) + np.float64(inter)
    return data


def write(path, array):
    """Writes `array` as a single-file little-endian NIfTI-1 volume with an identity affine and no scaling."""
    array = np.asarray(array)
    if array.dtype not in _CODES:
        raise NiftiError(f'unsupported dtype {array.dtype}')
    if not 1 <= array.ndim <= 7:
        raise NiftiError('1 to 7 dimensions are supported')
    hdr = bytearray(352)
    struct.pack_into('<i', hdr, 0, 348)
    dim = [array.ndim] + list(array.shape) + [1] * (7 - array.ndim)
    struct.pack_into('<8h', hdr, 40, *dim)
    struct.pack_into('<2h', hdr, 70, _CODES[array.dtype], array.dtype.itemsize * 8)
    struct.pack_into('<8f', hdr, 76, 1.0, *([1.0] * 7))                 # pixdim
    struct.pack_into('<3f', hdr, 108, 352.0, 0.0, 0.0)                  # vox_offset, scl_slope (0 = none), scl_inter
    struct.pack_into('<2h', hdr, 252, 0, 2)                             # qform_code 0, sform_code 2 (aligned)
    struct.pack_into('<4f', hdr, 280, 1.0, 0.0, 0.0, 0.0)               # srow_x/y/z = identity
    struct.pack_into('<4f', hdr, 296, 0.0, 1.0, 0.0, 0.0)
    struct.pack_into('<4f', hdr, 312, 0.0, 0.0, 1.0, 0.0)
    hdr[344:348] = b'n+1\0'
    with _open(path, 'wb') as f:
        f.write(bytes(hdr))
        f.write(np.asfortranarray(array.astype(array.dtype.newbyteorder('<'), copy=False)).tobytes(order='F'))

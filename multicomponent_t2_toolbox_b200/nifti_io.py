"""Minimal NIfTI-1 (.nii / .nii.gz) reader and writer — what motor_recon_met2 needs from nibabel
(motor/motor_recon_met2_real_data.py:167-173: nib.load(...).get_fdata(), img.affine; :474-503: nib.Nifti1Image(arr, affine)
+ nib.save).  nibabel is not installed in this image; SURVEY.md §8(f) row 1."""
import gzip
import os
import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_DTYPES = {2: np.uint8, 4: np.int16, 8: np.int32, 16: np.float32, 64: np.float64, 256: np.int8, 512: np.uint16,
           768: np.uint32, 1024: np.int64, 1280: np.uint64}
_CODES = {np.dtype(v).str[1:]: k for k, v in _DTYPES.items()}


class NiftiImage:
    def __init__(self, data, affine, header=None):
        self._data = data
        self.affine = np.asarray(affine, dtype=np.float64)
        self.header = header or {}

    @property
    def shape(self):
        return self._data.shape

    def get_fdata(self):
        """Floating-point array with the scl_slope / scl_inter scaling applied (nibabel semantics)."""
        # fresh writable C-order copy, like nibabel; the F-order -> C-order gather is split over the first axis
        src = self._data
        d = np.empty(src.shape, dtype=np.float64)
        if src.ndim >= 2 and src.shape[0] > 1 and src.size > (1 << 20):
            with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as pool:
                list(pool.map(lambda k: np.copyto(d[k], src[k], casting="unsafe"), range(src.shape[0])))
        else:
            d[...] = src
        slope = self.header.get("scl_slope", 0.0)
        inter = self.header.get("scl_inter", 0.0)
        if slope not in (0.0, 1.0) and np.isfinite(slope):
            d = d * slope
        if inter != 0.0 and np.isfinite(inter) and slope != 0.0:
            d = d + inter
        return d


def _quat_affine(h):
    b, c, d = h["quatern_b"], h["quatern_c"], h["quatern_d"]
    a = np.sqrt(max(0.0, 1.0 - (b * b + c * c + d * d)))
    R = np.array([[a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
                  [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
                  [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c]])
    qfac = -1.0 if h["pixdim"][0] < 0 else 1.0
    zooms = np.array([h["pixdim"][1], h["pixdim"][2], h["pixdim"][3] * qfac])
    A = np.eye(4)
    A[:3, :3] = R * zooms
    A[:3, 3] = [h["qoffset_x"], h["qoffset_y"], h["qoffset_z"]]
    return A


def load(path):
    opener = gzip.open if str(path).endswith(".gz") else open
    with opener(path, "rb") as fh:
        raw = fh.read()
    if len(raw) < 348:
        raise ValueError("%s: not a NIfTI-1 file" % path)
    end = "<"
    if struct.unpack("<i", raw[:4])[0] != 348:
        end = ">"
        if struct.unpack(">i", raw[:4])[0] != 348:
            raise ValueError("%s: bad sizeof_hdr" % path)
    magic = raw[344:348]
    if magic[:3] not in (b"n+1", b"ni1"):
        raise ValueError("%s: bad NIfTI magic %r" % (path, magic))
    dim = struct.unpack(end + "8h", raw[40:56])
    datatype, bitpix = struct.unpack(end + "hh", raw[70:74])
    pixdim = struct.unpack(end + "8f", raw[76:108])
    vox_offset, scl_slope, scl_inter = struct.unpack(end + "3f", raw[108:120])
    qform_code, sform_code = struct.unpack(end + "hh", raw[252:256])
    qb, qc, qd, qx, qy, qz = struct.unpack(end + "6f", raw[256:280])
    srow = np.array(struct.unpack(end + "12f", raw[280:328]), dtype=np.float64).reshape(3, 4)
    h = dict(dim=dim, datatype=datatype, bitpix=bitpix, pixdim=pixdim, vox_offset=vox_offset, scl_slope=scl_slope,
             scl_inter=scl_inter, qform_code=qform_code, sform_code=sform_code, quatern_b=qb, quatern_c=qc, quatern_d=qd,
             qoffset_x=qx, qoffset_y=qy, qoffset_z=qz)
    if datatype not in _DTYPES:
        raise ValueError("%s: unsupported NIfTI datatype %d" % (path, datatype))
    ndim = dim[0]
    shape = tuple(int(x) for x in dim[1:1 + ndim])
    dt = np.dtype(_DTYPES[datatype]).newbyteorder(end)
    off = int(vox_offset) if magic[:3] == b"n+1" else 0
    count = int(np.prod(shape))
    data = np.frombuffer(raw, dtype=dt, count=count, offset=max(off, 352 if magic[:3] == b"n+1" else 0))
    data = data.reshape(shape, order="F")
    if sform_code > 0:
        affine = np.vstack([srow, [0, 0, 0, 1]])
    elif qform_code > 0:
        affine = _quat_affine(h)
    else:
        affine = np.diag([pixdim[1] or 1.0, pixdim[2] or 1.0, pixdim[3] or 1.0, 1.0])
    return NiftiImage(data, affine, h)


def save(img_or_array, path, affine=None, compresslevel=1, threads=None):
    """Write a float64 (or the array's own dtype) single-file NIfTI-1 with the affine in the sform.  `.gz` paths are
    deflated by `threads` threads (default: all cores, at most 16) into one ordinary gzip member."""
    if isinstance(img_or_array, NiftiImage):
        arr, affine = img_or_array._data, img_or_array.affine
    else:
        arr = img_or_array
    arr = np.asarray(arr)
    key = arr.dtype.str[1:]
    if key not in _CODES:
        arr = arr.astype(np.float64)
        key = arr.dtype.str[1:]
    affine = np.eye(4) if affine is None else np.asarray(affine, dtype=np.float64)
    hdr = bytearray(348)
    struct.pack_into("<i", hdr, 0, 348)
    dim = [arr.ndim] + list(arr.shape) + [1] * (7 - arr.ndim)
    struct.pack_into("<8h", hdr, 40, *dim)
    struct.pack_into("<hh", hdr, 70, _CODES[key], arr.dtype.itemsize * 8)
    zooms = np.sqrt((affine[:3, :3] ** 2).sum(axis=0))
    pixdim = [1.0] + [float(z) for z in zooms] + [1.0] * 4
    struct.pack_into("<8f", hdr, 76, *pixdim)
    struct.pack_into("<3f", hdr, 108, 352.0, 1.0, 0.0)
    hdr[123] = 10  # xyzt_units: mm + sec
    struct.pack_into("<hh", hdr, 252, 0, 2)
    struct.pack_into("<12f", hdr, 280, *[float(x) for x in affine[:3, :].reshape(-1)])
    hdr[344:348] = b"n+1\x00"
    threads = threads or min(16, os.cpu_count() or 1)
    payload = np.empty(352 + arr.nbytes, dtype=np.uint8)
    payload[:348] = np.frombuffer(bytes(hdr), dtype=np.uint8)
    payload[348:352] = 0
    _fortran_copy(arr, payload[352:].view(arr.dtype.newbyteorder("<")), threads)
    with open(path, "wb") as fh:
        if str(path).endswith(".gz"):
            for piece in _gzip_pieces(payload, compresslevel, threads):
                fh.write(piece)
        else:
            fh.write(payload)


def _fortran_copy(arr, out_flat, threads):
    """out_flat (1-D, arr.size elements) <- arr in Fortran (first index fastest) element order, which is what NIfTI
    stores.  The strided gather is split over the last axis (slowest in the output) and run by `threads` threads
    (NumPy copies release the GIL): the 265 MB fsol_4D volume takes 1.1 s single-threaded."""
    if arr.ndim < 2 or arr.size == 0:
        out_flat[...] = arr.reshape(-1)
        return
    out = out_flat.reshape(arr.shape[::-1])     # C-order view with reversed axes == F-order of arr
    src = arr.transpose()                       # src[t, z, y, x] = arr[x, y, z, t]
    n = src.shape[0]
    if n == 1 and arr.ndim > 2:                 # split the next axis instead
        out, src, n = out[0], src[0], src.shape[1]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        list(pool.map(lambda k: np.copyto(out[k], src[k]), range(n)))


_GZ_CHUNK = 4 << 20


def _gzip_pieces(payload, compresslevel=1, threads=None):
    """ONE standard gzip member (RFC 1952) whose deflate stream is produced by several threads (SURVEY.md §8f row 1:
    gzip of the ~0.42 GB of outputs dominates the wall time of a run once the fit takes < 1 s).  The payload is cut
    into 4 MiB chunks; each chunk is deflated on its own (raw stream, no dictionary carried over) and closed with a
    sync flush — an empty stored block that ends on a byte boundary and does not set the final-block bit — so the
    pieces concatenate into a valid stream; the last chunk is finished normally.  zlib releases the GIL, so the
    chunks really run in parallel.  Any gzip reader (gzip, zlib, nibabel, FSL) sees an ordinary single-member file."""
    view = memoryview(payload)
    n = len(view)
    threads = threads or min(16, os.cpu_count() or 1)
    bounds = [(o, min(n, o + _GZ_CHUNK)) for o in range(0, n, _GZ_CHUNK)] or [(0, 0)]

    def deflate(k):
        lo, hi = bounds[k]
        c = zlib.compressobj(compresslevel, zlib.DEFLATED, -15)
        body = c.compress(view[lo:hi])
        return body + c.flush(zlib.Z_FINISH if k == len(bounds) - 1 else zlib.Z_SYNC_FLUSH)

    yield b"\x1f\x8b\x08\x00" + struct.pack("<I", 0) + (b"\x04" if compresslevel == 1 else b"\x00") + b"\xff"
    with ThreadPoolExecutor(max_workers=max(1, threads)) as pool:
        crc = pool.submit(zlib.crc32, view)
        for piece in pool.map(deflate, range(len(bounds))):
            yield piece
        yield struct.pack("<II", crc.result() & 0xFFFFFFFF, n & 0xFFFFFFFF)

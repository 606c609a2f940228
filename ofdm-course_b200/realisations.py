"""Shared realisations between a MATLAB/Octave run of the reference scripts and the GPU run (north_star:
"noise, channel taps, STO/CFO draws and payload bits can be imported from reference-generated files so both
sides see identical realisations"), plus the payload path of the scripts.

* ``savemat5`` / ``loadmat5``: Level-5 MAT-files without MATLAB or SciPy.  The writer emits uncompressed
  numeric arrays (``load`` in MATLAB/Octave reads them); the reader also accepts what ``save`` writes by default
  (v7: zlib-compressed elements, integer-valued doubles stored in a narrower type).
* ``export_realisation`` / ``import_realisation``: one stream's draws under the variable names the reference
  scripts use (``input_bits``, ``Time_Delay``, ``Freq_Shift``, ``channel_taps``, ``SNR_dB``) plus
  ``noise_normals`` (L x 2: column 1 = the first ``normrnd`` call / real part, column 2 = the second / imaginary
  part, `Task 5/Noise.m:7-8`) and ``h_t`` (an impulse response as returned by `lteFadingChannel` on a unit impulse,
  `Task 5/Task5_part2.m:154`).
* ``file_reader`` / ``display_pic``: `Task 5/file_reader.m:2-12`, `display_pic.m:2-15` (baseline TIFF reader and
  Otsu ``imbinarize`` included; host-side I/O, no device work).
"""
from __future__ import annotations

import struct
import time
import zlib

import numpy as np

# MAT-file data types / array classes (MAT-File Format, Level 5)
_MI = {1: np.int8, 2: np.uint8, 3: np.int16, 4: np.uint16, 5: np.int32, 6: np.uint32, 7: np.float32, 9: np.float64, 12: np.int64, 13: np.uint64}
_MI_OF = {np.dtype(v): k for k, v in _MI.items()}
_MX = {6: np.float64, 7: np.float32, 8: np.int8, 9: np.uint8, 10: np.int16, 11: np.uint16, 12: np.int32, 13: np.uint32, 14: np.int64, 15: np.uint64}
_MX_OF = {np.dtype(v): k for k, v in _MX.items()}
_MI_MATRIX, _MI_COMPRESSED, _MI_UTF8, _MX_CHAR = 14, 15, 16, 4


def _element(mi_type, payload: bytes) -> bytes:
    n = len(payload)
    if 0 < n <= 4:                                     # small data element format
        return struct.pack("<HH", mi_type, n) + payload.ljust(4, b"\0")
    return struct.pack("<II", mi_type, n) + payload + b"\0" * ((-n) % 8)


def _matrix(name: str, a) -> bytes:
    a = np.asarray(a)
    if a.dtype == np.bool_:
        a, logical = a.astype(np.uint8), True
    else:
        logical = False
    is_char = a.dtype.kind in "US"
    if is_char:                                        # one char row vector
        s = str(a.item()) if a.ndim == 0 else "".join(a.tolist())
        dims = (1, len(s))
        flags = struct.pack("<II", _MX_CHAR, 0)
        body = _element(4, np.frombuffer(s.encode("utf-16-le"), dtype=np.uint16).tobytes())
    else:
        if a.ndim == 0:
            a = a.reshape(1, 1)
        elif a.ndim == 1:
            a = a.reshape(1, -1)                       # MATLAB row vector
        cplx = np.iscomplexobj(a)
        real_dtype = np.dtype(a.real.dtype) if cplx else np.dtype(a.dtype)
        if real_dtype not in _MX_OF:
            raise TypeError(f"unsupported dtype {a.dtype}")
        dims = a.shape
        flags = struct.pack("<II", _MX_OF[real_dtype] | (0x0800 if cplx else 0) | (0x0200 if logical else 0), 0)
        f = np.asfortranarray(a)
        body = _element(_MI_OF[real_dtype], np.ascontiguousarray(f.real if cplx else f).tobytes(order="F"))
        if cplx:
            body += _element(_MI_OF[real_dtype], np.ascontiguousarray(f.imag).tobytes(order="F"))
    sub = _element(6, flags) + _element(5, np.asarray(dims, dtype=np.int32).tobytes()) + _element(1, name.encode("ascii")) + body
    return struct.pack("<II", _MI_MATRIX, len(sub)) + sub


def savemat5(path, arrays: dict):
    """Write ``arrays`` (name -> scalar / ndarray / str) as an uncompressed Level-5 MAT-file."""
    text = f"MATLAB 5.0 MAT-file, Platform: ofdm-b200, Created on: {time.asctime()}".encode("ascii")
    head = text.ljust(116, b" ") + b"\0" * 8 + struct.pack("<H", 0x0100) + b"IM"
    with open(path, "wb") as fh:
        fh.write(head)
        for name, a in arrays.items():
            fh.write(_matrix(name, a))


def _read_elements(buf: bytes, pos: int, end: int):
    """Yield (mi_type, payload bytes) of the data elements in buf[pos:end]."""
    while pos + 8 <= end:
        t, n = struct.unpack_from("<II", buf, pos)
        if t >> 16:                                    # small element: type in the low half, size in the high half
            n, t = t >> 16, t & 0xFFFF
            yield t, buf[pos + 4: pos + 4 + n]
            pos += 8
        else:
            yield t, buf[pos + 8: pos + 8 + n]
            pos += 8 + n + ((-n) % 8 if t != _MI_COMPRESSED else 0)


def _parse_matrix(payload: bytes):
    it = list(_read_elements(payload, 0, len(payload)))
    (_, fl), (_, dm), (_, nm) = it[0], it[1], it[2]
    cls_flags = struct.unpack_from("<I", fl, 0)[0]
    cls, cplx, logical = cls_flags & 0xFF, bool(cls_flags & 0x0800), bool(cls_flags & 0x0200)
    dims = tuple(int(x) for x in np.frombuffer(dm, dtype=np.int32))
    name = nm.decode("ascii")
    if cls == _MX_CHAR:
        t, d = it[3]
        if t == _MI_UTF8:
            return name, d.decode("utf-8")
        raw = np.frombuffer(d, dtype=_MI[t]) if t in _MI else np.frombuffer(d, dtype=np.uint16)
        return name, "".join(chr(int(c)) for c in raw)
    if cls not in _MX:
        return name, None                              # cells, structs, sparse: not part of a realisation file
    def part(k):
        t, d = it[k]
        return np.frombuffer(d, dtype=_MI[t]).astype(_MX[cls])
    n = int(np.prod(dims)) if dims else 0
    re_ = part(3)[:n]
    a = (re_ + 1j * part(4)[:n]) if cplx else re_
    a = a.reshape(dims, order="F")
    return name, (a.astype(bool) if logical else a)


def loadmat5(path) -> dict:
    """Read the numeric / logical / char variables of a Level-5 MAT-file (v5, v6, v7-compressed)."""
    buf = open(path, "rb").read()
    if len(buf) < 128 or buf[126:128] != b"IM":
        raise ValueError("not a little-endian Level-5 MAT-file (v7.3/HDF5 files: re-save with '-v7' or '-v6')")
    out = {}
    for t, payload in _read_elements(buf, 128, len(buf)):
        if t == _MI_COMPRESSED:
            inner = zlib.decompress(payload)
            t2, n2 = struct.unpack_from("<II", inner, 0)
            t, payload = t2, inner[8: 8 + n2]
        if t == _MI_MATRIX:
            name, val = _parse_matrix(payload)
            if val is not None:
                out[name] = val
    return out


# ------------------------------------------------------------------------------------------ realisations
REALISATION_FIELDS = ("input_bits", "noise_normals", "channel_taps", "h_t", "Time_Delay", "Freq_Shift", "SNR_dB")


def export_realisation(path, **fields):
    """Write one stream's draws under the reference scripts' variable names (see module docstring).  Bits are
    stored as doubles 0/1 in a row vector, as `file_reader.m:8` produces them."""
    unknown = set(fields) - set(REALISATION_FIELDS) - {k for k in fields if k.startswith("ref_")}
    if unknown:
        raise KeyError(f"unknown realisation fields {sorted(unknown)}; extra reference outputs must be named ref_*")
    out = {}
    for k, v in fields.items():
        a = np.asarray(v)
        if k == "input_bits":
            a = a.astype(np.float64).reshape(1, -1)
        elif k == "noise_normals":
            a = np.asarray(a, dtype=np.float64)
            if a.ndim != 2 or a.shape[1] != 2:
                raise ValueError("noise_normals must be L x 2 (real block, imaginary block)")
        elif a.dtype.kind in "biuf":
            a = a.astype(np.float64)
        out[k] = a
    savemat5(path, out)


def import_realisation(path) -> dict:
    """Inverse of export_realisation; also reads a file MATLAB saved with the same variable names."""
    d = loadmat5(path)
    if "input_bits" in d:
        d["input_bits"] = np.asarray(d["input_bits"]).ravel(order="F").astype(np.uint8)
    if "noise_normals" in d:
        n = np.asarray(d["noise_normals"], dtype=np.float64)
        d["noise_normals"] = n if n.shape[1] == 2 else n.T
    for k in ("Time_Delay", "Freq_Shift", "SNR_dB"):
        if k in d:
            d[k] = float(np.asarray(d[k]).ravel()[0])
    return d


# ------------------------------------------------------------------------------------------ payload path
def read_tiff_gray8(path) -> np.ndarray:
    """Baseline TIFF, 8-bit grayscale, uncompressed strips (what `imread('eagle.tiff')` sees): H x W uint8."""
    b = open(path, "rb").read()
    bo = {b"II": "<", b"MM": ">"}.get(b[:2])
    if bo is None or struct.unpack(bo + "H", b[2:4])[0] != 42:
        raise ValueError("not a TIFF file")
    off = struct.unpack(bo + "I", b[4:8])[0]
    n = struct.unpack(bo + "H", b[off: off + 2])[0]
    sizes = {1: 1, 2: 1, 3: 2, 4: 4}
    tags = {}
    for i in range(n):
        tag, typ, cnt = struct.unpack(bo + "HHI", b[off + 2 + 12 * i: off + 10 + 12 * i])
        raw = b[off + 10 + 12 * i: off + 14 + 12 * i]
        sz = sizes.get(typ, 0) * cnt
        if sz == 0:
            continue
        data = raw[:sz] if sz <= 4 else b[struct.unpack(bo + "I", raw)[0]: struct.unpack(bo + "I", raw)[0] + sz]
        fmt = {1: "B", 2: "c", 3: "H", 4: "I"}[typ]
        tags[tag] = struct.unpack(bo + fmt * cnt, data)
    W, H = tags[256][0], tags[257][0]
    if tags.get(258, (1,))[0] != 8 or tags.get(259, (1,))[0] != 1 or tags.get(277, (1,))[0] != 1:
        raise ValueError("only uncompressed 8-bit single-channel TIFF is supported")
    img = b"".join(b[o: o + c] for o, c in zip(tags[273], tags[279]))
    a = np.frombuffer(img[: W * H], dtype=np.uint8).reshape(H, W)
    return 255 - a if tags.get(262, (1,))[0] == 0 else a          # WhiteIsZero


def imbinarize(gray: np.ndarray) -> np.ndarray:
    """MATLAB ``imbinarize`` default: Otsu's global threshold on the 256-bin histogram (``graythresh``: mean of
    the maximising bins, (idx-1)/255), pixels strictly above it are 1."""
    img = np.asarray(gray).astype(np.float64)
    counts = np.bincount(np.asarray(gray).astype(np.int64).ravel(), minlength=256).astype(np.float64)
    p = counts / counts.sum()
    omega = np.cumsum(p)
    mu = np.cumsum(p * np.arange(1, 257))
    with np.errstate(divide="ignore", invalid="ignore"):
        sigma_b2 = (mu[-1] * omega - mu) ** 2 / (omega * (1 - omega))
    sigma_b2[~np.isfinite(sigma_b2)] = -np.inf
    level = np.mean(np.nonzero(sigma_b2 == sigma_b2.max())[0]) / 255.0
    return (img / 255.0) > level


def file_reader(File, Size_Buffer) -> np.ndarray:
    """`Task 5/file_reader.m:2-12`: binarised image, column-major, first ``Size_Buffer`` bits (1 x N of 0/1)."""
    bw = imbinarize(read_tiff_gray8(File))
    bits = bw.ravel(order="F")
    if Size_Buffer > bits.size:
        raise IndexError("Index exceeds the number of array elements")       # as MATLAB would
    return bits[: int(Size_Buffer)].astype(np.uint8)


def display_pic(binaryImage, side=360) -> np.ndarray:
    """`Task 5/display_pic.m:2-15` without the ``imshow``: zero-pad to side*side bits, reshape column-major,
    scale to 0/255.  Returns the uint8 image."""
    b = np.asarray(binaryImage).ravel()
    full = np.zeros(side * side, dtype=np.uint8)
    full[: min(b.size, full.size)] = b[: full.size]
    return full.reshape((side, side), order="F") * np.uint8(255)

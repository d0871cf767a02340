"""Minimal OME-TIFF layer under the Filter / Label stages.

The reference reaches its intermediates through ``ImInfo.get_memmap`` / ``ImInfo.allocate_memory``
(nellie/im_info/verifier.py:967-1070): ``tifffile.imwrite(path, shape=, dtype=, bigtiff=True,
metadata={"axes": axes}, photometric="minisblack")`` creates an empty, *contiguous* BigTIFF with an OME-XML
description, and ``tifffile.memmap(path, mode="r+")`` maps its pixel block as one ``T[Z]YX`` array that the
stages fill frame by frame.  tifffile / ome_types are not part of this image, so this module writes and maps the
same kind of file itself:

* :func:`write_ome_bigtiff` — little-endian BigTIFF, one IFD per YX page, all pixel data in one aligned block,
  OME-XML (axes, sizes, physical pixel sizes, pixel type, description) in the first page's ImageDescription;
* :func:`memmap_ome_tiff` — ``numpy.memmap`` over the pixel block of any uncompressed TIFF / BigTIFF whose pages
  are stored back to back (what tifffile.memmap accepts), shaped by the OME-XML sizes;
* :func:`read_tiff` — assembles the pages of an uncompressed, possibly strip-interleaved TIFF (e.g. the
  reference's ``sample_data/*.ome.tif``) into an array, for the conversion step that the reference's
  ``FileInfo.save_ome_tiff`` performs (verifier.py:620-700);
* :class:`StackInfo` — the attributes and two methods of ``ImInfo`` that the hot path uses (§8b), with the
  reference's output naming (verifier.py:574-618, :805-828).

Only what the path needs: no compression, no tiles, one sample per pixel.
"""
from __future__ import annotations

import os
import re
import struct
from dataclasses import dataclass, field
from typing import Dict, Optional, Sequence
from xml.sax.saxutils import escape

import numpy as np

_TYPE_SIZE = {1: 1, 2: 1, 3: 2, 4: 4, 5: 8, 6: 1, 7: 1, 8: 2, 9: 4, 10: 8, 11: 4, 12: 8, 16: 8, 17: 8, 18: 8}
_TYPE_FMT = {1: "B", 3: "H", 4: "I", 16: "Q", 6: "b", 8: "h", 9: "i", 17: "q", 11: "f", 12: "d", 18: "Q"}
_OME_TYPES = {"uint8": "uint8", "uint16": "uint16", "uint32": "uint32", "int8": "int8", "int16": "int16",
              "int32": "int32", "float32": "float", "float64": "double"}
_OME_TO_NP = {v: k for k, v in _OME_TYPES.items()}
DATA_ALIGN = 4096


def _sample_format(dt: np.dtype) -> int:
    return {"u": 1, "i": 2, "f": 3}[dt.kind]


def ome_xml(shape: Sequence[int], dtype, axes: str, dim_res: Optional[dict] = None, description: str = "") -> str:
    """OME-XML of one image whose array is ``shape`` in ``axes`` order (subset of TCZYX ending in YX)."""
    dt = np.dtype(dtype)
    size = {ax: 1 for ax in "TCZYX"}
    for ax, n in zip(axes, shape):
        size[ax] = int(n)
    dim_res = dim_res or {}
    phys = ""
    for ax, key in (("X", "PhysicalSizeX"), ("Y", "PhysicalSizeY"), ("Z", "PhysicalSizeZ")):
        if dim_res.get(ax) is not None:
            phys += f' {key}="{float(dim_res[ax])!r}" {key}Unit="µm"'
    if dim_res.get("T") is not None:
        phys += f' TimeIncrement="{float(dim_res["T"])!r}" TimeIncrementUnit="s"'
    n_planes = size["T"] * size["C"] * size["Z"]
    return ('<?xml version="1.0" encoding="UTF-8"?>'
            '<OME xmlns="http://www.openmicroscopy.org/Schemas/OME/2016-06" '
            'xmlns:xsi="http://www.w3.org/2001/XMLSchema-instance" '
            'xsi:schemaLocation="http://www.openmicroscopy.org/Schemas/OME/2016-06 '
            'http://www.openmicroscopy.org/Schemas/OME/2016-06/ome.xsd" Creator="nellie_b200">'
            f'<Image ID="Image:0" Name="{escape(axes)}"><Description>{escape(description)}</Description>'
            f'<Pixels ID="Pixels:0" DimensionOrder="XYZCT" Type="{_OME_TYPES[dt.name]}" '
            f'SizeX="{size["X"]}" SizeY="{size["Y"]}" SizeZ="{size["Z"]}" SizeC="{size["C"]}" SizeT="{size["T"]}"'
            f'{phys} BigEndian="false"><TiffData IFD="0" PlaneCount="{n_planes}"/></Pixels></Image></OME>')


def write_ome_bigtiff(path: str, shape: Sequence[int], dtype, axes: str, dim_res: Optional[dict] = None,
                      description: str = "No description.", data: Optional[np.ndarray] = None) -> None:
    """Create a contiguous OME BigTIFF (zero-filled unless ``data`` is given); counterpart of
    ``tifffile.imwrite(path, shape=, dtype=, bigtiff=True, metadata={"axes": axes})`` (verifier.py:1033-1049)."""
    shape = tuple(int(s) for s in shape)
    if len(shape) != len(axes) or len(shape) < 2 or axes[-2:] != "YX":
        raise ValueError(f"shape {shape} does not match axes {axes!r} (must end in YX)")
    dt = np.dtype(dtype).newbyteorder("<")
    ny, nx = shape[-2:]
    n_pages = int(np.prod(shape[:-2], dtype=np.int64)) if len(shape) > 2 else 1
    page_bytes = ny * nx * dt.itemsize
    xml = ome_xml(shape, dt, axes, dim_res, description).encode("utf-8") + b"\0"
    data_off = -(-(16 + len(xml)) // DATA_ALIGN) * DATA_ALIGN
    ifd_off0 = data_off + n_pages * page_bytes
    ifd_off0 += (-ifd_off0) % 16

    def entry(tag, typ, count, value):
        return struct.pack("<HHQQ", tag, typ, count, value)

    ifds = bytearray()
    pos = ifd_off0
    for p in range(n_pages):
        tags = [entry(256, 4, 1, nx), entry(257, 4, 1, ny), entry(258, 3, 1, dt.itemsize * 8), entry(259, 3, 1, 1),
                entry(262, 3, 1, 1)]
        if p == 0:
            tags.append(entry(270, 2, len(xml), 16))
        tags += [entry(273, 16, 1, data_off + p * page_bytes), entry(277, 3, 1, 1), entry(278, 4, 1, ny),
                 entry(279, 16, 1, page_bytes), entry(339, 3, 1, _sample_format(dt))]
        size = 8 + 20 * len(tags) + 8
        nxt = pos + size if p + 1 < n_pages else 0
        ifds += struct.pack("<Q", len(tags)) + b"".join(tags) + struct.pack("<Q", nxt)
        pos += size
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    with open(path, "wb") as f:
        f.write(struct.pack("<2sHHHQ", b"II", 43, 8, 0, ifd_off0))
        f.write(xml)
        f.seek(ifd_off0)
        f.write(bytes(ifds))
        if data is None:
            f.truncate(max(f.tell(), ifd_off0 + len(ifds)))   # the pixel block is a hole: zero-filled, sparse on disk
    if data is not None:
        mm = np.memmap(path, dtype=dt, mode="r+", offset=data_off, shape=shape)
        mm[...] = np.asarray(data).reshape(shape)
        mm.flush()
        del mm


@dataclass
class _Page:
    width: int
    height: int
    bits: int
    compression: int
    sample_format: int
    samples: int
    rows_per_strip: int
    offsets: tuple
    counts: tuple
    description: Optional[str]


class TiffFile:
    """Parsed IFD chain of a classic TIFF or BigTIFF (either byte order)."""

    def __init__(self, path: str):
        self.path = path
        with open(path, "rb") as f:
            head = f.read(16)
            if head[:2] not in (b"II", b"MM"):
                raise ValueError(f"{path}: not a TIFF file")
            self.bo = "<" if head[:2] == b"II" else ">"
            version = struct.unpack(self.bo + "H", head[2:4])[0]
            if version == 42:
                self.big = False
                off = struct.unpack(self.bo + "I", head[4:8])[0]
            elif version == 43:
                self.big = True
                off = struct.unpack(self.bo + "Q", head[8:16])[0]
            else:
                raise ValueError(f"{path}: unknown TIFF version {version}")
            self.pages = []
            while off:
                off = self._read_ifd(f, off)

    def _values(self, f, typ, count, raw):
        size = _TYPE_SIZE.get(typ, 1) * count
        inline = 8 if self.big else 4
        if size <= inline:
            buf = raw[:size]
        else:
            pos = f.tell()
            f.seek(struct.unpack(self.bo + ("Q" if self.big else "I"), raw)[0])
            buf = f.read(size)
            f.seek(pos)
        if typ == 2:
            return buf.split(b"\0")[0].decode("utf-8", "replace")
        if typ in _TYPE_FMT:
            return struct.unpack(self.bo + _TYPE_FMT[typ] * count, buf)
        return buf

    def _read_ifd(self, f, off):
        f.seek(off)
        if self.big:
            n = struct.unpack(self.bo + "Q", f.read(8))[0]
            fmt, es = self.bo + "HHQ8s", 20
        else:
            n = struct.unpack(self.bo + "H", f.read(2))[0]
            fmt, es = self.bo + "HHI4s", 12
        raw = f.read(n * es)
        nxt = struct.unpack(self.bo + ("Q" if self.big else "I"), f.read(8 if self.big else 4))[0]
        tags = {}
        for i in range(n):
            tag, typ, count, val = struct.unpack(fmt, raw[i * es:(i + 1) * es])
            if tag in (256, 257, 258, 259, 270, 273, 277, 278, 279, 339):
                tags[tag] = self._values(f, typ, count, val)
        one = lambda t, d: (tags[t][0] if t in tags else d)       # noqa: E731
        height = one(257, 0)
        self.pages.append(_Page(one(256, 0), height, one(258, 1), one(259, 1), one(339, 1), one(277, 1),
                                min(one(278, height), height) or height, tuple(tags.get(273, ())),
                                tuple(tags.get(279, ())), tags.get(270)))
        return nxt

    # ------------------------------------------------------------------------------------------------------
    def dtype(self) -> np.dtype:
        p = self.pages[0]
        kind = {1: "u", 2: "i", 3: "f"}.get(p.sample_format, "u")
        return np.dtype(f"{self.bo}{kind}{p.bits // 8}")

    def _check(self):
        if not self.pages:
            raise ValueError(f"{self.path}: no pages")
        p0 = self.pages[0]
        for p in self.pages:
            if p.compression != 1 or p.samples != 1:
                raise ValueError(f"{self.path}: only uncompressed single-sample TIFFs are supported")
            if (p.width, p.height, p.bits, p.sample_format) != (p0.width, p0.height, p0.bits, p0.sample_format):
                raise ValueError(f"{self.path}: pages differ in shape or type")

    def ome_shape_axes(self):
        """(shape, axes) of the series from the OME-XML of page 0, squeezing singleton C/Z/T like the
        reference's necessities files (T kept when present); falls back to (pages, Y, X) / 'QYX'."""
        p0 = self.pages[0]
        n = len(self.pages)
        desc = p0.description or ""
        m = {k: re.search(rf'\bSize{k}="(\d+)"', desc) for k in "TCZYX"}
        if all(m.values()):
            size = {k: int(v.group(1)) for k, v in m.items()}
            order = re.search(r'DimensionOrder="([A-Z]+)"', desc)
            order = order.group(1) if order else "XYZCT"
            if size["T"] * size["C"] * size["Z"] == n and size["Y"] == p0.height and size["X"] == p0.width:
                lead = [ax for ax in reversed(order) if ax not in "YX"]      # slowest first
                axes = "".join(ax for ax in lead if size[ax] > 1 or ax == "T") + "YX"
                return tuple(size[ax] for ax in axes), axes
        if n == 1:
            return (p0.height, p0.width), "YX"
        return (n, p0.height, p0.width), "QYX"

    def contiguous_offset(self) -> Optional[int]:
        """Offset of the pixel block when all strips of all pages lie back to back, else None."""
        self._check()
        pos = None
        start = None
        for p in self.pages:
            for o, c in zip(p.offsets, p.counts):
                if pos is None:
                    start = pos = o
                if o != pos:
                    return None
                pos += c
        expect = len(self.pages) * self.pages[0].height * self.pages[0].width * (self.pages[0].bits // 8)
        return start if pos is not None and pos - start == expect else None

    def read(self) -> np.ndarray:
        """All pages as one native-endian array shaped by :meth:`ome_shape_axes`."""
        self._check()
        p0 = self.pages[0]
        dt = self.dtype()
        out = np.empty((len(self.pages), p0.height * p0.width), dtype=dt.newbyteorder("="))
        with open(self.path, "rb") as f:
            for i, p in enumerate(self.pages):
                buf = bytearray()
                for o, c in zip(p.offsets, p.counts):
                    f.seek(o)
                    buf += f.read(c)
                out[i] = np.frombuffer(bytes(buf), dtype=dt, count=p0.height * p0.width)
        shape, _ = self.ome_shape_axes()
        return out.reshape(shape)


def read_tiff(path: str) -> np.ndarray:
    return TiffFile(path).read()


def memmap_ome_tiff(path: str, mode: str = "r+") -> np.ndarray:
    """``tifffile.memmap(path, mode=mode)`` for contiguous uncompressed files (verifier.py:986)."""
    tf = TiffFile(path)
    off = tf.contiguous_offset()
    if off is None:
        raise ValueError(f"{path}: image data is not contiguous; cannot memory-map (convert it first)")
    shape, _ = tf.ome_shape_axes()
    return np.memmap(path, dtype=tf.dtype(), mode=mode, offset=off, shape=shape)


@dataclass
class StackInfo:
    """The part of ``nellie.im_info.verifier.ImInfo`` the Filter / Label stages touch:
    ``no_z, no_t, axes, shape, dim_res, im_path, pipeline_paths`` + ``get_memmap`` / ``allocate_memory``."""
    im_path: str
    axes: str
    shape: tuple
    dim_res: Dict[str, Optional[float]]
    output_dir: str
    name: str
    dtype: np.dtype = np.dtype("uint16")
    pipeline_paths: Dict[str, str] = field(default_factory=dict)

    PIPELINE = ("im_preprocessed", "im_instance_label", "im_skel", "im_skel_relabelled", "im_pixel_class",
                "im_obj_label_reassigned", "im_branch_label_reassigned", "im_marker", "im_distance", "im_border")

    def __post_init__(self):
        self.no_z = not ("Z" in self.axes and self.shape[self.axes.index("Z")] > 1)
        self.no_t = not ("T" in self.axes and self.shape[self.axes.index("T")] > 1)
        self.new_axes = None
        for key in self.PIPELINE:
            self.create_output_path(key)

    # -- naming (verifier.py:574-618, :805-828) ------------------------------------------------------------------
    @staticmethod
    def output_name(filename_no_ext: str, axes: str, dim_res: dict, ch: int = 0, t_start: int = 0,
                    t_end: Optional[int] = None) -> str:
        parts = []
        for ax in axes:
            if ax not in dim_res:
                continue
            r = dim_res[ax]
            parts.append(f"{ax}{('None' if r is None else str(round(r, 4))).replace('.', 'p')}")
        t_text = f"-t{t_start}_to_{t_end}" if "T" in axes else ""
        return f"{filename_no_ext}-{axes}-{'_'.join(parts)}-ch{ch}{t_text}"

    def create_output_path(self, pipeline_path: str, ext: str = ".ome.tif") -> str:
        base = os.path.join(self.output_dir, "nellie_necessities", self.name)
        self.pipeline_paths[pipeline_path] = f"{base}-{pipeline_path}{ext}"
        return self.pipeline_paths[pipeline_path]

    # -- construction -----------------------------------------------------------------------------------------------
    @classmethod
    def from_array(cls, data: np.ndarray, axes: str, dim_res: dict, output_dir: str, filename_no_ext: str = "stack",
                   ch: int = 0) -> "StackInfo":
        """Write ``data`` as the necessities OME-TIFF (what ``FileInfo.save_ome_tiff`` produces) and describe it.
        A T axis of length 1 is added when missing: the stages index ``memmap[t, ...]`` (§8b)."""
        data = np.asarray(data)
        if "T" not in axes:
            axes, data = "T" + axes, data[None]
        t_n = data.shape[axes.index("T")]
        name = cls.output_name(filename_no_ext, axes, dim_res, ch, 0, t_n - 1)
        path = os.path.join(output_dir, "nellie_necessities", name + ".ome.tif")
        write_ome_bigtiff(path, data.shape, data.dtype, axes, dim_res, "nellie necessities copy of the raw stack", data)
        return cls(path, axes, tuple(data.shape), dict(dim_res), output_dir, name, np.dtype(data.dtype))

    @classmethod
    def from_tiff(cls, path: str, dim_res: dict, output_dir: Optional[str] = None, axes: Optional[str] = None) -> "StackInfo":
        tf = TiffFile(path)
        data = tf.read()
        _, file_axes = tf.ome_shape_axes()
        axes = axes or file_axes.replace("Q", "Z" if data.ndim == 3 else "T")
        stem = os.path.basename(path)
        for ext in (".ome.tif", ".ome.tiff", ".tif", ".tiff"):
            if stem.endswith(ext):
                stem = stem[:-len(ext)]
                break
        out = output_dir or os.path.join(os.path.dirname(os.path.abspath(path)), "nellie_output")
        return cls.from_array(data, axes, dim_res, out, stem)

    # -- the two ImInfo methods of the hot path (verifier.py:967-1070) ----------------------------------------
    def get_memmap(self, file_path: str, read_mode: str = "r+") -> np.ndarray:
        mm = memmap_ome_tiff(file_path, read_mode)
        if mm.ndim == len(self.shape) - 1 and self.axes.startswith("T") and self.shape[0] == 1:
            mm = mm[None]                      # squeezed singleton T: the stages index memmap[t, ...]
        return mm

    def allocate_memory(self, output_path: str, dtype="float", data=None, description: str = "No description.",
                        return_memmap: bool = False, read_mode: str = "r+"):
        dt = np.dtype("float64" if dtype == "float" else dtype) if data is None else np.asarray(data).dtype
        write_ome_bigtiff(output_path, self.shape, dt, self.new_axes or self.axes, self.dim_res, description, data)
        if return_memmap:
            return self.get_memmap(output_path, read_mode=read_mode)
        return None

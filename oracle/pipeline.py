"""Numpy/scipy restatement of nellie's Filter + Label hot path (TEST INFRASTRUCTURE ONLY).

Every function cites the reference lines it restates (paths relative to the reference
checkout, aelefebv/nellie @ 54bf227).  The arithmetic of this path lives in numpy and
scipy.ndimage (third-party, pinned in the reference's ``uv.lock``: numpy 2.3.5, scipy
1.16.3; this image has numpy 2.3.5 / scipy 1.18.1).  The restatement therefore calls the
same library primitive at every reference call site, in the same dtype and the same
order, so that it is bit-identical to the reference by construction; that claim is then
*checked* against outputs of the executed reference (``tests/golden``, produced by
``oracle/make_golden.py``).

The functions are written as a flat, stateless pipeline (no ImInfo, no memmaps, no
device ladder): input is one frame ``(Z, Y, X)`` or ``(Y, X)`` plus a :class:`FrameSpec`.
Only the reference's full-volume (non-low-memory) branch is restated; the chunked
branches compute different numbers and are not a parity target.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np
import scipy.ndimage as ndi

F32 = np.float32


# --------------------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------------------
@dataclass
class FrameSpec:
    """Physical description of one frame; mirrors the ``im_info`` attributes the path reads
    (``no_z``, ``dim_res``) plus the Filter/Label constructor knobs
    (filtering.py:23-40, labelling.py:23-35)."""

    dim_res: dict
    no_z: bool = False
    min_radius_um: float = 0.25
    max_radius_um: float = 1.0
    alpha_sq: float = 0.5
    beta_sq: float = 0.5
    frob_thresh: Optional[float] = None
    frob_thresh_division: float = 2
    max_threshold_samples: int = 1_000_000
    truncate: float = 3.0
    remove_edges: bool = False
    sigmas: Optional[Sequence[float]] = None  # explicit override (BASELINE config #3)
    run_mask: bool = True  # Filter._run_frame(t, mask=...) (filtering.py:910): False = no Frobenius gate
    # Label knobs
    label_min_radius_um: float = 0.25
    threshold_sampling_pixels: int = 1_000_000
    histogram_nbins: int = 256
    otsu_thresh_intensity: bool = False
    threshold: Optional[float] = None

    def spacing(self):
        """filtering.py:265-275 (_get_spacing)."""
        y = float(self.dim_res.get("Y") or 1.0)
        x = float(self.dim_res.get("X") or 1.0)
        if self.no_z:
            return (y, x)
        z = float(self.dim_res.get("Z") or self.dim_res.get("X") or 1.0)
        return (z, y, x)

    def z_ratio(self):
        """filtering.py:75-78."""
        z_res = self.dim_res.get("Z") or self.dim_res.get("X") or 1.0
        x_res = self.dim_res.get("X") or 1.0
        return float(z_res) / float(x_res)


def sigma_schedule(spec: FrameSpec):
    """filtering.py:88-89, :288-316 (_set_default_sigmas)."""
    if spec.sigmas is not None:
        return sorted(float(s) for s in spec.sigmas)
    min_px = spec.min_radius_um / spec.dim_res["X"]
    max_px = spec.max_radius_um / spec.dim_res["X"]
    s_a, s_b = min_px / 2.0, max_px / 3.0
    s_lo, s_hi = min(s_a, s_b), max(s_a, s_b)
    if s_hi <= s_lo:
        s_hi = s_lo + 0.2
    step = max(0.2, (s_hi - s_lo) / 5.0)
    out = list(np.arange(s_lo, s_hi, step, dtype=float))
    out.sort()
    return [float(s) for s in out]


def sigma_vector(spec: FrameSpec, sigma: float):
    """filtering.py:277-286 (_get_sigma_vec)."""
    if spec.no_z:
        return (float(sigma), float(sigma))
    return (float(sigma) / spec.z_ratio(), float(sigma), float(sigma))


def delta_sigma_vector(spec: FrameSpec, prev_sigma: float, sigma: float):
    """filtering.py:816-825: per-axis incremental sigma of the cascade."""
    out = []
    for sp, sc in zip(sigma_vector(spec, prev_sigma), sigma_vector(spec, sigma)):
        out.append(float(np.sqrt(max(0.0, float(sc) ** 2 - float(sp) ** 2))))
    return tuple(out)


def gaussian_radius(delta_sigma: float, truncate: float = 3.0) -> int:
    """scipy.ndimage.gaussian_filter1d: ``lw = int(truncate * sd + 0.5)``."""
    return int(truncate * float(delta_sigma) + 0.5)


# --------------------------------------------------------------------------------------
# F2 / U1 / U2 : sampling lattice and histogram thresholds
# --------------------------------------------------------------------------------------
def sample_strides(shape, max_samples):
    """filtering.py:328-340 (_sample_strides)."""
    nd = len(shape)
    if max_samples is None or max_samples <= 0:
        return (1,) * nd
    total = int(np.prod(shape))
    if total <= max_samples:
        return (1,) * nd
    s0 = max(1, int(np.ceil((total / max_samples) ** (1.0 / nd))))
    st = [s0] * nd
    while int(np.prod([int(np.ceil(n / s)) for n, s in zip(shape, st)])) > max_samples:
        k = int(np.argmax([n / s for n, s in zip(shape, st)]))
        st[k] += 1
    return tuple(st)


def lattice_positive(arr, max_samples):
    """filtering.py:348-363 (_subsample_for_thresholds): strided lattice, keep > 0."""
    if arr.size == 0:
        return arr
    st = sample_strides(arr.shape, max_samples)
    sub = arr if all(s == 1 for s in st) else arr[tuple(slice(None, None, s) for s in st)]
    sub = sub[sub > 0]
    if sub.size > max_samples and sub.size > 0:  # never triggers; kept for fidelity
        sub = sub[:: max(1, sub.size // max_samples)]
    return sub


def otsu(values, nbins=256):
    """utils/gpu_functions.py:23-50 (otsu_threshold); returns the bin centre (f32)."""
    flat = values.reshape(-1)
    counts, edges = np.histogram(flat, bins=nbins, range=(flat.min(), flat.max()))
    centers = (edges[:-1] + edges[1:]) / 2.0
    p = counts / np.sum(counts)
    w_lo = np.cumsum(p)
    m_lo = np.cumsum(p * centers) / w_lo
    w_hi = np.cumsum(p[::-1])[::-1]
    m_hi = (np.cumsum((p * centers)[::-1]) / w_hi[::-1])[::-1]
    between = w_lo[:-1] * w_hi[1:] * (m_lo[:-1] - m_hi[1:]) ** 2
    return centers[np.argmax(between)]


def triangle(values, nbins=256):
    """utils/gpu_functions.py:53-94 (triangle_threshold); returns the bin centre (f32)."""
    flat = values.reshape(-1)
    hist, edges = np.histogram(flat, bins=nbins, range=(np.min(flat), np.max(flat)))
    centers = (edges[:-1] + edges[1:]) / 2.0
    hist = hist / np.sum(hist)
    i_peak = np.argmax(hist)
    h_peak = hist[i_peak]
    i_lo, i_hi = np.flatnonzero(hist)[[0, -1]]
    flipped = i_peak - i_lo < i_hi - i_peak
    if flipped:
        hist = np.flip(hist, axis=0)
        i_lo = nbins - i_hi - 1
        i_peak = nbins - i_peak - 1
    width = i_peak - i_lo
    xs = np.arange(width)
    ys = hist[xs + i_lo]
    nrm = np.sqrt(h_peak ** 2 + width ** 2)
    h_peak = h_peak / nrm
    width = width / nrm
    i_lvl = np.argmax(h_peak * xs - width * ys) + i_lo
    if flipped:
        i_lvl = nbins - i_lvl - 1
    return centers[i_lvl]


def gamma_of(gauss, spec: FrameSpec) -> float:
    """filtering.py:365-380 (_calculate_gamma)."""
    pos = lattice_positive(gauss, spec.max_threshold_samples)
    if pos.size == 0:
        return float(np.finfo(np.float32).eps)
    g = float(min(triangle(pos), otsu(pos)))
    if g <= 0:
        g = float(np.finfo(np.float32).eps)
    return g


# --------------------------------------------------------------------------------------
# F1 : cascaded Gaussian
# --------------------------------------------------------------------------------------
def gauss_step(gauss, spec: FrameSpec, prev_sigma: float, sigma: float):
    """filtering.py:827-835: in-place incremental blur prev_sigma -> sigma."""
    dvec = delta_sigma_vector(spec, prev_sigma, sigma)
    if any(s > 0 for s in dvec):
        ndi.gaussian_filter(gauss, sigma=dvec, output=gauss, mode="reflect", cval=0.0,
                            truncate=spec.truncate)
    return gauss


# --------------------------------------------------------------------------------------
# F4 / F5 : finite-difference Hessian, Frobenius mask
# --------------------------------------------------------------------------------------
def hessian(gauss, spec: FrameSpec):
    """filtering.py:446-562 (_compute_hessian, non-low-memory branch).

    Returns (components dict in the reference's naming, frob_sq, max_abs, frob)."""
    img = gauss.astype(F32, copy=False)
    sp = spec.spacing()
    if img.ndim == 2:
        d0, d1 = np.gradient(img, *sp)
        comp = {
            "hxx": np.gradient(d0, sp[0], axis=0).astype(F32, copy=False),
            "hxy": np.gradient(d0, sp[1], axis=1).astype(F32, copy=False),
            "hyy": np.gradient(d1, sp[1], axis=1).astype(F32, copy=False),
        }
        frob_sq = comp["hxx"] ** 2 + comp["hyy"] ** 2 + 2.0 * (comp["hxy"] ** 2)
    elif img.ndim == 3:
        d0, d1, d2 = np.gradient(img, *sp)
        comp = {
            "hxx": np.gradient(d0, sp[0], axis=0).astype(F32, copy=False),
            "hxy": np.gradient(d0, sp[1], axis=1).astype(F32, copy=False),
            "hxz": np.gradient(d0, sp[2], axis=2).astype(F32, copy=False),
            "hyy": np.gradient(d1, sp[1], axis=1).astype(F32, copy=False),
            "hyz": np.gradient(d1, sp[2], axis=2).astype(F32, copy=False),
            "hzz": np.gradient(d2, sp[2], axis=2).astype(F32, copy=False),
        }
        frob_sq = (comp["hxx"] ** 2 + comp["hyy"] ** 2 + comp["hzz"] ** 2
                   + 2.0 * (comp["hxy"] ** 2 + comp["hxz"] ** 2 + comp["hyz"] ** 2))
    else:
        raise ValueError("frame must be 2-D or 3-D")
    max_abs = 0.0
    for c in comp.values():
        if c.size > 0:
            max_abs = max(max_abs, float(np.max(np.abs(c))))
    if max_abs <= 0:
        max_abs = 1.0
    frob = np.sqrt(frob_sq) / max_abs
    return comp, frob_sq, max_abs, frob


def frob_threshold(frob, spec: FrameSpec) -> float:
    """filtering.py:432-441: min(triangle, otsu) of the positive lattice sample (or fixed)."""
    if spec.frob_thresh is not None:
        return float(spec.frob_thresh)
    pos = lattice_positive(frob, spec.max_threshold_samples)
    if pos.size == 0:
        return 0.0
    return float(min(triangle(pos), otsu(pos)))


def frob_mask(frob, spec: FrameSpec):
    """filtering.py:407-444 (_get_frob_mask). Returns (mask, threshold_before_division)."""
    inf = np.isinf(frob)
    if np.any(inf):
        fin = frob[~inf]
        top = float(np.max(fin)) if fin.size > 0 else 0.0
        frob = frob.copy()
        frob[inf] = top
    if not spec.frob_thresh_division:
        return frob > 0, 0.0
    thr = frob_threshold(frob, spec)
    return frob > (thr / spec.frob_thresh_division), thr


# --------------------------------------------------------------------------------------
# F7 / F8 : eigenvalues and vesselness
# --------------------------------------------------------------------------------------
def eig_sorted_3d(comp, where):
    """filtering.py:693-707 + :580-585: eigvalsh (f64 inside numpy, f32 result) sorted by |.|."""
    h = np.stack([
        np.stack([comp["hxx"][where], comp["hxy"][where], comp["hxz"][where]], axis=-1),
        np.stack([comp["hxy"][where], comp["hyy"][where], comp["hyz"][where]], axis=-1),
        np.stack([comp["hxz"][where], comp["hyz"][where], comp["hzz"][where]], axis=-1),
    ], axis=-2)
    ev = np.linalg.eigvalsh(h)
    return np.take_along_axis(ev, np.argsort(np.abs(ev), axis=1), axis=1)


def eig_sorted_2d(comp, where):
    """filtering.py:676-690: closed-form 2x2, all float32."""
    a, b, d = comp["hxx"][where], comp["hxy"][where], comp["hyy"][where]
    tr = a + d
    df = a - d
    root = np.sqrt(df * df + 4.0 * (b * b))
    lo = 0.5 * (tr - root)
    hi = 0.5 * (tr + root)
    swap = np.abs(lo) > np.abs(hi)
    return np.stack([np.where(swap, hi, lo), np.where(swap, lo, hi)], axis=1)


def vesselness(ev, spec: FrameSpec, gamma_sq: float):
    """filtering.py:717-767 (_filter_hessian)."""
    with np.errstate(all="ignore"):
        if spec.no_z:
            l1, l2 = ev[:, 0], ev[:, 1]
            rb_sq = (np.abs(l1) / (np.abs(l2) + 1e-12)) ** 2
            s_sq = l1 ** 2 + l2 ** 2
            v = np.exp(-(rb_sq / spec.beta_sq)) * (1.0 - np.exp(-(s_sq / gamma_sq)))
        else:
            l1, l2, l3 = ev[:, 0], ev[:, 1], ev[:, 2]
            ra_sq = (np.abs(l2) / (np.abs(l3) + 1e-12)) ** 2
            rb_sq = (np.abs(l2) / (np.sqrt(np.abs(l2 * l3)) + 1e-12)) ** 2
            s_sq = l1 ** 2 + l2 ** 2 + l3 ** 2
            v = ((1.0 - np.exp(-(ra_sq / spec.alpha_sq)))
                 * np.exp(-(rb_sq / spec.beta_sq))
                 * (1.0 - np.exp(-(s_sq / gamma_sq))))
    if not spec.no_z:
        v[ev[:, 2] > 0] = 0.0
    v[ev[:, 1] > 0] = 0.0
    return np.nan_to_num(v, nan=0.0, posinf=0.0, neginf=0.0)


def vesselness_volume(comp, mask, spec: FrameSpec, gamma_sq: float):
    """filtering.py:651-715 (_compute_vesselness_chunkwise); chunking does not change values."""
    where = np.where(mask)
    out = np.zeros_like(next(iter(comp.values())), dtype=F32)
    if where[0].size == 0:
        return out
    ev = eig_sorted_2d(comp, where) if spec.no_z else eig_sorted_3d(comp, where)
    out[where] = vesselness(ev, spec, gamma_sq).astype(F32, copy=False)
    return out


# --------------------------------------------------------------------------------------
# F9 / F10 : per-frame response
# --------------------------------------------------------------------------------------
def log_blobness(blurred, and_mask, spec: FrameSpec, sigmas):
    """filtering.py:772-795 (_filter_log) — 2-D only; ``blurred`` is the sigma_max-blurred
    frame because of the reference's aliasing."""
    frame = blurred.astype(F32, copy=False)
    acc = None
    for i, s in enumerate(sigmas):
        cur = -ndi.gaussian_laplace(frame, sigma_vector(spec, s)) * (float(s) ** 2)
        cur = cur * and_mask
        if i == 0:
            acc = cur
        else:
            better = cur > acc
            acc[better] = cur[better]
    acc[acc < 0] = 0.0
    top = np.max(acc)
    return (acc / (top + 1e-12)) / 10.0


def frangi_frame(frame, spec: FrameSpec, trace: Optional[list] = None):
    """filtering.py:806-853 (_compute_vesselness) + :924-933 (_run_frame, full-volume).

    ``trace`` (optional list) receives one dict per sigma with the intermediates the
    CUDA kernels are checked against."""
    sigmas = sigma_schedule(spec)
    gauss = np.array(frame, dtype=F32, copy=True)  # never mutate the caller's array (App. C-1)
    response = np.zeros_like(gauss, dtype=F32)
    alive = np.ones_like(gauss, dtype=bool)
    prev = 0.0
    for s in sigmas:
        gauss_step(gauss, spec, prev, s)
        prev = s
        gamma = gamma_of(gauss, spec)
        gamma_sq = 2.0 * (float(gamma) ** 2)
        comp, frob_sq, max_abs, frob = hessian(gauss, spec)
        m, thr = frob_mask(frob, spec)
        if not spec.run_mask:                 # filtering.py:563-566: h_mask = ones when mask=False
            m = np.ones_like(gauss, dtype=bool)
        rec = None
        if trace is not None:
            rec = dict(sigma=s, gauss=gauss.copy(), gamma=gamma, gamma_sq=gamma_sq, max_abs=max_abs,
                       frob_thr=thr, mask=m.copy(), comp={k: v.copy() for k, v in comp.items()},
                       skipped=not bool(np.any(m)))
            trace.append(rec)
        if not np.any(m):
            continue  # filtering.py:843-844 — note: the AND below is skipped too
        v = vesselness_volume(comp, m, spec, gamma_sq)
        if rec is not None:
            rec["vessel"] = v.copy()
        response = np.maximum(response, v)
        alive &= m
    out = response * alive
    if spec.no_z:
        blob = np.maximum(log_blobness(gauss, alive, spec, sigmas), 0)
        out = np.maximum(out, blob)
    if spec.remove_edges:
        out = remove_edges(out, spec)
    return out


def _bbox_rows(sl):
    rows = np.any(sl, axis=1)
    cols = np.any(sl, axis=0)
    if (not rows.any()) or (not cols.any()):
        return 0, 0
    r = np.where(rows)[0]
    return int(r[0]), int(r[-1])


def remove_edges(v, spec: FrameSpec):
    """filtering.py:969-1000 (_remove_edges): zero 15-row bands at the bbox top/bottom."""
    planes = [v] if spec.no_z else [v[z] for z in range(v.shape[0])]
    for sl in planes:
        if sl.size == 0:
            continue
        r0, r1 = _bbox_rows(sl)
        height = max(0, r1 - r0 + 1)
        if height <= 0:
            continue
        m = min(15, height)
        sl[r0:r0 + m, :] = 0
        sl[r1 - m + 1:r1 + 1, :] = 0
    return v


def finalize_mask(v, spec: FrameSpec):
    """filtering.py:952-967 (_mask_volume) guarded as in :1014-1018 (_run_filter)."""
    if not float(np.sum(v)) > 0.0:
        return v
    pos = lattice_positive(v, spec.max_threshold_samples)
    if pos.size == 0:
        return v
    thr = np.percentile(pos, 1)
    keep = ndi.binary_opening(v > thr)
    return v * keep


def filter_frame(frame, spec: FrameSpec, trace=None):
    """One iteration of filtering.py:1007-1031 (_run_filter) without the memmap write."""
    return finalize_mask(frangi_frame(frame, spec, trace=trace), spec)


# --------------------------------------------------------------------------------------
# Label (L1-L5)
# --------------------------------------------------------------------------------------
def label_min_radius_um(spec: FrameSpec) -> float:
    """labelling.py:95-97."""
    return max(float(spec.label_min_radius_um), float(spec.dim_res.get("X") or 1.0))


def label_min_area(spec: FrameSpec) -> int:
    """labelling.py:209-219 (_compute_min_area_pixels)."""
    x = spec.dim_res.get("X") or 1.0
    y = spec.dim_res.get("Y") or x
    r = label_min_radius_um(spec)
    if spec.no_z:
        return max(1, int(np.ceil(np.pi * (r ** 2) / (float(x) * float(y)))))
    z = spec.dim_res.get("Z") or x
    vol = (4.0 / 3.0) * np.pi * (r ** 3)
    return max(1, int(np.ceil(vol / (float(x) * float(y) * float(z)))))


def label_sample(frame, spec: FrameSpec, gate_frame=None, gate_thresh=None):
    """labelling.py:385-438 (_sample_nonzero)."""
    flat = frame.reshape(-1)
    if flat.size == 0:
        return flat
    gate = gate_frame.reshape(-1) if (gate_frame is not None and gate_thresh is not None) else None
    step = max(int(flat.size) // max(1, int(spec.threshold_sampling_pixels)), 1)
    offsets = (0, step // 2) if step > 1 and step // 2 > 0 else (0,)
    vals = flat[:0]
    for off in offsets:
        s = flat[off::step]
        if gate is not None:
            vals = s[(s > 0) & (gate[off::step] > gate_thresh)]
        else:
            vals = s[s > 0]
        if vals.size > 0 or step == 1:
            return vals
    if float(flat.max()) <= 0:
        return vals
    if gate is not None:
        return flat[(flat > 0) & (gate > gate_thresh)]
    return flat[flat > 0]


def label_frangi_threshold(frangi, spec: FrameSpec, gate_frame=None, gate_thresh=None):
    """labelling.py:440-455 (_compute_frangi_threshold)."""
    vals = label_sample(frangi, spec, gate_frame, gate_thresh)
    if vals.size == 0:
        return None
    lv = np.log10(vals)
    t = 10 ** triangle(lv, nbins=spec.histogram_nbins)
    o = 10 ** otsu(lv, nbins=spec.histogram_nbins)
    return min(t, o)


def label_thresholds(raw, frangi, spec: FrameSpec):
    """labelling.py:511-532 (_compute_frame_thresholds)."""
    it = None
    if spec.otsu_thresh_intensity:
        vals = label_sample(raw, spec)
        it = otsu(vals, nbins=spec.histogram_nbins) if vals.size else 0
    elif spec.threshold is not None:
        it = spec.threshold
    if it is not None:
        return it, label_frangi_threshold(frangi, spec, gate_frame=raw, gate_thresh=it)
    return it, label_frangi_threshold(frangi, spec)


def label_frame(frangi, spec: FrameSpec, frangi_thresh, raw=None, intensity_thresh=None,
                stages: Optional[dict] = None):
    """labelling.py:467-509 (_get_labels) + :546-556 (_run_frame_full_volume)."""
    frangi = np.asarray(frangi)
    if intensity_thresh is not None:
        frangi = frangi * (np.asarray(raw) > intensity_thresh)
    structure = np.ones((3,) * frangi.ndim, dtype=bool)
    mask = np.zeros_like(frangi, dtype=bool) if frangi_thresh is None else frangi > frangi_thresh
    if not spec.no_z:
        mask = ndi.binary_fill_holes(mask)
    if stages is not None:
        stages["filled"] = mask.copy()
    labels, _ = ndi.label(mask, structure=structure)
    if stages is not None:
        stages["labels_first"] = labels.copy()
    if labels.size == 0:
        return labels
    areas = np.bincount(labels.ravel())
    if areas.size <= 1:
        return labels
    areas[0] = 0
    keep = areas >= label_min_area(spec)
    mask = keep[labels]
    if stages is not None:
        stages["kept"] = mask.copy()
    mask = ndi.uniform_filter(mask.astype(F32), size=3) > 0.5
    if stages is not None:
        stages["smoothed"] = mask.copy()
    labels, _ = ndi.label(mask, structure=structure)
    return labels


def segment_frame(raw, spec: FrameSpec):
    """Filter then Label for one frame, as nellie.run does per timepoint (run.py:56-73)."""
    fr = filter_frame(raw, spec)
    it, ft = label_thresholds(raw, fr, spec)
    return fr, label_frame(fr, spec, ft, raw=raw, intensity_thresh=it)


# ---------------------------------------------------------------------------------------------
# Network stage, array kernels of its GPU backend — restated for the parity tests
# ---------------------------------------------------------------------------------------------
def network_pixel_class(skel, no_z: bool):
    """networking.py:669-680 (_get_pixel_class_impl)."""
    m = (np.asarray(skel) > 0).astype(np.uint8)
    w = np.ones((3, 3) if no_z else (3, 3, 3))
    s = ndi.convolve(m, weights=w, mode="constant", cval=0) * m
    s[s > 4] = 4
    return s


def network_branch_labels(pixel_class, no_z: bool):
    """networking.py:758-797 (_get_branch_skel_labels)."""
    pc = np.asarray(pixel_class)
    lab, _ = ndi.label((pc > 0) & (pc != 4), structure=np.ones((3, 3) if no_z else (3, 3, 3)))
    return lab


def network_remove_connected(labels, no_z: bool):
    """networking.py:261-296 (_remove_connected_label_pixels_impl)."""
    labels = np.asarray(labels)
    size = (3, 3) if no_z else (3, 3, 3)
    mx = ndi.maximum_filter(labels, size=size, mode="constant", cval=0)
    bg = int(labels.max()) + 1
    mn = ndi.minimum_filter(np.where(labels == 0, bg, labels), size=size, mode="constant", cval=bg)
    mn = np.where(mn == bg, 0, mn)
    amb = (labels > 0) & (mn > 0) & (mx > 0) & (mn != mx)
    inner = np.zeros(labels.shape, bool)
    inner[tuple(slice(1, -1) for _ in labels.shape)] = True
    return np.where(amb & inner, 0, labels)


# ---------------------------------------------------------------------------------------------
# Markers stage: full-volume branch of nellie/segmentation/mocap_marking.py
# ---------------------------------------------------------------------------------------------
@dataclass
class MarkerSpec:
    """Constructor knobs of ``Markers`` (mocap_marking.py:84-160) + the ``im_info`` attributes it reads."""

    dim_res: dict
    no_z: bool = False
    min_radius_um: float = 0.20
    max_radius_um: float = 1.0
    use_im: str = "distance"
    num_sigma: int = 5
    peak_min_distance: int = 2

    def x_res(self):
        return self.dim_res.get("X") or 1.0

    def z_ratio(self):
        """mocap_marking.py:124-130."""
        if self.no_z:
            return 1.0
        z_res = self.dim_res.get("Z") or self.x_res()
        return float(z_res) / float(self.x_res())

    def radii_px(self):
        """mocap_marking.py:132-135: (min_radius_px, max_radius_px)."""
        min_um = max(self.min_radius_um, float(self.x_res()))
        return min_um / float(self.x_res()), self.max_radius_um / float(self.x_res())


def marker_sigmas(spec: MarkerSpec):
    """mocap_marking.py:340-378 (_set_default_sigmas)."""
    min_px, max_px = spec.radii_px()
    sigma_min = min_px / 2.0
    sigma_max = max_px / 3.0
    sigma_range = sigma_max - sigma_min
    if sigma_range <= 0:
        return [sigma_min]
    step = max(0.2, sigma_range / max(spec.num_sigma, 1))
    sigmas = list(np.arange(sigma_min, sigma_max, step))
    return sigmas if sigmas else [sigma_min]


def marker_sigma_vec(spec: MarkerSpec, sigma):
    """mocap_marking.py:318-338 (_get_sigma_vec)."""
    return (sigma, sigma) if spec.no_z else (sigma / spec.z_ratio(), sigma, sigma)


def marker_distance(mask, spec: MarkerSpec):
    """mocap_marking.py:419-450 (_distance_im): (distance float32 clamped to 2*max_radius_px, border shell bool)."""
    mask = np.asarray(mask, dtype=bool)
    border = ndi.binary_dilation(mask, iterations=1) ^ mask
    distance = ndi.distance_transform_edt(mask)
    distance = distance.astype(F32, copy=False)
    np.minimum(distance, spec.radii_px()[1] * 2.0, out=distance)
    return distance, border


def marker_log_response(use_im, spec: MarkerSpec, sigma):
    """mocap_marking.py:488-494: scale-normalised negated LoG of one sigma, negatives clamped, float32."""
    sigma_val = float(sigma)
    resp = -ndi.gaussian_laplace(use_im, marker_sigma_vec(spec, sigma_val))
    resp = (resp * (sigma_val ** 2)).astype(F32, copy=False)
    resp[resp < 0] = 0
    return resp


def marker_peaks(use_im, mask, distance, spec: MarkerSpec, sigmas=None):
    """mocap_marking.py:452-512 (_local_max_peak, full-volume branch): bool peak mask (the reference returns
    ``argwhere`` of it) and the best response per voxel."""
    valid = np.asarray(mask, dtype=bool) & (distance > 0)
    best = np.zeros_like(use_im, dtype=F32)
    peak = np.zeros_like(use_im, dtype=bool)
    for s in (marker_sigmas(spec) if sigmas is None else sigmas):
        resp = marker_log_response(use_im, spec, s)
        local_max = resp == ndi.maximum_filter(resp, size=3, mode="nearest")
        local_max &= valid
        better = local_max & (resp > best)
        peak[better] = True
        best[better] = resp[better]
    return peak, best


def marker_nms(peak, intensity, spec: MarkerSpec):
    """mocap_marking.py:569-606 (_remove_close_peaks, full-volume branch): bool mask of the kept peaks."""
    coords = np.argwhere(peak)
    if coords.size == 0:
        return np.zeros(peak.shape, bool)
    score = np.zeros_like(intensity, dtype=F32)
    score[tuple(coords.T)] = intensity[tuple(coords.T)]
    size = 2 * int(spec.peak_min_distance) + 1
    mx = ndi.maximum_filter(score, size=size, mode="nearest")
    return (score == mx) & (score > 0)


def marker_frame(intensity, labels, spec: MarkerSpec, frangi=None, sigmas=None):
    """mocap_marking.py:648-703 (_run_frame_impl): (marker uint8, distance float32, border uint8)."""
    intensity = np.asarray(intensity)
    mask = (np.asarray(labels) > 0).astype(bool, copy=False)
    if not mask.any():
        return (np.zeros(intensity.shape, np.uint8), np.zeros(intensity.shape, F32), np.zeros(intensity.shape, np.uint8))
    distance, border = marker_distance(mask, spec)
    if spec.use_im == "distance":
        base = distance
    elif spec.use_im == "frangi":
        if frangi is None:
            raise RuntimeError("Frangi image requested for peak detection but not available.")
        base = np.asarray(frangi)
    else:
        raise ValueError(f"Unknown use_im value: {spec.use_im}")
    peak, _ = marker_peaks(base, mask, distance, spec, sigmas)
    keep = marker_nms(peak, intensity, spec)
    return keep.astype(np.uint8), distance, border.astype(np.uint8)


# ---------------------------------------------------------------------------------------------
# HuMomentTracking, per-frame feature extraction: nellie/tracking/hu_tracking.py:225-392, :585-750
# ---------------------------------------------------------------------------------------------
def hu_transform_frangi(frangi):
    """hu_tracking.py:604-612: log10 of the positive response, then the negative values shifted so their minimum is 0."""
    f = np.asarray(frangi).copy()
    pos = f > 0
    if np.any(pos):
        f[pos] = np.log10(f[pos])
    neg = f < 0
    if np.any(neg):
        f[neg] -= np.min(f[neg])
    return f


def hu_distance_max(distance):
    """hu_tracking.py:614-616: 3^d maximum filter of the distance frame, doubled."""
    d = np.asarray(distance).copy()
    ndi.maximum_filter(d, size=3, output=d)
    d *= 2
    return d


def hu_bounds(markers, distance_max, frame_shape):
    """hu_tracking.py:392-421 (_get_im_bounds): per axis (low, high) as integer arrays (the reference keeps floats and
    applies int() at the point of use)."""
    radii = distance_max[tuple(markers.T)]
    r = np.ceil(radii)
    out = []
    for a, size in enumerate(frame_shape):
        out.append(np.clip(markers[:, a] - r, 0, size).astype(np.int64))
        out.append(np.clip(markers[:, a] + (r + 1), 0, size).astype(np.int64))
    return out


def hu_mean_and_variance(images):
    """hu_tracking.py:341-390 (_calculate_mean_and_variance) for (N, ...) ROI stacks."""
    if images.size == 0:
        return np.zeros((0, 2), F32)
    feats = np.zeros((images.shape[0], 2), F32)
    mask = images != 0
    axis = tuple(range(1, images.ndim))
    count = np.sum(mask, axis=axis)
    safe = np.where(count == 0, 1, count)
    s = np.sum(images * mask, axis=axis)
    ss = np.sum((images * mask) ** 2, axis=axis)
    mean = s / safe
    var = (ss - (s ** 2) / safe) / safe
    feats[:, 0] = np.where(count == 0, 0.0, mean)
    feats[:, 1] = np.where(count == 0, 0.0, var)
    return feats


def hu_normalized_moments(images, real=np.float64):
    """hu_tracking.py:225-268 (_calculate_normalized_moments), images (N, H, W).  ``real``: the float type of the
    arithmetic (the reference: float64; tests pass np.longdouble to find the entries whose float64 value is rounding noise)."""
    n, h, w = images.shape
    ext = images[:, :, :, None, None]
    if real is not np.float64:
        ext = ext.astype(real)
    x, y = np.meshgrid(np.arange(w), np.arange(h))
    x = x[None, :, :, None, None]
    y = y[None, :, :, None, None]
    powers = np.arange(4)
    px = powers[None, None, None, :, None]
    py = powers[None, None, None, None, :]
    M = np.sum(ext * (x ** px) * (y ** py), axis=(1, 2))
    eps = real(1e-12)
    x_bar = (M[:, 1, 0] / (M[:, 0, 0] + eps))[:, None, None, None, None]
    y_bar = (M[:, 0, 1] / (M[:, 0, 0] + eps))[:, None, None, None, None]
    mu = np.sum(ext * (x - x_bar) ** px * (y - y_bar) ** py, axis=(1, 2))
    ipj = np.arange(4)[:, None] + np.arange(4)[None, :]
    denom = (M[:, 0, 0][:, None, None] ** ((ipj[None, :, :] + 2) / real(2.0))) + eps
    return mu / denom


def hu_moments(eta):
    """hu_tracking.py:270-312 (_calculate_hu_moments): first six Hu invariants."""
    hu = np.zeros((eta.shape[0], 6), dtype=eta.dtype)
    e20, e02, e11 = eta[:, 2, 0], eta[:, 0, 2], eta[:, 1, 1]
    e30, e12, e21, e03 = eta[:, 3, 0], eta[:, 1, 2], eta[:, 2, 1], eta[:, 0, 3]
    hu[:, 0] = e20 + e02
    hu[:, 1] = (e20 - e02) ** 2 + 4 * e11 ** 2
    hu[:, 2] = (e30 - 3 * e12) ** 2 + (3 * e21 - e03) ** 2
    hu[:, 3] = (e30 + e12) ** 2 + (e21 + e03) ** 2
    hu[:, 4] = ((e30 - 3 * e12) * (e30 + e12) * ((e30 + e12) ** 2 - 3 * (e21 + e03) ** 2) +
                (3 * e21 - e03) * (e21 + e03) * (3 * (e30 + e12) ** 2 - (e21 + e03) ** 2))
    hu[:, 5] = ((e20 - e02) * ((e30 + e12) ** 2 - (e21 + e03) ** 2) + 4 * e11 * (e30 + e12) * (e21 + e03))
    return hu


def hu_log(hu):
    """hu_tracking.py:314-325 (_log_hu)."""
    if hu.size == 0:
        return hu
    a = np.maximum(np.abs(hu), np.finfo(hu.dtype).tiny)
    out = -np.sign(hu) * np.log10(a)
    return np.where(np.isfinite(out), out, 0.0)


def hu_of_subvolumes(sub, no_z, real=np.float64):
    """hu_tracking.py:544-571 (_get_hu_moments): 6 invariants of a 2-D ROI, 18 of the three max projections of a 3-D ROI."""
    if no_z:
        return hu_moments(hu_normalized_moments(sub, real))
    return np.concatenate([hu_moments(hu_normalized_moments(np.max(sub, axis=a), real)) for a in (1, 2, 3)], axis=1)


def hu_frame_features(intensity, frangi, distance, marker, scaling, no_z, dense=True, real=np.float64):
    """hu_tracking.py:585-680 (_get_frame_features_impl): (coords_voxel, coords_phys, stats (N, 4) float32, log-Hu (N, 6 | 18)).
    dense=True: the batched zero-padded ROI cube of :641-657; dense=False: the per-ROI streaming path of :682-750."""
    intensity = np.asarray(intensity)
    fr = hu_transform_frangi(frangi)
    dmax = hu_distance_max(distance)
    marker_mask = np.asarray(marker) > 0
    coords = np.argwhere(marker_mask)
    dims = 2 if no_z else 3
    if coords.size == 0:
        return np.zeros((0, dims), int), np.zeros((0, dims), float), np.zeros((0, 0), F32), np.zeros((0, 0), F32)
    phys = coords * np.asarray(scaling, dtype=float)
    b = hu_bounds(coords, dmax, intensity.shape)
    n = coords.shape[0]
    R = int(np.ceil(np.max(dmax[marker_mask])).item()) * 2 + 1
    hu_dim = 6 if no_z else 18
    if dense:
        def gather(frame):
            sub = np.zeros((n,) + (R,) * dims, dtype=frame.dtype)
            for i in range(n):
                lo = [int(b[2 * a][i]) for a in range(dims)]
                hi = [int(b[2 * a + 1][i]) for a in range(dims)]
                if any(l >= h for l, h in zip(lo, hi)):
                    continue
                sub[(i,) + tuple(slice(0, h - l) for l, h in zip(lo, hi))] = frame[tuple(slice(l, h) for l, h in zip(lo, hi))]
            return sub
        isub, fsub = gather(intensity), gather(fr)
        stats = np.concatenate([hu_mean_and_variance(isub), hu_mean_and_variance(fsub)], axis=1)
        return coords.astype(int), phys, stats, hu_log(hu_of_subvolumes(isub, no_z, real))
    stats = np.zeros((n, 4), F32)
    log_hu = np.zeros((n, hu_dim), F32)
    for i in range(n):
        lo = [int(b[2 * a][i]) for a in range(dims)]
        hi = [int(b[2 * a + 1][i]) for a in range(dims)]
        if any(l >= h for l, h in zip(lo, hi)):
            continue
        sl = tuple(slice(l, h) for l, h in zip(lo, hi))
        iroi, froi = intensity[sl][None], fr[sl][None]
        stats[i] = np.concatenate((hu_mean_and_variance(iroi)[0], hu_mean_and_variance(froi)[0]), axis=0)
        log_hu[i] = hu_log(hu_of_subvolumes(iroi, no_z, real)[0])
    return coords.astype(int), phys, stats, log_hu


# ---------------------------------------------------------------------------------------------
# Network stage, host steps: networking.py:315-392, :485-577, :825-851
# ---------------------------------------------------------------------------------------------
def network_add_missing(skel, labels, frangi):
    """networking.py:315-392 (_add_missing_skeleton_labels): objects without a skeleton voxel get one at the position of
    their largest Frangi response."""
    skel = np.array(skel, copy=True)
    labels = np.asarray(labels)
    missing = np.setdiff1d(np.unique(labels), np.unique(skel))
    missing = missing[missing != 0]
    if missing.size == 0:
        return skel
    positions = ndi.maximum_position(np.asarray(frangi), labels=labels, index=missing)
    for lab, pos in zip(missing, positions):
        skel[tuple(int(p) for p in pos)] = lab
    return skel


def network_relabel_objects(branch, labels, scaling):
    """networking.py:485-577 (_relabel_objects): per object, every voxel takes the branch label of the nearest seed."""
    labels = np.asarray(labels).astype(np.int32, copy=False)
    branch = np.asarray(branch).astype(np.int32, copy=False)
    out = np.zeros_like(labels, dtype=np.uint32)
    max_label = int(labels.max())
    if max_label == 0:
        return out
    slices = ndi.find_objects(labels)
    for lab in range(1, max_label + 1):
        sl = slices[lab - 1]
        if sl is None:
            continue
        obj = labels[sl] == lab
        sub_branch = branch[sl]
        seeds = (sub_branch > 0) & obj
        if not seeds.any():
            continue
        idx = ndi.distance_transform_edt(np.logical_not(seeds), sampling=scaling, return_distances=False,
                                         return_indices=True)
        nearest = sub_branch[tuple(idx)]
        nearest[~obj] = 0
        sub = out[sl]
        sub[obj] = nearest[obj].astype(np.uint32, copy=False)
        out[sl] = sub
    return out


def network_frame(labels, frangi, skeleton_mask, scaling, no_z):
    """networking.py:825-851 (_run_frame_backend) given the skeleton mask the host thinning produced (:394-410:
    skel_frame = labels * skeletonize(labels > 0)): (branch_skel_labels, pixel_class, branch_labels)."""
    labels = np.asarray(labels)
    skel = labels * np.asarray(skeleton_mask).astype(bool)
    skel = network_remove_connected(skel, no_z)
    skel = network_add_missing(skel, labels, frangi)
    skel_pre = (skel > 0) * labels
    pixel_class = network_pixel_class(skel_pre, no_z)
    branch = network_branch_labels(pixel_class, no_z)
    return branch, pixel_class, network_relabel_objects(branch, labels, scaling)


def tie_free(frangi):
    """Test inputs only: the phantoms' Frangi frames hold exactly equal maxima inside one object (mirror-symmetric tubes), and
    scipy.ndimage.maximum_position picks among equal values by an unstable sort (arbitrary, not reproducible).  Scaling every
    voxel by a factor that depends on its index makes the maxima unique without changing the character of the data."""
    f = np.asarray(frangi, dtype=F32)
    idx = np.arange(f.size, dtype=np.uint64).reshape(f.shape)
    h = ((idx * np.uint64(2654435761)) % np.uint64(1 << 20)).astype(np.float64) / float(1 << 20)
    return (f * (1.0 + h / 64.0)).astype(F32)

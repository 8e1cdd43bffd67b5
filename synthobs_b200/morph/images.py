"""synthobs.morph.images — same public API as the reference module
(reference: synthobs/morph/images.py) for the particle -> image hot path:
rebin, core, observed(...).particle, particle.  The per-particle Python loops of
the reference are CUDA kernels here (include/synthobs_b200.h); there is no CPU
path.

Deliberate behaviours kept from the reference:
  * observed.particle recentres the caller's X, Y IN PLACE (images.py:183-187);
  * particle(..., offsets=True) draws from the global np.random (images.py:265-267);
  * particle() passes yoffset_pix = xoffset_pix (images.py:282);
  * img.data IS img.simple when there is no smoothing (images.py:103);
  * pixel pitch is width/(ndim-1) because G = linspace(-w/2, w/2, ndim) (images.py:53).
Difference, stated: the adaptive Gaussian is truncated at SUPPORT_SIGMAS sigma
(default 7: the dropped tail is < 2.3e-11 of a particle's peak) and normalised
over that truncated box, so total flux is conserved exactly as in the reference;
SUPPORT_SIGMAS = 0 evaluates every pixel like the reference does.  ('adaptive', 1)
raises IndexError in the reference (images.py:89 on a scalar); here it means
FWHM = 0.01 for every particle.
"""

from ctypes import c_double, c_int64

import numpy as np
import torch

from .. import _cabi
from .. import device as dv

try:
    import flare.observatories as _obs   # type: ignore
    filter_info = _obs.filter_info
except Exception:
    filter_info = {}     # offline: callers register {'<filter>': {'pixel_scale': arcsec}}


class _PixelScaleView:
    """flare.filters.pixel_scale[filter] (used by Sersic / point, images.py:300,323) read
    through filter_info so offline callers register a filter once."""

    def __getitem__(self, f):
        return filter_info[f]['pixel_scale']


try:
    import flare.filters as _flt   # type: ignore
    pixel_scale_of = _flt.pixel_scale
except Exception:
    pixel_scale_of = _PixelScaleView()

SUPPORT_SIGMAS = 7.0
MAX_PAIRS = 1 << 30        # (particle, sub-tile) pairs per binning pass; larger jobs are split by particle halves


class empty():
    pass


def _report(title, rows):
    """the verbose=True listing of the reference (images.py:159-178, 343-358): '***** TITLE', then 'name: value' lines
    ('{:.2f}' unless another format is given), None = a rule of dashes"""
    print('*' * 5, title)
    for row in rows:
        if row is None:
            print('-' * 10)
        else:
            name, value, fmt = row
            print(name + ': ' + (fmt or '{0:.2f}').format(value))


# --------------------------------------------------------------------------
# building blocks on device tensors
# --------------------------------------------------------------------------

def pixel_grid(resolution, ndim):
    width = ndim * resolution                                   # images.py:39
    return width, np.linspace(-width / 2., width / 2., ndim)    # images.py:53


_grid_cache = {}


def pixel_grid_device(resolution, ndim):
    """(width, G, G on the device).  The device copy is cached per (resolution, ndim, device): a pageable
    host-to-device copy synchronises the stream, which would stall every call behind the previous one."""
    dev = dv.device()
    key = (float(resolution), int(ndim), dev.index)
    hit = _grid_cache.get(key)
    if hit is None:
        if len(_grid_cache) >= 64:
            _grid_cache.clear()
        width, G = pixel_grid(resolution, ndim)
        hit = _grid_cache[key] = (width, G, dv.to_dev(G))
    return hit


def rebin_device(img, n_out):
    """[..., n*s, n*s] float32 CUDA -> [..., n, n] block sums (K11)."""
    _cabi.require_device()
    shape = img.shape
    nin = shape[-1]
    s = nin // n_out
    if shape[-2] != nin or s * n_out != nin:
        raise ValueError('cannot rebin %s to (%d, %d)' % (tuple(shape), n_out, n_out))
    if img.dim() == 3 and img.dtype == torch.float32 and img.stride(-1) == 1 and not img.is_contiguous():
        # a window of a larger buffer (padded deposit, uncropped convolution): read it in place
        out = torch.empty((shape[0], n_out, n_out), dtype=torch.float32, device=img.device)
        _cabi.check(_cabi.lib.so_img_rebin_ld(_cabi.ptr(img), img.stride(-2), img.stride(-3), _cabi.ptr(out), shape[0], n_out,
                                              s, _cabi.stream()), 'so_img_rebin_ld')
        return out
    x = img.reshape(-1, nin, nin).contiguous()
    out = torch.empty((x.shape[0], n_out, n_out), dtype=torch.float32, device=x.device)
    _cabi.check(_cabi.lib.so_img_rebin(_cabi.ptr(x), _cabi.ptr(out), x.shape[0], n_out, s, _cabi.stream()),
                'so_img_rebin')
    return out.reshape(shape[:-2] + (n_out, n_out))


def rebin(arr, new_shape):
    """reference: images.py:19-22 — block SUM to new_shape."""
    if new_shape[0] != new_shape[1] or arr.shape[0] != arr.shape[1]:
        raise ValueError('rebin: square images only')
    if dv.is_device_tensor(arr):
        return rebin_device(arr.to(torch.float32), new_shape[0])
    return dv.to_host(rebin_device(dv.to_dev(arr, torch.float32), new_shape[0]))


KERNEL_CROP_EPS = 1e-14     # kernel elements below this fraction of the peak are outside the cropped support


def _smooth7(m):
    for f in (2, 3, 5, 7):
        while m % f == 0:
            m //= f
    return m == 1


def _fft_len(n):
    """Transform length >= n with radices 2, 3, 5, 7 only (the ones cuFFT runs at full speed).  Among the
    candidates within 6 % of the smallest one, lengths with a large power of two are preferred: measured on
    B200 for 8 images of 1024^2 plus a 57^2 kernel, rfft2 + product + irfft2 take 0.259 ms at 1080 = 2^3 3^3 5
    and 0.221 ms at 1120 = 2^5 5 7.  The choice is a fixed rule (not a run-time timing), so results stay
    reproducible."""
    n = max(int(n), 1)
    first = n
    while not _smooth7(first):
        first += 1
    best, best_cost = first, None
    for c in range(first, int(first * 1.06) + 1):
        if not _smooth7(c):
            continue
        cost = float(c) * c * (1.0 if c % 32 == 0 else 1.1 if c % 16 == 0 else 1.25)
        if best_cost is None or cost < best_cost:
            best, best_cost = c, cost
    return best


class KernelSpectrum:
    """A convolution kernel prepared for an image size: normalised (float64), cropped to the box
    that holds every element above KERNEL_CROP_EPS of its peak (a PSF sampled on the full
    super-sampled grid, images.py:198-200, is mostly exact or negligible zeros: the dropped mass
    is < 1e-8 of the flux and the transforms shrink from 2n to ~n), zero-padded, transformed
    once.  Re-usable across calls (the PSF image of an observed() set-up only depends on its
    geometry)."""

    def __init__(self, kernel, n_y, n_x):
        dev = dv.device()
        if isinstance(kernel, torch.Tensor):
            k = kernel.to(dev, torch.float64)
        else:
            k = dv.to_dev(np.asarray(kernel, dtype=np.float64))
        k = k / k.sum(dim=(-2, -1), keepdim=True)
        self.per_image = k.dim() == 3
        self.n_kernels = k.shape[0] if self.per_image else 1
        cy, cx = k.shape[-2] // 2, k.shape[-1] // 2            # astropy centres the kernel on element n//2
        a = k.abs().reshape(-1, k.shape[-2], k.shape[-1])
        big = (a > KERNEL_CROP_EPS * a.amax(dim=(-2, -1), keepdim=True)).any(dim=0)
        rows = torch.nonzero(big.any(dim=1)).reshape(-1)
        cols = torch.nonzero(big.any(dim=0)).reshape(-1)
        r0, r1 = min(int(rows[0]), cy), max(int(rows[-1]), cy) + 1   # the box always contains the centre element
        c0, c1 = min(int(cols[0]), cx), max(int(cols[-1]), cx) + 1
        k = k[..., r0:r1, c0:c1].to(torch.float32).contiguous()
        self.n_y, self.n_x = n_y, n_x
        self.py = _fft_len(n_y + k.shape[-2] - 1)
        self.px = _fft_len(n_x + k.shape[-1] - 1)
        self.oy, self.ox = cy - r0, cx - c0
        self.K = torch.fft.rfft2(k, s=(self.py, self.px)).contiguous()   # 2-D rfft2 returns a strided view


def convolve_fft_device(img, kernel, padded=None, crop_copy=True):
    """astropy.convolution.convolve_fft with its defaults (zero fill, kernel
    normalised to unit sum and centred on element n//2, output cropped to the
    image; see oracle/ref_loader.py for the restated identity) as a batched
    cuFFT step (K10).  img: [..., n, n] float32 CUDA; kernel: [nk, nk] (shared by
    all images) or [C, nk, nk] (one per image) as host float64 array or CUDA
    tensor, or a prepared KernelSpectrum.  padded: the zero-padded buffer
    [C, py, px] that already holds img in its top-left corner (no pad copy);
    crop_copy=False returns the result as a window of the full transform
    (no crop copy)."""
    _cabi.require_device()
    n_y, n_x = img.shape[-2], img.shape[-1]
    ks = kernel if isinstance(kernel, KernelSpectrum) else KernelSpectrum(kernel, n_y, n_x)
    if (ks.n_y, ks.n_x) != (n_y, n_x):
        raise ValueError('kernel spectrum prepared for another image size')
    x = img.reshape(-1, n_y, n_x)
    if ks.per_image and ks.n_kernels != x.shape[0]:
        raise ValueError('one kernel per image expected')
    if padded is not None:
        if tuple(padded.shape[-2:]) != (ks.py, ks.px) or not padded.is_contiguous():
            raise ValueError('padded buffer does not match the kernel spectrum')
        A = torch.fft.rfft2(padded.reshape(-1, ks.py, ks.px)).contiguous()
    else:
        A = torch.fft.rfft2(x, s=(ks.py, ks.px)).contiguous()
    _cabi.check(_cabi.lib.so_complex_mul_bcast(_cabi.ptr(torch.view_as_real(A)), _cabi.ptr(torch.view_as_real(ks.K)),
                                               _cabi.ptr(torch.view_as_real(A)), A.shape[-2] * A.shape[-1],
                                               A.shape[0], 1 if ks.per_image else 0, _cabi.stream()),
                'so_complex_mul_bcast')
    full = torch.fft.irfft2(A, s=(ks.py, ks.px))
    out = full[:, ks.oy:ks.oy + n_y, ks.ox:ks.ox + n_x]
    if not crop_copy and img.dim() == 3:
        return out
    return out.contiguous().reshape(img.shape)


def convolve_fft(array, kernel):
    if dv.is_device_tensor(array):
        return convolve_fft_device(array.to(torch.float32), kernel)
    return dv.to_host(convolve_fft_device(dv.to_dev(array, torch.float32), kernel))


def ngp_index_device(Xd, Yd, Gd, ndim, half_width):
    n = Xd.numel()
    ix = torch.empty(n, dtype=torch.int32, device=Xd.device)
    iy = torch.empty(n, dtype=torch.int32, device=Xd.device)
    _cabi.check(_cabi.lib.so_img_ngp_index(_cabi.ptr(Xd), _cabi.ptr(Yd), n, _cabi.ptr(Gd), ndim, float(half_width),
                                           _cabi.ptr(ix), _cabi.ptr(iy), _cabi.stream()), 'so_img_ngp_index')
    return ix, iy


FWHM_FLOOR = 0.01           # images.py:89: FWHM = max(d_k, 0.01)


def knn_device(Xd, Yd, ix, k, half_width, img=None, n_img=1, floor=0.0, part=0, n_parts=1, want_order=False):
    """distance to the k-th neighbour (self = first) among the selected particles (K7); with `img` [n] int32 /
    n_img, among the selected particles of the particle's own image.  floor > 0: particles that already have k
    neighbours within `floor` get an upper bound <= floor instead of the exact distance (enough for
    max(dk, floor), and their tree walk is skipped).  part / n_parts: answer only that share of the selected
    particles in Morton order (then the result is (dk, pids of the share))."""
    n = Xd.numel()
    dk = (torch.zeros if n_parts > 1 else torch.empty)(n, dtype=torch.float64, device=Xd.device)
    order = torch.empty(n, dtype=torch.int32, device=Xd.device) if want_order else None
    rng = torch.zeros(3, dtype=torch.int64, device=Xd.device) if want_order else None
    wsb = _cabi.lib.so_knn_batch_workspace_bytes(n, int(n_img), int(k))
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_knn_kth_distance_ex(_cabi.ptr(Xd), _cabi.ptr(Yd), _cabi.ptr(ix), _cabi.ptr(img), n,
                                                 int(n_img), int(k), float(half_width), int(part), int(n_parts),
                                                 float(floor), _cabi.ptr(dk), _cabi.ptr(order), _cabi.ptr(rng),
                                                 _cabi.ptr(ws), wsb, _cabi.stream()), 'so_knn_kth_distance_ex')
    if want_order:
        b, e, _ = (int(v) for v in rng.cpu())
        return dk, order[b:e].to(torch.int64)
    return dk


def knn_device_part(Xd, Yd, ix, k, half_width, part, n_parts, floor=0.0):
    """k-th neighbour distances for share `part` of `n_parts` of the selected particles in Morton
    order (searching all positions).  Returns (dk [n] valid for the share only, pids [m] int64:
    the particles of the share)."""
    return knn_device(Xd, Yd, ix, k, half_width, floor=floor, part=part, n_parts=n_parts, want_order=True)


def knn_device_masked(Xd, Yd, ix, k, half_width, query_mask, floor=0.0):
    """k-th neighbour distance (self = first) among ALL selected particles, answered only for the particles whose
    query_mask byte has bit 1 set (uint8 CUDA tensor; the others keep 0): so_knn_kth_distance_q."""
    n = Xd.numel()
    dk = torch.zeros(n, dtype=torch.float64, device=Xd.device)
    wsb = _cabi.lib.so_knn_batch_workspace_bytes(n, 1, int(k))
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_knn_kth_distance_q(_cabi.ptr(Xd), _cabi.ptr(Yd), _cabi.ptr(ix), n, int(k), float(half_width),
                                                _cabi.ptr(query_mask), float(floor), _cabi.ptr(dk), _cabi.ptr(ws), wsb,
                                                _cabi.stream()), 'so_knn_kth_distance_q')
    return dk


def share_grid_bits(n, n_parts=1):
    """coarse grid (2^cb cells per axis) of the particle shares: ~1500 particles per cell (256 x 256 cells for 1e8
    particles: a ring of two cells around a share of 1/8 adds ~10 % halo) but at least ~500 cells per share (the
    shares are balanced to one cell), 8 <= cells per axis <= 512"""
    cb = int(np.floor(np.log2(max(np.sqrt(max(n, 1) / 1500.0), 1.0)) + 0.5))
    cb = max(cb, int(np.ceil(0.5 * np.log2(500.0 * max(int(n_parts), 1)))))
    return int(min(max(cb, 3), 9))


def share_hist_device(Xd, Yd, half_width, cb, begin=0, end=None, out=None):
    """counts of the selected particles [begin, end) per coarse cell (Morton order), int32 CUDA tensor [4^cb]"""
    n = Xd.numel()
    end = n if end is None else int(end)
    hist = torch.empty(1 << (2 * cb), dtype=torch.int32, device=Xd.device) if out is None else out
    _cabi.check(_cabi.lib.so_share_hist(_cabi.ptr(Xd), _cabi.ptr(Yd), int(begin), end, float(half_width), int(cb), _cabi.ptr(hist),
                                        _cabi.stream()), 'so_share_hist')
    return hist


def share_plan_device(Xd, Yd, half_width, k, part, n_parts, cb=None, hist=None, bounds=None):
    """Flags of share `part` of `n_parts` of one huge image (so_share_plan): uint8 CUDA tensor, bit 0 = local (member
    of the share's search tree: own + halo), bit 1 = own; plus the device counts [local, own].  hist = the coarse
    histogram of ALL particles (share_hist_device; summed over the ranks when each counted a part).  bounds = the
    n_parts + 1 cumulative work fractions of the shares (default: equal)."""
    n = Xd.numel()
    cb = share_grid_bits(n, n_parts) if cb is None else int(cb)
    if hist is None:
        hist = share_hist_device(Xd, Yd, half_width, cb)
    lo, hi = (part / n_parts, (part + 1) / n_parts) if bounds is None else (float(bounds[part]), float(bounds[part + 1]))
    if part == n_parts - 1:
        hi = 2.0            # the last share takes everything up to the end
    flags = torch.empty(n, dtype=torch.uint8, device=Xd.device)
    counts = torch.zeros(2, dtype=torch.int64, device=Xd.device)
    wsb = _cabi.lib.so_share_plan_workspace_bytes(cb)
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_share_plan(_cabi.ptr(Xd), _cabi.ptr(Yd), n, float(half_width), cb, _cabi.ptr(hist), int(k),
                                        float(FWHM_FLOOR), lo, hi, _cabi.ptr(flags), _cabi.ptr(counts), _cabi.ptr(ws), wsb,
                                        _cabi.stream()), 'so_share_plan')
    return flags, counts


def share_compact_device(Xd, Yd, flags, bit_mask, m):
    """The m particles whose flag byte has a bit of bit_mask set, in their original order (so_share_compact):
    (index int32 [m], X [m], Y [m], flags [m]) as CUDA tensors."""
    n = Xd.numel()
    dev = Xd.device
    idx = torch.empty(m, dtype=torch.int32, device=dev)
    oX = torch.empty(m, dtype=torch.float64, device=dev)
    oY = torch.empty(m, dtype=torch.float64, device=dev)
    oF = torch.empty(m, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(1, dtype=torch.int64, device=dev)
    wsb = _cabi.lib.so_share_compact_workspace_bytes(n)
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_share_compact(_cabi.ptr(Xd), _cabi.ptr(Yd), _cabi.ptr(flags), n, int(bit_mask), int(m), _cabi.ptr(idx),
                                           _cabi.ptr(oX), _cabi.ptr(oY), _cabi.ptr(oF), _cabi.ptr(cnt), _cabi.ptr(ws), wsb,
                                           _cabi.stream()), 'so_share_compact')
    return idx, oX, oY, oF


def deposit_device(Xd, Yd, Ld, resolution, ndim, adaptive_k=None, want_data=True, want_ngp=True,
                   support_sigmas=None, dk=None, return_pairs=False, L_event=None, img=None, n_img=1, exact_dk=False,
                   pad_to=None, out=None, push=None):
    """images.core on device tensors.  Xd, Yd: [n] float64 CUDA; Ld: [C, n] float64
    CUDA (C images sharing positions).  Returns dict of float32 CUDA tensors:
    data [C, ndim, ndim] (adaptive only), simple [C, ndim, ndim], hist [ndim, ndim],
    plus ix, iy, dk and the stats (pairs, deposits, selected, dropped).  pad_to=(py, px): `data` is the
    top-left window of a zero-padded buffer [C, py, px] (returned as 'data_padded') ready for the PSF transform.
    out: a float32 CUDA tensor [C, ndim, ndim] to deposit into (every pixel is written) instead of a new one.
    push: a _cabi.PushDesc -- one image shared by several GPUs, the exchange fused into the deposit (dist.PeerImage;
    `out` is then the peer-mapped partial image and the result is read from the descriptor's result buffer)."""
    _cabi.require_device()
    if Ld.dim() == 1:
        Ld = Ld.reshape(1, -1)
    Ld = Ld.contiguous()
    C, n = Ld.shape
    width, G, Gd = pixel_grid_device(resolution, ndim)
    ix, iy = ngp_index_device(Xd, Yd, Gd, ndim, width / 2.)
    adaptive = adaptive_k is not None
    if adaptive and dk is None:
        # only max(dk, FWHM_FLOOR) enters the image: particles settled below the floor skip the exact search
        dk = knn_device(Xd, Yd, ix, adaptive_k, width / 2., img=img, n_img=n_img, floor=0.0 if exact_dk else FWHM_FLOOR)
    R = SUPPORT_SIGMAS if support_sigmas is None else support_sigmas
    pws_b = _cabi.lib.so_img_plan_workspace_bytes(n)
    pws = dv.workspace(pws_b)
    n_pairs = c_int64(0)
    stats = (c_double * 4)()
    _cabi.check(_cabi.lib.so_img_plan(_cabi.ptr(ix), _cabi.ptr(iy), _cabi.ptr(Xd), _cabi.ptr(Yd),
                                      _cabi.ptr(dk) if adaptive else None, n, _cabi.ptr(Gd), ndim, float(R),
                                      _cabi.ptr(pws), pws_b, n_pairs, stats, _cabi.stream()), 'so_img_plan')
    npairs = n_pairs.value
    if push is not None and npairs > MAX_PAIRS:
        raise ValueError('a shared (push) deposit takes at most %d (particle, sub-tile) pairs per rank' % MAX_PAIRS)
    if npairs > MAX_PAIRS and n > 1:
        # more (particle, sub-tile) pairs than one binning pass indexes (int32) or should hold in memory:
        # deposit the two halves of the particle list separately and add the images (the sum is associative)
        h = n // 2
        sub = lambda t, a, b: None if t is None else t[..., a:b]   # noqa: E731
        r1 = deposit_device(Xd[:h], Yd[:h], Ld[:, :h], resolution, ndim, adaptive_k, want_data, want_ngp,
                            support_sigmas, sub(dk, 0, h), False, img=sub(img, 0, h), n_img=n_img, pad_to=pad_to, out=out)
        r2 = deposit_device(Xd[h:], Yd[h:], Ld[:, h:], resolution, ndim, adaptive_k, want_data, want_ngp,
                            support_sigmas, sub(dk, h, n), False, img=sub(img, h, n), n_img=n_img, pad_to=pad_to)
        for key in ('data', 'simple', 'hist'):
            if r1[key] is not None:
                r1[key] += r2[key]
        r1['stats'] = {q: r1['stats'][q] + r2['stats'][q] for q in r1['stats']}
        r1['ix'], r1['iy'], r1['dk'], r1['tiles'], r1['pids'] = ix, iy, dk, None, None
        return r1
    dev = Xd.device
    lead = () if img is None else (int(n_img),)     # a batch of images adds a leading axis
    padded = None
    if adaptive and want_data and pad_to is not None and img is None:
        padded = torch.zeros((C, int(pad_to[0]), int(pad_to[1])), dtype=torch.float32, device=dev)
        data = padded[:, :ndim, :ndim]
    elif out is not None and adaptive and want_data and img is None:
        if tuple(out.shape) != (C, ndim, ndim) or out.dtype != torch.float32 or not out.is_contiguous():
            raise ValueError('out must be a contiguous float32 tensor [C, ndim, ndim]')
        data = out
    else:
        data = torch.empty(lead + (C, ndim, ndim), dtype=torch.float32, device=dev) if (adaptive and want_data) else None
    simple = torch.empty(lead + (C, ndim, ndim), dtype=torch.float32, device=dev) if want_ngp else None
    hist = torch.empty(lead + (ndim, ndim), dtype=torch.float32, device=dev) if want_ngp else None
    tiles = torch.empty(npairs, dtype=torch.int32, device=dev) if return_pairs else None
    pids = torch.empty(npairs, dtype=torch.int32, device=dev) if return_pairs else None
    wsb = _cabi.lib.so_img_deposit_batch_workspace_bytes(n, npairs, ndim, C, int(n_img))
    ws = dv.workspace(wsb)
    if L_event is not None:        # the weights may still be arriving on a copy stream (k-NN and binning need positions only)
        torch.cuda.current_stream().wait_event(L_event)
    ld, plane = (padded.shape[-1], padded.shape[-1] * padded.shape[-2]) if padded is not None else (ndim, ndim * ndim)
    from ctypes import byref
    _cabi.check(_cabi.lib.so_img_deposit_push(_cabi.ptr(pws), _cabi.ptr(Ld), _cabi.ptr(img), n, C, int(n_img), ndim,
                                              npairs, _cabi.ptr(padded if padded is not None else data), ld, plane,
                                              _cabi.ptr(simple), _cabi.ptr(hist), _cabi.ptr(tiles), _cabi.ptr(pids),
                                              _cabi.ptr(ws), wsb, None if push is None else byref(push), _cabi.stream()),
                'so_img_deposit_push')
    return {'data': data, 'data_padded': padded, 'simple': simple, 'hist': hist, 'ix': ix, 'iy': iy, 'dk': dk, 'G': G,
            'stats': {'pairs': stats[0], 'deposits': stats[1], 'selected': stats[2], 'dropped': stats[3]},
            'tiles': tiles, 'pids': pids}


def _parse_smoothing(smoothing):
    if smoothing:
        return smoothing[0], smoothing[1]
    return False, None


def core_device(Xd, Yd, Ld, resolution=0.1, ndim=100, smoothing=False, want_ngp=True, img=None, n_img=1):
    """images.core for [C, n] weights on the device; returns (data, simple, hist, info).  With `img`
    [n] int32 / n_img the particles belong to n_img separate images (leading axis on every product)."""
    method, scale = _parse_smoothing(smoothing)
    if method == 'adaptive':
        r = deposit_device(Xd, Yd, Ld, resolution, ndim, adaptive_k=int(scale), want_ngp=want_ngp, img=img, n_img=n_img)
        return r['data'], r['simple'], r['hist'], r
    r = deposit_device(Xd, Yd, Ld, resolution, ndim, adaptive_k=None, want_ngp=True, img=img, n_img=n_img)
    if method == 'convolved_gaussian':
        width, G = pixel_grid(resolution, ndim)
        Gx, Gy = np.meshgrid(G, G)
        sigma = scale / 2.355                                              # images.py:74
        gauss = np.exp(-((Gx ** 2 + Gy ** 2) / (2.0 * sigma ** 2)))       # images.py:75
        gauss /= np.sum(gauss)
        data = convolve_fft_device(r['simple'], gauss)                    # images.py:78
    else:
        data = r['simple']                                                 # images.py:103
    return data, r['simple'], r['hist'], r


# --------------------------------------------------------------------------
# reference API
# --------------------------------------------------------------------------

def core(X, Y, L, resolution=0.1, ndim=100, smoothing=False, verbose=False):
    """reference: images.py:28-105.  X, Y, resolution and the smoothing scale in
    the same units (default physical kpc).  Returns an object with .hist,
    .simple, .data ([ndim, ndim], row = y, column = x)."""
    host = not any(dv.is_device_tensor(a) for a in (X, Y, L))
    Xd, Yd, Ld = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L)
    data, simple, hist, info = core_device(Xd, Yd, Ld, resolution, ndim, smoothing)
    if verbose:
        method, _ = _parse_smoothing(smoothing)
        print('*' * 5, 'CORE')
        print('width: {0:.2f} kpc'.format(ndim * resolution))
        print('ndim: {0:.2f} '.format(ndim))
        print('resolution: {0:.2f} '.format(resolution))
        print('N_particles: {0:.2f} '.format(info['stats']['selected']))
        print('smoothing method: {0} '.format(method))
    img = empty()
    same = data is simple
    if host:
        img.hist = dv.to_host(hist)
        img.simple = dv.to_host(simple[0])
        img.data = img.simple if same else dv.to_host(data[0])
    else:
        img.hist, img.simple = hist, simple[0]
        img.data = img.simple if same else data[0]
    img.info = info['stats']
    return img


def _image_ids(offsets, dev):
    off = dv.to_dev(offsets, torch.int64)
    G = off.numel() - 1
    ids = torch.repeat_interleave(torch.arange(G, device=dev, dtype=torch.int32), (off[1:] - off[:-1]))
    return off, ids, G


def core_batch(X, Y, L, offsets, resolution=0.1, ndim=100, smoothing=False):
    """images.core for a batch of objects in ONE pass (BASELINE config 4; no upstream equivalent, the
    reference loops over objects, examples/core_manytestobjects.py:17-22).  X, Y, L: concatenated
    particles, `offsets` = CSR [G+1]; L may be [C, n] for C images per object sharing positions.
    Returns an object with .hist [G, ndim, ndim], .simple and .data [G, (C,) ndim, ndim], equal to a loop
    of core() over the objects (the k-NN of an object only sees its own particles)."""
    host = not any(dv.is_device_tensor(a) for a in (X, Y, L))
    Xd, Yd, Ld = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L)
    one = Ld.dim() == 1
    off, ids, G = _image_ids(offsets, Xd.device)
    data, simple, hist, info = core_device(Xd, Yd, Ld, resolution, ndim, smoothing, img=ids, n_img=G)
    pick = (lambda t: t[:, 0]) if one else (lambda t: t)
    fin = (lambda t: dv.to_host(pick(t))) if host else pick
    img = empty()
    img.hist = dv.to_host(hist) if host else hist
    img.simple = fin(simple)
    img.data = img.simple if data is simple else fin(data)
    img.info = info['stats']
    return img


def median_segmented_device(x, off):
    """np.median of every CSR segment of a CUDA float64 tensor -> [G] CUDA tensor"""
    G = off.numel() - 1
    out = torch.empty(G, dtype=torch.float64, device=x.device)
    _cabi.check(_cabi.lib.so_median_segmented(_cabi.ptr(x), _cabi.ptr(off), G, _cabi.ptr(out), _cabi.stream()),
                'so_median_segmented')
    return out


def _recentre_segmented(x, off, shift):
    med = median_segmented_device(x, off)
    _cabi.check(_cabi.lib.so_recentre_segmented(_cabi.ptr(x), _cabi.ptr(off), off.numel() - 1, _cabi.ptr(med),
                                                float(shift), _cabi.stream()), 'so_recentre_segmented')
    return med


def median_device(x):
    """np.median of a CUDA float64 tensor (mean of the two middle values for even n) as a 0-d
    CUDA tensor; [m, n] input -> [m] medians of the rows.  Radix selection (so_median_f64), no sort."""
    _cabi.require_device()
    x2 = x.reshape(1, -1) if x.dim() == 1 else x
    x2 = x2.to(torch.float64).contiguous()
    m, n = x2.shape
    out = torch.empty(m, dtype=torch.float64, device=x2.device)
    wsb = _cabi.lib.so_median_workspace_bytes(n, m)
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_median_f64(_cabi.ptr(x2), n, m, x2.stride(0), _cabi.ptr(out), _cabi.ptr(ws), wsb,
                                        _cabi.stream()), 'so_median_f64')
    return out[0] if x.dim() == 1 else out


def median_device_pair(a, b):
    """np.median of two CUDA float64 vectors of equal length in ONE selection (the two arrays are
    addressed as rows of a strided 2-row matrix: one set of launches for both)."""
    if (a.dim() != 1 or b.dim() != 1 or a.numel() != b.numel() or a.dtype != torch.float64 or b.dtype != torch.float64
            or not a.is_contiguous() or not b.is_contiguous() or a.data_ptr() == b.data_ptr()):
        return median_device(a), median_device(b)
    lo, hi, swap = (a, b, False) if a.data_ptr() < b.data_ptr() else (b, a, True)
    diff = hi.data_ptr() - lo.data_ptr()
    n = a.numel()
    if diff % 8 or diff // 8 < n:
        return median_device(a), median_device(b)
    out = torch.empty(2, dtype=torch.float64, device=a.device)
    wsb = _cabi.lib.so_median_workspace_bytes(n, 2)
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_median_f64(_cabi.ptr(lo), n, 2, diff // 8, _cabi.ptr(out), _cabi.ptr(ws), wsb,
                                        _cabi.stream()), 'so_median_f64')
    return (out[1], out[0]) if swap else (out[0], out[1])


_psf_spec_cache = {}


def clear_caches():
    """Drop the cached PSF spectra and pixel grids (PSF providers are treated as immutable objects: a provider
    edited in place after its first use needs this)."""
    _psf_spec_cache.clear()
    _grid_cache.clear()


def _group_spectrum(obs_list):
    """KernelSpectrum of the PSFs of several observed() set-ups sharing one geometry.  The PSF image only depends on
    the provider object and the geometry (images.py:198-200), so it is evaluated and transformed once per
    (provider, geometry) and kept: images.particle() builds fresh observed() objects on every call."""
    o0 = obs_list[0]
    key = (dv.device().index,) + tuple((id(o.PSF), o.ndim_super, o.width_arcsec, o.native_pixel_scale) for o in obs_list)
    hit = _psf_spec_cache.get(key)
    if hit is None:
        if len(_psf_spec_cache) >= 16:
            _psf_spec_cache.clear()
        spec = KernelSpectrum(np.stack([o.psf_kernel() for o in obs_list]), o0.ndim_super, o0.ndim_super)
        hit = _psf_spec_cache[key] = (spec, [o.PSF for o in obs_list])      # the providers stay alive with their ids
    return hit[0]


def median_pair_mid(a, b):
    """np.median of two CUDA float64 vectors of equal length plus their two middle elements: (med [2], mid [4] =
    a_lo, a_hi, b_lo, b_hi) as CUDA tensors, one selection for both arrays when they sit in one allocation order."""
    n = a.numel()
    out = torch.empty(2, dtype=torch.float64, device=a.device)
    mid = torch.empty(4, dtype=torch.float64, device=a.device)
    diff = b.data_ptr() - a.data_ptr()
    if a.is_contiguous() and b.is_contiguous() and diff > 0 and diff % 8 == 0 and diff // 8 >= n:
        wsb = _cabi.lib.so_median_workspace_bytes(n, 2)
        ws = dv.workspace(wsb)
        _cabi.check(_cabi.lib.so_median_f64_mid(_cabi.ptr(a), n, 2, diff // 8, _cabi.ptr(out), _cabi.ptr(mid), _cabi.ptr(ws),
                                                wsb, _cabi.stream()), 'so_median_f64_mid')
        return out, mid
    wsb = _cabi.lib.so_median_workspace_bytes(n, 1)
    ws = dv.workspace(wsb)
    for i, t in enumerate((a.contiguous(), b.contiguous())):
        _cabi.check(_cabi.lib.so_median_f64_mid(_cabi.ptr(t), n, 1, n, _cabi.ptr(out[i:]), _cabi.ptr(mid[2 * i:]),
                                                _cabi.ptr(ws), wsb, _cabi.stream()), 'so_median_f64_mid')
    return out, mid


def recentre_chain_device(Xd, Yd, off_x, off_y, out_steps):
    """The positions after steps `out_steps` (1-based, ascending) of the reference's per-filter recentring
    X_f = (X_{f-1} - np.median(X_{f-1})) + off_x[f] (images.py:183-187 inside the loop of :277-284), bit-identical
    to the repeated in-place updates; Xd, Yd are read once (so_recentre_chain).  Returns [(X_t, Y_t), ...]."""
    from ctypes import c_int32, c_void_p
    n = Xd.numel()
    _, mid = median_pair_mid(Xd, Yd)
    outs = [(torch.empty_like(Xd), torch.empty_like(Yd)) for _ in out_steps]
    nst = len(off_x)
    _cabi.check(_cabi.lib.so_recentre_chain(
        _cabi.ptr(Xd), _cabi.ptr(Yd), n, _cabi.ptr(mid), nst, (c_double * nst)(*[float(v) for v in off_x]),
        (c_double * nst)(*[float(v) for v in off_y]), len(out_steps), (c_int32 * len(out_steps))(*[int(v) for v in out_steps]),
        (c_void_p * len(outs))(*[o[0].data_ptr() for o in outs]), (c_void_p * len(outs))(*[o[1].data_ptr() for o in outs]),
        _cabi.stream()), 'so_recentre_chain')
    return outs


def particle_device_multi(obs_list, Xd, Yd, Ld, L_event=None, offsets_applied=False):
    """Several observed() set-ups that share one geometry (same pixel scale, width,
    super-sampling, offsets, smoothing) imaged in one pass: one k-NN, one tile
    binning, a C-channel deposit, one batched FFT with a PSF per channel, one
    rebin.  Xd, Yd already recentred; Ld: [C, n].  The reference repeats the k-NN
    and the deposit per filter on identical positions (images.py:277-284)."""
    o0 = obs_list[0]
    for o in obs_list[1:]:
        same = (o.ndim_super == o0.ndim_super and o.resolution == o0.resolution and o.xoffset == o0.xoffset
                and o.yoffset == o0.yoffset and o.smoothing == o0.smoothing and o.ndim == o0.ndim)
        if not same:
            raise ValueError('particle_device_multi needs identical geometry; group the filters by pixel scale')
    method, scale = _parse_smoothing(o0.smoothing)
    X = Xd + o0.xoffset if (o0.xoffset and not offsets_applied) else Xd
    Y = Yd + o0.yoffset if (o0.yoffset and not offsets_applied) else Yd
    spec = _group_spectrum(obs_list)
    if method == 'adaptive':
        # deposit straight into the zero-padded buffer of the PSF transform; the convolved images stay windows of
        # the full inverse transform (rebin reads them in place): no pad copy, no crop copy
        r = deposit_device(X, Y, Ld, o0.resolution, o0.ndim_super, adaptive_k=int(scale), want_ngp=False,
                           L_event=L_event, pad_to=(spec.py, spec.px))
        no_psf = r['data']
        conv = convolve_fft_device(no_psf, spec, padded=r['data_padded'], crop_copy=False)
    else:
        if L_event is not None:
            torch.cuda.current_stream().wait_event(L_event)
        no_psf, _, _, r = core_device(X, Y, Ld, o0.resolution, o0.ndim_super, o0.smoothing)
        conv = convolve_fft_device(no_psf, spec)
    return {'super_no_PSF': no_psf, 'super_data': conv, 'no_PSF': rebin_device(no_psf, o0.ndim),
            'data': rebin_device(conv, o0.ndim), 'stats': r['stats']}


_copy_stream = {}


def _side_stream(key, dev):
    st = _copy_stream.get(key)
    if st is None:
        st = _copy_stream[key] = torch.cuda.Stream(device=dev)
    return st


def particle_host_multi(obs_list, Xh, Yh, Lh):
    """particle_device_multi for HOST inputs (NumPy arrays or CPU tensors, ideally pinned): the
    positions go up first, the [C, n] weights follow on a copy stream while the k-NN and the tile
    binning run, the deposit waits for them; recentring by the medians as observed.particle does
    (images.py:183-184) on the device copies (the caller's arrays are left alone).  Returns the
    final images [C, ndim, ndim] as a float32 CPU tensor plus the run statistics."""
    dev = dv.device()
    main = torch.cuda.current_stream()
    side = _side_stream(dev.index, dev)
    as_t = lambda a: a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a))   # noqa: E731
    Xd = as_t(Xh).to(dev, non_blocking=True)
    Yd = as_t(Yh).to(dev, non_blocking=True)
    xy_up = torch.cuda.Event()
    xy_up.record(main)
    with torch.cuda.stream(side):
        side.wait_event(xy_up)      # the positions get the whole link first: the k-NN can start while the weights follow
        Ld = as_t(Lh).to(dev, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(side)
    Ld.record_stream(main)
    mx, my = median_device_pair(Xd, Yd)
    Xd = Xd - mx
    Yd = Yd - my
    r = particle_device_multi(obs_list, Xd, Yd, Ld if Ld.dim() == 2 else Ld.reshape(1, -1), L_event=ev)
    # page-locked result buffer (torch caches the host block): a pageable .cpu() copy runs at a fraction of the link rate
    out = torch.empty(r['data'].shape, dtype=r['data'].dtype, pin_memory=True)
    out.copy_(r['data'], non_blocking=True)
    main.synchronize()
    return out, r['stats']


def _median_inplace_shift(A, offset):
    """A -= median(A); A += offset, in place on the caller's array (images.py:183-187)."""
    if isinstance(A, torch.Tensor):
        A -= median_device(A)
        A += offset
    else:
        A -= np.median(A)
        A += offset


class observed():
    """reference: images.py:112-213 (geometry + .particle)."""

    def __init__(self, filter, cosmo, z, target_width_arcsec, resampling_factor=False, pixel_scale=False,
                 smoothing=False, PSF=False, super_sampling=4, verbose=False, xoffset_pix=0.0, yoffset_pix=0.0):

        self.filter = filter
        self.target_width_arcsec = target_width_arcsec
        self.resampling_factor = resampling_factor
        self.pixel_scale = pixel_scale
        self.smoothing = smoothing
        self.PSF = PSF
        self.verbose = verbose
        self.super_sampling = super_sampling

        self.native_pixel_scale = filter_info[self.filter]['pixel_scale']

        if self.resampling_factor:
            self.pixel_scale = self.native_pixel_scale / self.resampling_factor
        elif self.pixel_scale:
            self.resampling_factor = self.native_pixel_scale / self.pixel_scale
        else:
            self.pixel_scale = self.native_pixel_scale
            self.resampling_factor = 1.0

        self.arcsec_per_proper_kpc = cosmo.arcsec_per_kpc_proper(z).value
        self.pixel_scale_kpc = self.pixel_scale / self.arcsec_per_proper_kpc
        self.ndim = int(self.target_width_arcsec / self.pixel_scale)
        self.width_arcsec = self.ndim * self.pixel_scale
        self.width = self.width_arcsec / self.arcsec_per_proper_kpc
        self.resolution = self.pixel_scale_kpc / self.super_sampling
        self.ndim_super = self.ndim * self.super_sampling
        self.xoffset = xoffset_pix * self.pixel_scale_kpc
        self.yoffset = yoffset_pix * self.pixel_scale_kpc

        if verbose:
            _report('OBSERVED', [('z', z, '{0}'), ('"/kpc', self.arcsec_per_proper_kpc, None), None,
                                 ('target width/"', self.target_width_arcsec, None), ('width/"', self.width_arcsec, None),
                                 ('width/kpc', self.width, None), None,
                                 ('native pixel scale/"', self.native_pixel_scale, None), ('pixel scale/"', self.pixel_scale, None),
                                 ('pixel scale/kpc', self.pixel_scale_kpc, None), ('super sampling', self.super_sampling, None),
                                 ('resolution/kpc', self.resolution, None), ('ndim', self.ndim, None), None])

    def psf_kernel(self):
        """PSF image on the super-sampled grid; depends only on the geometry fixed at
        construction, so it is evaluated once per observed() object."""
        if getattr(self, '_psf_image', None) is None:
            xx = yy = np.linspace(-(self.width_arcsec / self.native_pixel_scale / 2.),
                                  (self.width_arcsec / self.native_pixel_scale / 2.), self.ndim_super)   # images.py:198
            self._psf_image = self.PSF.f(xx, yy)                                                        # images.py:200
        return self._psf_image

    def psf_spectrum(self):
        if getattr(self, '_psf_spec', None) is None:
            self._psf_spec = KernelSpectrum(self.psf_kernel(), self.ndim_super, self.ndim_super)
        return self._psf_spec

    def particle_device(self, Xd, Yd, Ld, dk=None):
        """Device pipeline after recentring: deposit -> PSF FFT convolution -> rebin.
        Ld: [C, n].  Returns dict of float32 CUDA tensors."""
        method, scale = _parse_smoothing(self.smoothing)
        spec = self.psf_spectrum()
        if method == 'adaptive':
            r = deposit_device(Xd, Yd, Ld, self.resolution, self.ndim_super, adaptive_k=int(scale), want_ngp=False,
                               dk=dk, pad_to=(spec.py, spec.px))
            no_psf = r['data']
            conv = convolve_fft_device(no_psf, spec, padded=r['data_padded'], crop_copy=False)      # images.py:202
        else:
            no_psf, _, _, r = core_device(Xd, Yd, Ld, self.resolution, self.ndim_super, self.smoothing)
            conv = convolve_fft_device(no_psf, spec)                                                # images.py:202
        return {'super_no_PSF': no_psf, 'super_data': conv,
                'no_PSF': rebin_device(no_psf, self.ndim), 'data': rebin_device(conv, self.ndim),  # images.py:210-211
                'info': r}

    def particle(self, X, Y, L):
        host = not any(dv.is_device_tensor(a) for a in (X, Y, L))
        _median_inplace_shift(X, self.xoffset)     # mutates the caller's arrays, as the reference does
        _median_inplace_shift(Y, self.yoffset)
        Xd, Yd, Ld = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L)
        r = self.particle_device(Xd, Yd, Ld)
        conv = (lambda t: dv.to_host(t[0])) if host else (lambda t: t[0])
        imgs = empty()
        imgs.super = empty()
        imgs.super.no_PSF = conv(r['super_no_PSF'])
        imgs.super.data = conv(r['super_data'])
        imgs.img = empty()
        imgs.img.pixel_scale = self.pixel_scale
        imgs.img.pixel_scale_kpc = self.pixel_scale_kpc
        imgs.img.no_PSF = conv(r['no_PSF'])
        imgs.img.data = conv(r['data'])
        return imgs


def _sersic_device(G, p, x0, y0, Ld):
    """L[c] * Sersic2D / sum(Sersic2D) on the grid G x G -> [C, n, n] float32 CUDA (so_img_sersic)."""
    from ..providers import Sersic2D
    _cabi.require_device()
    ndim = int(G.size)
    Gd = dv.to_dev(G)
    C = Ld.numel()
    out = torch.empty((C, ndim, ndim), dtype=torch.float32, device=Gd.device)
    wsb = _cabi.lib.so_img_sersic_workspace_bytes(ndim)
    ws = dv.workspace(wsb)
    _cabi.check(_cabi.lib.so_img_sersic(_cabi.ptr(Gd), ndim, float(p['r_eff']), float(p['n']), float(x0), float(y0),
                                        float(p['ellip']), float(p['theta']), Sersic2D.bn(float(p['n'])),
                                        _cabi.ptr(Ld), C, _cabi.ptr(out), _cabi.ptr(ws), wsb, _cabi.stream()),
                'so_img_sersic')
    return out


def _observed_Sersic(self, L, p):
    """reference: images.py:219-257 -- analytic Sersic source through the same back half as
    .particle (PSF convolution on the super-sampled grid, rebin).  L: scalar (or [C] array /
    CUDA tensor for C channels sharing the profile: then every product has a leading C axis)."""
    scalar = np.ndim(L) == 0 and not isinstance(L, torch.Tensor)
    host = not dv.is_device_tensor(L)
    Ld = dv.to_dev(np.atleast_1d(np.asarray(L, dtype=np.float64))) if host else L.reshape(-1).to(torch.float64)
    g = np.linspace(-self.width / 2., self.width / 2., self.ndim_super)                         # images.py:224
    no_psf = _sersic_device(g, p, self.xoffset, self.yoffset, Ld)                               # images.py:228-238
    # NB the reference builds this PSF grid in units of pixel_scale, not native_pixel_scale (images.py:242)
    if getattr(self, '_psf_spec_sersic', None) is None:
        half = self.width_arcsec / self.pixel_scale / 2.
        xx = yy = np.linspace(-half, half, self.ndim_super)
        self._psf_spec_sersic = KernelSpectrum(self.PSF.f(xx, yy), self.ndim_super, self.ndim_super)
    conv = convolve_fft_device(no_psf, self._psf_spec_sersic)                                   # images.py:246
    pick = (lambda t: t[0]) if scalar else (lambda t: t)
    fin = (lambda t: dv.to_host(pick(t))) if host else pick
    imgs = empty()
    imgs.super = empty()
    imgs.super.pixel_scale = self.pixel_scale / self.super_sampling
    imgs.super.pixel_scale_kpc = self.pixel_scale_kpc / self.super_sampling
    imgs.super.no_PSF = fin(no_psf)
    imgs.super.data = fin(conv)
    imgs.img = empty()
    imgs.img.pixel_scale = self.pixel_scale
    imgs.img.pixel_scale_kpc = self.pixel_scale_kpc
    imgs.img.no_PSF = fin(rebin_device(no_psf, self.ndim))                                      # images.py:255-256
    imgs.img.data = fin(rebin_device(conv, self.ndim))
    return imgs


observed.Sersic = _observed_Sersic


def _observed_particle_batch(self, X, Y, L, offsets):
    """observed.particle for a batch of objects in one pass (CSR `offsets` [G+1]): every object is
    recentred on its own medians (images.py:183-187; on device copies, the caller's arrays are left
    alone), imaged on its own super-sampled grid, convolved with the PSF and rebinned.  L: [n] or
    [C, n].  Returns dict of arrays [G, (C,) ndim, ndim]: 'no_PSF', 'data' (+ the super-sampled pair),
    NumPy float64 for host inputs, CUDA float32 otherwise."""
    host = not any(dv.is_device_tensor(a) for a in (X, Y, L))
    Xd, Yd, Ld = dv.to_dev(X).clone(), dv.to_dev(Y).clone(), dv.to_dev(L)
    one = Ld.dim() == 1
    off, ids, G = _image_ids(offsets, Xd.device)
    _recentre_segmented(Xd, off, self.xoffset)
    _recentre_segmented(Yd, off, self.yoffset)
    no_psf, _, _, r = core_device(Xd, Yd, Ld, self.resolution, self.ndim_super, self.smoothing, want_ngp=False,
                                  img=ids, n_img=G)
    conv = convolve_fft_device(no_psf, self.psf_spectrum())
    pick = (lambda t: t[:, 0]) if one else (lambda t: t)
    fin = (lambda t: dv.to_host(pick(t))) if host else pick
    return {'super_no_PSF': fin(no_psf), 'super_data': fin(conv), 'no_PSF': fin(rebin_device(no_psf, self.ndim)),
            'data': fin(rebin_device(conv, self.ndim)), 'stats': r['stats']}


observed.particle_batch = _observed_particle_batch


def _as_cpu_tensor(a):
    return a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64))


def particle(X, Y, L, filters, cosmo, z, target_width_arcsec, resampling_factor=False, pixel_scale=False,
             smoothing=False, PSFs=False, super_sampling=10, verbose=False, offsets=False):
    """reference: images.py:263-286 -- one observed image per filter; returns {filter: img} with
    img.pixel_scale, img.pixel_scale_kpc, img.no_PSF, img.data.

    The reference runs observed(...).particle(X, Y, L[filter]) filter by filter, i.e. it repeats the median
    recentring, the k-NN and the deposit on identical positions (images.py:277-284).  Here filters whose
    geometry is identical (same pixel scale, hence same grid, offsets and smoothing) are imaged TOGETHER: one
    recentring, one k-NN, one tile binning, a multi-channel deposit, one batched PSF FFT (particle_device_multi).
    X, Y, L[f]: NumPy arrays (as upstream), CPU tensors (pinned ones upload asynchronously) or CUDA tensors;
    host inputs give NumPy float64 images, CUDA inputs CUDA float32 tensors.
    The caller's X and Y are left exactly as upstream leaves them (recentred once per filter; all those steps are
    applied in one pass, so_recentre_chain).  Deviation, stated: a group is imaged at the positions its FIRST
    filter sees; upstream re-takes the median for every filter, which moves the later filters' positions by the
    rounding residue of the previous shift (<= 1 ulp-level of the coordinates, exactly 0 for odd N)."""
    if offsets:
        xoffset_pix_base = np.random.random() - 0.5
        yoffset_pix_base = np.random.random() - 0.5
    else:
        xoffset_pix_base = yoffset_pix_base = 0.0

    max_pixel_scale = np.max([filter_info[filter]['pixel_scale'] for filter in filters])

    obs = {}
    for filter in filters:
        xoffset_pix = xoffset_pix_base * (max_pixel_scale / filter_info[filter]['pixel_scale'])
        yoffset_pix = yoffset_pix_base * (max_pixel_scale / filter_info[filter]['pixel_scale'])  # noqa: F841 (unused upstream too)
        obs[filter] = observed(filter, cosmo, z, target_width_arcsec, resampling_factor=resampling_factor,
                               pixel_scale=pixel_scale, smoothing=smoothing, PSF=PSFs[filter],
                               super_sampling=super_sampling, verbose=verbose, xoffset_pix=xoffset_pix,
                               yoffset_pix=xoffset_pix)               # images.py:282 passes xoffset_pix twice
    groups = {}
    for filter in filters:
        o = obs[filter]
        key = (o.ndim, o.ndim_super, o.resolution, o.xoffset, o.yoffset, repr(o.smoothing))
        groups.setdefault(key, []).append(filter)

    host = not any(dv.is_device_tensor(a) for a in [X, Y] + [L[f] for f in filters])
    dev = dv.device()
    main = torch.cuda.current_stream()
    n = len(X)
    if host:
        # positions first (the k-NN and the binning only need them), the weights follow on a copy stream
        side = _side_stream(dev.index, dev)
        XY = torch.empty((2, n), dtype=torch.float64, device=dev)     # one block: both medians in one selection
        XY[0].copy_(_as_cpu_tensor(X), non_blocking=True)
        XY[1].copy_(_as_cpu_tensor(Y), non_blocking=True)
        Xd, Yd = XY[0], XY[1]
        xy_up = torch.cuda.Event()
        xy_up.record(main)
        Ld, L_ev = {}, {}
        with torch.cuda.stream(side):
            side.wait_event(xy_up)
            for key, fs in groups.items():
                buf = torch.empty((len(fs), n), dtype=torch.float64, device=dev)
                for c, f in enumerate(fs):
                    buf[c].copy_(_as_cpu_tensor(L[f]), non_blocking=True)
                buf.record_stream(main)
                Ld[key] = buf
                L_ev[key] = torch.cuda.Event()
                L_ev[key].record(side)
    else:
        Xd, Yd = dv.to_dev(X), dv.to_dev(Y)
        Ld = {key: torch.stack([dv.to_dev(L[f]) for f in fs]) for key, fs in groups.items()}
        L_ev = {key: None for key in groups}
    # the reference's per-filter in-place recentring (images.py:183-187 once per filter), all steps in one pass;
    # a group is imaged at the positions its FIRST filter sees
    first_step = {key: filters.index(fs[0]) + 1 for key, fs in groups.items()}
    out_steps = sorted(set(first_step.values()) | {len(filters)})
    pos = dict(zip(out_steps, recentre_chain_device(Xd, Yd, [obs[f].xoffset for f in filters],
                                                    [obs[f].yoffset for f in filters], out_steps)))

    # the caller's arrays end up exactly as the reference leaves them (recentred once per filter).  For host inputs the
    # recentred positions go back on a second copy stream as soon as they exist, under the k-NN and the deposit (the
    # device-to-host direction of the link is idle until the images are ready)
    Xf, Yf = pos[len(filters)]
    back, back_done = [], None
    if host:
        down = _side_stream((dev.index, 'down'), dev)
        rc_done = torch.cuda.Event()
        rc_done.record(main)
        with torch.cuda.stream(down):
            down.wait_event(rc_done)
            for A, Af in ((X, Xf), (Y, Yf)):
                Af.record_stream(down)
                if isinstance(A, torch.Tensor) and A.is_pinned() and A.dtype == torch.float64 and A.is_contiguous():
                    A.copy_(Af, non_blocking=True)      # straight into the caller's page-locked buffer
                else:
                    Ah = torch.empty(n, dtype=torch.float64, pin_memory=True)
                    Ah.copy_(Af, non_blocking=True)
                    back.append((A, Ah))
            back_done = torch.cuda.Event()
            back_done.record(down)

    IMGs = {}
    pending = []
    for key, fs in groups.items():
        Xg, Yg = pos[first_step[key]]
        r = particle_device_multi([obs[f] for f in fs], Xg, Yg, Ld[key], L_event=L_ev[key], offsets_applied=True)
        if host:
            # one page-locked float64 block per group: [2, C, ndim, ndim] (no_PSF, data); widened on the device (a
            # host-side astype of the block costs more than the whole imaging step)
            out = torch.empty((2,) + tuple(r['data'].shape), dtype=torch.float64, pin_memory=True)
            out[0].copy_(r['no_PSF'].to(torch.float64), non_blocking=True)
            out[1].copy_(r['data'].to(torch.float64), non_blocking=True)
            pending.append((fs, out))
        else:
            for c, f in enumerate(fs):
                img = empty()
                img.pixel_scale, img.pixel_scale_kpc = obs[f].pixel_scale, obs[f].pixel_scale_kpc
                img.no_PSF, img.data = r['no_PSF'][c], r['data'][c]
                IMGs[f] = img
    if host:
        main.synchronize()
        back_done.synchronize()
        for A, Ah in back:
            if isinstance(A, torch.Tensor):
                A.copy_(Ah)
            else:
                A[...] = Ah.numpy()
        for fs, out in pending:
            arr = out.numpy()          # views of the page-locked block
            for c, f in enumerate(fs):
                img = empty()
                img.pixel_scale, img.pixel_scale_kpc = obs[f].pixel_scale, obs[f].pixel_scale_kpc
                img.no_PSF, img.data = arr[0, c], arr[1, c]
                IMGs[f] = img
    else:
        if isinstance(X, torch.Tensor) and X.is_cuda and X.dtype == torch.float64:
            X.copy_(Xf)
        if isinstance(Y, torch.Tensor) and Y.is_cuda and Y.dtype == torch.float64:
            Y.copy_(Yf)
    return {f: IMGs[f] for f in filters}


def Sersic(L, p, filters, cosmo, z, target_width_arcsec, resampling_factor=False, pixel_scale=False, smoothing=False,
           PSFs=False, super_sampling=10, verbose=False, offsets=False):
    """reference: images.py:290-313 -- one observed Sersic image per filter."""
    if offsets:
        xoffset_pix_base = np.random.random() - 0.5
        yoffset_pix_base = np.random.random() - 0.5
    else:
        xoffset_pix_base = yoffset_pix_base = 0.0

    max_pixel_scale = np.max([pixel_scale_of[filter] for filter in filters])

    IMGs = {}
    for filter in filters:
        xoffset_pix = xoffset_pix_base * (max_pixel_scale / pixel_scale_of[filter])
        yoffset_pix = yoffset_pix_base * (max_pixel_scale / pixel_scale_of[filter])  # noqa: F841 (unused upstream too)
        imgs = observed(filter, cosmo, z, target_width_arcsec, resampling_factor=resampling_factor,
                        pixel_scale=pixel_scale, smoothing=smoothing, PSF=PSFs[filter],
                        super_sampling=super_sampling, verbose=verbose, xoffset_pix=xoffset_pix,
                        yoffset_pix=xoffset_pix).Sersic(L[filter], p)
        IMGs[filter] = imgs.img
    return IMGs


def point(flux, filter, target_width_arcsec, resampling_factor=False, pixel_scale=False, PSF=False, verbose=False,
          super_sampling=5, xoffset_pix=0.0, yoffset_pix=0.0):
    """reference: images.py:320-414 -- a point source: one NGP deposit on the super-sampled
    grid, PSF convolution (PSF normalised on the image, times the PSF fraction enclosed by the
    image footprint), rebin, renormalisation of the final image to `flux`."""
    native_pixel_scale = pixel_scale_of[filter]

    if resampling_factor:
        pixel_scale = native_pixel_scale / resampling_factor
    elif pixel_scale:
        resampling_factor = native_pixel_scale / pixel_scale
    else:
        pixel_scale = native_pixel_scale
        resampling_factor = 1.0

    ndim = int(target_width_arcsec / pixel_scale)
    ndim_super = ndim * super_sampling
    width_arcsec = ndim * pixel_scale
    xoffset_arcsec = xoffset_pix * pixel_scale
    yoffset_arcsec = yoffset_pix * pixel_scale

    if verbose:
        _report('POINT', [None, ('target width/"', target_width_arcsec, None), ('width/"', width_arcsec, None), None,
                          ('native pixel scale/"', native_pixel_scale, None), ('pixel scale/"', pixel_scale, None),
                          ('super sampling', super_sampling, None), ('ndim', ndim, None), None])

    host = not dv.is_device_tensor(flux)
    fd = dv.to_dev(np.atleast_1d(np.asarray(flux, dtype=np.float64))) if host else flux.reshape(-1).to(torch.float64)
    # the single source lands on its nearest super-pixel (images.py:370-373): one index pair, computed
    # with the reference's own expression; the images live on the device from here on
    g = np.linspace(-width_arcsec / 2., width_arcsec / 2., ndim_super)
    i, j = (np.abs(g - xoffset_arcsec)).argmin(), (np.abs(g - yoffset_arcsec)).argmin()
    hist = torch.zeros((ndim_super, ndim_super), dtype=torch.float32, device=fd.device)
    simple = torch.zeros((fd.numel(), ndim_super, ndim_super), dtype=torch.float32, device=fd.device)
    hist[j, i] += 1
    simple[:, j, i] += fd.to(torch.float32)

    xx = yy = np.linspace(-(width_arcsec / native_pixel_scale) / 2, width_arcsec / native_pixel_scale / 2, ndim_super)
    psf = PSF.f(xx, yy)                                                                         # images.py:379
    psf = psf / np.sum(psf)                                                                     # images.py:381

    d = int(PSF.ndim * width_arcsec / PSF.width / 2)                                            # images.py:387-389
    c = PSF.ndim // 2
    psf_data = np.sum(PSF.data[c - d:c + d + 1, c - d:c + d + 1]) / np.sum(PSF.data)

    if verbose:
        print('sum(PSF width): {0:.2f}'.format(PSF.width))
        print('sum(psf_data): {0:.2f}'.format(psf_data))
        print('sum(PSF): {0:.2f}'.format(np.sum(psf)))

    data_super = convolve_fft_device(simple, psf) * float(psf_data)                             # images.py:396
    no_psf = rebin_device(simple, ndim)                                                         # images.py:407
    data = rebin_device(data_super, ndim)                                                       # images.py:408
    tot = data.to(torch.float64).sum(dim=(-2, -1), keepdim=True)
    data = (data.to(torch.float64) / tot * fd.reshape(-1, 1, 1)).to(torch.float32)              # images.py:410-411

    scalar = np.ndim(flux) == 0 and not isinstance(flux, torch.Tensor)
    pick = (lambda t: t[0]) if scalar else (lambda t: t)
    fin = (lambda t: dv.to_host(pick(t))) if host else pick
    imgs = empty()
    imgs.super = empty()
    imgs.super.ndim = ndim_super
    imgs.super.pixel_scale = pixel_scale / super_sampling
    imgs.super.hist = dv.to_host(hist) if host else hist
    imgs.super.simple = fin(simple)
    imgs.super.data = fin(data_super)
    imgs.img = empty()
    imgs.img.ndim = ndim
    imgs.img.pixel_scale = pixel_scale
    imgs.img.no_PSF = fin(no_psf)
    imgs.img.data = fin(data)
    return imgs


def points(fluxes, filters, width_arcsec, resampling_factor=False, pixel_scale=False, PSFs=False, verbose=False):
    """reference: images.py:419-432 -- one point-source image per filter at a common random
    sub-pixel offset (global np.random, as upstream)."""
    xoffset_pix = np.random.random() - 0.5
    yoffset_pix = np.random.random() - 0.5

    IMGs = {}
    for filter in filters:
        imgs = point(fluxes[filter], filter, width_arcsec, resampling_factor=resampling_factor,
                     pixel_scale=pixel_scale, PSF=PSFs[filter], verbose=verbose, xoffset_pix=xoffset_pix,
                     yoffset_pix=yoffset_pix)
        IMGs[filter] = imgs.img
    return IMGs

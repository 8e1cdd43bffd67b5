"""TEST INFRASTRUCTURE — CPU (NumPy, float64) restatement of the reference's
particle->image hot path (synthobs/morph/images.py).  Not shipped, not a
fallback (see oracle/sed_oracle.py header for who may import this).

Parity status: pinned only by running the unmodified reference in this
container (tests/golden/make_golden.py -> tests/golden/*.npz; checked in
tests/test_oracle_golden.py).  astropy.convolution.convolve_fft is not
installed: its semantics are restated from recall in `convolve_fft_same`
("parity unpinned" at that boundary).  scipy's cKDTree is available and used
as-is for the k-th neighbour distance.

Two forms of the adaptive deposit are given: `adaptive_literal` keeps the
reference's per-particle full-image loop (small cases), `adaptive_separable`
uses the exact factorisation exp(-(dx^2+dy^2)/2s^2) = exp(-dx^2/2s^2) *
exp(-dy^2/2s^2) (equal to the loop to ~1e-15) so parity tests
can run at 10^4-10^5 particles in seconds.  `adaptive_window` / `adaptive_tiled`
evaluate the same float64 sum only where its terms are not exactly 0.0 (within
40 sigma of a particle), which takes the oracle to the benchmarked sizes
(1e6 particles on 1024^2; windows of an 8192^2 image); they are pinned against
adaptive_separable and the golden images in tests/test_oracle_golden.py.
"""

import numpy as np


def rebin(arr, new_shape):
    """synthobs/morph/images.py:19-22 — block SUM."""
    shape = (new_shape[0], arr.shape[0] // new_shape[0],
             new_shape[1], arr.shape[1] // new_shape[1])
    return arr.reshape(shape).sum(-1).sum(1)


def convolve_fft_same(array, kernel):
    """astropy convolve_fft defaults restated (see oracle/ref_loader.py
    convolve_fft for the derivation): zero fill, kernel normalised to sum 1 and
    centred on element n//2, output cropped to the array."""
    import scipy.fft as sfft
    a = np.asarray(array, dtype=np.float64)
    k = np.asarray(kernel, dtype=np.float64)
    k = k / k.sum()
    shape = [sfft.next_fast_len(a.shape[d] + k.shape[d] - 1, real=True) for d in range(2)]
    full = sfft.irfft2(sfft.rfft2(a, shape) * sfft.rfft2(k, shape), shape)
    oy, ox = k.shape[0] // 2, k.shape[1] // 2
    return full[oy:oy + a.shape[0], ox:ox + a.shape[1]].copy()


def pixel_grid(resolution, ndim):
    """images.py:39,53 — width = ndim*resolution; G = linspace(-w/2, w/2, ndim)
    (pitch is width/(ndim-1), not `resolution`)."""
    width = ndim * resolution
    return width, np.linspace(-width / 2., width / 2., ndim)


def select(X, Y, width):
    """images.py:47 — strict |X|<w/2 & |Y|<w/2."""
    return (np.fabs(X) < width / 2.) & (np.fabs(Y) < width / 2.)


def ngp_pixels(G, x, chunk=16384):
    """images.py:67 — literal first-minimum argmin |G - x| in float64."""
    x = np.asarray(x, dtype=np.float64)
    out = np.empty(x.shape[0], dtype=np.int64)
    for s in range(0, x.shape[0], chunk):
        e = min(x.shape[0], s + chunk)
        out[s:e] = np.abs(G[None, :] - x[s:e, None]).argmin(axis=1)
    return out


def kth_neighbour_distance(X, Y, k):
    """images.py:83-85 — distance to the k-th neighbour counting the particle
    itself as the first (nndist[-1]); inf when k exceeds the particle count."""
    from scipy.spatial import cKDTree
    pts = np.column_stack([X, Y])
    tree = cKDTree(pts, leafsize=16, compact_nodes=True, copy_data=False, balanced_tree=True)
    nndists, _ = tree.query(pts, k=int(k), workers=-1)
    if int(k) == 1:
        return np.asarray(nndists, dtype=np.float64).reshape(-1)
    return nndists[:, -1]


def adaptive_literal(G, X, Y, L, dk):
    """images.py:87-97 — per-particle Gaussian over the FULL image, normalised
    by its sum over the image; dropped when the sum is 0."""
    Gx, Gy = np.meshgrid(G, G)
    data = np.zeros((G.size, G.size))
    for x, y, l, d in zip(X, Y, L, dk):
        FWHM = np.max([d, 0.01])
        sigma = FWHM / 2.355
        with np.errstate(under='ignore'):
            gauss = np.exp(-(((Gx - x) ** 2 + (Gy - y) ** 2) / (2.0 * sigma ** 2)))
        sgauss = np.sum(gauss)
        if sgauss > 0:
            data += l * gauss / sgauss
    return data


def adaptive_separable(G, X, Y, L, dk, chunk=2048):
    """Same sum via the exact per-axis factorisation.  The reference drops a
    particle when every float64 exp underflows (sgauss == 0); that is decided
    here on the largest 2-D term exactly as the reference would see it."""
    n = X.shape[0]
    data = np.zeros((G.size, G.size))
    for s in range(0, n, chunk):
        e = min(n, s + chunk)
        sigma = np.maximum(dk[s:e], 0.01) / 2.355
        with np.errstate(under='ignore', over='ignore', invalid='ignore'):
            inv = 1.0 / (2.0 * sigma ** 2)
            ax2 = (G[None, :] - X[s:e, None]) ** 2
            ay2 = (G[None, :] - Y[s:e, None]) ** 2
            # shift by the per-axis minimum so narrow particles survive float
            # range; cancels exactly in the normalised product
            mx = ax2.min(axis=1)
            my = ay2.min(axis=1)
            ex = np.exp(-(ax2 - mx[:, None]) * inv[:, None])
            ey = np.exp(-(ay2 - my[:, None]) * inv[:, None])
            # reference: largest term exp(-(mx+my)*inv) == 0.0 in float64 -> dropped
            alive = np.exp(-((mx + my) / (2.0 * sigma ** 2))) > 0
        w = np.where(alive, L[s:e], 0.0) / (ex.sum(axis=1) * ey.sum(axis=1))
        data += (ey * w[:, None]).T @ ex
    return data


CUTOFF_SIGMAS = 40.0     # exp(-x) is exactly 0.0 in float64 for x > 745.2; 40 sigma -> x = 800


def _nearest_index(G, x):
    """index of a grid node nearest to x (any of two at a midpoint): only used to centre evaluation windows"""
    i = np.clip(np.searchsorted(G, x), 1, G.size - 1)
    return np.where(np.abs(G[i - 1] - x) <= np.abs(G[i] - x), i - 1, i)


def _axis_sums(G, x, inv, sigma, budget=1 << 22):
    """Per particle: m = min_i (G_i - x)^2 and S = sum_i exp(-((G_i - x)^2 - m) * inv) over ALL grid nodes, evaluated
    only on the nodes within CUTOFF_SIGMAS sigma (+2 nodes) of the particle: every other term is exactly 0.0 in
    float64, so S is the full-grid sum of adaptive_separable up to the order of the float64 additions.  Particles are
    grouped by support width (powers of two) so each group is one dense [n_g, 2W+1] evaluation."""
    n, ndim = x.size, G.size
    pitch = (G[-1] - G[0]) / (ndim - 1) if ndim > 1 else 1.0
    near = _nearest_index(G, x) if ndim > 1 else np.zeros(n, dtype=np.int64)
    with np.errstate(over='ignore', invalid='ignore'):
        Wf = np.ceil(CUTOFF_SIGMAS * sigma / pitch) + 2
    W = np.where(np.isfinite(Wf), np.minimum(Wf, ndim), ndim).astype(np.int64)
    bucket = np.ceil(np.log2(np.maximum(W, 1))).astype(np.int64)
    m = np.empty(n)
    S = np.empty(n)
    for b in np.unique(bucket):
        sel = np.nonzero(bucket == b)[0]
        Wb = 1 << int(b)
        width = min(2 * Wb + 1, ndim)
        step = max(1, budget // width)
        for s in range(0, sel.size, step):
            q = sel[s:s + step]
            if width == ndim:
                g = np.broadcast_to(G[None, :], (q.size, ndim))
                valid = None
            else:
                idx = near[q, None] + np.arange(-Wb, Wb + 1)[None, :]
                valid = (idx >= 0) & (idx < ndim)
                g = G[np.clip(idx, 0, ndim - 1)]
            d2 = (g - x[q, None]) ** 2
            if valid is not None:
                d2 = np.where(valid, d2, np.inf)
            mq = d2.min(axis=1)
            with np.errstate(under='ignore', over='ignore', invalid='ignore'):
                t = np.exp(-(d2 - mq[:, None]) * inv[q, None])
            if valid is not None:
                t = np.where(valid, t, 0.0)
            m[q] = mq
            S[q] = t.sum(axis=1)
    return m, S


def adaptive_norms(G, X, Y, dk):
    """Per-particle quantities of images.py:89-97 that do not depend on the pixel: sigma, 1/(2 sigma^2), the
    per-axis shifts and full-image factor sums, and the reference's `sgauss > 0` test (see adaptive_separable)."""
    sigma = np.maximum(dk, 0.01) / 2.355
    with np.errstate(under='ignore', over='ignore', invalid='ignore', divide='ignore'):
        inv = 1.0 / (2.0 * sigma ** 2)
        mx, Sx = _axis_sums(G, X, inv, sigma)
        my, Sy = _axis_sums(G, Y, inv, sigma)
        alive = np.exp(-((mx + my) / (2.0 * sigma ** 2))) > 0
    return {'sigma': sigma, 'inv': inv, 'mx': mx, 'my': my, 'Sx': Sx, 'Sy': Sy, 'alive': alive}


def adaptive_window(G, X, Y, L, dk, rows, cols, norms=None, chunk=4096):
    """images.py:87-97 restricted to the pixel window rows = (r0, r1), cols = (c0, c1) of the image: the
    same float64 sum as adaptive_separable (normalisation over ALL image pixels), but only the particles
    within CUTOFF_SIGMAS sigma of the window are evaluated -- the others contribute exactly 0.0 in float64.
    L: [n] or [C, n] (C images sharing positions).  Lets the oracle reach 8192^2 images / 1e6+ particles.
    `norms` = adaptive_norms of ALL particles when several windows share them; otherwise they are computed
    for the particles that reach the window only."""
    r0, r1 = rows
    c0, c1 = cols
    L2 = np.atleast_2d(np.asarray(L, dtype=np.float64))
    pitch = (G[-1] - G[0]) / (G.size - 1) if G.size > 1 else 1.0
    with np.errstate(over='ignore', invalid='ignore'):
        rad = CUTOFF_SIGMAS * (np.maximum(dk, 0.01) / 2.355) + pitch
        rad = np.where(np.isfinite(rad), rad, np.inf)
    gx, gy = G[c0:c1], G[r0:r1]
    hit = np.nonzero((X >= gx[0] - rad) & (X <= gx[-1] + rad) & (Y >= gy[0] - rad) & (Y <= gy[-1] + rad))[0]
    if norms is None:
        nm = adaptive_norms(G, X[hit], Y[hit], dk[hit])
    else:
        nm = {k: v[hit] for k, v in norms.items()}
    Xh, Yh, Lh = X[hit], Y[hit], L2[:, hit]
    data = np.zeros((L2.shape[0], gy.size, gx.size))
    for s in range(0, hit.size, chunk):
        q = slice(s, min(hit.size, s + chunk))
        inv = nm['inv'][q, None]
        with np.errstate(under='ignore', over='ignore', invalid='ignore'):
            ex = np.exp(-((gx[None, :] - Xh[q, None]) ** 2 - nm['mx'][q, None]) * inv)
            ey = np.exp(-((gy[None, :] - Yh[q, None]) ** 2 - nm['my'][q, None]) * inv)
            w = np.where(nm['alive'][q][None, :], Lh[:, q], 0.0) / (nm['Sx'][q] * nm['Sy'][q])[None, :]
        for c in range(L2.shape[0]):
            data[c] += (ey * w[c][:, None]).T @ ex
    return data if np.ndim(L) == 2 else data[0]


def adaptive_tiled(G, X, Y, L, dk, block=128):
    """The whole image as a mosaic of adaptive_window blocks: equals adaptive_separable (to the order of the
    float64 additions) at a cost proportional to the particles' supports instead of N x ndim."""
    nm = adaptive_norms(G, X, Y, dk)
    n = G.size
    L2 = np.atleast_2d(np.asarray(L, dtype=np.float64))
    out = np.zeros((L2.shape[0], n, n))
    for r0 in range(0, n, block):
        for c0 in range(0, n, block):
            r1, c1 = min(n, r0 + block), min(n, c0 + block)
            out[:, r0:r1, c0:c1] = adaptive_window(G, X, Y, L2, dk, (r0, r1), (c0, c1), norms=nm)
    return out if np.ndim(L) == 2 else out[0]


def core(X, Y, L, resolution=0.1, ndim=100, smoothing=False, literal=False, return_aux=False, tiled=False):
    """synthobs/morph/images.py:28-105 — images.core.  Returns dict with
    hist, simple, data ([ndim, ndim], row = y, col = x); with return_aux also
    the selection mask, NGP pixel indices and (adaptive) k-th neighbour
    distances so index parity can be asserted bit-exactly."""
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    L = np.asarray(L, dtype=np.float64)
    method, scale = smoothing if smoothing else (False, None)
    width, G = pixel_grid(resolution, ndim)
    sel = select(X, Y, width)
    Xs, Ys, Ls = X[sel], Y[sel], L[sel]
    i = ngp_pixels(G, Xs)
    j = ngp_pixels(G, Ys)
    hist = np.zeros((ndim, ndim))
    simple = np.zeros((ndim, ndim))
    np.add.at(simple, (j, i), Ls)                                   # :68
    np.add.at(hist, (j, i), 1.0)                                    # :69
    aux = {'sel': sel, 'i': i, 'j': j, 'G': G}
    if method == 'convolved_gaussian':                              # :72-78
        Gx, Gy = np.meshgrid(G, G)
        sigma = scale / 2.355
        gauss = np.exp(-((Gx ** 2 + Gy ** 2) / (2.0 * sigma ** 2)))
        gauss /= np.sum(gauss)
        data = convolve_fft_same(simple, gauss)
    elif method == 'adaptive':                                      # :81-97
        if Xs.size:
            dk = kth_neighbour_distance(Xs, Ys, scale)
        else:
            dk = np.zeros(0)
        aux['dk'] = dk
        fn = adaptive_literal if literal else (adaptive_tiled if tiled else adaptive_separable)
        data = fn(G, Xs, Ys, Ls, dk)
    else:
        data = simple                                               # :103
    out = {'hist': hist, 'simple': simple, 'data': data}
    if return_aux:
        out['aux'] = aux
    return out


def observed_geometry(native_pixel_scale, arcsec_per_proper_kpc, target_width_arcsec,
                      resampling_factor=False, pixel_scale=False, super_sampling=4,
                      xoffset_pix=0.0, yoffset_pix=0.0):
    """images.py:125-157 — observed.__init__ scalar geometry."""
    g = {}
    if resampling_factor:
        g['pixel_scale'] = native_pixel_scale / resampling_factor
        g['resampling_factor'] = resampling_factor
    elif pixel_scale:
        g['pixel_scale'] = pixel_scale
        g['resampling_factor'] = native_pixel_scale / pixel_scale
    else:
        g['pixel_scale'] = native_pixel_scale
        g['resampling_factor'] = 1.0
    g['native_pixel_scale'] = native_pixel_scale
    g['pixel_scale_kpc'] = g['pixel_scale'] / arcsec_per_proper_kpc
    g['ndim'] = int(target_width_arcsec / g['pixel_scale'])
    g['width_arcsec'] = g['ndim'] * g['pixel_scale']
    g['width'] = g['width_arcsec'] / arcsec_per_proper_kpc
    g['resolution'] = g['pixel_scale_kpc'] / super_sampling
    g['ndim_super'] = g['ndim'] * super_sampling
    g['xoffset'] = xoffset_pix * g['pixel_scale_kpc']
    g['yoffset'] = yoffset_pix * g['pixel_scale_kpc']
    return g


def observed_particle(g, X, Y, L, psf_f, smoothing=False):
    """images.py:181-213 — observed.particle on COPIES of X, Y (the reference
    recentres the caller's arrays in place, :183-187)."""
    X = X - np.median(X)
    Y = Y - np.median(Y)
    X = X + g['xoffset']
    Y = Y + g['yoffset']
    sup = core(X, Y, L, resolution=g['resolution'], ndim=g['ndim_super'], smoothing=smoothing)
    no_PSF = sup['data']
    half = g['width_arcsec'] / g['native_pixel_scale'] / 2.
    xx = yy = np.linspace(-half, half, g['ndim_super'])                  # :198
    psf = psf_f(xx, yy)                                                  # :200
    data = convolve_fft_same(no_PSF, psf)                                # :202
    nd = g['ndim']
    return {'super_no_PSF': no_PSF, 'super_data': data, 'psf': psf, 'X': X, 'Y': Y,
            'no_PSF': rebin(no_PSF, (nd, nd)), 'data': rebin(data, (nd, nd))}


def observed_sersic(g, L, p, psf_f, sersic2d):
    """images.py:219-257 -- observed.Sersic: unit-sum Sersic2D profile on the super-sampled
    grid times L, PSF convolution, rebin.  `sersic2d` is the Sersic2D class (third party:
    astropy; restated in synthobs_b200/providers.py, parity unpinned at that boundary)."""
    gg = np.linspace(-g['width'] / 2., g['width'] / 2., g['ndim_super'])                         # :224
    xx, yy = np.meshgrid(gg, gg)
    mod = sersic2d(amplitude=1, r_eff=p['r_eff'], n=p['n'], x_0=g['xoffset'], y_0=g['yoffset'],
                   ellip=p['ellip'], theta=p['theta'])                                           # :228
    sersic = mod(xx, yy)
    sersic /= np.sum(sersic)                                                                     # :233
    no_PSF = L * sersic                                                                          # :238
    half = g['width_arcsec'] / g['pixel_scale'] / 2.
    xa = np.linspace(-half, half, g['ndim_super'])                                               # :242 (pixel_scale, not native)
    psf = psf_f(xa, xa)
    data = convolve_fft_same(no_PSF, psf)                                                        # :246
    nd = g['ndim']
    return {'super_no_PSF': no_PSF, 'super_data': data, 'no_PSF': rebin(no_PSF, (nd, nd)),
            'data': rebin(data, (nd, nd))}


def point_source(flux, native_pixel_scale, target_width_arcsec, PSF, resampling_factor=False, pixel_scale=False,
                 super_sampling=5, xoffset_pix=0.0, yoffset_pix=0.0):
    """images.py:320-414 -- point()."""
    if resampling_factor:                                                                        # :325-333
        pixel_scale = native_pixel_scale / resampling_factor
    elif pixel_scale:
        resampling_factor = native_pixel_scale / pixel_scale
    else:
        pixel_scale = native_pixel_scale
    ndim = int(target_width_arcsec / pixel_scale)                                                # :337
    ndim_super = ndim * super_sampling
    width_arcsec = ndim * pixel_scale
    xo, yo = xoffset_pix * pixel_scale, yoffset_pix * pixel_scale
    hist = np.zeros((ndim_super, ndim_super))
    simple = np.zeros((ndim_super, ndim_super))
    gg = np.linspace(-width_arcsec / 2., width_arcsec / 2., ndim_super)                          # :370
    i, j = (np.abs(gg - xo)).argmin(), (np.abs(gg - yo)).argmin()
    hist[j, i] += 1
    simple[j, i] += flux
    xx = np.linspace(-(width_arcsec / native_pixel_scale) / 2, width_arcsec / native_pixel_scale / 2, ndim_super)
    psf = PSF.f(xx, xx)                                                                          # :379
    psf = psf / np.sum(psf)
    d = int(PSF.ndim * width_arcsec / PSF.width / 2)                                             # :387
    c = PSF.ndim // 2
    psf_data = np.sum(PSF.data[c - d:c + d + 1, c - d:c + d + 1]) / np.sum(PSF.data)
    sdata = convolve_fft_same(simple, psf) * psf_data                                            # :396
    no_PSF = rebin(simple, (ndim, ndim))
    data = rebin(sdata, (ndim, ndim))
    data = data / np.sum(data) * flux                                                            # :410-411
    return {'super_hist': hist, 'super_simple': simple, 'super_data': sdata, 'no_PSF': no_PSF, 'data': data,
            'ndim': ndim, 'pixel_scale': pixel_scale}


def support_halfwidth(dk, pitch, R):
    """Product-side rule restated for the tile-assignment parity test: a
    particle's truncated support is +-W pixels around its NGP pixel,
    W = ceil(R * sigma / pitch), sigma = max(dk, 0.01)/2.355; W is unbounded
    (clipped to the image by the caller) when sigma is inf."""
    sigma = np.maximum(dk, 0.01) / 2.355
    with np.errstate(over='ignore', invalid='ignore'):
        w = np.ceil(R * sigma / pitch)
    return w


def tile_pairs(ip, jp, W, ndim, T):
    """Product-side rule restated: particle p overlaps tile (ty, tx) when the
    pixel box [ip-W, ip+W] x [jp-W, jp+W], clipped to the image, intersects the
    TxT tile.  Returns (tile_id, particle_id) pairs sorted by tile then
    particle — the order the device CSR must have."""
    ntile = (ndim + T - 1) // T
    Wc = np.minimum(W, float(ndim)).astype(np.int64)
    x0 = np.clip(ip - Wc, 0, ndim - 1) // T
    x1 = np.clip(ip + Wc, 0, ndim - 1) // T
    y0 = np.clip(jp - Wc, 0, ndim - 1) // T
    y1 = np.clip(jp + Wc, 0, ndim - 1) // T
    tiles, pids = [], []
    for p in range(ip.shape[0]):
        ty, tx = np.meshgrid(np.arange(y0[p], y1[p] + 1), np.arange(x0[p], x1[p] + 1), indexing='ij')
        t = (ty * ntile + tx).ravel()
        tiles.append(t)
        pids.append(np.full(t.shape, p, dtype=np.int64))
    if not tiles:
        return np.zeros(0, np.int64), np.zeros(0, np.int64)
    tiles = np.concatenate(tiles)
    pids = np.concatenate(pids)
    order = np.lexsort((pids, tiles))
    return tiles[order], pids[order]

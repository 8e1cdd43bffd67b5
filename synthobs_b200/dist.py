"""Multi-GPU partitioning of the two hot paths (one process per GPU, torch.distributed).

SURVEY.md §8e: (i) batches of galaxies are independent units -> shard galaxies over ranks,
balanced by particle count, NO collective on the compute path, one gather of the small
[G, F] result; (ii) a single huge image -> particle blocks per rank, each rank deposits onto
a private full image, one sum all-reduce of the image over NVLink (NCCL).

The partitioning / gather logic is backend-agnostic (it is exercised with gloo on CPU in
tests/test_dist_gloo.py); the compute callbacks are the CUDA paths.
"""

import os

import numpy as np
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend=None):
    """torchrun-style initialisation (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
    ws = int(os.environ.get('WORLD_SIZE', '1'))
    if ws <= 1 or dist.is_initialized():
        return world()
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        torch.cuda.set_device(local)
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend)
    return world()


def shard_galaxies(sizes, n_ranks):
    """Longest-processing-time-first assignment of galaxies to ranks, balanced by particle
    count.  Returns a list (per rank) of ascending galaxy indices; deterministic."""
    sizes = np.asarray(sizes, dtype=np.int64)
    order = np.argsort(-sizes, kind='stable')
    load = np.zeros(n_ranks, dtype=np.int64)
    out = [[] for _ in range(n_ranks)]
    for g in order:
        r = int(np.argmin(load))          # first minimum: deterministic tie-break
        out[r].append(int(g))
        load[r] += sizes[g]
    return [sorted(o) for o in out]


_gather_plans = {}


def _gather_plan(all_index, n_total, device):
    """index bookkeeping of gather_rows for one sharding (cached: the same catalogue is gathered every step)"""
    key = (int(n_total), str(device), hash(b'|'.join(np.asarray(i, dtype=np.int64).tobytes() for i in all_index)))
    hit = _gather_plans.get(key)
    if hit is None:
        if len(_gather_plans) > 32:
            _gather_plans.clear()
        cmax = max(1, max(len(i) for i in all_index))
        src = np.concatenate([r * cmax + np.arange(len(i), dtype=np.int64) for r, i in enumerate(all_index)]) \
            if n_total else np.zeros(0, np.int64)
        dst = np.concatenate([np.asarray(i, dtype=np.int64) for i in all_index]) if n_total else np.zeros(0, np.int64)
        order = np.empty(n_total, dtype=np.int64)
        order[dst] = src                       # row g of the result = row order[g] of the gathered padded blocks
        hit = _gather_plans[key] = (cmax, torch.as_tensor(order, device=device))
    return hit


def gather_rows(local_rows, local_index, n_total, device=None, all_index=None):
    """Every rank holds rows for the galaxy indices `local_index`; returns the full [n_total, ...] array on every
    rank.  ONE all_gather of the padded blocks -- the only collective of the galaxy-sharded path, a few MB at most.
    all_index = the index lists of ALL ranks (the sharding is deterministic, so every rank knows them: no size
    exchange and no host synchronisation); without it the sizes are exchanged first."""
    rank, ws = world()
    rows = torch.as_tensor(local_rows)
    if ws == 1:
        out = torch.empty((n_total,) + tuple(rows.shape[1:]), dtype=rows.dtype, device=rows.device)
        out[torch.as_tensor(local_index, dtype=torch.long, device=rows.device)] = rows
        return out
    dev = rows.device if device is None else device
    if all_index is None:
        counts = torch.zeros(ws, dtype=torch.long, device=dev)
        counts[rank] = len(local_index)
        dist.all_reduce(counts)
        cmax = int(counts.max().item())
        pad_idx = torch.full((cmax,), -1, dtype=torch.long, device=dev)
        pad_idx[:len(local_index)] = torch.as_tensor(local_index, dtype=torch.long, device=dev)
        all_idx = [torch.empty_like(pad_idx) for _ in range(ws)]
        dist.all_gather(all_idx, pad_idx)
        all_index = [t[:int(c)].cpu().numpy() for t, c in zip(all_idx, counts.cpu())]
    cmax, order = _gather_plan(all_index, n_total, dev)
    F = tuple(rows.shape[1:])
    pad_rows = torch.empty((cmax,) + F, dtype=rows.dtype, device=dev)        # the padding rows are never read back
    pad_rows[:len(local_index)] = rows.to(dev)
    gathered = torch.empty((ws * cmax,) + F, dtype=rows.dtype, device=dev)
    dist.all_gather_into_tensor(gathered, pad_rows)
    return gathered[order]


def sharded_galaxy_map(compute, offsets, n_filters_hint=None):
    """Run `compute(galaxy_indices) -> [len(indices), F] tensor` on this rank's share of the
    galaxies described by CSR `offsets` and return the full [G, F] result on every rank."""
    rank, ws = world()
    offsets = np.asarray(offsets, dtype=np.int64)
    sizes = np.diff(offsets)
    shards = shard_galaxies(sizes, ws)
    mine = shards[rank]
    rows = compute(mine)
    return gather_rows(rows, mine, sizes.size, all_index=shards)


def subset_csr(arrays, offsets, indices):
    """Concatenate the particle ranges of the chosen galaxies: (dict of arrays, new offsets)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    idx = np.asarray(indices, dtype=np.int64)
    lens = offsets[idx + 1] - offsets[idx]
    new_off = np.zeros(idx.size + 1, dtype=np.int64)
    np.cumsum(lens, out=new_off[1:])
    sel = np.concatenate([np.arange(offsets[g], offsets[g + 1]) for g in idx]) if idx.size else np.zeros(0, np.int64)
    return {k: v[sel] for k, v in arrays.items()}, new_off


def particle_block(n, rank=None, n_ranks=None):
    """Contiguous block [begin, end) of n particles owned by a rank."""
    r, ws = world()
    rank = r if rank is None else rank
    n_ranks = ws if n_ranks is None else n_ranks
    per = (n + n_ranks - 1) // n_ranks
    b = min(n, rank * per)
    return b, min(n, b + per)


def reduce_image(img):
    """Sum the ranks' partial images in place (NCCL all-reduce over NVLink on GPUs).  FP32
    summation order differs between rank counts (~1e-7 relative), inside the tolerance."""
    rank, ws = world()
    if ws > 1:
        dist.all_reduce(img, op=dist.ReduceOp.SUM)
    return img


class _DevicePtr:
    """a cudaMalloc'ed buffer seen as a CUDA array (torch.as_tensor maps it without a copy)"""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {'shape': tuple(int(v) for v in shape), 'typestr': '<f4', 'data': (int(ptr), False),
                                         'version': 3, 'strides': None}


class PeerImage:
    """Peer-mapped buffers of ONE image [C, ndim, ndim] shared by the GPUs of a node (so_img_deposit_push): the partial
    image, two alternating result images and tile-mask arrays, and the barrier flags -- allocated with cudaMalloc,
    exported with cudaIpc and mapped by every other rank over NVLink.  Created collectively, once per image shape."""

    _cache = {}

    @classmethod
    def get(cls, n_planes, ndim):
        from . import device as dv
        rank, ws = world()
        key = (int(n_planes), int(ndim), ws, dv.device().index)
        if key not in cls._cache:
            cls._cache[key] = cls(n_planes, ndim)
        return cls._cache[key]

    def __init__(self, n_planes, ndim):
        from ctypes import byref, c_void_p, create_string_buffer
        from . import _cabi, device as dv
        self.rank, self.ws = world()
        if self.ws > _cabi.SO_MAX_PEERS:
            raise ValueError('at most %d ranks share an image' % _cabi.SO_MAX_PEERS)
        self.n_planes, self.ndim = int(n_planes), int(ndim)
        self.dev = dv.device()
        img = 4 * self.n_planes * self.ndim * self.ndim
        ntiles = ((self.ndim + 15) // 16) ** 2
        sizes = {'partial': img, 'result0': img, 'result1': img, 'masks0': self.ws * ntiles, 'masks1': self.ws * ntiles,
                 'flags': 8 * (_cabi.SO_MAX_PEERS + 2)}
        self.names = list(sizes)
        own = {}
        for k, nbytes in sizes.items():
            p = c_void_p()
            _cabi.check(_cabi.lib.so_peer_alloc(nbytes, byref(p)), 'so_peer_alloc')
            own[k] = p.value
        handles = {}
        for k, p in own.items():
            h = create_string_buffer(64)
            _cabi.check(_cabi.lib.so_ipc_get_handle(c_void_p(p), h), 'so_ipc_get_handle')
            handles[k] = h.raw
        every = [None] * self.ws
        dist.all_gather_object(every, handles)
        self.ptrs = {k: [None] * self.ws for k in sizes}           # [buffer][rank]
        for r in range(self.ws):
            for k in sizes:
                if r == self.rank:
                    self.ptrs[k][r] = own[k]
                else:
                    q = c_void_p()
                    _cabi.check(_cabi.lib.so_ipc_open(create_string_buffer(every[r][k], 64), byref(q)), 'so_ipc_open')
                    self.ptrs[k][r] = q.value
        shape = (self.n_planes, self.ndim, self.ndim)
        self.partial = torch.as_tensor(_DevicePtr(own['partial'], shape), device=self.dev)
        self.results = [torch.as_tensor(_DevicePtr(own['result%d' % b], shape), device=self.dev) for b in range(2)]
        flags = torch.as_tensor(_DevicePtr(own['flags'], (2 * (_cabi.SO_MAX_PEERS + 2),)), device=self.dev)   # uint64 seen as 2 floats
        flags.zero_()
        self.waits = flags.view(torch.int64)[_cabi.SO_MAX_PEERS:]      # [2] ns this rank waited at the barriers of the last call
        torch.cuda.synchronize()
        self.calls = 0
        dist.barrier()

    def push_desc(self):
        """so_push_t of the next call (alternating result / mask buffers, barrier generation)"""
        from . import _cabi
        b = self.calls % 2
        d = _cabi.PushDesc()
        d.n_ranks, d.rank, d.epoch = self.ws, self.rank, self.calls + 1
        for r in range(self.ws):
            d.partial[r] = self.ptrs['partial'][r]
            d.result[r] = self.ptrs['result%d' % b][r]
            d.masks[r] = self.ptrs['masks%d' % b][r]
            d.flags[r] = self.ptrs['flags'][r]
        return d

    def finish(self):
        """the result image of the call just made (valid until the call after the next one writes it again)"""
        out = self.results[self.calls % 2]
        self.calls += 1
        return out


_share_balance = {}      # (n, ndim, k, ranks) -> work fractions of the shares, adapted to the measured busy times
SHARE_ADAPT_CALLS = 6    # calls of one (catalogue size, image, k, ranks) after which the fractions are frozen: reading the
                         # busy time of the previous call is a host synchronisation per call (measured: ~0.4 ms of a 12 ms step)


def peer_reduce_enabled():
    """SO_PEER_REDUCE=0 selects the plain NCCL all-reduce of the whole image"""
    return os.environ.get('SO_PEER_REDUCE', 'auto') != '0'


def reduce_mode():
    """How image_huge sums the partial images: 'peer' (SO_PEER_REDUCE=1: the block exchange over peer memory fused into
    the deposit, csrc/peer.cu), 'nccl' (SO_PEER_REDUCE=0: one ncclAllReduce of the whole image) or 'auto' (default):
    both are timed on the device in the first calls of a configuration (after the share fractions have settled) and
    the faster one - the maximum over the ranks decides, so every rank agrees - is kept.  Measured on 8 x B200 with
    NVSwitch the in-switch all-reduce of the 268 MB image costs ~0.5 ms and wins by 2-4 % at 2, 4 and 8 ranks."""
    v = os.environ.get('SO_PEER_REDUCE', 'auto')
    return 'nccl' if v == '0' else ('peer' if v == '1' else 'auto')


REDUCE_TRIALS = ('peer', 'peer', 'nccl', 'nccl')      # after SHARE_ADAPT_CALLS calls; the first of each pair is a warm-up


SED_KEYS = ('Masses', 'Ages', 'Metallicities', 'tauVs_ISM', 'tauVs_BC')
MAX_PASS_PARTICLES = 40000000      # particles per images.core_batch pass of a share (core_batch_sharded)


_last_shards = [None]


def shard_plan(offsets, rank=None, n_ranks=None):
    """This rank's galaxies of a catalogue with CSR `offsets`: (ascending galaxy indices, their local CSR offsets)."""
    r, ws = world()
    rank = r if rank is None else rank
    n_ranks = ws if n_ranks is None else n_ranks
    offsets = np.asarray(offsets, dtype=np.int64)
    sizes = np.diff(offsets)
    shards = shard_galaxies(sizes, n_ranks)
    mine = np.asarray(shards[rank], dtype=np.int64)
    local = np.zeros(mine.size + 1, dtype=np.int64)
    np.cumsum(sizes[mine], out=local[1:])
    _last_shards[0] = shards
    return mine, local


def local_share(batch, keys=SED_KEYS):
    """Extract this rank's share of a host catalogue (dict of arrays + 'offsets') ONCE: {'indices', 'offsets'
    (local CSR), 'n_total', arrays...}.  The sharded entry points take it as `share=` so a resident share is
    not re-extracted (or re-uploaded, if its arrays are CUDA tensors) per call."""
    mine, _ = shard_plan(batch['offsets'])
    sub, off = subset_csr({k: np.asarray(batch[k]) for k in keys if k in batch}, batch['offsets'], mine)
    sub.update(indices=mine, offsets=off, n_total=int(np.asarray(batch['offsets']).size - 1), all_indices=_last_shards[0])
    return sub


def _gather_share(rows, share):
    return gather_rows(rows, share['indices'], share['n_total'], all_index=share.get('all_indices'))


def generate_Fnu_sharded(model, F, batch, kind='Fnu', share=None, **kw):
    """Broadband fluxes of a batch of galaxies, sharded by galaxy over the ranks (LPT by particle count, no
    collective on the compute path, one all_gather of the [G, F] result).  `batch` is a dict with Masses, Ages,
    Metallicities, tauVs_ISM, tauVs_BC, offsets (host arrays); `share` = local_share(batch) (or a dict of the same
    layout whose arrays already live on the device) skips the extraction."""
    from . import device as dv
    from .sed import models
    sh = local_share(batch) if share is None else share
    if len(sh['indices']) == 0:      # fewer galaxies than ranks: this rank only takes part in the gather
        rows = torch.zeros((0, len(F['filters'])), dtype=torch.float64, device=dv.device())
    else:
        out = models.generate_Lnu_batch(model, F, sh['Masses'], sh['Ages'], sh['Metallicities'], sh['offsets'],
                                        tauVs_ISM=sh.get('tauVs_ISM', False), tauVs_BC=sh.get('tauVs_BC', False),
                                        kind=kind, **kw)
        rows = torch.as_tensor(out).to(dv.device())
    return _gather_share(rows, sh)


def generate_SED_sharded(model, batch, share=None, **kw):
    """The five spectra of generate_SED for a batch of galaxies, sharded by galaxy (config 4): returns a dict of
    [G, n_lam] float32 CUDA tensors, complete on every rank (one gather of the [g, 5, n_lam] block)."""
    from . import device as dv
    from .sed import models
    names = ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC')
    sh = local_share(batch) if share is None else share
    if len(sh['indices']) == 0:
        rows = torch.zeros((0, len(names), model.lam.size), dtype=torch.float32, device=dv.device())
    else:
        out = models.sed_spectra_device(model, sh['Masses'], sh['Ages'], sh['Metallicities'],
                                        sh.get('tauVs_ISM', False), sh.get('tauVs_BC', False),
                                        kw.get('fesc', 0.0), kw.get('log10t_BC', 7.), offsets=sh['offsets'])
        rows = torch.stack([out[k] for k in names], dim=1)          # [g, 5, n_lam]
    full = _gather_share(rows, sh)
    return {k: full[:, i] for i, k in enumerate(names)}


def core_batch_sharded(X, Y, L, offsets, resolution=0.1, ndim=100, smoothing=False, share=None, gather=True):
    """images.core for a batch of objects, sharded by object over the ranks (config 4): every rank images its
    share in one pass (images.core_batch) and the [G, ndim, ndim] stack is gathered on every rank (gather=False:
    the local [g, ndim, ndim] block and its galaxy indices are returned instead -- 1e4 images of 100 x 100 are
    400 MB per rank once gathered).  `share` = local_share({'X', 'Y', 'L', 'offsets'}, keys=('X', 'Y', 'L'))."""
    from . import device as dv
    from .morph import images
    sh = local_share({'X': X, 'Y': Y, 'L': L, 'offsets': offsets}, keys=('X', 'Y', 'L')) if share is None else share
    if len(sh['indices']) == 0:
        rows = torch.zeros((0, ndim, ndim), dtype=torch.float32, device=dv.device())
    else:
        # passes of at most MAX_PASS_PARTICLES particles: bounded workspace whatever the share holds
        off = np.asarray(sh['offsets'], dtype=np.int64)
        Xd, Yd, Ld = dv.to_dev(sh['X']), dv.to_dev(sh['Y']), dv.to_dev(sh['L'])
        parts, g0 = [], 0
        while g0 < off.size - 1:
            g1 = int(np.searchsorted(off, off[g0] + MAX_PASS_PARTICLES, side='right')) - 1
            g1 = min(max(g1, g0 + 1), off.size - 1)
            a, b = int(off[g0]), int(off[g1])
            data, _, _, _ = images.core_device(Xd[a:b], Yd[a:b], Ld[a:b], resolution, ndim, smoothing,
                                               img=images._image_ids(off[g0:g1 + 1] - off[g0], dv.device())[1],
                                               n_img=g1 - g0)
            parts.append(data[:, 0])
            g0 = g1
        rows = torch.cat(parts) if len(parts) > 1 else parts[0]
    if not gather:
        return rows, sh['indices']
    return _gather_share(rows, sh)


def image_huge(X, Y, L, resolution, ndim, k, support_sigmas=None, stages=None, rank=None, n_ranks=None, reduce=True):
    """One huge adaptive-smoothed image split over the ranks (BASELINE config 5).  Every rank holds all positions,
    but works on a spatially compact SHARE of the particles (a contiguous Morton range of a coarse grid over the
    image, equal counts): it builds the search tree only for its share plus the halo that makes the k-th neighbour
    distances of the share exact (so_share_plan), answers the k-NN queries of the share, deposits the share onto a
    private image, and the partial images are summed with one all-reduce (NCCL over NVLink).  Work per rank is
    ~1/N (+ the halo, + two passes over all positions for the coarse histogram and the flags).
    Returns the full [C, ndim, ndim] (or [ndim, ndim] for 1-D L) float32 CUDA image on every rank.
    rank / n_ranks override the process group (reduce=False: the partial image of that share, e.g. to add the shares
    on one GPU); `stages` (dict) is filled with the device time [ms] of the stages of this call (CUDA events)."""
    from . import device as dv
    from .morph import images
    r0, w0 = world()
    rank = r0 if rank is None else int(rank)
    ws = w0 if n_ranks is None else int(n_ranks)
    marks = []

    def mark(name):
        if stages is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            marks.append((name, e))
    Xd, Yd, Ld = dv.to_dev(X), dv.to_dev(Y), dv.to_dev(L)
    one = Ld.dim() == 1
    Ld = Ld.reshape(1, -1) if one else Ld
    mark('start')
    width, G, Gd = images.pixel_grid_device(resolution, ndim)
    hw = width / 2.
    if ws == 1:
        Xo, Yo, Lo = Xd, Yd, Ld
        ix, iy = images.ngp_index_device(Xo, Yo, Gd, ndim, hw)
        mark('nearest-pixel index')
        dko = images.knn_device(Xo, Yo, ix, int(k), hw, floor=images.FWHM_FLOOR)
        mark('k-NN')
    else:
        n = Xd.numel()
        cb = images.share_grid_bits(n, ws)
        bounds = None
        bal = None
        if w0 == ws and w0 > 1:
            # every rank counts its 1/N of the particles; the small table is summed over the ranks (exact integers).
            # The same all-reduce carries every rank's busy time of the previous call (below)
            bal = _share_balance.setdefault((n, int(ndim), int(k), ws), {'frac': np.full(ws, 1.0 / ws), 'pending': None, 'calls': 0, 'hist': []})
            cells = 1 << (2 * cb)
            buf = torch.zeros(cells + ws, dtype=torch.int32, device=Xd.device)
            b, e = particle_block(n, rank, ws)
            images.share_hist_device(Xd, Yd, hw, cb, b, e, out=buf[:cells])
            if bal['pending'] is not None:
                e0, e1, waits = bal['pending']
                e1.synchronize()
                busy_us = max(1.0, e0.elapsed_time(e1) * 1e3 - float(waits.sum()) * 1e-3)
                buf[cells + rank] = int(busy_us)
            dist.all_reduce(buf, op=dist.ReduceOp.SUM)
            hist = buf[:cells]
            if bal['pending'] is not None:
                t = buf[cells:].cpu().numpy().astype(np.float64)
                if np.all(t > 0):
                    bal['hist'].append((float(t.max()), bal['frac'].copy()))     # what the previous call's fractions cost
                    if bal['calls'] < SHARE_ADAPT_CALLS:
                        # shares proportional to the measured speed of their rank, damped; identical on every rank
                        f = bal['frac'] * (t.mean() / t) ** 0.7
                        f = np.clip(f / f.sum(), 0.4 / ws, 2.5 / ws)
                        bal['frac'] = f / f.sum()
                    elif bal['hist']:
                        # last measurement: freeze on the best fractions seen (the damped iteration may still be swinging)
                        bal['frac'] = min(bal['hist'], key=lambda h: h[0])[1]
            bounds = np.concatenate([[0.0], np.cumsum(bal['frac'])])
        else:
            hist = images.share_hist_device(Xd, Yd, hw, cb)
        flags, counts = images.share_plan_device(Xd, Yd, hw, int(k), rank, ws, cb=cb, hist=hist, bounds=bounds)
        if bal is not None:
            e0 = torch.cuda.Event(enable_timing=True)
            e0.record()
        n_loc, n_own = (int(v) for v in counts.cpu())        # the one host synchronisation: the sizes of the share
        mark('share plan (coarse histogram, halo flags)')
        loc, Xl, Yl, fl = images.share_compact_device(Xd, Yd, flags, 1, n_loc)      # local = own + halo
        mark('compaction of the local particles')
        ix, iy = images.ngp_index_device(Xl, Yl, Gd, ndim, hw)
        mark('nearest-pixel index')
        dkl = images.knn_device_masked(Xl, Yl, ix, int(k), hw, fl, floor=images.FWHM_FLOOR)
        mark('k-NN (tree of the local particles, queries of the share)')
        own, Xo, Yo, _ = images.share_compact_device(Xl, Yl, fl, 2, n_own)
        dko = dkl.index_select(0, own)
        Lo = Ld.index_select(1, loc.index_select(0, own))
        mark('compaction of the own particles')
    peer = None
    mode, trial = reduce_mode(), None
    if reduce and ws > 1 and ws == w0 and mode == 'auto':
        bal = _share_balance.get((Xd.numel(), int(ndim), int(k), ws))
        c = bal['calls'] - SHARE_ADAPT_CALLS
        if bal.get('mode') is not None:
            mode = bal['mode']
        elif c < 0:
            mode = 'peer'
        elif c < len(REDUCE_TRIALS):
            mode = trial = REDUCE_TRIALS[c]
        else:
            # both timed: the slower rank decides, identically everywhere
            t = torch.tensor([min(a.elapsed_time(b) for m, a, b in bal['trials'] if m == w and (b.synchronize() or True))
                              for w in ('peer', 'nccl')], dtype=torch.float64, device=Xd.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t = t.cpu().numpy()
            mode = bal['mode'] = 'peer' if t[0] <= t[1] else 'nccl'
            bal['trial_ms'] = {'peer': float(t[0]), 'nccl': float(t[1])}
        if trial is not None:
            t_start = torch.cuda.Event(enable_timing=True)
            t_start.record()
    if reduce and ws > 1 and ws == w0 and mode == 'peer':
        try:
            peer = PeerImage.get(Ld.shape[0], ndim)
        except Exception as e:      # noqa: BLE001  (no peer access between the GPUs of this node: NCCL does the sum)
            import warnings
            warnings.warn('synthobs_b200.dist: peer-memory image reduce unavailable (%r); using ncclAllReduce' % (e,))
            os.environ['SO_PEER_REDUCE'] = '0'
    r = images.deposit_device(Xo, Yo, Lo, resolution, ndim, adaptive_k=int(k), want_ngp=False,
                              support_sigmas=support_sigmas, dk=dko, out=None if peer is None else peer.partial,
                              push=None if peer is None else peer.push_desc())
    img = r['data']
    if ws > 1 and ws == w0 and reduce:
        bal = _share_balance.get((Xd.numel(), int(ndim), int(k), ws))
        if bal is not None:
            bal['calls'] += 1
        if bal is not None and peer is not None and bal['calls'] <= SHARE_ADAPT_CALLS:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            bal['pending'] = (e0, e1, peer.waits.clone())
        elif bal is not None:
            bal['pending'] = None
    if reduce and ws > 1:
        if peer is not None:
            img = peer.finish()
            mark('plan + binning + deposit fused with the exchange over peer memory')
        else:
            mark('plan + binning + deposit of the share')
            img = reduce_image(img)
            mark('all-reduce of the image')
    else:
        mark('plan + binning + deposit of the share')
    if trial is not None:
        t_end = torch.cuda.Event(enable_timing=True)
        t_end.record()
        # (the second call of each pair counts: the first follows a change of mode)
        if c % 2 == 1:
            _share_balance[(Xd.numel(), int(ndim), int(k), ws)].setdefault('trials', []).append((trial, t_start, t_end))
    if stages is not None:
        torch.cuda.synchronize()
        for (_, a), (name, b) in zip(marks[:-1], marks[1:]):
            stages[name] = stages.get(name, 0.0) + a.elapsed_time(b)
    return img[0] if one else img


def generate_SED_huge(model, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0, log10t_BC=7.):
    """generate_SED of ONE giant galaxy, particle-block sharded: every rank computes the five
    spectra of its contiguous block of particles, the partial spectra add (all-reduce of 5 x n_lam floats over
    NVLink).  Inputs are the full host arrays on every rank (each uploads only its block).  Returns dict of
    [n_lam] float32 CUDA tensors, identical on all ranks."""
    from .sed import models
    rank, ws = world()
    n = len(Masses)
    b, e = particle_block(n, rank, ws)
    opt = lambda t: t[b:e] if isinstance(t, (np.ndarray, torch.Tensor)) else t   # noqa: E731
    out = models.sed_spectra_device(model, Masses[b:e], Ages[b:e], Metallicities[b:e], opt(tauVs_ISM), opt(tauVs_BC),
                                    fesc, log10t_BC)
    keys = ('stellar', 'intrinsic', 'BC', 'total', 'no_HII_no_BC')
    stack = torch.stack([out[k] for k in keys])
    if ws > 1:
        dist.all_reduce(stack, op=dist.ReduceOp.SUM)
    return {k: stack[i] for i, k in enumerate(keys)}


def generate_Fnu_huge(model, F, Masses, Ages, Metallicities, tauVs_ISM=False, tauVs_BC=False, fesc=0.0, log10t_BC=7.,
                      kind='Fnu'):
    """Broadband fluxes of one giant galaxy, particle-block sharded: per-rank sums over its block (float64),
    all-reduce of [F] doubles.  Returns {filter: value} like generate_Fnu."""
    from .sed import models
    rank, ws = world()
    n = len(Masses)
    b, e = particle_block(n, rank, ws)
    opt = lambda t: t[b:e] if isinstance(t, (np.ndarray, torch.Tensor)) else t   # noqa: E731
    _, sums, _ = models._broadband(model, kind, F, F['filters'], Masses[b:e], Ages[b:e], Metallicities[b:e],
                                   opt(tauVs_ISM), opt(tauVs_BC), fesc, log10t_BC)
    s = sums[0].clone()
    if ws > 1:
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    s = s.cpu().numpy()
    return {f: s[i] for i, f in enumerate(F['filters'])}

"""Multi-GPU plumbing: one ice sheet per rank, A/E-grid contributions merged across sheets.

Reference semantics: every sheet contributes sparse (index, weight = wM, vals = M_unscaled x) entries on the A (or E)
grid (IceCoupler.cpp:445-457); the GCM side sums them and divides by the summed weights
(VectorMultivec::to_dense_scale / to_dense, multivec.cpp:35-81).  Sheets are independent until this merge
(GCMCoupler.cpp:491-497), so ranks exchange only the rows of cells that more than one sheet touches.

torch.distributed is used for the collectives only (NCCL over NVLink on GPUs, gloo in the CPU tests); pack / unpack
run in the library's own CUDA kernels (csrc/merge.cu).
"""
import numpy as np
import torch
import torch.distributed as dist


def plan_shared(local_ids, n_sparse, group=None):
    """All ranks pass the sparse ids of their dense rows (dims[0] of AvI / EvI).  Returns
    (shared_ids, loc, slot_of_row): shared_ids = ascending sparse ids present on >= 2 ranks (identical on all ranks),
    loc[k] = local dense row of shared_ids[k] or -1, slot_of_row[r] = k if local row r is shared else -1.
    Works on any backend (CPU tensors with gloo, CUDA tensors with nccl)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    local_ids = np.asarray(local_ids, dtype=np.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if (dist.is_initialized() and dist.get_backend(group) == "nccl") else torch.device("cpu")
    cnt = torch.zeros(n_sparse, dtype=torch.int32, device=dev)
    cnt[torch.from_numpy(local_ids).to(dev)] = 1
    if world > 1:
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM, group=group)
    shared_ids = torch.nonzero(cnt >= 2).flatten().cpu().numpy().astype(np.int64)
    pos = {int(s): r for r, s in enumerate(local_ids)}
    loc = np.array([pos.get(int(s), -1) for s in shared_ids], dtype=np.int32)
    slot_of_row = np.full(len(local_ids), -1, dtype=np.int32)
    for k, r in enumerate(loc):
        if r >= 0:
            slot_of_row[r] = k
    return shared_ids, loc, slot_of_row


def partition_cellsA(ex_iA, nA, world):
    """Domain decomposition of ONE sheet's build over `world` ranks by GCM cells (the reference shards global_ec by ranges of
    coarse-grid cells and concatenates, modele/global_ec.cpp:818-893): GCM cells in ascending index order are cut into
    `world` contiguous runs holding (nearly) equal numbers of exchange cells.  Returns keep[world, nA] bool masks for
    GCMRegridder.filter_cellsA; every (A, I) and (E, I) matrix row then lives on exactly one rank, so AvI / EvI of the ranks
    concatenate, while I-side matrices of ice cells that straddle a cut are summed (unscaled) across the two ranks."""
    ex_iA = np.asarray(ex_iA, dtype=np.int64)
    cnt = np.bincount(ex_iA, minlength=nA)
    used = np.nonzero(cnt)[0]
    cum = np.cumsum(cnt[used])
    total = int(cum[-1]) if len(cum) else 0
    keep = np.zeros((world, nA), bool)
    lo = 0
    for r in range(world):
        hi = int(np.searchsorted(cum, total * (r + 1) / world, side="left")) + 1 if r + 1 < world else len(used)
        hi = max(hi, lo)
        keep[r, used[lo:hi]] = True
        lo = hi
    return keep


def merge_reference(contribs, n_sparse, fill=np.nan):
    """Numpy statement of the merge for tests: contribs = [(ids, wM, num[nrow, nvar]), ...] -> dense [n_sparse, nvar]."""
    nvar = contribs[0][2].shape[1]
    num = np.zeros((n_sparse, nvar))
    w = np.zeros(n_sparse)
    touched = np.zeros(n_sparse, bool)
    for ids, wM, vals in contribs:
        np.add.at(num, ids, vals)
        np.add.at(w, ids, wM)
        touched[ids] = True
    out = np.full((n_sparse, nvar), fill)
    with np.errstate(divide="ignore", invalid="ignore"):
        out[touched] = num[touched] * (1.0 / w[touched])[:, None]
    return out


class SheetMerge:
    """Per-rank merge state for a set of (unscaled) matrices whose rows live on the A or E grid.

    begin(i, Y) packs the shared rows of result i and starts their all-reduce asynchronously (torch runs NCCL on its
    own stream, ordered after the work already queued on the current stream), so the exchange of one result overlaps
    the SpMM that produces the next; finish(Ys) waits and unpacks, dividing by the summed weights."""

    def __init__(self, spaces, nvar, device, group=None):
        """spaces: list of (n_sparse, Weighted)."""
        from ._lib import check, lib
        self._lib, self._check = lib(), check
        self.nvar, self.group = nvar, group
        self.parts = []
        self._multi = dist.is_initialized() and dist.get_world_size(group) > 1
        for n_sparse, W in spaces:
            shared_ids, loc, slot = plan_shared(W.dim(0), n_sparse, group)
            ns = len(shared_ids)
            part = dict(nrow=W.shape[0], nshared=ns, shared_ids=shared_ids,
                        loc=torch.from_numpy(loc).to(device), slot=torch.from_numpy(slot).to(device),
                        wlocal=torch.from_numpy(W.get_weights(0)).to(device),
                        buf=torch.zeros((max(ns, 1), nvar), dtype=torch.float64, device=device),
                        wtot=torch.zeros(max(ns, 1), dtype=torch.float64, device=device), work=None)
            # summed weights of the shared cells (they change only when the matrices are rebuilt)
            if ns:
                have = part["loc"] >= 0
                part["wtot"][:ns][have] = part["wlocal"][part["loc"][have].long()]
                if self._multi:
                    dist.all_reduce(part["wtot"], group=group)
            self.parts.append(part)
        self.nshared = sum(p["nshared"] for p in self.parts)
        # one contiguous exchange buffer for all parts: run() then needs a single all-reduce
        self.buf_all = torch.zeros((max(self.nshared, 1), nvar), dtype=torch.float64, device=device)
        off = 0
        for p in self.parts:
            p["buf"] = self.buf_all[off: off + max(p["nshared"], 0)] if p["nshared"] else p["buf"]
            off += p["nshared"]

    def _on(self, stream):
        """The pack / unpack kernels and the collective must be ordered on ONE stream: torch orders NCCL against its
        current stream, so everything runs under `with torch.cuda.stream(s)`.  `stream`: None (the current stream), a
        torch.cuda.Stream, or a raw handle that must BE the current stream."""
        dev = self.buf_all.device
        if stream is None:
            s = torch.cuda.current_stream(dev)
        elif isinstance(stream, torch.cuda.Stream):
            s = stream
        else:
            s = torch.cuda.current_stream(dev)
            if int(stream) != s.cuda_stream:
                raise ValueError("SheetMerge: pass a torch.cuda.Stream (or None); a raw handle must be torch's current stream, "
                                 "otherwise the all-reduce is not ordered against the pack / unpack kernels")
        return torch.cuda.stream(s), s.cuda_stream

    def begin(self, i, Y, stream=None):
        p, nvar = self.parts[i], self.nvar
        ctx, h = self._on(stream)
        with ctx:
            if p["nshared"]:
                self._check(self._lib.ib_merge_pack_dev(Y.data_ptr(), nvar, nvar, p["loc"].data_ptr(), p["nshared"], p["buf"].data_ptr(), h))
                if self._multi:
                    p["work"] = dist.all_reduce(p["buf"], group=self.group, async_op=True)     # NCCL sum over NVLink

    def finish(self, Ys, stream=None):
        """Ys[i] [nrow_i, nvar] holds this sheet's unscaled sums; on return they hold the merged, scaled values."""
        nvar = self.nvar
        ctx, h = self._on(stream)
        with ctx:
            for p, Y in zip(self.parts, Ys):
                if p["work"] is not None:
                    p["work"].wait()               # the stream waits for the collective
                    p["work"] = None
                self._check(self._lib.ib_merge_unpack_dev(Y.data_ptr(), nvar, nvar, p["nrow"], p["slot"].data_ptr(), p["buf"].data_ptr(),
                                                          p["wtot"].data_ptr(), p["wlocal"].data_ptr(), h))

    def run(self, Ys, stream=None):
        """pack every part, ONE all-reduce of the shared rows of all parts, unpack."""
        L, nvar = self._lib, self.nvar
        ctx, h = self._on(stream)
        with ctx:
            for p, Y in zip(self.parts, Ys):
                if p["nshared"]:
                    self._check(L.ib_merge_pack_dev(Y.data_ptr(), nvar, nvar, p["loc"].data_ptr(), p["nshared"], p["buf"].data_ptr(), h))
            if self.nshared and self._multi:
                dist.all_reduce(self.buf_all, group=self.group)
            for p, Y in zip(self.parts, Ys):
                self._check(L.ib_merge_unpack_dev(Y.data_ptr(), nvar, nvar, p["nrow"], p["slot"].data_ptr(), p["buf"].data_ptr(),
                                                  p["wtot"].data_ptr(), p["wlocal"].data_ptr(), h))


class PeerMerge:
    """The merge of SheetMerge without NCCL: ranks publish the rows of their shared cells in peer-mapped memory and gather
    their owners' rows over NVLink inside the library's own kernels (csrc/xchg.cu, ib_xchg_*): two kernel launches per
    step, no collective call, capturable in a CUDA graph.  Results are identical on every rank (owners are added in
    ascending rank order).

    One instance per rank (process / GPU); `torch.distributed` is used ONCE, at construction, to gather the ranks' cell ids,
    weights and IPC handles -- any host-side all-gather (MPI in the reference, GCMCoupler_ModelE.cpp:761-793) would do.
    PeerMerge.local() builds the contexts of several ranks inside one process (tests; a host driving several GPUs)."""

    def __init__(self, spaces, nvar, device, group=None, scaled=False):
        """spaces: list of (n_sparse, Weighted) -- the EvI / AvI of this rank's sheet: unscaled (the coupler's matrices; every
        row is divided by its weight in the merge) or, with scaled=True, scaled (only the shared cells are touched)."""
        self.scaled = bool(scaled)
        world = dist.get_world_size(group) if dist.is_initialized() else 1
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        mine = [(int(n), np.asarray(W.dim(0), dtype=np.int64), np.asarray(W.get_weights(0), dtype=np.float64)) for n, W in spaces]
        if world > 1:
            every = [None] * world
            dist.all_gather_object(every, mine, group=group)
        else:
            every = [mine]
        self._init(rank, world, every, nvar)
        handle = self._handle()
        if world > 1:
            handles = [None] * world
            dist.all_gather_object(handles, handle, group=group)
            self._check(self._lib.ib_xchg_open(self._h, b"".join(handles)))
        else:
            self._check(self._lib.ib_xchg_open(self._h, handle))
        self._set_parts()

    @classmethod
    def local(cls, spaces_per_rank, nvar, scaled=False):
        """Contexts of all ranks in this process: spaces_per_rank[r] = [(n_sparse, Weighted), ...] of rank r."""
        import ctypes as C
        from ._lib import vp
        every = [[(int(n), np.asarray(W.dim(0), dtype=np.int64), np.asarray(W.get_weights(0), dtype=np.float64)) for n, W in sp]
                 for sp in spaces_per_rank]
        objs = []
        for r in range(len(every)):
            o = cls.__new__(cls)
            o.scaled = bool(scaled)
            o._init(r, len(every), every, nvar)
            objs.append(o)
        arr = (vp * len(objs))(*[o._h for o in objs])
        for o in objs:
            o._check(o._lib.ib_xchg_open_local(o._h, arr))
            o._set_parts()
        return objs

    def _init(self, rank, world, every, nvar):
        import ctypes as C
        from ._lib import check, lib, vp
        self._lib, self._check = lib(), check
        self.rank, self.world, self.nvar = rank, world, nvar
        nparts = len(every[rank])
        self.parts = []
        i64p = C.POINTER(C.c_int64)
        for p in range(nparts):
            n_sparse = every[rank][p][0]
            ids = [np.ascontiguousarray(every[r][p][1]) for r in range(world)]
            ptrs = (vp * world)(*[a.ctypes.data_as(vp) for a in ids])
            cnt = (C.c_int64 * world)(*[len(a) for a in ids])
            ns = self._lib.ib_xchg_plan(world, ptrs, cnt, n_sparse, rank, None, None, None)
            check(ns)
            shared = np.zeros(ns, np.int64); loc = np.zeros(ns, np.int32); owners = np.zeros(ns, np.uint64)
            if ns:
                check(self._lib.ib_xchg_plan(world, ptrs, cnt, n_sparse, rank, shared.ctypes.data_as(vp), loc.ctypes.data_as(vp),
                                             owners.ctypes.data_as(vp)))
            # summed weights of the shared cells over their owners, added in rank order
            wtot = np.zeros(ns)
            for r in range(world):
                pos = {int(s): i for i, s in enumerate(ids[r])}
                w = every[r][p][2]
                for k, s in enumerate(shared):
                    i = pos.get(int(s))
                    if i is not None:
                        wtot[k] += w[i]
            self.parts.append(dict(n_sparse=n_sparse, nrow=len(ids[rank]), nshared=int(ns), shared_ids=shared, loc=loc, owners=owners,
                                   wtot=wtot, wlocal=np.ascontiguousarray(every[rank][p][2])))
        self.nshared = sum(p["nshared"] for p in self.parts)
        h = vp()
        ns_arr = (C.c_int64 * nparts)(*[p["nshared"] for p in self.parts])
        check(self._lib.ib_xchg_create(C.byref(h), rank, world, nparts, ns_arr, int(nvar)))
        self._h = h

    def _handle(self):
        import ctypes as C
        buf = C.create_string_buffer(self._lib.ib_xchg_handle_bytes())
        self._check(self._lib.ib_xchg_get_handle(self._h, buf))
        return buf.raw

    def _set_parts(self):
        from ._lib import vp
        for i, p in enumerate(self.parts):
            self._check(self._lib.ib_xchg_set_part(self._h, i, p["nrow"], p["loc"].ctypes.data_as(vp), p["owners"].ctypes.data_as(vp),
                                                   p["wtot"].ctypes.data_as(vp), p["wlocal"].ctypes.data_as(vp), int(self.scaled)))

    def _args(self, Ys):
        import ctypes as C
        from ._lib import vp
        n = len(self.parts)
        return (vp * n)(*[y.data_ptr() for y in Ys]), (C.c_int64 * n)(*[y.stride(0) for y in Ys])

    def args(self, Ys):
        """Pre-built C argument arrays for merge_raw (the per-step Python cost matters at a 60 us step)."""
        return self._args(Ys)

    def merge(self, Ys, stream=0):
        """Ys[p]: [nrow_p, nvar] unscaled sums in, merged and scaled values out (device tensors, in place)."""
        yp, ld = self._args(Ys)
        self._check(self._lib.ib_xchg_merge_dev(self._h, yp, ld, Ys[0].shape[1], stream or None))

    def merge_raw(self, yp, ld, nvar, stream):
        self._check(self._lib.ib_xchg_merge_dev(self._h, yp, ld, nvar, stream))

    def publish_raw(self, yp, ld, nvar, stream):
        self._check(self._lib.ib_xchg_publish_dev(self._h, yp, ld, nvar, stream))

    def gather_raw(self, yp, ld, nvar, stream):
        self._check(self._lib.ib_xchg_gather_dev(self._h, yp, ld, nvar, stream))

    def publish(self, Ys, stream=0):
        yp, ld = self._args(Ys)
        self._check(self._lib.ib_xchg_publish_dev(self._h, yp, ld, Ys[0].shape[1], stream or None))

    def gather(self, Ys, stream=0):
        yp, ld = self._args(Ys)
        self._check(self._lib.ib_xchg_gather_dev(self._h, yp, ld, Ys[0].shape[1], stream or None))

    def status(self):
        self._check(self._lib.ib_xchg_status(self._h))

    def steps_done(self):
        return int(self._lib.ib_xchg_step(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.ib_xchg_free(self._h)
                self._h = None
        except Exception:
            pass

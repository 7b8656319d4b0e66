"""z-slab sharding of the transfer path across the GPUs of one node.

One process per GPU.  Rank r owns the cell layers [k0, k1) = gfs_slab_range(K, world, r): its particles are
those whose cell lies in them.  Grids are allocated whole on every rank (they are small next to the particles:
0.4 GB of fields at 256^3 against 180 GB of HBM) and indexed globally, so a layer keeps the same index on both
sides of a cut; each rank's grid kernels only touch its own layers plus one halo layer.  Per substep:

  sort -> P2G splat
       -> C1: add the neighbour's partial accumulators on the two node layers either side of each cut
              (64-bit integers: the merged sums are bit-identical to the single-GPU ones), and copy the
              neighbour's freshly classified boundary material layer
       -> P2G normalise + assemble on the owned layers
       -> C2: copy H halo layers of the NEW and SAVED fields from each neighbour (what a sharded pressure solve
              would have to hand over; H = stencil radius + displacement in cells)
       -> PIC/FLIP + RK4
       -> C3: migrate the particles whose new cell layer left the slab (count-prefixed send/recv)

All exchanges are neighbour-only.  Transport is either torch.distributed point-to-point batches (NCCL over NVLink
on GPUs; gloo in the CPU tests, which drive this same code with a numpy backend) or, for single-process tests, an
in-process loopback between several drivers.  `backend` is duck-typed:

  owned = (k0, k1);  K;  sort(); p2g_begin(); p2g_end(); g2p_advect(dt)
  layer_bytes(what) -> int
  pack(what, k_first, k_count) -> uint8 tensor;   unpack(what, k_first, k_count, uint8 tensor, add)
  extract(k_lo, k_hi) -> (down, up) float32 tensors [n,6];   append(float32 tensor [n,6]);   num_particles
"""
import torch
import torch.distributed as dist

ACC = (10, 11, 12)
MATERIAL = 9
NEW = (0, 1, 2)
SAVED = (3, 4, 5)
INT_MIN, INT_MAX = -2 ** 31, 2 ** 31 - 1
SIDES = ("down", "up")


def slab_ranges(K, world):
    """[(k0,k1)] for every rank -- same arithmetic as gfs_slab_range."""
    return [(K * r // world, K * (r + 1) // world) for r in range(world)]


class SlabDriver:
    """The per-rank half of every exchange: what to send to each neighbour, what to do with what arrives."""

    def __init__(self, backend, rank, world, halo=2):
        self.b, self.rank, self.world, self.halo = backend, rank, world, int(halo)
        self.k0, self.k1 = backend.owned
        self.K = backend.K
        self.peer = {"down": rank - 1 if rank > 0 else None, "up": rank + 1 if rank + 1 < world else None}
        assert world == 1 or self.k1 - self.k0 >= max(self.halo, 2), "slab thinner than the halo"

    def sides(self):
        return [s for s in SIDES if self.peer[s] is not None]

    def _cut_layers(self, side):
        """node layers either side of the cut shared with the neighbour on `side`: (k-1, k) for the cut at cell k."""
        return (self.k0 if side == "down" else self.k1) - 1, 2

    # ---- C1: P2G partial sums + boundary material ----------------------------------------------------------
    def partials_send(self):
        out = {}
        for side in self.sides():
            first, count = self._cut_layers(side)
            parts = [self.b.pack(what, first, count) for what in ACC]
            parts.append(self.b.pack(MATERIAL, self.k0 if side == "down" else self.k1 - 1, 1))   # my boundary layer
            out[side] = torch.cat(parts)
        return out

    def partials_recv_sizes(self):
        return {side: sum(self.b.layer_bytes(w) * 2 for w in ACC) + self.b.layer_bytes(MATERIAL) for side in self.sides()}

    def partials_recv(self, recv):
        for side, buf in recv.items():
            first, count = self._cut_layers(side)
            sizes = [self.b.layer_bytes(w) * count for w in ACC] + [self.b.layer_bytes(MATERIAL)]
            chunks = torch.split(buf, sizes)
            for what, chunk in zip(ACC, chunks[:3]):
                self.b.unpack(what, first, count, chunk, add=True)
            # the neighbour's boundary layer is my halo layer
            self.b.unpack(MATERIAL, self.k0 - 1 if side == "down" else self.k1, 1, chunks[3], add=False)

    # ---- C2: field halos -------------------------------------------------------------------------------------
    def _halo_ranges(self, side):
        H = self.halo
        if side == "down":
            return (self.k0, min(H, self.k1 - self.k0)), (max(self.k0 - H, 0), self.k0 - max(self.k0 - H, 0))
        return (max(self.k1 - H, self.k0), self.k1 - max(self.k1 - H, self.k0)), (self.k1, min(self.k1 + H, self.K) - self.k1)

    def halos_send(self, whats=NEW + SAVED):
        out = {}
        for side in self.sides():
            (first, count), _ = self._halo_ranges(side)
            out[side] = torch.cat([self.b.pack(what, first, count) for what in whats])
        return out

    def halos_recv_sizes(self, whats=NEW + SAVED):
        return {side: sum(self.b.layer_bytes(w) * self._halo_ranges(side)[1][1] for w in whats) for side in self.sides()}

    def halos_recv(self, recv, whats=NEW + SAVED):
        for side, buf in recv.items():
            _, (first, count) = self._halo_ranges(side)
            for what, chunk in zip(whats, torch.split(buf, [self.b.layer_bytes(w) * count for w in whats])):
                self.b.unpack(what, first, count, chunk, add=False)

    # ---- C3: particle migration --------------------------------------------------------------------------------
    def migrate_send(self):
        k_lo = self.k0 if self.peer["down"] is not None else INT_MIN
        k_hi = self.k1 if self.peer["up"] is not None else INT_MAX
        down, up = self.b.extract(k_lo, k_hi)
        return {s: t for s, t in (("down", down), ("up", up)) if self.peer[s] is not None}

    def migrate_recv(self, recv):
        n = 0
        for side, t in recv.items():
            if t.shape[0] > 0:
                self.b.append(t)
                n += t.shape[0]
        return n


class DistTransport:
    """Neighbour exchange over torch.distributed (NCCL or gloo): one batch of isend/irecv per exchange."""

    def __init__(self, group=None):
        self.group = group
        self.bytes_sent = 0

    def exchange(self, drv, send, recv):
        ops = []
        for side in drv.sides():
            if send[side].numel() > 0:
                ops.append(dist.P2POp(dist.isend, send[side], drv.peer[side], self.group))
                self.bytes_sent += send[side].numel() * send[side].element_size()
            if recv[side].numel() > 0:
                ops.append(dist.P2POp(dist.irecv, recv[side], drv.peer[side], self.group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()
        return recv

    def fixed(self, drv, send, sizes):
        some = next(iter(send.values())) if send else None
        recv = {s: torch.empty(sizes[s], dtype=torch.uint8, device=some.device) for s in send}
        return self.exchange(drv, send, recv)

    def variable(self, drv, send):
        """count-prefixed exchange of [n,6] float32 particle blocks"""
        if not send:
            return {}
        dev = next(iter(send.values())).device
        cnt_s = {s: torch.tensor([send[s].shape[0]], dtype=torch.int64, device=dev) for s in send}
        cnt_r = self.exchange(drv, cnt_s, {s: torch.zeros(1, dtype=torch.int64, device=dev) for s in send})
        recv = {s: torch.empty((int(cnt_r[s].item()), 6), dtype=torch.float32, device=dev) for s in send}
        flat = self.exchange(drv, {s: send[s].reshape(-1) for s in send}, {s: recv[s].reshape(-1) for s in recv})
        return {s: flat[s].reshape(-1, 6) for s in flat}


def substep(drv, transport, dt, exchange_fields=True):
    """One sharded substep of one rank.  Returns (particles sent away, particles received)."""
    b = drv.b
    b.sort()
    b.p2g_begin()
    drv.partials_recv(transport.fixed(drv, drv.partials_send(), drv.partials_recv_sizes()))
    b.p2g_end()
    if exchange_fields:
        drv.halos_recv(transport.fixed(drv, drv.halos_send(), drv.halos_recv_sizes()))
    b.g2p_advect(dt)
    out = drv.migrate_send()
    return sum(t.shape[0] for t in out.values()), drv.migrate_recv(transport.variable(drv, out))


class LoopbackWorld:
    """Several slabs stepped in lockstep inside ONE process (all on one device): what rank r sends 'up' is what rank
    r+1 receives from 'down'.  Lets the sharded code path be checked bit for bit against the unsharded one on a
    single GPU."""

    def __init__(self, drivers):
        self.drv = list(drivers)

    def _swap(self, sends):
        recv = [dict() for _ in self.drv]
        for r, s in enumerate(sends):
            if "up" in s:
                recv[r + 1]["down"] = s["up"].clone()
            if "down" in s:
                recv[r - 1]["up"] = s["down"].clone()
        return recv

    def substep(self, dt, exchange_fields=True):
        for d in self.drv:
            d.b.sort()
            d.b.p2g_begin()
        for d, rcv in zip(self.drv, self._swap([d.partials_send() for d in self.drv])):
            d.partials_recv(rcv)
        for d in self.drv:
            d.b.p2g_end()
        if exchange_fields:
            for d, rcv in zip(self.drv, self._swap([d.halos_send() for d in self.drv])):
                d.halos_recv(rcv)
        for d in self.drv:
            d.b.g2p_advect(dt)
        moved = 0
        for d, rcv in zip(self.drv, self._swap([d.migrate_send() for d in self.drv])):
            moved += d.migrate_recv(rcv)
        return moved


class CudaSlabBackend:
    """The C-ABI context as a slab backend.  Comm buffers are torch tensors; the library packs / unpacks."""

    def __init__(self, ctx, dims, owned, interp, arith=0, order=4, migrate_cap=None):
        self.ctx, self.K, self.owned = ctx, dims[2], tuple(owned)
        self.interp, self.arith, self.order = interp, arith, order
        self.dev = torch.device("cuda", torch.cuda.current_device())
        ctx.set_owned_layers(*owned)
        self.cap = migrate_cap
        self._bufs = None

    @property
    def num_particles(self):
        return self.ctx.num_particles

    def sort(self):
        self.ctx.sort_unstable()

    def p2g_begin(self):
        self.ctx.p2g_begin(self.arith)

    def p2g_end(self):
        self.ctx.p2g_end()

    def g2p_advect(self, dt):
        self.ctx.g2p_advect(dt, order=self.order, interp=self.interp, arith=self.arith)

    def layer_bytes(self, what):
        return self.ctx.layer_bytes(what)

    def pack(self, what, k_first, k_count):
        t = torch.empty(self.layer_bytes(what) * k_count, dtype=torch.uint8, device=self.dev)
        if k_count > 0:
            self.ctx.pack_layers(what, k_first, k_count, t.data_ptr())
        return t

    def unpack(self, what, k_first, k_count, t, add):
        if k_count > 0:
            t = t.contiguous()
            self.ctx.unpack_layers(what, k_first, k_count, t.data_ptr(), add)
            t.record_stream(torch.cuda.current_stream())

    def extract(self, k_lo, k_hi):
        n = self.ctx.num_particles
        cap = self.cap if self.cap is not None else max(1024, n // 4)
        if self._bufs is None or self._bufs[0].shape[0] < cap:
            self._bufs = (torch.empty((cap, 6), dtype=torch.float32, device=self.dev),
                          torch.empty((cap, 6), dtype=torch.float32, device=self.dev))
        down, up = self._bufs
        nd, nu = self.ctx.extract_particles(k_lo, k_hi, down.data_ptr(), up.data_ptr(), cap)
        return down[:nd], up[:nu]

    def append(self, t):
        t = t.contiguous()
        self.ctx.append_particles_device(t.data_ptr(), t.shape[0])
        t.record_stream(torch.cuda.current_stream())

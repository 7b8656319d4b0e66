"""z-slab sharding of the transfer path across the GPUs of one node.

One process per GPU.  Rank r owns the cell layers [k0, k1) = gfs_slab_range(K, world, r): its particles are
those whose cell lies in them.  Grids are allocated whole on every rank (they are small next to the particles:
0.4 GB of fields at 256^3 against 180 GB of HBM) and indexed globally, so a layer keeps the same index on both
sides of a cut; each rank's grid kernels only touch its own layers plus one halo layer.  Per substep:

  sort -> P2G splat
       -> C1: add the neighbour's partial accumulators on the two node layers either side of each cut
              (64-bit integers: the merged sums are bit-identical to the single-GPU ones), and copy the
              neighbour's freshly classified boundary material layer
          C2: copy H halo layers of the NEW and SAVED fields from each neighbour (what a sharded pressure solve
              would have to hand over; H = stencil radius + displacement in cells) -- same message as C1 when
              nothing sits between P2G and G2P (the synthetic substep), its own message otherwise
       -> P2G normalise + assemble on the owned layers
       -> PIC/FLIP + RK4
       -> C3: migrate the particles whose new cell layer left the slab: counts travel device-to-device, are read
              back with ONE host synchronisation per substep, then the payloads

All exchanges are neighbour-only, one batch of isend/irecv each, out of persistent buffers the backend packs into.
Transport is torch.distributed (NCCL over NVLink on GPUs; gloo in the CPU tests, which drive this same code with a
numpy backend) or, for single-process tests, an in-process loopback between several drivers.  `backend` is duck-typed:

  owned = (k0, k1);  K;  device;  sort(); p2g_begin(); p2g_end(); g2p_advect(dt);  num_particles
  layer_bytes(what) -> int
  pack_batch([(what, k_first, k_count, offset, _)], uint8 tensor);   unpack_batch([(what, k_first, k_count, offset, add)], uint8 tensor)
  extract_async(k_lo, k_hi) -> (down [cap,6] f32, up [cap,6] f32, counts int32[4] = kept, n_down, n_up, spare)
  extract_commit(n_kept);  append(float32 tensor [n,6])
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


def slab_ranges_weighted(layer_counts, world, min_layers=4):
    """[(k0,k1)] for every rank with cuts that balance the PARTICLES, not the layers: layer_counts[k] = particles in
    cell layer k (efficiency is governed by particles per slab).  Cut r is placed after the layer where the
    running count reaches r/world of the total; every slab keeps at least min_layers layers (the halo)."""
    counts = [float(c) for c in layer_counts]
    K, total = len(counts), sum(counts)
    if world == 1 or total <= 0 or K < world * min_layers:
        return slab_ranges(K, world)
    cuts, run, k = [0], 0.0, 0
    for r in range(1, world):
        target = total * r / world
        while k < K and run + counts[k] <= target:
            run += counts[k]
            k += 1
        # the layer straddling the target goes to whichever side leaves the smaller error
        if k < K and (run + counts[k] - target) < (target - run):
            run += counts[k]
            k += 1
        k = max(k, cuts[-1] + min_layers)
        k = min(k, K - (world - r) * min_layers)
        run = sum(counts[:k])
        cuts.append(k)
    cuts.append(K)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


class SlabDriver:
    """The per-rank half of every exchange: what to send to each neighbour, what to do with what arrives."""

    def __init__(self, backend, rank, world, halo=2):
        self.b, self.rank, self.world, self.halo = backend, rank, world, int(halo)
        self.k0, self.k1 = backend.owned
        self.K = backend.K
        self.peer = {"down": rank - 1 if rank > 0 else None, "up": rank + 1 if rank + 1 < world else None}
        assert world == 1 or self.k1 - self.k0 >= max(self.halo, 2), "slab thinner than the halo"
        self._plans, self._bufs = {}, {}

    def sides(self):
        return [s for s in SIDES if self.peer[s] is not None]

    # ---- what travels: lists of (what, send_first, send_count, recv_first, recv_count, add) per side --------------
    def _items(self, side, phase):
        if phase == "partials":                  # C1
            cut = self.k0 if side == "down" else self.k1
            items = [(w, cut - 1, 2, cut - 1, 2, True) for w in ACC]       # node layers (k-1, k) either side of the cut
            mine, theirs = (self.k0, self.k0 - 1) if side == "down" else (self.k1 - 1, self.k1)
            items.append((MATERIAL, mine, 1, theirs, 1, False))            # my boundary layer -> its halo layer
            return items
        if phase == "halos":                     # C2
            H = self.halo
            if side == "down":
                s, r = (self.k0, min(H, self.k1 - self.k0)), (max(self.k0 - H, 0), self.k0 - max(self.k0 - H, 0))
            else:
                s, r = (max(self.k1 - H, self.k0), self.k1 - max(self.k1 - H, self.k0)), (self.k1, min(self.k1 + H, self.K) - self.k1)
            return [(w, s[0], s[1], r[0], r[1], False) for w in NEW + SAVED]
        raise ValueError(phase)

    def plan(self, phases):
        """{side: (items, send offsets, send bytes, recv offsets, recv bytes)}, cached"""
        key = tuple(phases)
        if key not in self._plans:
            out = {}
            for side in self.sides():
                items = [it for ph in phases for it in self._items(side, ph)]
                so, ro, sb, rb = [], [], 0, 0
                for (w, sf, sc, rf, rc, add) in items:
                    so.append(sb); ro.append(rb)
                    sb += self.b.layer_bytes(w) * sc
                    rb += self.b.layer_bytes(w) * rc
                out[side] = (items, so, sb, ro, rb)
            self._plans[key] = out
        return self._plans[key]

    def buffers(self, phases):
        key = tuple(phases)
        if key not in self._bufs:
            dev = self.b.device
            self._bufs[key] = ({s: torch.empty(p[2], dtype=torch.uint8, device=dev) for s, p in self.plan(phases).items()},
                               {s: torch.empty(p[4], dtype=torch.uint8, device=dev) for s, p in self.plan(phases).items()})
        return self._bufs[key]

    def pack(self, phases):
        send, _ = self.buffers(phases)
        for side, (items, so, sb, ro, rb) in self.plan(phases).items():
            self.b.pack_batch([(w, sf, sc, off, False) for (w, sf, sc, rf, rc, add), off in zip(items, so)], send[side])
        return send

    def unpack(self, phases, recv):
        for side, (items, so, sb, ro, rb) in self.plan(phases).items():
            self.b.unpack_batch([(w, rf, rc, off, add) for (w, sf, sc, rf, rc, add), off in zip(items, ro)], recv[side])

    # ---- C3: particle migration --------------------------------------------------------------------------------
    def migrate_begin(self):
        k_lo = self.k0 if self.peer["down"] is not None else INT_MIN
        k_hi = self.k1 if self.peer["up"] is not None else INT_MAX
        return self.b.extract_async(k_lo, k_hi)

    def migrate_end(self, recv):
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

    def agree(self, drv):
        """max |v| over all ranks -> every rank's fixed-point scale (see gfs_device_ptr(16)); the one collective."""
        word = drv.b.scale_word() if hasattr(drv.b, "scale_word") else None
        if word is not None and drv.world > 1:
            dist.all_reduce(word, op=dist.ReduceOp.MAX, group=self.group)

    def layers(self, drv, phases):
        send = drv.pack(phases)
        _, recv = drv.buffers(phases)
        drv.unpack(phases, self.exchange(drv, send, recv))

    def migrate(self, drv):
        down, up, counts = drv.migrate_begin()
        sides = drv.sides()
        if not sides:
            kept = int(counts.cpu()[0])
            drv.b.extract_commit(kept)
            return 0, 0
        slot = {"down": 1, "up": 2}
        theirs = torch.zeros(4, dtype=counts.dtype, device=counts.device)
        self.exchange(drv, {s: counts[slot[s]:slot[s] + 1] for s in sides}, {s: theirs[slot[s]:slot[s] + 1] for s in sides})
        host = torch.cat([counts, theirs]).cpu().tolist()          # the one host synchronisation of the substep
        kept, n_out, n_in = host[0], {"down": host[1], "up": host[2]}, {"down": host[5], "up": host[6]}
        out = {"down": down, "up": up}
        for s_ in SIDES:
            if n_out[s_] > out[s_].shape[0]:
                raise RuntimeError("migration buffer too small: %d leavers %s, capacity %d" % (n_out[s_], s_, out[s_].shape[0]))
        drv.b.extract_commit(kept)
        send = {s: out[s][: n_out[s]].reshape(-1) for s in sides}
        recv = {s: torch.empty((n_in[s], 6), dtype=torch.float32, device=counts.device) for s in sides}
        self.exchange(drv, send, {s: recv[s].reshape(-1) for s in sides})
        return sum(n_out[s] for s in sides), drv.migrate_end(recv)


SIDE_ID = {"down": 0, "up": 1}


class PeerTransport:
    """Neighbour exchange through peer memory (CUDA IPC over NVLink/NVSwitch), no NCCL in the data path: every rank
    packs its layers / leavers straight into the neighbour's comm block with its own kernels and raises a flag
    there; waits are device-side spins (gfs_comm_*).  torch.distributed only ships the 64-byte IPC handles once."""

    @staticmethod
    def layer_bytes(drv, phases=(("partials", "halos"), ("partials",), ("halos",))):
        """Bytes of the largest layer message this driver sends or receives on one side."""
        return max([p[2] for ph in phases for p in drv.plan(ph).values()] +
                   [p[4] for ph in phases for p in drv.plan(ph).values()] + [256])

    def __init__(self, drv, group=None, particle_cap=None, connect=True, layer_bytes=None):
        self.ctx = drv.b.ctx
        self.bytes_sent = 0
        self._plan_key, self._plan_bytes = None, {}
        layer_bytes = max(int(layer_bytes or 0), self.layer_bytes(drv))
        # both ends of a link must agree on the sizes: take the maxima over all ranks
        cap = particle_cap if particle_cap is not None else max(4096, drv.b.num_particles // 6)
        if connect and drv.world > 1:
            sizes = [None] * drv.world
            dist.all_gather_object(sizes, (int(layer_bytes), int(cap)), group=group)
            layer_bytes, cap = max(s[0] for s in sizes), max(s[1] for s in sizes)
        self.ctx.comm_alloc(layer_bytes, cap)
        if connect and drv.world > 1:
            mine = (self.ctx.comm_export(0), self.ctx.comm_export(1))
            everyone = [None] * drv.world
            dist.all_gather_object(everyone, mine, group=group)
            for side in drv.sides():          # the neighbour on my `side` exposes its block for the opposite side
                self.ctx.comm_connect(SIDE_ID[side], everyone[drv.peer[side]][1 - SIDE_ID[side]])
            self.ctx.comm_world_alloc(drv.rank, drv.world)
            tables = [None] * drv.world
            dist.all_gather_object(tables, self.ctx.comm_world_export(), group=group)
            for r, h in enumerate(tables):
                if r != drv.rank:
                    self.ctx.comm_world_connect(r, h)
            dist.barrier(group=group)

    def agree(self, drv):
        if drv.world > 1:
            self.ctx.comm_allmax_scale()

    def substep(self, drv, dt):
        """The whole substep as ONE native call (gfs_comm_substep): same sequence as slabs.substep()."""
        key = (id(drv), drv.b.interp)
        if self._plan_key != key:
            for side, (items, so, sb, ro, rb) in drv.plan(("partials", "halos")).items():
                self.ctx.comm_set_plan(SIDE_ID[side],
                                       [(w, sf, sc, off, False) for (w, sf, sc, rf, rc, add), off in zip(items, so)],
                                       [(w, rf, rc, off, add) for (w, sf, sc, rf, rc, add), off in zip(items, ro)])
                self._plan_bytes[side] = sb
            self._plan_key = key
        b = drv.b
        sent, got = self.ctx.comm_substep(dt, drv.peer["down"] is not None, drv.peer["up"] is not None,
                                          order=b.order, interp=b.interp, arith=b.arith)
        self.bytes_sent += sent * 24 + sum(self._plan_bytes[s] for s in drv.sides())
        return sent, got

    def layers(self, drv, phases):
        plan = drv.plan(phases)
        for side, (items, so, sb, ro, rb) in plan.items():
            self.ctx.comm_push_layers(SIDE_ID[side], [(w, sf, sc, off, False) for (w, sf, sc, rf, rc, add), off in zip(items, so)])
            self.bytes_sent += sb
        for side, (items, so, sb, ro, rb) in plan.items():
            self.ctx.comm_pull_layers(SIDE_ID[side], [(w, rf, rc, off, add) for (w, sf, sc, rf, rc, add), off in zip(items, ro)])

    def migrate(self, drv):
        self.ctx.comm_migrate_begin(drv.peer["down"] is not None, drv.peer["up"] is not None)
        sent, got = self.ctx.comm_migrate_finish()
        self.bytes_sent += sent * 24
        return sent, got

    def advect_and_migrate(self, drv, dt):
        """G2P + RK with the migration fused into the kernel (gfs_comm_g2p_advect), then the one host sync."""
        b = drv.b
        self.ctx.comm_g2p_advect(dt, drv.peer["down"] is not None, drv.peer["up"] is not None,
                                 order=b.order, interp=b.interp, arith=b.arith)
        sent, got = self.ctx.comm_migrate_finish()
        self.bytes_sent += sent * 24
        return sent, got


class PeerLoopbackWorld:
    """Several slabs of ONE process exchanging through the peer-memory path (comm blocks connected in-process): every
    context enqueues its pushes before anyone's pull is enqueued, so the device-side waits always find their data."""

    def __init__(self, drivers, particle_cap=None):
        self.drv = list(drivers)
        cap = particle_cap or max(4096, max(d.b.num_particles for d in self.drv))
        self.tr = [PeerTransport(d, particle_cap=cap, connect=False) for d in self.drv]
        for r, d in enumerate(self.drv):
            for side in d.sides():
                d.b.ctx.comm_connect_local(SIDE_ID[side], self.drv[d.peer[side]].b.ctx)
            d.b.ctx.comm_world_alloc(r, len(self.drv))
        for r, d in enumerate(self.drv):
            for q, e in enumerate(self.drv):
                if q != r:
                    d.b.ctx.comm_world_connect_local(q, e.b.ctx)

    def _layers(self, phases):
        plans = [d.plan(phases) for d in self.drv]
        for d, plan in zip(self.drv, plans):
            for side, (items, so, sb, ro, rb) in plan.items():
                d.b.ctx.comm_push_layers(SIDE_ID[side], [(w, sf, sc, off, False) for (w, sf, sc, rf, rc, add), off in zip(items, so)])
        for d, plan in zip(self.drv, plans):
            for side, (items, so, sb, ro, rb) in plan.items():
                d.b.ctx.comm_pull_layers(SIDE_ID[side], [(w, rf, rc, off, add) for (w, sf, sc, rf, rc, add), off in zip(items, ro)])

    def substep(self, dt, pressure_solve_between=False, fused=True):
        for d in self.drv:
            d.b.sort()
        for d in self.drv:
            d.b.ctx.comm_allmax_scale()
        for d in self.drv:
            d.b.p2g_begin()
        if pressure_solve_between:
            self._layers(("partials",))
            for d in self.drv:
                d.b.p2g_end()
            self._layers(("halos",))
        else:
            self._layers(("partials", "halos"))
            for d in self.drv:
                d.b.p2g_end()
        if fused:
            for d in self.drv:
                d.b.ctx.comm_g2p_advect(dt, d.peer["down"] is not None, d.peer["up"] is not None,
                                        order=d.b.order, interp=d.b.interp, arith=d.b.arith)
        else:
            for d in self.drv:
                d.b.g2p_advect(dt)
            for d in self.drv:
                d.b.ctx.comm_migrate_begin(d.peer["down"] is not None, d.peer["up"] is not None)
        return sum(d.b.ctx.comm_migrate_finish()[1] for d in self.drv)


def substep(drv, transport, dt, pressure_solve_between=False):
    """One sharded substep of one rank.  Returns (particles sent away, particles received)."""
    if hasattr(transport, "substep") and not pressure_solve_between:
        return transport.substep(drv, dt)          # the same sequence, natively, in one call
    b = drv.b
    b.sort()
    transport.agree(drv)                           # the fixed-point scale of the partial sums: max |v| over all ranks
    b.p2g_begin()
    if pressure_solve_between:
        transport.layers(drv, ("partials",))
        b.p2g_end()
        transport.layers(drv, ("halos",))          # the fields a solver produced after P2G
    else:
        transport.layers(drv, ("partials", "halos"))
        b.p2g_end()
    if hasattr(transport, "advect_and_migrate"):
        return transport.advect_and_migrate(drv, dt)
    b.g2p_advect(dt)
    return transport.migrate(drv)


class LoopbackWorld:
    """Several slabs stepped in lockstep inside ONE process (all on one device): what rank r sends 'up' is what rank
    r+1 receives from 'down'.  Lets the sharded code path be checked bit for bit against the unsharded one on a
    single GPU."""

    def __init__(self, drivers):
        self.drv = list(drivers)

    @staticmethod
    def _swap(sends):
        recv = [dict() for _ in sends]
        for r, s in enumerate(sends):
            if "up" in s:
                recv[r + 1]["down"] = s["up"].clone()
            if "down" in s:
                recv[r - 1]["up"] = s["down"].clone()
        return recv

    def _layers(self, phases):
        for d, rcv in zip(self.drv, self._swap([d.pack(phases) for d in self.drv])):
            d.unpack(phases, rcv)

    def substep(self, dt, pressure_solve_between=False):
        for d in self.drv:
            d.b.sort()
        words = [d.b.scale_word() for d in self.drv if hasattr(d.b, "scale_word")]
        if words:                                   # max |v| over all slabs -> every slab's fixed-point scale
            top = torch.stack(words).max()
            for w in words:
                w.fill_(top)
            torch.cuda.current_stream().synchronize()
        for d in self.drv:
            d.b.p2g_begin()
        if pressure_solve_between:
            self._layers(("partials",))
            for d in self.drv:
                d.b.p2g_end()
            self._layers(("halos",))
        else:
            self._layers(("partials", "halos"))
            for d in self.drv:
                d.b.p2g_end()
        for d in self.drv:
            d.b.g2p_advect(dt)
        sends = []
        for d in self.drv:
            down, up, counts = d.migrate_begin()
            c = counts.cpu().tolist()
            d.b.extract_commit(c[0])
            sends.append({s: t[:n] for s, t, n in (("down", down, c[1]), ("up", up, c[2])) if d.peer[s] is not None})
        moved = 0
        for d, rcv in zip(self.drv, self._swap(sends)):
            moved += d.migrate_end(rcv)
        return moved


class CudaSlabBackend:
    """The C-ABI context as a slab backend.  Comm buffers are torch tensors; the library packs / unpacks."""

    def __init__(self, ctx, dims, owned, interp, arith=0, order=4, migrate_cap=None, shared_stream=False):
        """shared_stream: the context was created on torch's current CUDA stream (gfs_create(device, stream)), so the
        library's kernels and torch's / NCCL's work are already ordered.  Otherwise every hand-over between the two
        is fenced with a stream synchronisation (correct, slower: what the single-process tests use)."""
        self.shared_stream = bool(shared_stream)
        self.ctx, self.K, self.owned = ctx, dims[2], tuple(owned)
        self.interp, self.arith, self.order = interp, arith, order
        self.device = torch.device("cuda", torch.cuda.current_device())
        ctx.set_owned_layers(*owned)
        self.cap = migrate_cap
        self._bufs = None
        self._counts = torch.zeros(4, dtype=torch.int32, device=self.device)

    @property
    def num_particles(self):
        return self.ctx.num_particles

    def sort(self):
        if self.arith == 0:
            self.ctx.sort_index()          # fast arithmetic is order independent: no physical reorder
        else:
            self.ctx.sort_unstable()

    def scale_word(self):
        """The device word holding max |v| (float bits) as an int32 tensor (bit patterns of floats >= 0 order like ints)."""
        class _Word:
            pass
        w = _Word()
        w.__cuda_array_interface__ = {"shape": (1,), "typestr": "<i4", "data": (int(self.ctx.device_ptr(16)), False), "version": 2}
        self._torch_done()
        self._lib_done()
        return torch.as_tensor(w, device=self.device)

    def p2g_begin(self):
        self.ctx.p2g_begin(self.arith)

    def p2g_end(self):
        self.ctx.p2g_end()

    def g2p_advect(self, dt):
        self.ctx.g2p_advect(dt, order=self.order, interp=self.interp, arith=self.arith)

    def layer_bytes(self, what):
        return self.ctx.layer_bytes(what)

    def _lib_done(self):          # library work must be visible to torch
        if not self.shared_stream:
            self.ctx.sync()

    def _torch_done(self):        # torch work must be visible to the library
        if not self.shared_stream:
            torch.cuda.current_stream().synchronize()

    def pack_batch(self, items, buf):
        self._torch_done()
        self.ctx.copy_layers_batch(0, items, buf.data_ptr())
        self._lib_done()

    def unpack_batch(self, items, buf):
        self._torch_done()
        self.ctx.copy_layers_batch(1, items, buf.data_ptr())
        self._lib_done()

    def extract_async(self, k_lo, k_hi):
        n = self.ctx.num_particles
        cap = self.cap if self.cap is not None else max(1024, n // 4)
        if self._bufs is None or self._bufs[0].shape[0] < cap:
            self._bufs = (torch.empty((cap, 6), dtype=torch.float32, device=self.device),
                          torch.empty((cap, 6), dtype=torch.float32, device=self.device))
        down, up = self._bufs
        self._torch_done()
        self.ctx.extract_particles_async(k_lo, k_hi, down.data_ptr(), up.data_ptr(), down.shape[0], self._counts.data_ptr())
        self._lib_done()
        return down, up, self._counts

    def extract_commit(self, n_kept):
        self.ctx.extract_commit(n_kept)

    def append(self, t):
        t = t.contiguous()
        self._torch_done()
        self.ctx.append_particles_device(t.data_ptr(), t.shape[0])
        self._lib_done()
        t.record_stream(torch.cuda.current_stream())

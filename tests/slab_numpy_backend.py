"""A numpy/oracle stand-in for the CUDA slab backend, used by the CPU (gloo) tests of gridfluidsim3d_b200.slabs:
same interface, same global indexing, the compute done by the oracle.  TEST INFRASTRUCTURE."""
import numpy as np
import torch

from gridfluidsim3d_b200 import synth


class NumpySlabBackend:
    def __init__(self, oracle, scene, owned, mode=0):
        self.o, self.dims, self.dx, self.owned = oracle, scene["dims"], scene["dx"], tuple(owned)
        self.K = self.dims[2]
        self.device = torch.device("cpu")
        self.mode = mode
        self.material = scene["material"].copy()
        self.new = [a.copy() for a in scene["new"]]
        self.saved = [a.copy() for a in scene["saved"]]
        self.fdims = synth.face_dims(self.dims)
        self.p2g = [np.zeros(a * b * c, np.float32) for a, b, c in self.fdims]
        self.acc = [[np.zeros(a * b * c, np.float32), np.zeros(a * b * c, np.float32)] for a, b, c in self.fdims]
        ijk = oracle.cell_index(scene["pos"], self.dx)
        mine = (ijk[:, 2] >= owned[0]) & (ijk[:, 2] < owned[1])
        self.pos, self.vel = scene["pos"][mine].copy(), scene["vel"][mine].copy()

    @property
    def num_particles(self):
        return len(self.pos)

    def _work(self):
        return max(self.owned[0] - 1, 0), min(self.owned[1] + 1, self.K)

    def sort(self):
        pass

    def p2g_begin(self):
        I, J, K = self.dims
        lo, hi = self._work()
        tmp = self.material.copy()
        self.o.classify(self.pos, self.dims, self.dx, tmp)
        self.material.reshape(K, J, I)[lo:hi] = tmp.reshape(K, J, I)[lo:hi]
        for comp in range(3):
            self.o.splat_component(self.pos, self.vel, comp, self.dims, self.dx, *self.acc[comp])

    def p2g_end(self):
        for comp in range(3):
            f, w = self.acc[comp]
            self.p2g[comp] = self.o.finish_component(f, w, comp, self.dims, self.dx, self.material)
            f[:] = 0
            w[:] = 0

    def g2p_advect(self, dt):
        if len(self.pos):
            self.pos, self.vel, _ = self.o.g2p_advect(self.pos, self.vel, self.new, self.saved, self.dims, self.dx, dt,
                                                      mode=self.mode, material=self.material)

    # ---- layers
    def _array(self, what):
        if what < 3:
            return self.new[what], self.fdims[what]
        if what < 6:
            return self.saved[what - 3], self.fdims[what - 3]
        if what < 9:
            return self.p2g[what - 6], self.fdims[what - 6]
        if what == 9:
            return self.material, self.dims
        raise ValueError(what)

    def layer_bytes(self, what):
        if what >= 10:
            a, b, _ = self.fdims[what - 10]
            return a * b * 8                      # float32 sum + float32 weight
        arr, (a, b, _) = self._array(what)
        return a * b * arr.itemsize

    def pack_batch(self, items, buf):
        for what, k_first, k_count, offset, _ in items:
            t = self.pack(what, k_first, k_count)
            buf[offset: offset + t.numel()] = t

    def unpack_batch(self, items, buf):
        for what, k_first, k_count, offset, add in items:
            n = self.layer_bytes(what) * k_count
            self.unpack(what, k_first, k_count, buf[offset: offset + n].clone(), add)

    def pack(self, what, k_first, k_count):
        if what >= 10:
            a, b, c = self.fdims[what - 10]
            f, w = self.acc[what - 10]
            sl = slice(a * b * k_first, a * b * (k_first + k_count))
            raw = np.concatenate([f[sl], w[sl]]).view(np.uint8)
        else:
            arr, (a, b, c) = self._array(what)
            raw = arr[a * b * k_first: a * b * (k_first + k_count)].view(np.uint8)
        return torch.from_numpy(raw.copy())

    def unpack(self, what, k_first, k_count, t, add):
        raw = t.numpy()
        if what >= 10:
            a, b, c = self.fdims[what - 10]
            f, w = self.acc[what - 10]
            vals = raw.view(np.float32)
            sl = slice(a * b * k_first, a * b * (k_first + k_count))
            assert add
            f[sl] += vals[: a * b * k_count]
            w[sl] += vals[a * b * k_count:]
        else:
            arr, (a, b, c) = self._array(what)
            arr[a * b * k_first: a * b * (k_first + k_count)] = raw.view(arr.dtype)

    # ---- particles
    def extract_async(self, k_lo, k_hi):
        k = self.o.cell_index(self.pos, self.dx)[:, 2] if len(self.pos) else np.zeros(0, np.int32)
        down, up = k < k_lo, k >= k_hi
        out = []
        for m in (down, up):
            out.append(torch.from_numpy(np.concatenate([self.pos[m], self.vel[m]], 1).astype(np.float32).reshape(-1, 6)))
        keep = ~(down | up)
        self._kept = (self.pos[keep], self.vel[keep])
        counts = torch.tensor([int(keep.sum()), int(down.sum()), int(up.sum()), 0], dtype=torch.int32)
        return out[0], out[1], counts

    def extract_commit(self, n_kept):
        assert n_kept == len(self._kept[0])
        self.pos, self.vel = self._kept

    def append(self, t):
        a = t.numpy()
        self.pos = np.concatenate([self.pos, a[:, :3]])
        self.vel = np.concatenate([self.vel, a[:, 3:]])

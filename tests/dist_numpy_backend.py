"""TEST-ONLY local-compute backend for DistributedEngine: the float64 oracle on the per-step compact tables.  It lets
the routing / sharding / planning logic of relation_autoencoder_b200.dist run on CPU under gloo (world_size 2).  Where
the CUDA backend reads its peers' memory directly (fetch of table rows, pull of gradient rows), this backend all-gathers
the peers' buffers first - the plans, slots and summation orders under test are the production ones."""
import numpy as np
import torch
import torch.distributed as dist

from oracle import rae_oracle as O
from relation_autoencoder_b200.dist import GranularStep


class NumpyBackend(GranularStep):
    dtype = torch.float64

    def __init__(self, de):
        self.de = de
        self.model = O.MODEL_ALIASES[de.model]
        self.order = [n for n in ("C", "C1", "C2", "Wb") if n in de.names]
        self.dense_grad = None
        self._cost = torch.zeros(1, dtype=torch.float64)
        self.shards = {}
        self.compact = {}
        self.grads = {}

    # -- peers' buffers (emulation of the IPC mapping)
    def _all(self, t):
        if self.de.world == 1:
            return [t]
        n = torch.tensor([t.shape[0]])
        sizes = [torch.zeros_like(n) for _ in range(self.de.world)]
        dist.all_gather(sizes, n, group=self.de.group)
        mx = max(int(s) for s in sizes)
        pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype)
        pad[: t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(self.de.world)]
        dist.all_gather(parts, pad, group=self.de.group)
        return [parts[r][: int(sizes[r])] for r in range(self.de.world)]

    def alloc_tables(self, shapes):
        self.shards = {n: torch.zeros(shp, dtype=torch.float64) for n, shp in shapes.items()}
        return self.shards

    def setup(self, f_cap, n_cap):
        de = self.de
        f64 = dict(dtype=torch.float64)
        self.compact = {"W": torch.zeros(f_cap, de.K, **f64), "A": torch.zeros(n_cap, de.d, **f64), "Ab": torch.zeros(n_cap, **f64)}
        self.grads = {"W": torch.zeros(f_cap, de.K, **f64), "A": torch.zeros(n_cap, de.d, **f64), "Ab": torch.zeros(n_cap, **f64)}

    def bind_dense(self, dense, dense_acc):
        self.dense, self.dense_acc = dense, dense_acc
        self.dense_grad = torch.zeros(sum(dense[n].numel() for n in self.order), dtype=torch.float64)

    def bind_train_csr(self, indptr, indices_compact, a1, a2):
        self.ip, self.ixc = indptr.numpy().astype(np.int64), indices_compact.numpy().astype(np.int64)

    def fetch(self, name, ids):
        tabs = self._all(self.shards[name])
        w = self.de.world
        idl = ids.long()
        out = self.compact[name]
        for r in range(w):
            m = (idl % w) == r
            out[: idl.numel()][m] = tabs[r][idl[m] // w]

    def local_step(self, b, a1c, a2c, n1c, n2c, neg_ld):
        B = self.de.B
        lo, hi = self.ip[b * B], self.ip[(b + 1) * B]
        indptr = self.ip[b * B:(b + 1) * B + 1] - lo
        indices = self.ixc[lo:hi]
        p = {"W": self.compact["W"].numpy(), "A": self.compact["A"].numpy(), "Ab": self.compact["Ab"].numpy()}
        for n in self.order:
            p[n] = self.dense[n].numpy()
        n1 = n1c[:, :B].numpy().astype(np.int64)
        n2 = n2c[:, :B].numpy().astype(np.int64)
        cost, _, g = O.cost_and_grads(self.model, p, indptr, indices, a1c.numpy().astype(np.int64), a2c.numpy().astype(np.int64),
                                      n1, n2, self.de.alpha, z_total=self.de.z_total)
        self._cost[0] = cost
        for n in ("W", "A", "Ab"):
            self.grads[n].copy_(torch.from_numpy(g[n]))
        off = 0
        for n in self.order:
            k = g[n].size
            self.dense_grad[off:off + k] = torch.from_numpy(g[n].reshape(-1))
            off += k

    def dense_apply(self):
        off = 0
        for n in self.order:
            t, a = self.dense[n].numpy(), self.dense_acc[n].numpy()
            g = self.dense_grad[off:off + t.size].numpy().reshape(t.shape)
            off += t.size
            if self.de.optimizer == "adagrad":
                a += g * g
                t -= self.de.lr * g / (np.sqrt(a) + 1e-6)
            else:
                t -= self.de.lr * g

    def pull_apply(self, name, table, acc, rows_local, ent_off, ent_src, ent_slot):
        grads = self._all(self.grads[name])          # collective: called by every rank even when it owns no touched row
        t, a = table.numpy(), acc.numpy()
        rl, eo = rows_local.numpy(), ent_off.numpy()
        es, el = ent_src.numpy(), ent_slot.numpy()
        for i in range(len(rl)):
            g = 0.0
            for e in range(eo[i], eo[i + 1]):        # rank order
                g = g + grads[es[e]][el[e]].numpy()
            r = rl[i]
            if self.de.optimizer == "adagrad":
                a[r] = a[r] + g * g
                t[r] = t[r] - self.de.lr * g / (np.sqrt(a[r]) + 1e-6)
            else:
                t[r] = t[r] - self.de.lr * g

    def local_cost_tensor(self):
        return self._cost.clone()

    def label(self, indptr, indices_compact):
        return O.label_batch(self.compact["W"].numpy(), self.dense["Wb"].numpy(), indptr.numpy(), indices_compact.numpy())

    def close(self):
        pass

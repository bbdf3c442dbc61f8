"""TEST-ONLY local-compute backend for DistributedEngine: the float64 oracle on the per-step compact tables.  It lets
the routing / sharding / collective logic of relation_autoencoder_b200.dist run on CPU under gloo (world_size 2)."""
import numpy as np
import torch

from oracle import rae_oracle as O


class NumpyBackend:
    def __init__(self, de, f_cap, n_cap):
        self.de = de
        self.model = O.MODEL_ALIASES[de.model]
        f64 = dict(dtype=torch.float64)
        self.compact = {"W": torch.zeros(f_cap, de.K, **f64), "A": torch.zeros(n_cap, de.d, **f64), "Ab": torch.zeros(n_cap, **f64)}
        self.grads = {"W": torch.zeros(f_cap, de.K, **f64), "A": torch.zeros(n_cap, de.d, **f64), "Ab": torch.zeros(n_cap, **f64)}
        self.order = [n for n in ("C", "C1", "C2", "Wb") if n in de.names]
        self.dense_grad = None
        self._cost = 0.0

    def bind_dense(self, dense, dense_acc):
        self.dense, self.dense_acc = dense, dense_acc
        self.dense_grad = torch.zeros(sum(dense[n].numel() for n in self.order), dtype=torch.float64)

    def gather_rows(self, table, rows):
        return table[rows.long()].reshape(rows.numel(), -1).clone()

    def rows_apply(self, table, acc, rows, grads):
        t, a = table.numpy(), acc.numpy()
        r = rows.numpy().astype(np.int64)
        g = grads.numpy().reshape((len(r),) + t.shape[1:])
        uniq = np.unique(r)
        gsum = np.zeros((len(uniq),) + t.shape[1:])
        np.add.at(gsum, np.searchsorted(uniq, r), g)
        if self.de.optimizer == "adagrad":
            a[uniq] = a[uniq] + gsum * gsum
            t[uniq] = t[uniq] - self.de.lr * gsum / (np.sqrt(a[uniq]) + 1e-6)
        else:
            t[uniq] = t[uniq] - self.de.lr * gsum

    def local_step(self, indptr, indices, nnz, a1, a2, n1, n2, neg_ld):
        p = {"W": self.compact["W"].numpy(), "A": self.compact["A"].numpy(), "Ab": self.compact["Ab"].numpy()}
        for n in self.order:
            p[n] = self.dense[n].numpy()
        cost, _, g = O.cost_and_grads(self.model, p, indptr.numpy(), indices.numpy(), a1.numpy().astype(np.int64),
                                      a2.numpy().astype(np.int64), n1.numpy().astype(np.int64), n2.numpy().astype(np.int64),
                                      self.de.alpha, z_total=self.de.z_total)
        self._cost = cost
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

    def local_cost(self):
        return self._cost

    def label(self, indptr, indices):
        return O.label_batch(self.compact["W"].numpy(), self.dense["Wb"].numpy(), indptr.numpy(), indices.numpy())

    def synchronize(self):
        pass

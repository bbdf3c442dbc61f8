"""Data-parallel training step over the GPUs of one node.  The reference is single-process and has no
counterpart; the partitioning follows from its objective: examples are independent given the parameters and the cost is
a sum over examples divided by the global Z (learning/OieModel.py:90).

* batch: every rank owns B examples of each global batch of world*B; Z and adj are global.
* dense parameters C/R, C1, C2, Wb: replicated; local gradients are summed with ONE all-reduce per step (flat buffer),
  then every rank applies the identical optimiser step (Optimizers.py:29-32).
* sparse tables W[F,K], A[N,d], Ab[N] (+ AdaGrad accumulators): row-sharded, owner(row) = row mod world, local index =
  row // world.  Per step a rank (1) receives the rows its batch touches from their owners (all-to-all) into compact
  tables, (2) runs the fused step on the compact tables with RAE_FLAG_EMIT_ONLY (one reduced gradient row per touched
  row, duplicates inside the rank already summed), (3) returns the gradient rows to the owners (all-to-all), and (4) each
  owner stable-sorts the received (row, gradient) pairs - input order is source-rank major - segment-reduces them and
  applies ONE optimiser RMW per row.  Fixed orders everywhere: bitwise reproducible for a given world size.
* routing (which rows, to whom) depends only on the ids: plans are built once per batch at bind time (features) / once
  per epoch (entities), so the step itself needs no host synchronisation.

torch.distributed is the plumbing (NCCL on GPUs, gloo in the CPU tests); the numerical work is done by a backend:
:class:`CudaBackend` (librae.so) in production.  The CPU tests inject a NumPy backend to check the routing logic.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.distributed as dist

SPARSE = ("W", "A", "Ab")


@dataclass
class RowPlan:
    """Routing of one batch's distinct rows of one table."""
    ids_sorted: torch.Tensor     # int64 [U] distinct global row ids, ascending
    inv: torch.Tensor            # int64 [U] position in ids_sorted -> compact row (owner-major order)
    send_counts: List[int]       # rows requested from each owner
    recv_counts: List[int]       # rows each rank requests from me
    recv_rows: torch.Tensor      # int32 [sum recv] local row index (id // world) of the requested rows, source-rank major
    U: int

    def remap(self, x: torch.Tensor) -> torch.Tensor:
        return self.inv[torch.searchsorted(self.ids_sorted, x.to(torch.int64))].to(torch.int32)


def build_plan(ids: torch.Tensor, world: int, group=None) -> RowPlan:
    """Collective: every rank calls it with the distinct ids it needs (any order)."""
    ids = torch.unique(ids.to(torch.int64), sorted=True)
    owner = ids % world
    order = torch.argsort(owner, stable=True)
    send_ids = ids[order].contiguous()
    counts = torch.bincount(owner, minlength=world)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    sc, rc = counts.tolist(), recv_counts.tolist()
    recv_ids = torch.empty(int(sum(rc)), dtype=torch.int64, device=ids.device)
    dist.all_to_all_single(recv_ids, send_ids, rc, sc, group=group)
    inv = torch.empty_like(order)
    inv[order] = torch.arange(order.numel(), device=ids.device)
    return RowPlan(ids, inv, sc, rc, (recv_ids // world).to(torch.int32).contiguous(), int(ids.numel()))


def shard_rows(full: np.ndarray, rank: int, world: int) -> np.ndarray:
    return np.ascontiguousarray(full[rank::world])


class CudaBackend:
    """Local numerical work on one GPU through librae.so (emit-only engine on compact tables)."""

    def __init__(self, model, K, d, S, B, F_cap, N_cap, n_train, lr, alpha, optimizer, device, z_total, adj):
        from . import _lib as L
        from .engine import Engine
        self.L = L
        self.eng = Engine(model, K, d, S, B, F_cap, N_cap, n_train, lr=lr, alpha=alpha, optimizer=optimizer, device=device,
                          flags=L.RAE_FLAG_EMIT_ONLY | L.RAE_FLAG_NO_FEATURE_CACHE, z_total=z_total, adj=adj)
        self.lib, self.h = self.eng.lib, self.eng._h
        self.device = self.eng.device
        self.K, self.d = K, d
        f32 = dict(dtype=torch.float32, device=self.device)
        self.compact = {"W": torch.zeros(F_cap, K, **f32), "A": torch.zeros(N_cap, d, **f32), "Ab": torch.zeros(N_cap, **f32)}
        self.grads = {"W": torch.zeros(F_cap, K, **f32), "A": torch.zeros(N_cap, d, **f32), "Ab": torch.zeros(N_cap, **f32)}
        self.dense_grad = torch.zeros(int(self.lib.rae_dense_grad_size(self.h)), **f32)
        self.dense: Dict[str, torch.Tensor] = {}
        self.dense_acc: Dict[str, torch.Tensor] = {}
        self._cost = C.c_double(0.0)

    def _p(self, t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    def bind_dense(self, dense: Dict[str, torch.Tensor], dense_acc: Dict[str, torch.Tensor]):
        self.dense, self.dense_acc = dense, dense_acc
        g = lambda m, n: self._p(m.get(n))
        e = self.eng
        e._check(self.lib.rae_bind_params(self.h, self._p(self.compact["W"]), g(dense, "Wb"), self._p(self.compact["A"]),
                                          self._p(self.compact["Ab"]), g(dense, "C"), g(dense, "C1"), g(dense, "C2")), "rae_bind_params")
        e._check(self.lib.rae_bind_accumulators(self.h, None, g(dense_acc, "Wb"), None, None, g(dense_acc, "C"),
                                                g(dense_acc, "C1"), g(dense_acc, "C2")), "rae_bind_accumulators")
        e._check(self.lib.rae_bind_grad_buffers(self.h, self._p(self.grads["W"]), self._p(self.grads["A"]),
                                                self._p(self.grads["Ab"]), self._p(self.dense_grad)), "rae_bind_grad_buffers")

    def gather_rows(self, table: torch.Tensor, rows: torch.Tensor) -> torch.Tensor:
        width = 1 if table.dim() == 1 else table.shape[1]
        out = torch.empty((rows.numel(), width), dtype=torch.float32, device=self.device)
        self.eng._check(self.lib.rae_gather_rows(self.h, self._p(table), width, self._p(rows), rows.numel(), self._p(out),
                                                 self.eng._stream), "rae_gather_rows")
        return out

    def rows_apply(self, table, acc, rows, grads):
        width = 1 if table.dim() == 1 else table.shape[1]
        self.eng._check(self.lib.rae_sparse_rows_apply(self.h, self._p(table), self._p(acc), width, self._p(rows),
                                                       self._p(grads), rows.numel(), table.shape[0], self.eng._stream),
                        "rae_sparse_rows_apply")

    def local_step(self, indptr, indices, nnz, a1, a2, n1, n2, neg_ld):
        self.eng._check(self.lib.rae_train_step_begin_explicit(self.h, self._p(indptr), self._p(indices), int(nnz), self._p(a1),
                                                               self._p(a2), self._p(n1), self._p(n2), int(neg_ld),
                                                               self.eng._stream), "rae_train_step_begin_explicit")

    def dense_apply(self):
        self.eng._check(self.lib.rae_train_step_end(self.h, self.eng._stream), "rae_train_step_end")

    def local_cost(self) -> float:
        self.eng._check(self.lib.rae_read_cost(self.h, C.byref(self._cost), self.eng._stream), "rae_read_cost")
        return float(self._cost.value)

    def label(self, indptr, indices):
        # emit-only engines label through the explicit encoder on the compact W (bound as the 'test' split)
        self.eng.bind_split("test", indptr, indices)
        return self.eng.label("test", 0)

    def synchronize(self):
        torch.cuda.synchronize(self.device)


class DistributedEngine:
    """Same calls as :class:`relation_autoencoder_b200.engine.Engine`, one instance per rank."""

    def __init__(self, model: str, K: int, d: int, S: int, B: int, F: int, N: int, n_train: int, lr: float = 0.1,
                 l1: float = 0.0, l2: float = 0.0, alpha: float = 1.0, optimizer: str = "adagrad", ext_reg: bool = True,
                 device=0, rank: Optional[int] = None, world: Optional[int] = None, backend_factory=None, group=None):
        if l1 != 0.0 or l2 != 0.0:
            raise NotImplementedError("row-sharded multi-GPU training supports l1 = l2 = 0 only (the regulariser makes dW dense)")
        from .engine import MODEL_IDS, MODEL_PARAMS
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.group = group
        self.model, self.K, self.d, self.S, self.B, self.F, self.N = model, K, d, S, B, F, N
        self.n_train = n_train
        self.lr, self.alpha, self.optimizer = lr, alpha, optimizer
        self.names = list(MODEL_PARAMS[MODEL_IDS[model]])
        self.dense_names = [n for n in self.names if n not in SPARSE]
        self.z_total = self.world * (4 * B + 2 * B * S)                       # global Z (OieModel.py:90)
        self.adj = float(self.world * B) / float(max(1, n_train))              # OieInduction.py:131 with the global batch
        self.backend_factory = backend_factory
        self.device_index = device
        self.backend = None
        self.dev = torch.device("cpu") if backend_factory is not None else torch.device("cuda", device)
        self.shard: Dict[str, torch.Tensor] = {}
        self.shard_acc: Dict[str, torch.Tensor] = {}
        self.dense: Dict[str, torch.Tensor] = {}
        self.dense_acc: Dict[str, torch.Tensor] = {}
        self.fplans: List[RowPlan] = []
        self.fcsr = []
        self.eplans: List[RowPlan] = []
        self.eidx = []
        self._pending = None
        self._last_cost = None

    # ------------------------------------------------------------------ parameters
    def set_params_numpy(self, params: Dict[str, np.ndarray], acc: Optional[Dict[str, np.ndarray]] = None):
        """``params`` are the FULL tables (every rank passes the same arrays); each rank keeps its own rows."""
        dt = torch.float64 if self.backend_factory is not None else torch.float32
        for n in self.names:
            a = params[n] if acc is None else acc[n]
            if n in SPARSE:
                self.shard[n] = torch.as_tensor(shard_rows(np.asarray(params[n]), self.rank, self.world)).to(self.dev, dt).contiguous()
                self.shard_acc[n] = (torch.zeros_like(self.shard[n]) if acc is None else
                                     torch.as_tensor(shard_rows(np.asarray(a), self.rank, self.world)).to(self.dev, dt).contiguous())
            else:
                self.dense[n] = torch.as_tensor(np.ascontiguousarray(params[n])).to(self.dev, dt).contiguous()
                self.dense_acc[n] = (torch.zeros_like(self.dense[n]) if acc is None else
                                     torch.as_tensor(np.ascontiguousarray(a)).to(self.dev, dt).contiguous())
        if self.backend is not None:
            self.backend.bind_dense(self.dense, self.dense_acc)

    def _gather_full(self, shards: Dict[str, torch.Tensor], dense: Dict[str, torch.Tensor]) -> Dict[str, np.ndarray]:
        out = {}
        for n in self.names:
            if n in SPARSE:
                total = self.F if n == "W" else self.N
                width = shards[n].shape[1:] if shards[n].dim() > 1 else ()
                full = torch.zeros((total,) + tuple(width), dtype=shards[n].dtype, device=self.dev)
                rows = (total + self.world - 1) // self.world
                pad = torch.zeros((rows,) + tuple(width), dtype=shards[n].dtype, device=self.dev)
                pad[: shards[n].shape[0]] = shards[n]
                parts = [torch.empty_like(pad) for _ in range(self.world)]
                dist.all_gather(parts, pad, group=self.group)
                for r in range(self.world):
                    cnt = len(range(r, total, self.world))
                    full[r::self.world] = parts[r][:cnt]
                out[n] = full.cpu().numpy()
            else:
                out[n] = dense[n].cpu().numpy()
        return out

    def get_params_numpy(self) -> Dict[str, np.ndarray]:
        return self._gather_full(self.shard, self.dense)

    def get_acc_numpy(self) -> Dict[str, np.ndarray]:
        return self._gather_full(self.shard_acc, self.dense_acc)

    # ------------------------------------------------------------------ data
    def _i64(self, x):
        if isinstance(x, torch.Tensor):
            return x.to(self.dev, torch.int64)
        return torch.as_tensor(np.ascontiguousarray(x)).to(self.dev, torch.int64)

    def bind_split(self, split: str, indptr, indices, args1=None, args2=None):
        """Local rows of the split.  For 'train': builds the per-batch feature routing plans (collective)."""
        ip, ix = self._i64(indptr), self._i64(indices)
        if split != "train":
            self._label_split = getattr(self, "_label_split", {})
            self._label_split[split] = (ip, ix)
            return
        self.ip, self.ix = ip, ix
        self.a1, self.a2 = self._i64(args1), self._i64(args2)
        nb = (ip.numel() - 1) // self.B
        nbt = torch.tensor([nb], device=self.dev)
        dist.all_reduce(nbt, op=dist.ReduceOp.MIN, group=self.group)
        self.nb = int(nbt.item())                                              # every rank steps the same number of batches
        self.fplans, self.fcsr = [], []
        for b in range(self.nb):
            lo, hi = int(ip[b * self.B]), int(ip[(b + 1) * self.B])
            feats = ix[lo:hi]
            plan = build_plan(feats, self.world, self.group)
            self.fplans.append(plan)
            self.fcsr.append(((ip[b * self.B:(b + 1) * self.B + 1] - lo).to(torch.int32).contiguous(),
                              plan.remap(feats).contiguous(), hi - lo))
        if self.backend is None:
            f_cap = max([p.U for p in self.fplans] + [1])
            n_cap = (2 + 2 * self.S) * self.B
            if self.backend_factory is not None:
                self.backend = self.backend_factory(self, f_cap, n_cap)
            else:
                self.backend = CudaBackend(self.model, self.K, self.d, self.S, self.B, f_cap, n_cap, self.n_train, self.lr,
                                           self.alpha, self.optimizer, self.device_index, self.z_total, self.adj)
            self.backend.bind_dense(self.dense, self.dense_acc)

    def n_batches(self, split: str = "train") -> int:
        return self.nb

    def _entity_plan(self, b, n1, n2):
        """n1, n2: int64 [S, B] negatives of batch b (this rank's columns)."""
        a1 = self.a1[b * self.B:(b + 1) * self.B]
        a2 = self.a2[b * self.B:(b + 1) * self.B]
        plan = build_plan(torch.cat([a1, a2, n1.reshape(-1), n2.reshape(-1)]), self.world, self.group)
        idx = (plan.remap(a1).contiguous(), plan.remap(a2).contiguous(),
               plan.remap(n1.reshape(-1)).reshape(n1.shape).contiguous(), plan.remap(n2.reshape(-1)).reshape(n2.shape).contiguous())
        return plan, idx

    def bind_epoch_negatives(self, neg1, neg2):
        """The epoch's negatives [S, n_local] (OieInduction.py:183-184): builds the entity routing plans (collective)."""
        n1, n2 = self._i64(neg1), self._i64(neg2)
        self.eplans, self.eidx = [], []
        for b in range(self.nb):
            plan, idx = self._entity_plan(b, n1[:, b * self.B:(b + 1) * self.B], n2[:, b * self.B:(b + 1) * self.B])
            self.eplans.append(plan)
            self.eidx.append(idx)

    # ------------------------------------------------------------------ the step
    def _fetch(self, plan: RowPlan, name: str, out: torch.Tensor):
        rows = self.backend.gather_rows(self.shard[name], plan.recv_rows)
        view = out[: plan.U].reshape(plan.U, -1)
        dist.all_to_all_single(view, rows, plan.send_counts, plan.recv_counts, group=self.group)

    def _return_and_apply(self, plan: RowPlan, name: str, grads: torch.Tensor):
        view = grads[: plan.U].reshape(plan.U, -1)
        recv = torch.empty((int(sum(plan.recv_counts)), view.shape[1]), dtype=view.dtype, device=self.dev)
        dist.all_to_all_single(recv, view.contiguous(), plan.recv_counts, plan.send_counts, group=self.group)
        self.backend.rows_apply(self.shard[name], self.shard_acc[name], plan.recv_rows, recv)

    def _step(self, b, eplan, eidx, want_cost):
        fplan = self.fplans[b]
        ipc, ixc, nnz = self.fcsr[b]
        bk = self.backend
        self._fetch(fplan, "W", bk.compact["W"])
        self._fetch(eplan, "A", bk.compact["A"])
        self._fetch(eplan, "Ab", bk.compact["Ab"])
        a1c, a2c, n1c, n2c = eidx
        bk.local_step(ipc, ixc, nnz, a1c, a2c, n1c, n2c, self.B)
        dist.all_reduce(bk.dense_grad, group=self.group)          # sum of the ranks' dense gradients (C | C1 | C2 | Wb)
        bk.dense_apply()
        self._return_and_apply(fplan, "W", bk.grads["W"])
        self._return_and_apply(eplan, "A", bk.grads["A"])
        self._return_and_apply(eplan, "Ab", bk.grads["Ab"])
        if want_cost:
            c = torch.tensor([bk.local_cost()], dtype=torch.float64, device=self.dev)
            dist.all_reduce(c, group=self.group)
            return float(c.item())
        return None

    def train_device(self, batch_index: int, want_cost: bool = True):
        return self._step(batch_index, self.eplans[batch_index], self.eidx[batch_index], want_cost)

    def train(self, batch_index: int, neg1, neg2) -> float:
        """func['train'](batch_index, neg1, neg2) with this rank's host negatives [S,B]; returns the GLOBAL batch cost."""
        plan, idx = self._entity_plan(batch_index, self._i64(neg1), self._i64(neg2))
        return self._step(batch_index, plan, idx, True)

    def label(self, split: str, batch_index: int):
        if split == "train":
            ip, ix = self.ip, self.ix
        else:
            ip, ix = self._label_split[split]
        lo, hi = int(ip[batch_index * self.B]), int(ip[(batch_index + 1) * self.B])
        feats = ix[lo:hi]
        plan = build_plan(feats, self.world, self.group)
        self._fetch(plan, "W", self.backend.compact["W"])
        return self.backend.label((ip[batch_index * self.B:(batch_index + 1) * self.B + 1] - lo).to(torch.int32), plan.remap(feats))

    # ------------------------------------------------------------------ misc
    def stats(self) -> dict:
        st = self.backend.eng.stats() if hasattr(self.backend, "eng") else {}
        return st

    def set_profiling(self, on: bool):
        pass

    def close(self):
        if self.backend is not None and hasattr(self.backend, "eng"):
            self.backend.eng.close()

"""Host-side glue of a training run around the drop-in modules — the callers on either side of the hot path.

Stand-ins for what the reference takes from PyG and for its ``TrainModel`` step recipe, so that a training script written against the
reference (``*/main.py`` + ``*/training_class.py``) runs on the B200 path with its structure unchanged:

* :class:`Batch` / :func:`Batch.from_data_list` and :class:`DataListLoader` — ``torch_geometric.data.Batch`` /
  ``torch_geometric.loader.DataListLoader`` as used at ``dirichlet/psignn/main.py:70-77`` (lists of per-mesh ``Data`` objects,
  collated by concatenation with node-index offsets);
* :class:`DataParallel` — ``torch_geometric.nn.DataParallel`` as used at ``main.py:106``: a wrapper with a ``.module`` attribute that
  accepts a *list* of graphs.  Here it is SPMD (one process per GPU under ``torch.distributed``): the wrapper collates the rank's
  list, keeps the collated device batch **and its native re-layout** in an LRU cache keyed by the items of the list (an epoch loop
  with ``shuffle=False`` — the Dirichlet reader — pays collation, H2D copy and the SELL build once per batch, not once per step),
  and the gradient all-reduce happens in :meth:`TrainModel.train_loop`;
* :class:`TrainModel` — the step recipe of ``dirichlet/psignn/training_class.py:147-166`` (two Adam optimisers, launch-script loss
  combination, gradient clipping, ReduceLROnPlateau) and its checkpoint layout (``:297-307``), without the plotting and CSV logging.

Nothing here is on the hot path; it only has to keep it fed.
"""
from __future__ import annotations

import os
import time
from collections import OrderedDict
from typing import Iterable, List, Optional, Sequence

import torch
import torch.nn as nn

from . import parallel
from .synthetic import GraphData, collate


class Batch(GraphData):
    """``torch_geometric.data.Batch`` stand-in: a :class:`GraphData` with ``num_graphs``, ``ptr`` and ``batch``."""

    @staticmethod
    def from_data_list(data_list: Sequence[GraphData]) -> "Batch":
        out = Batch()
        out.__dict__.update(collate(list(data_list)).__dict__)
        return out


class DataListLoader:
    """yields *lists* of dataset items (``torch_geometric.loader.DataListLoader``); under ``world > 1`` every rank iterates over its
    own contiguous share of each global batch (the split PyG's ``DataParallel.scatter`` makes inside one process)"""

    def __init__(self, dataset: Sequence[GraphData], batch_size: int = 1, shuffle: bool = False, rank: int = 0, world: int = 1, seed: int = 0):
        self.dataset, self.batch_size, self.shuffle = dataset, int(batch_size), bool(shuffle)
        self.rank, self.world, self.seed, self.epoch = int(rank), int(world), int(seed), 0

    def __len__(self):
        return (len(self.dataset) + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        n = len(self.dataset)
        order = list(range(n))
        if self.shuffle:
            g = torch.Generator().manual_seed(self.seed + self.epoch)
            order = torch.randperm(n, generator=g).tolist()
        self.epoch += 1
        for s in range(0, n, self.batch_size):
            ids = order[s:s + self.batch_size]
            if self.world > 1:
                per = (len(ids) + self.world - 1) // self.world
                ids = ids[self.rank * per:(self.rank + 1) * per]
            if ids:
                items = [self.dataset[i] for i in ids]
                for i, it in zip(ids, items):
                    if getattr(it, "_psi_item_id", None) is None:
                        it._psi_item_id = i
                yield items


class DataParallel(nn.Module):
    """``torch_geometric.nn.DataParallel`` stand-in (see module docstring)."""

    def __init__(self, module: nn.Module, device=None, cache_batches: int = 64):
        super().__init__()
        self.module = module
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.cache_batches = int(cache_batches)
        self._cache: "OrderedDict[tuple, Batch]" = OrderedDict()
        self.cache_hits = 0

    def collate(self, data_list) -> Batch:
        if isinstance(data_list, GraphData):
            return data_list if data_list.edge_index.device == self.device else data_list.to(self.device)
        key = tuple(getattr(d, "_psi_item_id", id(d)) for d in data_list)
        hit = self._cache.get(key)
        if hit is not None:
            self._cache.move_to_end(key)
            self.cache_hits += 1
            return hit
        b = Batch.from_data_list(data_list).pin_memory().to(self.device, non_blocking=True)
        out = Batch()
        out.__dict__.update(b.__dict__)
        if self.cache_batches > 0:
            self._cache[key] = out          # the native graph of the batch is cached ON the batch object by graph_of()
            while len(self._cache) > self.cache_batches:
                self._cache.popitem(last=False)
        return out

    def forward(self, data_list):
        return self.module(self.collate(data_list))

    def inference(self, data_list):
        return self.module.inference(self.collate(data_list))


class TrainModel:
    """Step recipe and checkpoint layout of the reference's ``TrainModel`` (dirichlet/psignn/training_class.py:20-66, 147-166, 225-240,
    297-307).  ``config`` takes the reference's keys: loader_train, loader_val, model (a :class:`DataParallel`), config_model, lr_deq,
    lr_ae, sched_step_deq, sched_step_ae, path_ckpt, max_epochs, gradient_clip, jac_weight, min_loss_save (sup_weight is accepted and
    unused, as in the reference)."""

    LOSS_KEYS = ("loss", "residual_loss", "jacobian_loss", "encoder_loss", "autoencoder_loss", "mse_loss")

    def __init__(self, config):
        self.loader_train, self.loader_val = config["loader_train"], config.get("loader_val")
        self.model, self.config_model = config["model"], config.get("config_model", {})
        self.lr_deq, self.lr_ae = config["lr_deq"], config["lr_ae"]
        self.sched_step_deq, self.sched_step_ae = config.get("sched_step_deq", 0.5), config.get("sched_step_ae", 0.5)
        self.path_ckpt = config.get("path_ckpt")
        self.min_loss_save = config.get("min_loss_save", float("inf"))
        self.max_epochs = config.get("max_epochs", 1)
        self.gradient_clip = config["gradient_clip"]
        self.jac_weight = config["jac_weight"]
        self.training_time = 0.0
        self.hist_train = {k: [] for k in self.LOSS_KEYS}
        self.hist_val = {k: [] for k in self.LOSS_KEYS}
        self.createOptimizerAndScheduler()

    def createOptimizerAndScheduler(self):
        m = self.model.module
        self.opt_deq = torch.optim.Adam(m.deqdss.parameters(), lr=self.lr_deq)
        self.sched_deq = torch.optim.lr_scheduler.ReduceLROnPlateau(self.opt_deq, mode='min', factor=self.sched_step_deq)
        self.opt_ae = torch.optim.Adam(m.autoencoder.parameters(), lr=self.lr_ae)
        self.sched_ae = torch.optim.lr_scheduler.ReduceLROnPlateau(self.opt_ae, mode='min', factor=self.sched_step_ae)

    def _loss(self, loss_dic):
        return (loss_dic["residual_loss"].mean() + self.jac_weight * loss_dic["jacobian_loss"].mean()
                + loss_dic["encoder_loss"].mean() + loss_dic["autoencoder_loss"].mean())

    def train_step(self, train_batch):
        """one batch of training_class.py:147-166 (+ the gradient all-reduce that replaces DataParallel's ReduceAddCoalesced)"""
        self.opt_ae.zero_grad()
        self.opt_deq.zero_grad()
        U_sol, loss_dic = self.model(train_batch)
        loss = self._loss(loss_dic)
        loss.backward()
        params = list(self.model.module.parameters())
        parallel.allreduce_gradients(params)
        torch.nn.utils.clip_grad_norm_(params, self.gradient_clip)
        self.opt_deq.step()
        self.opt_ae.step()
        return loss, loss_dic

    def train_loop(self, current_epoch=0):
        self.model.train()
        acc = {k: 0.0 for k in self.LOSS_KEYS}
        n = 0
        for train_batch in self.loader_train:
            loss, ld = self.train_step(train_batch)
            acc["loss"] += loss.item()
            for k in self.LOSS_KEYS[1:]:
                acc[k] += ld[k].mean().item()
            n += 1
        for k in self.LOSS_KEYS:
            self.hist_train[k].append(acc[k] / max(n, 1))
        return self.hist_train["loss"][-1]

    def val_loop(self, current_epoch=0):
        """validation under no_grad (training_class.py:225-240); the drop-in DeepEquilibrium then takes the eval branch of the
        reference (Jacobian estimate + 150 power iterations on the native VJP)"""
        if self.loader_val is None:
            return None
        self.model.eval()
        acc = {k: 0.0 for k in self.LOSS_KEYS}
        n = 0
        with torch.no_grad():
            for val_batch in self.loader_val:
                _, ld = self.model(val_batch)
                acc["loss"] += self._loss(ld).item()
                for k in self.LOSS_KEYS[1:]:
                    acc[k] += ld[k].mean().item()
                n += 1
        for k in self.LOSS_KEYS:
            self.hist_val[k].append(acc[k] / max(n, 1))
        return self.hist_val["residual_loss"][-1]

    def state(self, epoch):
        """the reference's checkpoint dictionary (training_class.py:297-307)"""
        return {"epoch": epoch, "hyperparameters": {k: v for k, v in self.config_model.items() if k != "solver"},
                "state_dict": self.model.module.state_dict(), "hist_train": self.hist_train, "hist_val": self.hist_val,
                "opt_deq": self.opt_deq.state_dict(), "opt_ae": self.opt_ae.state_dict(),
                "sched_deq": self.sched_deq.state_dict(), "sched_ae": self.sched_ae.state_dict(), "training_time": self.training_time}

    def save_model(self, state, dirName=None, model_name=None):
        os.makedirs(dirName, exist_ok=True)
        torch.save(state, os.path.join(dirName, "{}.pt".format(model_name)))

    def load_model(self, path):
        ck = torch.load(path, map_location=self.model.device, weights_only=False)
        self.model.module.load_state_dict(ck["state_dict"])
        self.opt_deq.load_state_dict(ck["opt_deq"]); self.opt_ae.load_state_dict(ck["opt_ae"])
        self.sched_deq.load_state_dict(ck["sched_deq"]); self.sched_ae.load_state_dict(ck["sched_ae"])
        self.hist_train, self.hist_val = ck["hist_train"], ck["hist_val"]
        return ck["epoch"]

    def train_model(self):
        best = float("inf")
        t0 = time.time()
        for epoch in range(self.max_epochs):
            self.train_loop(epoch)
            val = self.val_loop(epoch)
            self.training_time = time.time() - t0
            metric = val if val is not None else self.hist_train["residual_loss"][-1]
            self.sched_deq.step(metric)
            self.sched_ae.step(metric)
            if self.path_ckpt:
                self.save_model(self.state(epoch), self.path_ckpt, "running_model")
                if metric < best and metric < self.min_loss_save:
                    best = metric
                    self.save_model(self.state(epoch), self.path_ckpt, "best_model")
            if self.opt_deq.param_groups[0]["lr"] <= 1e-7 and self.opt_ae.param_groups[0]["lr"] <= 1e-7:
                break
        if self.path_ckpt:
            self.save_model(self.state(epoch), self.path_ckpt, "final_model")
        return self.hist_train, self.hist_val

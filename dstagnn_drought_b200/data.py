"""Data path of the reference's driver, B200 side: graph CSVs, the prepared ``*_dstagnn.npz`` signal file,
batch loaders.

Mirrors, by name and return layout, what ``train_DSTAGNN_my.py:51-73`` calls:

  * ``load_graphdata_channel1``  (lib/utils1.py:294-343)  npz -> tensors + train / val / test loaders
  * ``get_adjacency_matrix2``    (lib/utils1.py:92-145)   edge-list CSV -> {0,1} adjacency
  * ``load_weighted_adjacency_matrix`` / ``load_PA`` / ``load_weighted_adjacency_matrix2``
                                 (lib/dataloader.py:5-23)  STAG / STRG CSVs -> binarised matrices

Differences from the reference, all on the plumbing side:
  * The loaders are ``DeviceLoader`` objects, not ``torch.utils.data.DataLoader``: a PEMS-size data set
    (PEMS07: 1.2 GB) is a rounding error of 180 GB of HBM, so it is kept RESIDENT on the device and a batch is
    one index-gather kernel -- no per-step host work, no H2D copy.  A data set that does not fit the budget
    stays in pinned host memory and the next batch is copied on a side stream while the current one trains.
  * Batches are sharded over data-parallel ranks (the reference's 8-core launch does not shard):
    every rank draws the same permutation and takes its own contiguous slice of each global batch.
"""
from __future__ import annotations

import csv
import os
from typing import Iterator, Optional, Tuple

import numpy as np
import torch


# ----------------------------------------------------------------------------------------------------
# graph files
# ----------------------------------------------------------------------------------------------------
def _read_matrix_csv(path: str) -> np.ndarray:
    """Header-less numeric CSV -> 2-D float64 array (pd.read_csv(header=None).to_numpy() in the reference)."""
    return np.atleast_2d(np.loadtxt(path, delimiter=",", dtype=np.float64))


def load_weighted_adjacency_matrix(file_path: str, num_v: Optional[int] = None) -> np.ndarray:
    """STAG CSV -> float64 {0,1} (lib/dataloader.py:5-9; the weights are binarised)."""
    return np.float64(_read_matrix_csv(file_path) > 0)


def load_PA(file_path: str) -> np.ndarray:
    """STRG CSV -> float64 {0,1} (lib/dataloader.py:11-15)."""
    return np.float64(_read_matrix_csv(file_path) > 0)


def load_weighted_adjacency_matrix2(file_path: str, num_v: int) -> np.ndarray:
    """Adjacency CSV of the non-PEMS data sets: binarised, identity subtracted (lib/dataloader.py:17-23)."""
    return np.int64(_read_matrix_csv(file_path) > 0) - np.identity(num_v)


def get_adjacency_matrix2(distance_df_filename: str, num_of_vertices: int, type_: str = "connectivity",
                          id_filename: Optional[str] = None) -> np.ndarray:
    """Edge list ``from,to,cost`` -> float32 adjacency (lib/utils1.py:92-145): directed connectivity
    (``A[i, j] = 1``) by default, symmetric when an id file remaps the node ids."""
    n = int(num_of_vertices)
    A = np.zeros((n, n), dtype=np.float32)
    id_dict = None
    if id_filename:
        with open(id_filename, "r") as f:
            id_dict = {int(i): idx for idx, i in enumerate(f.read().strip().split("\n"))}
    with open(distance_df_filename, "r") as f:
        f.readline()
        for row in csv.reader(f):
            if len(row) != 3:
                continue
            i, j, distance = int(row[0]), int(row[1]), float(row[2])
            if id_dict is not None:
                A[id_dict[i], id_dict[j]] = 1
                A[id_dict[j], id_dict[i]] = 1
            elif type_ == "connectivity":
                A[i, j] = 1
            elif type_ == "distance":
                A[i, j] = 1 / distance
                A[j, i] = 1 / distance
            else:
                raise ValueError("type_ error, must be connectivity or distance!")
    return A


# ----------------------------------------------------------------------------------------------------
# loaders
# ----------------------------------------------------------------------------------------------------
class DeviceLoader:
    """Iterable of ``(x, target)`` batches on ``device``; ``len()`` = number of (global) batches, as
    ``len(DataLoader)`` in the reference's loop (train_DSTAGNN_my.py:59,170).

    ``x`` (S, N, F, T) and ``target`` (S, N, P) are either resident on the device (``resident=True``) or in
    pinned host memory with one batch of look-ahead on a side stream.  With ``world > 1`` a global batch of
    ``batch_size * world`` samples is split into contiguous per-rank slices: the ranks together see every
    sample once per epoch, and the single-process loss over the global batch equals the mean of the ranks'.
    """

    def __init__(self, x: torch.Tensor, target: torch.Tensor, batch_size: int, shuffle: bool, device,
                 rank: int = 0, world: int = 1, seed: int = 1, resident: Optional[bool] = None,
                 in_channels: Optional[int] = None, drop_ragged: bool = False):
        assert x.shape[0] == target.shape[0]
        self.device = torch.device(device)
        self.batch_size, self.shuffle, self.rank, self.world, self.seed = batch_size, shuffle, rank, world, seed
        self.in_channels = in_channels
        self.n = x.shape[0]
        self.epoch = 0
        self.drop_ragged = drop_ragged
        on_gpu = self.device.type == "cuda"
        if resident is None:
            nbytes = (x.numel() + target.numel()) * 4
            budget = 0.2 * torch.cuda.get_device_properties(self.device).total_memory if on_gpu else float("inf")
            resident = nbytes <= budget
        self.resident = bool(resident)
        if self.resident:
            self.x, self.target = x.to(self.device), target.to(self.device)
        else:
            self.x, self.target = x.pin_memory(), target.pin_memory()
            self._copy_stream = torch.cuda.Stream(self.device)

    def __len__(self) -> int:
        g = self.batch_size * self.world
        return self.n // g if self.drop_ragged else (self.n + g - 1) // g

    def set_epoch(self, epoch: int):
        self.epoch = epoch

    def _order(self) -> torch.Tensor:
        if not self.shuffle:
            return torch.arange(self.n)
        g = torch.Generator().manual_seed(self.seed * 1000003 + self.epoch)      # same permutation on every rank
        return torch.randperm(self.n, generator=g)

    def _slices(self):
        order = self._order()
        g = self.batch_size * self.world
        for b in range(len(self)):
            idx = order[b * g:(b + 1) * g]
            per = (idx.numel() + self.world - 1) // self.world               # ragged last batch: near-equal shares
            mine = idx[self.rank * per:(self.rank + 1) * per]
            if mine.numel():
                yield mine

    def _select(self, t: torch.Tensor, idx: torch.Tensor, is_x: bool) -> torch.Tensor:
        out = t.index_select(0, idx)
        if is_x and self.in_channels is not None and out.shape[2] != self.in_channels:
            out = out[:, :, :self.in_channels].contiguous()                     # keep feature 0
        return out

    def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
        if self.resident:
            for idx in self._slices():
                idx = idx.to(self.device)
                yield self._select(self.x, idx, True), self._select(self.target, idx, False)
            return
        # host-resident data: gather on the host, copy on a side stream one batch ahead
        main = torch.cuda.current_stream(self.device)
        nxt = None

        def stage(idx):
            xb = self._select(self.x, idx, True).pin_memory()
            tb = self._select(self.target, idx, False).pin_memory()
            with torch.cuda.stream(self._copy_stream):
                xd, td = xb.to(self.device, non_blocking=True), tb.to(self.device, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(self._copy_stream)
            return xd, td, ev, (xb, tb)

        for idx in self._slices():
            cur, nxt = nxt, stage(idx)
            if cur is not None:
                main.wait_event(cur[2])
                cur[0].record_stream(main), cur[1].record_stream(main)
                yield cur[0], cur[1]
        if nxt is not None:
            main.wait_event(nxt[2])
            nxt[0].record_stream(main), nxt[1].record_stream(main)
            yield nxt[0], nxt[1]


def dstagnn_npz_path(graph_signal_matrix_filename: str, num_of_hours: int, num_of_days: int, num_of_weeks: int) -> str:
    """``<dir>/<file>_r{h}_d{d}_w{w}_dstagnn.npz`` (lib/utils1.py:295-297; written by prepareData.py:137-140)."""
    file = os.path.basename(graph_signal_matrix_filename).split(".")[0]
    dirpath = os.path.dirname(graph_signal_matrix_filename)
    return os.path.join(dirpath, f"{file}_r{num_of_hours}_d{num_of_days}_w{num_of_weeks}_dstagnn.npz")


def load_graphdata_channel1(graph_signal_matrix_filename, num_of_hours, num_of_days, num_of_weeks, DEVICE, batch_size,
                            shuffle=True, rank: int = 0, world: int = 1, in_channels: Optional[int] = None,
                            resident: Optional[bool] = None, verbose: bool = True):
    """Same arguments and the same 11-tuple as the reference (lib/utils1.py:294-343):

        train_x, train_loader, train_target, val_x, val_loader, val_target, test_x, test_loader, test_target, mean, std

    with ``x`` (S, N, F, T) / ``target`` (S, N, P) float32 tensors and ``DeviceLoader`` loaders that yield batches
    on ``DEVICE``.  The returned tensors are the loaders' own storage (on the device when resident)."""
    filename = dstagnn_npz_path(graph_signal_matrix_filename, num_of_hours, num_of_days, num_of_weeks)
    if verbose and rank == 0:
        print("load file:", filename[:-4])
    d = np.load(filename)
    out = []
    for split, shuf in (("train", shuffle), ("val", False), ("test", False)):
        x = torch.from_numpy(d[f"{split}_x"]).type(torch.FloatTensor)
        t = torch.from_numpy(d[f"{split}_target"]).type(torch.FloatTensor)
        loader = DeviceLoader(x, t, batch_size, shuf, DEVICE, rank, world, in_channels=in_channels, resident=resident)
        out += [loader.x, loader, loader.target]
        if verbose and rank == 0:
            print(f"{split}:", tuple(x.shape), tuple(t.shape), "(device resident)" if loader.resident else "(pinned host)")
    return (*out, d["mean"], d["std"])


def write_synthetic_dataset(graph_signal_matrix_filename: str, num_of_hours: int, num_of_days: int, num_of_weeks: int,
                            num_of_vertices: int, len_input: int, num_for_predict: int, num_features: int = 1,
                            samples: int = 96, seed: int = 0) -> str:
    """A ``*_dstagnn.npz`` of the right layout filled with N(0,1) signals (real inputs are z-scored,
    prepareData.py:149-161) and a 60/20/20 split (prepareData.py:108-113): lets a conf run end to end when the
    PEMS / GAMBIA signal files are not available (they are not shipped with the reference, README.md:24-25)."""
    rng = np.random.RandomState(seed)
    path = dstagnn_npz_path(graph_signal_matrix_filename, num_of_hours, num_of_days, num_of_weeks)
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    n1, n2 = int(samples * 0.6), int(samples * 0.8)
    x = rng.standard_normal((samples, num_of_vertices, num_features, len_input)).astype(np.float32)
    # a learnable target: a smoothed continuation of feature 0, so that a short run shows the loss falling
    base = x[:, :, 0, -1:]
    t = (base + 0.1 * rng.standard_normal((samples, num_of_vertices, num_for_predict))).astype(np.float32)
    ts = np.arange(samples, dtype=np.int64)[:, None]
    np.savez_compressed(path[:-4], train_x=x[:n1], train_target=t[:n1], train_timestamp=ts[:n1],
                        val_x=x[n1:n2], val_target=t[n1:n2], val_timestamp=ts[n1:n2],
                        test_x=x[n2:], test_target=t[n2:], test_timestamp=ts[n2:],
                        mean=np.zeros((1, 1, num_features, 1), np.float32), std=np.ones((1, 1, num_features, 1), np.float32))
    return path


def synthetic_graph_files(dirpath: str, dataset: str, num_of_vertices: int, seed: int = 0, density: float = 0.01):
    """STAG / STRG / edge-list CSVs of a synthetic 1 %-dense graph in the formats the loaders above read
    (mimics data/STAG_gen.py:108-116: the top ``density * N`` entries per row are kept)."""
    rng = np.random.RandomState(seed)
    n = num_of_vertices
    m = max(1, int(density * n))
    tmd = np.zeros((n, n), np.float32)
    pa = np.zeros((n, n), np.float32)
    for i in range(n):
        cols = rng.permutation(n)[: m + 1]
        tmd[i, cols[:m]] = 1.0
        pa[i, cols] = 1.0
    os.makedirs(dirpath, exist_ok=True)
    stag = os.path.join(dirpath, f"stag_001_{dataset}.csv")
    strg = os.path.join(dirpath, f"strg_001_{dataset}.csv")
    adj = os.path.join(dirpath, f"{dataset}.csv")
    np.savetxt(stag, tmd, fmt="%g", delimiter=",")
    np.savetxt(strg, pa, fmt="%g", delimiter=",")
    with open(adj, "w") as f:
        f.write("from,to,cost\n")
        for i, j in zip(*np.nonzero(tmd)):
            f.write(f"{i},{j},1.0\n")
    return adj, stag, strg

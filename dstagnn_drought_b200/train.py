"""Config-driven training driver: the CUDA / NCCL replacement of the reference's TPU script.

    python -m dstagnn_drought_b200.train --config configurations/PEMS07_dstagnn.conf [--synthetic]
    torchrun --nproc-per-node 8 -m dstagnn_drought_b200.train --config ...        # batch-sharded, NCCL

Follows ``train_DSTAGNN_my.py:36-191`` step for step -- INI ``[Data]`` / ``[Training]`` sections, seed 1, the
prepared ``*_dstagnn.npz`` signal file, STAG / STRG / adjacency CSVs, ``make_model`` in the reference's argument
order, SmoothL1 + Adam, one validation pass per epoch, ``epoch_{e}.params`` whenever the validation loss improves,
final test pass on the best checkpoint -- without torch_xla / tensorboardX.  What differs, and why:

  * D1/D2: the model is built on the CUDA device directly (the reference builds on 'cpu' and moves
    to the XLA device; ``ctx`` is ignored there as well).
  * D3: ``graph`` defaults to ``AG`` (the PEMS03/07/08 confs have no such key; the reference raises KeyError).
  * D4: ``in_channels`` other than 1 cannot run through the reference model (it raises at
    model/DSTAGNN_my.py:252).  The GAMBIA conf (in_channels = 4) is run with in_channels = 1 on feature 0 of the
    signal, with a warning -- the only configuration of that data set the reference model can execute.
  * D5: ``xm.optimizer_step`` is called twice per iteration in the reference (lines 148 and 158), applying each
    gradient twice; here the optimizer steps once per batch.
  * D6: the 8-replica launch of the reference does not shard the data; here every rank takes its own slice of
    each global batch and gradients are averaged over NVLink (``parallel.GradBuckets``).
"""
from __future__ import annotations

import argparse
import configparser
import json
import os
import random
import sys
import warnings
from dataclasses import dataclass
from time import time
from typing import Optional

import numpy as np
import torch

from . import data as D
from .metrics import predict_and_save_results

PEMS = ("PEMS04", "PEMS08", "PEMS07", "PEMS03")


@dataclass
class Config:
    # [Data]
    adj_filename: str
    graph_signal_matrix_filename: str
    stag_filename: str
    strg_filename: str
    id_filename: Optional[str]
    num_of_vertices: int
    points_per_hour: int
    num_for_predict: int
    len_input: int
    dataset_name: str
    # [Training]
    in_channels: int
    conf_in_channels: int
    nb_block: int
    n_heads: int
    K: int
    d_k: int
    d_model: int
    nb_chev_filter: int
    nb_time_filter: int
    batch_size: int
    graph: str
    model_name: str
    num_of_weeks: int
    num_of_days: int
    num_of_hours: int
    start_epoch: int
    epochs: int
    learning_rate: float


def read_config(path: str) -> Config:
    """Parse a reference ``configurations/*.conf`` (train_DSTAGNN_my.py:41-45 and the keys read at 51-104,
    118-124, 138-139)."""
    cp = configparser.ConfigParser()
    if not cp.read(path):
        raise FileNotFoundError(path)
    d, t = cp["Data"], cp["Training"]
    conf_in = int(t["in_channels"])
    in_channels = conf_in
    if conf_in != 1:
        warnings.warn(f"{os.path.basename(path)}: in_channels = {conf_in} cannot run through the reference model "
                      "(the residual add of model/DSTAGNN_my.py:252 raises for it, SURVEY D4); running with "
                      "in_channels = 1 on feature 0 of the signal")
        in_channels = 1
    return Config(
        adj_filename=d["adj_filename"], graph_signal_matrix_filename=d["graph_signal_matrix_filename"],
        stag_filename=d["stag_filename"], strg_filename=d["strg_filename"], id_filename=d.get("id_filename"),
        num_of_vertices=int(d["num_of_vertices"]), points_per_hour=int(d["points_per_hour"]),
        num_for_predict=int(d["num_for_predict"]), len_input=int(d["len_input"]), dataset_name=d["dataset_name"],
        in_channels=in_channels, conf_in_channels=conf_in, nb_block=int(t["nb_block"]), n_heads=int(t["n_heads"]),
        K=int(t["K"]), d_k=int(t["d_k"]), d_model=int(t["d_model"]), nb_chev_filter=int(t["nb_chev_filter"]),
        nb_time_filter=int(t["nb_time_filter"]), batch_size=int(t["batch_size"]),
        graph=t.get("graph", "AG"),                                   # D3: key absent in PEMS03/07/08 confs
        model_name=t["model_name"], num_of_weeks=int(t["num_of_weeks"]), num_of_days=int(t["num_of_days"]),
        num_of_hours=int(t["num_of_hours"]), start_epoch=int(t["start_epoch"]), epochs=int(t["epochs"]),
        learning_rate=float(t["learning_rate"]))


# Hyper-parameters of the five shipped configurations (configurations/{PEMS03,PEMS04,PEMS07,PEMS08,GAMBIA}_dstagnn.conf):
#   num_of_vertices, len_input, points_per_hour, period, in_channels, nb_block, n_heads, K, d_model, batch_size,
#   num_of_hours, has a `graph` key, epochs
SHIPPED = {
    "PEMS03": (358, 12, 12, 288, 1, 4, 3, 3, 512, 32, 1, False, 90),
    "PEMS04": (307, 12, 12, 288, 1, 4, 3, 3, 512, 32, 1, True, 110),
    "PEMS07": (883, 12, 12, 288, 1, 4, 4, 3, 512, 12, 1, False, 30),
    "PEMS08": (170, 12, 12, 288, 1, 4, 3, 3, 512, 64, 1, False, 100),
    "GAMBIA": (2139, 144, 12, 12, 4, 2, 2, 2, 64, 4, 12, True, 110),
}


def write_conf(path: str, dataset: str, data_dir: str = "./data") -> str:
    """Write a configuration file with the keys and values of the reference's ``configurations/<dataset>_dstagnn.conf``
    (the files themselves live in the reference checkout; `--config` takes those verbatim).  Used by ``--dataset``
    and by the tests, which have no reference checkout on the GPU box."""
    n, t, pph, period, fin, nb, h, k, dm, bs, hours, has_graph, epochs = SHIPPED[dataset]
    cp = configparser.ConfigParser()
    cp["Data"] = {"adj_filename": f"{data_dir}/{dataset}/{dataset}.csv",
                  "graph_signal_matrix_filename": f"{data_dir}/{dataset}/{dataset}.npz",
                  "stag_filename": f"{data_dir}/{dataset}/stag_001_{dataset}.csv",
                  "strg_filename": f"{data_dir}/{dataset}/strg_001_{dataset}.csv",
                  "num_of_vertices": str(n), "period": str(period), "points_per_hour": str(pph),
                  "num_for_predict": "12", "len_input": str(t), "dataset_name": dataset}
    if dataset == "PEMS03":                       # the only data set whose sensor ids are remapped
        cp["Data"]["id_filename"] = f"{data_dir}/{dataset}/{dataset}.txt"
    tr = {"ctx": "0", "in_channels": str(fin), "nb_block": str(nb), "n_heads": str(h), "K": str(k), "d_k": "32",
          "d_model": str(dm), "nb_chev_filter": "32", "nb_time_filter": "32", "batch_size": str(bs)}
    if has_graph:
        tr["graph"] = "AG"
    tr.update({"model_name": "dstagnn", "dataset_name": dataset, "num_of_weeks": "0", "num_of_days": "0",
               "num_of_hours": str(hours), "start_epoch": "0", "epochs": str(epochs), "learning_rate": "0.0001"})
    cp["Training"] = tr
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    with open(path, "w") as f:
        cp.write(f)
    return path


def seed_torch(seed: int = 1):
    """train_DSTAGNN_my.py:20-29."""
    random.seed(seed)
    os.environ["PYTHONHASHSEED"] = str(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def load_graphs(cfg: Config):
    """adj_mx / adj_TMD / adj_pa as FloatTensors (train_DSTAGNN_my.py:63-78)."""
    if cfg.dataset_name in PEMS:
        adj_mx = D.get_adjacency_matrix2(cfg.adj_filename, cfg.num_of_vertices, id_filename=cfg.id_filename)
    else:
        adj_mx = D.load_weighted_adjacency_matrix2(cfg.adj_filename, cfg.num_of_vertices)
    adj_TMD = D.load_weighted_adjacency_matrix(cfg.stag_filename, cfg.num_of_vertices)
    adj_pa = D.load_PA(cfg.strg_filename)
    return torch.FloatTensor(adj_mx), torch.FloatTensor(adj_TMD), torch.FloatTensor(adj_pa)


def build_model(cfg: Config, device, graphs=None):
    """``make_model`` with the call site's argument order (train_DSTAGNN_my.py:85-104)."""
    from .modules import make_model
    adj_mx, adj_TMD, adj_pa = graphs if graphs is not None else load_graphs(cfg)
    adj_merge = adj_mx if cfg.graph == "G" else adj_TMD
    return make_model(device, cfg.in_channels, cfg.nb_block, cfg.in_channels, cfg.K, cfg.nb_chev_filter,
                      cfg.nb_time_filter, 1, adj_merge, adj_pa, adj_TMD, cfg.num_for_predict, cfg.len_input,
                      cfg.num_of_vertices, cfg.d_model, cfg.d_k, cfg.d_k, cfg.n_heads)


def params_dir(cfg: Config, root: str = "myexperiments") -> str:
    """``myexperiments/<dataset>/<model>_{h}h{d}d{w}w_channel{F}_{lr}`` (train_DSTAGNN_my.py:118-126; F is
    the conf's in_channels, so checkpoints land where the reference would put them)."""
    folder = "{}_{}h{}d{}w_channel{}_{}".format(cfg.model_name, cfg.num_of_hours, cfg.num_of_days, cfg.num_of_weeks,
                                                 cfg.conf_in_channels, float(cfg.learning_rate))
    return os.path.join(root, cfg.dataset_name, folder)


def redirect_to_synthetic(cfg: Config, workdir: str, samples: int, rank: int = 0):
    """--synthetic: write graph CSVs and a signal npz of the conf's shapes under ``workdir`` (rank 0) and point
    the config at them.  The real PEMS / GAMBIA files are not shipped with the reference (README.md:24-25)."""
    base = os.path.join(workdir, cfg.dataset_name)
    cfg.adj_filename = os.path.join(base, f"{cfg.dataset_name}.csv")
    cfg.stag_filename = os.path.join(base, f"stag_001_{cfg.dataset_name}.csv")
    cfg.strg_filename = os.path.join(base, f"strg_001_{cfg.dataset_name}.csv")
    cfg.graph_signal_matrix_filename = os.path.join(base, f"{cfg.dataset_name}.npz")
    cfg.id_filename = None
    npz = D.dstagnn_npz_path(cfg.graph_signal_matrix_filename, cfg.num_of_hours, cfg.num_of_days, cfg.num_of_weeks)
    if rank == 0:
        if not os.path.exists(cfg.stag_filename):
            D.synthetic_graph_files(base, cfg.dataset_name, cfg.num_of_vertices)
            if cfg.dataset_name not in PEMS:          # dense 0/1 adjacency CSV for load_weighted_adjacency_matrix2
                tmd = D.load_weighted_adjacency_matrix(cfg.stag_filename)
                np.savetxt(cfg.adj_filename, tmd + np.identity(cfg.num_of_vertices), fmt="%g", delimiter=",")
        if not os.path.exists(npz):
            D.write_synthetic_dataset(cfg.graph_signal_matrix_filename, cfg.num_of_hours, cfg.num_of_days,
                                      cfg.num_of_weeks, cfg.num_of_vertices, cfg.len_input, cfg.num_for_predict,
                                      num_features=cfg.conf_in_channels, samples=samples)


def fit(cfg: Config, device, out_root: str = "myexperiments", epochs: Optional[int] = None,
        max_batches: Optional[int] = None, rank: int = 0, world: int = 1, log=print, precision: Optional[str] = None,
        cuda_graph: bool = True):
    """The loop of train_DSTAGNN_my.py:132-191.  Returns a summary dict (losses, best epoch, checkpoint path).

    The step itself -- zero-grad, forward, SmoothL1, backward, gradient averaging, Adam (lines 154-158) -- is
    ``step.TrainStep``: every kernel is this library's, and on one GPU the step is captured in a CUDA graph."""
    import torch.distributed as dist

    from .contraction import set_precision
    from .step import TrainStep, smooth_l1_loss
    if precision:
        set_precision(precision)
    seed_torch(1)
    (_, train_loader, _, _, val_loader, _, _, test_loader, test_target, mean, std) = D.load_graphdata_channel1(
        cfg.graph_signal_matrix_filename, cfg.num_of_hours, cfg.num_of_days, cfg.num_of_weeks, device, cfg.batch_size,
        rank=rank, world=world, in_channels=cfg.in_channels, verbose=rank == 0)
    if rank == 0:
        log(f"Train loader: {len(train_loader)} batches")
    net = build_model(cfg, device)
    path = params_dir(cfg, out_root)
    if rank == 0:
        os.makedirs(path, exist_ok=True)
        log(f"Params path: {path}")
    criterion = smooth_l1_loss
    train_step = TrainStep(net, cfg.learning_rate, graph=cuda_graph and torch.device(device).type == "cuda")

    def mean_loss(loader, limit):
        net.eval()
        tot = torch.zeros((), device=device)
        n = 0
        with torch.no_grad():
            for i, (x, y) in enumerate(loader):
                if limit is not None and i >= limit:
                    break
                tot += criterion(net(x), y)
                n += 1
        if world > 1:                                   # every rank holds a slice of each batch: average them
            dist.all_reduce(tot)
            tot /= world
        return float(tot) / max(n, 1)

    best_val, best_epoch, history = float("inf"), None, []
    n_epochs = cfg.epochs if epochs is None else min(cfg.epochs, cfg.start_epoch + epochs)
    for epoch in range(cfg.start_epoch, n_epochs):
        net.train()
        train_loader.set_epoch(epoch)
        t0 = time()
        total = torch.zeros((), device=device)
        nb = 0
        for batch_idx, (x, y) in enumerate(train_loader):
            if max_batches is not None and batch_idx >= max_batches:
                break
            loss = train_step(x, y)
            total += loss                               # no host sync per batch (the reference's loss.item(), line 160)
            nb += 1
            if batch_idx % 100 == 0 and rank == 0:
                log(f"Epoch {epoch} Batch {batch_idx} Loss: {float(loss):.4f}")
        val = mean_loss(val_loader, max_batches)
        train_loss = float(total) / max(nb, 1)
        history.append((epoch, train_loss, val))
        if rank == 0:
            log(f"Epoch {epoch} Train Loss: {train_loss:.4f} Val Loss: {val:.4f} Time: {time() - t0:.2f}s")
            if val < best_val:
                best_val, best_epoch = val, epoch
                torch.save(net.state_dict(), os.path.join(path, f"epoch_{epoch}.params"))
    summary = {"history": history, "best_epoch": best_epoch, "best_val": best_val, "params_path": path}
    if world > 1:
        obj = [summary if rank == 0 else None]
        dist.broadcast_object_list(obj, src=0)
        summary = obj[0]
        best_epoch = summary["best_epoch"]
    if best_epoch is not None:
        ckpt = os.path.join(path, f"epoch_{best_epoch}.params")
        if world > 1:
            dist.barrier()
        net.load_state_dict(torch.load(ckpt, map_location=device))
        summary["test_loss"] = mean_loss(test_loader, max_batches)
        summary["checkpoint"] = ckpt
        if rank == 0:
            log(f"Final Test Loss: {summary['test_loss']:.4f}")
            if world == 1 and max_batches is None:
                summary["metrics"] = predict_and_save_results(net, test_loader, test_target, best_epoch, mean, std,
                                                              path, "test")
    return summary


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("--config", default=None, type=str,
                    help="configuration file path (a reference configurations/*.conf, taken verbatim)")
    ap.add_argument("--dataset", default=None, choices=sorted(SHIPPED),
                    help="instead of --config: write and use a conf with the shipped hyper-parameters of this data set")
    ap.add_argument("--synthetic", action="store_true",
                    help="generate graph + signal files of the conf's shapes (the real data sets are not shipped)")
    ap.add_argument("--synthetic-samples", type=int, default=96)
    ap.add_argument("--workdir", default="synthetic_data", help="where --synthetic writes its files")
    ap.add_argument("--out", default="myexperiments", help="checkpoint root (train_DSTAGNN_my.py:126)")
    ap.add_argument("--epochs", type=int, default=None, help="run at most this many epochs (default: the conf's)")
    ap.add_argument("--max-batches", type=int, default=None, help="limit batches per epoch / evaluation (smoke runs)")
    ap.add_argument("--batch-size", type=int, default=None, help="override the conf's per-GPU batch size")
    ap.add_argument("--precision", default=None, choices=["fp32", "tf32x3", "tf32"])
    ap.add_argument("--no-cuda-graph", action="store_true", help="run every step eagerly (default: captured on 1 GPU)")
    args = ap.parse_args(argv)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("dstagnn_drought_b200.train needs a CUDA device (there is no CPU path)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        from .step import multi_gpu_graph_env
        multi_gpu_graph_env()                  # the captured step contains the NCCL all-reduces
        dist.init_process_group("nccl", device_id=device)
    if (args.config is None) == (args.dataset is None):
        raise SystemExit("give exactly one of --config FILE or --dataset NAME")
    conf_path = args.config
    if conf_path is None:
        conf_path = os.path.join(args.workdir, f"{args.dataset}_dstagnn.conf")
        if rank == 0:
            write_conf(conf_path, args.dataset)
        if world > 1:
            dist.barrier()
    cfg = read_config(conf_path)
    if args.batch_size:
        cfg.batch_size = args.batch_size
    if args.synthetic:
        redirect_to_synthetic(cfg, args.workdir, args.synthetic_samples, rank)
        if world > 1:
            dist.barrier()
    summary = fit(cfg, device, args.out, args.epochs, args.max_batches, rank, world, precision=args.precision,
                  cuda_graph=not args.no_cuda_graph)
    if world > 1:
        from .step import shutdown_process_group
        if rank == 0:
            print(json.dumps({k: v for k, v in summary.items() if isinstance(v, (int, float, str))}), flush=True)
        shutdown_process_group(graphs_captured=not args.no_cuda_graph)
    return summary


if __name__ == "__main__":
    main(sys.argv[1:])

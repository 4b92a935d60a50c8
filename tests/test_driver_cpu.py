"""Host side of the config-driven driver (dstagnn_drought_b200.train / data / metrics), no GPU:
conf parsing with the D3 / D4 rules, graph-file readers, the npz loader (layout, rank sharding, epoch
shuffling), the evaluation metrics, and -- through the numpy emulation of the C ABI -- the training loop itself
with its checkpoint, which must load into the UNMODIFIED reference model when that is staged (oracle/_ref)."""
import configparser
import os
import warnings

import numpy as np
import pytest
import torch

from dstagnn_drought_b200 import data as D
from dstagnn_drought_b200 import metrics as M
from dstagnn_drought_b200 import train as TR

REF_CONF = "/root/reference/configurations"


def test_written_confs_parse_with_reference_rules(tmp_path):
    for ds, row in TR.SHIPPED.items():
        path = TR.write_conf(str(tmp_path / f"{ds}_dstagnn.conf"), ds)
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            cfg = TR.read_config(path)
        assert cfg.graph == "AG"                                     # D3: default when the key is absent
        assert cfg.in_channels == 1                                  # D4: the only value the model can run
        assert cfg.conf_in_channels == row[4]
        assert (len(w) == 1) == (row[4] != 1)                        # ... and the GAMBIA conf warns about it
        assert (cfg.num_of_vertices, cfg.len_input, cfg.nb_block, cfg.n_heads, cfg.K, cfg.d_model, cfg.batch_size) == \
            (row[0], row[1], row[5], row[6], row[7], row[8], row[9])
        assert TR.params_dir(cfg).endswith(f"{ds}/dstagnn_{row[10]}h0d0w_channel{row[4]}_0.0001")


@pytest.mark.skipif(not os.path.isdir(REF_CONF), reason="reference checkout not present")
def test_written_confs_equal_the_shipped_ones():
    """Every hyper-parameter of write_conf() equals the reference's own configurations/*.conf."""
    import tempfile
    skip = ("adj_filename", "graph_signal_matrix_filename", "stag_filename", "strg_filename", "use_tpu")
    for ds in TR.SHIPPED:
        with tempfile.TemporaryDirectory() as td:
            a = configparser.ConfigParser(); a.read(TR.write_conf(os.path.join(td, "c.conf"), ds))
        b = configparser.ConfigParser(); b.read(os.path.join(REF_CONF, f"{ds}_dstagnn.conf"))
        for sec in ("Data", "Training"):
            for k, v in b[sec].items():
                if k not in skip:
                    assert a[sec].get(k) == v, (ds, sec, k)
            assert not (set(a[sec]) - set(b[sec])), ds
        cfg = TR.read_config(os.path.join(REF_CONF, f"{ds}_dstagnn.conf"))       # the shipped file, verbatim
        assert cfg.graph == "AG" and cfg.in_channels == 1


@pytest.mark.skipif(not os.path.isdir("/root/reference/data/PEMS04"), reason="reference checkout not present")
def test_graph_readers_match_reference_on_the_shipped_pems04_graph():
    import sys
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.ticker"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.path.insert(0, "/root/reference")
    try:
        from lib import dataloader as RD
        from lib.utils1 import get_adjacency_matrix2 as ref_adj2
    finally:
        sys.path.remove("/root/reference")
    base = "/root/reference/data/PEMS04"
    stag, strg, adj = (os.path.join(base, f) for f in ("stag_001_PEMS04.csv", "strg_001_PEMS04.csv", "PEMS04.csv"))
    assert np.array_equal(D.load_weighted_adjacency_matrix(stag, 307), RD.load_weighted_adjacency_matrix(stag, 307))
    assert np.array_equal(D.load_PA(strg), RD.load_PA(strg))
    assert np.array_equal(D.load_weighted_adjacency_matrix2(stag, 307), RD.load_weighted_adjacency_matrix2(stag, 307))
    assert np.array_equal(D.get_adjacency_matrix2(adj, 307), ref_adj2(adj, 307))
    for k in [m for m in list(sys.modules) if m == "lib" or m.startswith("lib.")]:
        del sys.modules[k]


def _dataset(tmp_path, n=13, N=6, F=2, T=5, P=3):
    path = D.write_synthetic_dataset(str(tmp_path / "X.npz"), 1, 0, 0, N, T, P, num_features=F, samples=n)
    return path, str(tmp_path / "X.npz")


def test_loader_layout_and_reference_tuple(tmp_path):
    path, sig = _dataset(tmp_path)
    assert path.endswith("X_r1_d0_w0_dstagnn.npz")                   # lib/utils1.py:295-297
    out = D.load_graphdata_channel1(sig, 1, 0, 0, "cpu", 4, verbose=False)
    assert len(out) == 11
    tx, tl, tt, vx, vl, vt, sx, sl, st, mean, std = out
    assert tx.shape[1:] == (6, 2, 5) and tt.shape[1:] == (6, 3) and tx.dtype == torch.float32
    assert tx.shape[0] + vx.shape[0] + sx.shape[0] == 13
    assert len(tl) == (tx.shape[0] + 3) // 4
    xs = torch.cat([x for x, _ in vl])
    assert torch.equal(xs, vx)                                       # unshuffled loaders keep the order
    seen = torch.cat([y for _, y in tl])
    assert sorted(seen.flatten().tolist()) == sorted(tt.flatten().tolist())     # shuffled: a permutation
    # in_channels = 1 on a multi-feature file keeps feature 0
    tl1 = D.load_graphdata_channel1(sig, 1, 0, 0, "cpu", 4, in_channels=1, verbose=False)[1]
    x1, _ = next(iter(tl1))
    assert x1.shape[1:] == (6, 1, 5)


def test_loader_shards_every_global_batch_across_ranks(tmp_path):
    _, sig = _dataset(tmp_path, n=40)
    world = 3
    loaders = [D.load_graphdata_channel1(sig, 1, 0, 0, "cpu", 2, rank=r, world=world, verbose=False)[1]
               for r in range(world)]
    full = D.load_graphdata_channel1(sig, 1, 0, 0, "cpu", 2 * world, verbose=False)[1]
    assert len({len(l) for l in loaders}) == 1 and len(loaders[0]) == len(full)
    for ep in range(2):
        for l in loaders + [full]:
            l.set_epoch(ep)
        per_rank = [list(l) for l in loaders]
        for b, (gx, gy) in enumerate(full):
            parts = [per_rank[r][b][1] for r in range(world) if b < len(per_rank[r])]
            assert torch.equal(torch.cat(parts), gy)                  # the ranks' slices tile the global batch
    first = [next(iter(full))[1].clone()]
    full.set_epoch(5)
    assert not torch.equal(first[0], next(iter(full))[1])             # a new permutation every epoch


def test_metrics_match_sklearn_and_reference_formula():
    rng = np.random.RandomState(0)
    y = rng.rand(7, 5, 4).astype(np.float32) * 10
    y[0, 0, :] = 0.0                                                  # masked entries (null value 0)
    p = y + rng.randn(7, 5, 4).astype(np.float32)
    from sklearn.metrics import mean_absolute_error, mean_squared_error
    assert abs(M.mean_absolute_error(y, p) - mean_absolute_error(y.reshape(-1, 1), p.reshape(-1, 1))) < 1e-6
    assert abs(M.root_mean_squared_error(y, p) - mean_squared_error(y.reshape(-1, 1), p.reshape(-1, 1)) ** 0.5) < 1e-6
    mask = y != 0
    want = np.mean(np.abs((p[mask] - y[mask]) / y[mask])) * 100
    assert abs(M.masked_mape_np(y, p, 0) - want) < 1e-3
    res = M.evaluate_predictions(y, p)
    assert len(res["per_step"]) == 4 and set(res["all"]) == {"mae", "rmse", "mape"}


def test_training_loop_end_to_end_on_the_cpu_emulation(tmp_path, monkeypatch):
    """conf -> synthetic files -> make_model -> 2 epochs x 2 batches -> validation -> epoch_{e}.params -> test pass,
    with the kernels replaced by tests/cpu_emulation.py; the checkpoint must load into the reference model."""
    import cpu_emulation
    cpu_emulation.install(monkeypatch)
    conf = configparser.ConfigParser()
    conf.read(TR.write_conf(str(tmp_path / "c.conf"), "PEMS08"))
    conf["Data"]["num_of_vertices"] = "9"
    conf["Training"].update({"nb_block": "2", "d_model": "16", "n_heads": "2", "batch_size": "3", "epochs": "2"})
    with open(tmp_path / "c.conf", "w") as f:
        conf.write(f)
    cfg = TR.read_config(str(tmp_path / "c.conf"))
    TR.redirect_to_synthetic(cfg, str(tmp_path / "syn"), samples=20)
    build = TR.build_model

    def build_without_dropout(*a, **k):                               # the emulation has no Philox dropout
        net = build(*a, **k)
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = 0.0
        return net
    monkeypatch.setattr(TR, "build_model", build_without_dropout)
    logs = []
    summary = TR.fit(cfg, "cpu", str(tmp_path / "exp"), max_batches=2, log=logs.append)
    assert len(summary["history"]) == 2 and all(np.isfinite(v) for _, tr, va in summary["history"] for v in (tr, va))
    assert summary["best_epoch"] in (0, 1) and os.path.exists(summary["checkpoint"])
    assert os.path.basename(summary["checkpoint"]) == f"epoch_{summary['best_epoch']}.params"
    assert np.isfinite(summary["test_loss"])
    sd = torch.load(summary["checkpoint"])
    from oracle.build_ref import import_reference_make_model
    make_ref = import_reference_make_model()
    if make_ref is not None:
        adj_mx, tmd, pa = TR.load_graphs(cfg)
        ref = make_ref("cpu", 1, 2, 1, cfg.K, 32, 32, 1, tmd, pa, tmd, 12, 12, 9, 16, 32, 32, 2)
        missing, unexpected = ref.load_state_dict(sd, strict=True)
        assert not missing and not unexpected

"""The config-driven driver end to end on the GPU: conf -> synthetic files -> device-resident loaders ->
training epochs -> validation -> checkpoint -> test pass -> metrics (train_DSTAGNN_my.py:36-191)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_driver_runs_a_shipped_configuration_on_synthetic_data(tmp_path):
    from dstagnn_drought_b200 import train as TR
    summary = TR.main(["--dataset", "PEMS08", "--synthetic", "--synthetic-samples", "200", "--epochs", "2",
                       "--workdir", str(tmp_path / "syn"), "--out", str(tmp_path / "exp"), "--batch-size", "16"])
    hist = summary["history"]
    assert len(hist) == 2 and all(np.isfinite(v) for _, tr, va in hist for v in (tr, va))
    assert hist[1][1] < hist[0][1]                                     # the training loss falls
    assert os.path.exists(summary["checkpoint"]) and np.isfinite(summary["test_loss"])
    m = summary["metrics"]["all"]
    assert all(np.isfinite(m[k]) for k in ("mae", "rmse", "mape"))
    assert os.path.exists(os.path.join(summary["params_path"], f"output_epoch_{summary['best_epoch']}_test.npz"))
    # the checkpoint is a plain state_dict with the reference's keys: it loads into the reference model
    sd = torch.load(summary["checkpoint"], map_location="cpu")
    from oracle.build_ref import import_reference_make_model
    make_ref = import_reference_make_model()
    if make_ref is not None:
        cfg = TR.read_config(str(tmp_path / "syn" / "PEMS08_dstagnn.conf"))
        TR.redirect_to_synthetic(cfg, str(tmp_path / "syn"), 200)
        adj_mx, tmd, pa = TR.load_graphs(cfg)
        ref = make_ref("cpu", 1, cfg.nb_block, 1, cfg.K, 32, 32, 1, tmd, pa, tmd, 12, 12, cfg.num_of_vertices,
                       cfg.d_model, 32, 32, cfg.n_heads)
        ref.load_state_dict(sd, strict=True)
        # ... and the reference, run on the CPU with those weights, reproduces the GPU predictions
        from dstagnn_drought_b200 import data as D
        loaders = D.load_graphdata_channel1(cfg.graph_signal_matrix_filename, 1, 0, 0, "cpu", 4, verbose=False)
        x = loaders[6][:4]
        ref.eval()
        with torch.no_grad():
            y_ref = ref(x)
        net = TR.build_model(cfg, torch.device("cuda:0"))
        net.load_state_dict(sd)
        net.eval()
        with torch.no_grad():
            y = net(x.cuda())
        assert float((y.cpu() - y_ref).abs().max() / y_ref.abs().max()) < 1e-4


def test_host_resident_loader_prefetches_on_a_side_stream(tmp_path):
    from dstagnn_drought_b200 import data as D
    sig = str(tmp_path / "X.npz")
    D.write_synthetic_dataset(sig, 1, 0, 0, 11, 12, 12, samples=60)
    a = D.load_graphdata_channel1(sig, 1, 0, 0, "cuda:0", 8, shuffle=False, resident=False, verbose=False)[1]
    b = D.load_graphdata_channel1(sig, 1, 0, 0, "cuda:0", 8, shuffle=False, resident=True, verbose=False)[1]
    assert not a.resident and b.resident and len(a) == len(b)
    for (xa, ya), (xb, yb) in zip(a, b):
        assert xa.is_cuda and torch.equal(xa, xb) and torch.equal(ya, yb)

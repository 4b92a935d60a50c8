"""Evaluation metrics of the reference (lib/metrics.py:4-16, lib/utils1.py:381-506), as plain numpy / torch
reductions -- no scikit-learn dependency on the training box."""
from __future__ import annotations

import os
from typing import Dict, List, Optional

import numpy as np
import torch


def masked_mape_np(y_true, y_pred, null_val=np.nan) -> float:
    """Mean absolute percentage error over the entries whose truth is not ``null_val`` (lib/metrics.py:4-16):
    the mask is renormalised by its mean, so the result is the mean over the unmasked entries, in percent."""
    y_true, y_pred = np.asarray(y_true), np.asarray(y_pred)
    with np.errstate(divide="ignore", invalid="ignore"):
        mask = ~np.isnan(y_true) if np.isnan(null_val) else np.not_equal(y_true, null_val)
        mask = mask.astype("float32")
        mask /= np.mean(mask)
        mape = np.abs(np.divide(np.subtract(y_pred, y_true).astype("float32"), y_true))
        mape = np.nan_to_num(mask * mape)
        return float(np.mean(mape) * 100)


def mean_absolute_error(y_true, y_pred) -> float:
    """sklearn.metrics.mean_absolute_error as called at lib/utils1.py:453,492: mean over all entries."""
    return float(np.mean(np.abs(np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64))))


def root_mean_squared_error(y_true, y_pred) -> float:
    """``mean_squared_error(...) ** 0.5`` of lib/utils1.py:454,493."""
    d = np.asarray(y_true, np.float64) - np.asarray(y_pred, np.float64)
    return float(np.sqrt(np.mean(d * d)))


def re_normalization(x, mean, std):
    """lib/utils1.py:17-19."""
    return x * std + mean


@torch.no_grad()
def predict(net, loader) -> np.ndarray:
    """All predictions of ``loader`` (in order), as one (S, N, P) array."""
    was_training = net.training
    net.train(False)                       # dropout off (lib/utils1.py:394,447)
    outs = [net(x).detach().cpu().numpy() for x, _ in loader]
    net.train(was_training)
    return np.concatenate(outs, 0)


def evaluate_predictions(target: np.ndarray, prediction: np.ndarray) -> Dict[str, object]:
    """Per-horizon and overall MAE / RMSE / masked MAPE (null value 0) exactly as
    ``predict_and_save_results_mstgcn`` prints them (lib/utils1.py:486-506)."""
    assert target.shape == prediction.shape, (target.shape, prediction.shape)
    per_step: List[Dict[str, float]] = []
    for i in range(prediction.shape[2]):
        per_step.append({"mae": mean_absolute_error(target[:, :, i], prediction[:, :, i]),
                         "rmse": root_mean_squared_error(target[:, :, i], prediction[:, :, i]),
                         "mape": masked_mape_np(target[:, :, i], prediction[:, :, i], 0)})
    overall = {"mae": mean_absolute_error(target.reshape(-1, 1), prediction.reshape(-1, 1)),
               "rmse": root_mean_squared_error(target.reshape(-1, 1), prediction.reshape(-1, 1)),
               "mape": masked_mape_np(target.reshape(-1, 1), prediction.reshape(-1, 1), 0)}
    return {"per_step": per_step, "all": overall}


def predict_and_save_results(net, loader, target_tensor, global_step, mean, std, params_path: Optional[str],
                             kind: str, rank: int = 0, verbose: bool = True) -> Dict[str, object]:
    """``predict_and_save_results_mstgcn`` (lib/utils1.py:434-506): run the loader, save
    ``output_epoch_{step}_{kind}.npz`` (input feature 0 de-normalised, prediction, target) and return the
    metric table.  Works on the loader of ONE rank (evaluation is not sharded: pass a world-1 loader)."""
    prediction = predict(net, loader)
    target = target_tensor.detach().cpu().numpy()[: prediction.shape[0]]
    res = evaluate_predictions(target, prediction)
    if params_path and rank == 0:
        xs = loader.x.detach().cpu().numpy()[: prediction.shape[0], :, 0:1]
        np.savez(os.path.join(params_path, f"output_epoch_{global_step}_{kind}"),
                 input=re_normalization(xs, np.asarray(mean)[..., 0:1, :], np.asarray(std)[..., 0:1, :]),
                 prediction=prediction, data_target_tensor=target)
    if verbose and rank == 0:
        a = res["all"]
        print(f"all MAE: {a['mae']:.2f}\nall RMSE: {a['rmse']:.2f}\nall MAPE: {a['mape']:.2f}")
    return res

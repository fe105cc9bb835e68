"""Drop-in for the scoring helpers of lidbox/util.py that sit directly behind the model:

    predict_with_model(model, ds, predict_fn=None)      # util.py:23-38  batched inference over {"id", "input"} batches
    merge_chunk_predictions(chunk_predictions)          # util.py:41-57  mean over the chunks of every parent utterance
    average_detection_cost(true_sparse, pred_dense, n)  # util.py:76-82  the C_avg part of classification_report

The model runs batch after batch without a host synchronisation (outputs stay on the device until the loop ends);
the chunk mean is one kernel over all groups (lbx_group_mean_f32).  sklearn reports / confusion matrices of the
reference's classification_report are host-side bookkeeping and not part of the path.
"""
import numpy as np
import pandas as pd
import torch

from . import _lib, metrics


def predictions_to_dataframe(ids, predictions):
    """util.py:17-20."""
    df = pd.DataFrame.from_dict({"id": ids, "prediction": predictions}).set_index("id", drop=True)
    if not df.index.is_unique:                      # the reference's verify_integrity=True
        raise ValueError("Index has duplicate keys: %s" % list(df.index[df.index.duplicated()].unique()))
    return df.sort_index()


def predict_with_model(model, ds, predict_fn=None):
    """Map the callable model over all batches of ds (an iterable of dicts with keys "id" and "input")."""
    if predict_fn is None:
        def predict_fn(x):
            return x["id"], model(x["input"], training=False)
    ids, outs = [], []
    for batch in ds:
        batch_ids, pred = predict_fn(batch)
        ids.extend(i.decode("utf-8") if isinstance(i, bytes) else str(i) for i in batch_ids)
        outs.append(pred if isinstance(pred, torch.Tensor) else torch.as_tensor(np.asarray(pred)))
    if not outs:
        return predictions_to_dataframe([], [])
    pred = torch.cat([o.reshape(o.shape[0], -1) for o in outs]).cpu().numpy()      # the only device->host copy
    return predictions_to_dataframe(ids, list(pred))


def chunk_parent_id(chunk_id):
    """util.py:41-42."""
    return chunk_id.rsplit('-', 1)[0]


def group_mean(predictions, groups):
    """Mean of the rows of predictions [R, D] within each group; groups: list of row-index lists. Returns [G, D]."""
    pred = predictions if isinstance(predictions, torch.Tensor) else torch.as_tensor(np.asarray(predictions))
    dev = _lib.require_cuda()
    pred = pred.to(dev, torch.float32).contiguous()
    if pred.dim() != 2:
        raise ValueError("predictions must be [rows, D]")
    offsets = np.zeros(len(groups) + 1, np.int64)
    offsets[1:] = np.cumsum([len(g) for g in groups])
    index = np.concatenate([np.asarray(g, np.int64) for g in groups]) if groups else np.zeros(0, np.int64)
    if index.size and (index.min() < 0 or index.max() >= pred.shape[0]):
        raise IndexError("row index out of range")
    out = torch.empty((len(groups), pred.shape[1]), dtype=torch.float32, device=dev)
    if len(groups):
        idx_d, off_d = torch.as_tensor(index).to(dev), torch.as_tensor(offsets).to(dev)
        _lib.check(_lib.lib().lbx_group_mean_f32(_lib.ptr(pred), _lib.ptr(idx_d), _lib.ptr(off_d), len(groups),
                                                 pred.shape[1], _lib.ptr(out), _lib.stream_ptr(dev)))
    return out


def merge_chunk_predictions(chunk_predictions, merge_rows_fn=None):
    """util.py:47-57: chunk_predictions is a DataFrame indexed by chunk id with a "prediction" column; rows whose ids
    share the parent id (everything before the last '-') are averaged.  A custom merge_rows_fn runs on the host as in
    the reference."""
    by_parent = {}
    for row, cid in enumerate(chunk_predictions.index):
        by_parent.setdefault(chunk_parent_id(cid), []).append(row)
    ids = sorted(by_parent)
    if merge_rows_fn is not None:
        values = chunk_predictions.prediction.values
        return predictions_to_dataframe(ids, [merge_rows_fn(values[by_parent[i]]) for i in ids])
    if not ids:
        return predictions_to_dataframe([], [])
    pred = np.stack(chunk_predictions.prediction.values).astype(np.float32)
    merged = group_mean(pred.reshape(pred.shape[0], -1), [by_parent[i] for i in ids]).cpu().numpy()
    return predictions_to_dataframe(ids, list(merged.reshape((len(ids),) + pred.shape[1:])))


def average_detection_cost(true_sparse, pred_dense, num_labels, num_cavg_thresholds=100):
    """util.py:76-82: thresholds = linspace(min score, max score, num_cavg_thresholds), SparseAverageDetectionCost."""
    pred = np.asarray(pred_dense.cpu() if isinstance(pred_dense, torch.Tensor) else pred_dense)
    thresholds = np.linspace(pred.min(), pred.max(), num_cavg_thresholds)
    cavg = metrics.SparseAverageDetectionCost(num_labels, thresholds)
    cavg.update_state(true_sparse, pred_dense)
    return float(cavg.result().cpu())

"""GPU parity tests (-m gpu) of the callers either side of the path: signal chunking
(bit-exact copies), chunk-score merging (fp32 mean, 1e-6) and C_avg (counters exact, cost 1e-6) against the oracle."""
import numpy as np
import pandas as pd
import pytest
import torch

from oracle import lidbox_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,sr,length_ms,step_ms,pad_ms", [
    (16000, 16000, 500, 250, 0), (16001, 16000, 500, 250, 0), (20000, 16000, 1000, 1000, 800),
    (20000, 16000, 1000, 1000, 700), (100, 16000, 500, 250, 0), (100, 16000, 500, 250, 500),
    (7999, 8000, 2000, 500, 10), (48000, 44100, 30, 10, 5), (160000, 16000, 2000, 1500, 1999)])
def test_signal_chunks_bit_exact(built_lib, n, sr, length_ms, step_ms, pad_ms):
    from lidbox_b200.data import steps
    rng = np.random.default_rng(n)
    sig = rng.standard_normal((3, n)).astype(np.float32)
    got = steps.signal_chunks(sig, sr, length_ms, step_ms, pad_ms).cpu().numpy()
    for b in range(3):
        want = O.create_signal_chunks(sig[b], sr, length_ms, step_ms, pad_ms)
        assert got[b].shape == want.shape
        np.testing.assert_array_equal(got[b], want)
    one = steps.signal_chunks(sig[0], sr, length_ms, step_ms, pad_ms).cpu().numpy()
    np.testing.assert_array_equal(one, got[0])


def test_create_signal_chunks_elements(built_lib):
    from lidbox_b200.data import steps
    rng = np.random.default_rng(1)
    elements = [dict(id="a", signal=rng.standard_normal(40000).astype(np.float32), sample_rate=16000, duration=2.5,
                     label="fi"),
                dict(id="b", signal=rng.standard_normal(9000).astype(np.float32), sample_rate=16000, label="sv")]
    out = list(steps.create_signal_chunks(elements, 1000, 500, max_pad_ms=600))
    want_a = O.create_signal_chunks(elements[0]["signal"], 16000, 1000, 500, 600)
    want_b = O.create_signal_chunks(elements[1]["signal"], 16000, 1000, 500, 600)
    assert [o["id"] for o in out] == ["a-%06d" % i for i in range(1, len(want_a) + 1)] + \
        ["b-%06d" % i for i in range(1, len(want_b) + 1)]
    assert len(want_b) == 1 and all(o["label"] == ("fi" if o["id"][0] == "a" else "sv") for o in out)
    for o, w in zip(out, list(want_a) + list(want_b)):
        np.testing.assert_array_equal(o["signal"].cpu().numpy(), w)
    assert out[0]["duration"] == 1.0 and "duration" not in out[-1]
    with pytest.raises(ValueError):
        list(steps.create_signal_chunks(elements[:1], 1000, 500, max_num_chunks_per_signal=4))


def test_chunks_feed_logmel_like_separate_utterances(built_lib):
    # chunk -> feature stage: the log-mel of the chunk tensor equals the log-mel of each chunk taken by itself
    from lidbox_b200.data import steps
    from lidbox_b200.features import audio
    sig = np.random.default_rng(2).standard_normal((2, 48000)).astype(np.float32) * 0.1
    chunks = steps.signal_chunks(sig, 16000, 2000, 500)
    B, C, L = chunks.shape
    feats = audio.logmelspectrograms(chunks.reshape(B * C, L), 16000).cpu().numpy()
    want = O.logmel(chunks.reshape(B * C, L).cpu().numpy()[:3].astype(np.float64), 16000)
    np.testing.assert_allclose(feats[:3], want, rtol=1e-3, atol=5e-3)


def test_merge_chunk_predictions(built_lib):
    from lidbox_b200 import util
    rng = np.random.default_rng(3)
    ids, rows = [], []
    for u in range(37):
        for c in range(1, int(rng.integers(1, 9)) + 1):
            ids.append("utt-%02d-%06d" % (u, c))
            rows.append(rng.standard_normal(50).astype(np.float32))
    perm = rng.permutation(len(ids))
    ids, rows = [ids[i] for i in perm], [rows[i] for i in perm]
    df = util.predictions_to_dataframe(ids, rows)
    merged = util.merge_chunk_predictions(df)
    parents, want = O.merge_chunk_predictions(list(df.index), list(df.prediction.values))
    assert list(merged.index) == parents
    np.testing.assert_allclose(np.stack(merged.prediction.values), want, rtol=1e-6, atol=1e-6)
    custom = util.merge_chunk_predictions(df, merge_rows_fn=lambda v: np.stack(v).max(axis=0))
    assert list(custom.index) == parents
    assert util.chunk_parent_id("a-b-000003") == "a-b"


def test_predict_with_model_batches(built_lib):
    from lidbox_b200 import util
    from lidbox_b200.models import xvector
    rng = np.random.default_rng(4)
    m = xvector.create((None, 40), 5)
    x = rng.standard_normal((10, 60, 40)).astype(np.float32)
    ds = [dict(id=[b"u%02d" % i for i in range(0, 4)], input=x[0:4]),
          dict(id=["u%02d" % i for i in range(4, 10)], input=x[4:10])]
    df = util.predict_with_model(m, ds)
    assert list(df.index) == ["u%02d" % i for i in range(10)]
    ref = O.xvector_forward({k: v.astype(np.float64) for k, v in m.get_weights().items()}, x.astype(np.float64))
    np.testing.assert_allclose(np.stack(df.prediction.values), ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("B,N,Th", [(8, 3, 4), (1000, 10, 100), (5000, 50, 100), (300, 2, 1), (70000, 7, 33),
                                    (513, 160, 200), (2000, 100, 150)])
def test_cavg_matches_oracle(built_lib, B, N, Th):
    from lidbox_b200 import metrics
    rng = np.random.default_rng(B + N)
    y = rng.integers(0, N, B)
    scores = rng.standard_normal((B, N)).astype(np.float32)
    scores[np.arange(B), y] += 1.5
    logp = scores - np.log(np.exp(scores).sum(axis=1, keepdims=True))
    thresholds = np.linspace(logp.min(), logp.max(), Th).astype(np.float32)
    onehot = np.eye(N, dtype=np.float32)[y]
    ref = O.AverageDetectionCost(N, thresholds, C_miss=1.0, C_fa=2.0, P_tar=0.3)
    ref.update_state(onehot, logp)
    dense = metrics.AverageDetectionCost(N, thresholds, C_miss=1.0, C_fa=2.0, P_tar=0.3)
    sparse = metrics.SparseAverageDetectionCost(N, thresholds, C_miss=1.0, C_fa=2.0, P_tar=0.3)
    assert float(dense.result()) == 0.0
    dense.update_state(onehot, logp)
    half = B // 2
    sparse.update_state(y[:half], logp[:half])            # two updates accumulate
    sparse.update_state(y[half:], logp[half:])
    for m in (dense, sparse):
        np.testing.assert_array_equal(m.tp.cpu().numpy(), ref.tp)          # trial counts: exact
        np.testing.assert_array_equal(m.fn.cpu().numpy(), ref.fn)
        np.testing.assert_array_equal(m.fp_pairs.cpu().numpy(), ref.fp_pairs)
        np.testing.assert_array_equal(m.tn_pairs.cpu().numpy(), ref.tn_pairs)
        per_t = m.result_per_threshold().cpu().numpy()
        np.testing.assert_allclose(per_t[:Th], ref.result_per_threshold(), rtol=2e-6, atol=1e-7)
        assert per_t[Th] == per_t[:Th].min()
        assert abs(float(m.result()) - float(ref.result())) <= 2e-6 * float(ref.result()) + 1e-7
    dense.reset_states()
    assert float(dense.result()) == 0.0 and not dense.fp_pairs.any()


def test_cavg_reference_selftest_inputs_and_edge_cases(built_lib):
    from lidbox_b200 import metrics, util
    from test_oracle import _CAVG_PROB, _CAVG_TRUE
    with np.errstate(divide="ignore"):
        pred = np.log(_CAVG_PROB)                            # contains -inf scores, as the reference self-test does
    thresholds = np.log(np.array([0.05, 0.4, 0.6, 0.95], np.float32))
    ref = O.AverageDetectionCost(3, thresholds)
    ref.update_state(_CAVG_TRUE, pred)
    m = metrics.AverageDetectionCost(3, thresholds)
    m.update_state(_CAVG_TRUE, pred)
    assert abs(float(m.result()) - float(ref.result())) < 1e-7
    # NaN scores count neither as accepted nor as rejected; out-of-range sparse labels act as tf.one_hot's zero row
    pred2 = pred.copy()
    pred2[0, 1] = np.nan
    y = _CAVG_TRUE.argmax(axis=1)
    y[3] = 7
    ref = O.SparseAverageDetectionCost(3, thresholds)
    ref.update_state(y, pred2)
    m = metrics.SparseAverageDetectionCost(3, thresholds)
    m.update_state(y, pred2)
    np.testing.assert_array_equal(m.fp_pairs.cpu().numpy(), ref.fp_pairs)
    np.testing.assert_array_equal(m.tn_pairs.cpu().numpy(), ref.tn_pairs)
    np.testing.assert_array_equal(m.tp.cpu().numpy(), ref.tp)
    assert abs(float(m.result()) - float(ref.result())) < 1e-7
    with pytest.raises(ValueError):
        metrics.AverageDetectionCost(1, thresholds)
    with pytest.raises(ValueError):
        metrics.AverageDetectionCost(3, [[0.1]])
    # util.py:76-82 call site
    rng = np.random.default_rng(9)
    yy = rng.integers(0, 4, 200)
    sc = rng.standard_normal((200, 4)).astype(np.float32)
    sc[np.arange(200), yy] += 2
    ref = O.SparseAverageDetectionCost(4, np.linspace(sc.min(), sc.max(), 100))
    ref.update_state(yy, sc)
    assert abs(util.average_detection_cost(yy, sc, 4) - float(ref.result())) < 1e-6


def test_cavg_golden(built_lib):
    import os
    from lidbox_b200 import metrics
    c = np.load(os.path.join(os.path.dirname(__file__), "golden", "cavg.npz"))
    m = metrics.AverageDetectionCost(3, c["thr"])
    m.update_state(c["onehot"], c["pred"])
    for name in ("tp", "fn", "fp_pairs", "tn_pairs"):
        np.testing.assert_array_equal(getattr(m, name).cpu().numpy(), c[name])
    np.testing.assert_allclose(m.result_per_threshold().cpu().numpy()[:4], c["cavg"], rtol=2e-6, atol=1e-7)
    m6 = metrics.SparseAverageDetectionCost(6, c["thr6"], C_miss=1.0, C_fa=2.0, P_tar=0.3)
    m6.update_state(c["y6"], c["s6"])
    np.testing.assert_array_equal(m6.fp_pairs.cpu().numpy(), c["fp_pairs6"])
    np.testing.assert_allclose(m6.result_per_threshold().cpu().numpy()[:25], c["cavg6"], rtol=2e-6, atol=1e-7)

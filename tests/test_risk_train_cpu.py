"""Risk-network training (train_risk_network.py): batch composition, metrics, a run end to end on a
small synthetic data set in the reference's directory layout, and the exported weights against a NumPy forward pass."""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import mlp  # noqa: E402
from safemotionsrisk_b200 import risk_data, risk_train  # noqa: E402


def _data(n, rng, d=23, a=7):
    x = rng.uniform(-1, 1, (n, d + a)).astype(np.float32)
    y = ((x[:, 0] + 0.5 * x[:, d] - 0.3 * x[:, 3] * x[:, 5]) > 0.55).astype(np.float32)   # ~ 20 % risky
    return x, y


def test_batches_follow_the_reference_generator():
    rng = np.random.default_rng(0)
    x, y = _data(1050, rng)
    b = risk_train.RiskBatches(x, y, 100)                         # no shuffle: consecutive batches, remainder dropped
    assert len(b) == 10 and b.x.shape[0] == 1000
    assert np.array_equal(b.batch_indices(3), np.arange(300, 400))
    b = risk_train.RiskBatches(x, y, 5000)                        # batch bigger than the data set
    assert len(b) == 1 and b.batch_size == 1050
    b = risk_train.RiskBatches(x, y, 100, shuffle=True, rng=np.random.default_rng(1))
    i0, i1 = b.batch_indices(0), b.batch_indices(0)
    assert len(set(i0)) == 100 and not np.array_equal(i0, i1)     # a fresh random subset of the whole set every time
    b = risk_train.RiskBatches(x, y, 100, risky_state_rebalancing_fraction=0.37, shuffle=True, rng=np.random.default_rng(2))
    idx = b.batch_indices(0)
    assert (y[idx[:63]] == 0).all() and (y[idx[63:]] == 1).all()  # int(0.37 * 100) risky rows at the end of the batch
    with pytest.raises(ValueError, match="requires shuffle"):
        risk_train.RiskBatches(x, y, 100, risky_state_rebalancing_fraction=0.5)
    with pytest.raises(ValueError, match="Not enough risky datapoints"):
        risk_train.RiskBatches(x, y, 1000, risky_state_rebalancing_fraction=0.9, shuffle=True)
    with pytest.raises(ValueError, match="Not enough unrisky datapoints"):
        risk_train.RiskBatches(x[y == 1], y[y == 1], 10, risky_state_rebalancing_fraction=0.5, shuffle=True)


def test_metrics_by_hand():
    p = np.array([0.9, 0.8, 0.3, 0.04, 0.6, 0.02])
    y = np.array([1, 0, 1, 0, 1, 0])
    m = risk_train.binary_metrics(p, y)
    assert m["tp"] == 2 and m["fp"] == 1 and m["fn"] == 1 and m["tn"] == 2 and abs(m["accuracy"] - 4 / 6) < 1e-12
    assert abs(m["precision_0.5"] - 2 / 3) < 1e-12 and abs(m["recall_0.5"] - 2 / 3) < 1e-12
    assert m["recall_0.03"] == 1.0 and abs(m["precision_0.03"] - 3 / 5) < 1e-12
    assert 0.7 < m["auc"] <= 1.0 and 0.0 < m["prc"] <= 1.0
    assert risk_train.binary_metrics(y.astype(float), y)["auc"] > 0.99          # a perfect ranking


def test_training_run_end_to_end(tmp_path):
    rng = np.random.default_rng(3)
    root = tmp_path / "risk_data"
    for split, n in (("train", 3000), ("test", 600)):
        x, y = _data(n, rng)
        for k in range(2):   # two files per split, like two worker processes
            sl = slice(k * n // 2, (k + 1) * n // 2)
            risk_data.write_risk_csv(str(root), x[sl, :23], y[sl], action=x[sl, 23:], first_episode=k * 1000, pid=100 + k,
                                     risk_config={"observation_size": 23, "risk_check_next_state": 2})
        os.rename(root / "state_action_risk", root / split)
    data_dir = tmp_path / "set" / "data"
    os.makedirs(data_dir.parent)
    os.rename(root, data_dir)
    os.rename(data_dir / "risk_config.json", data_dir.parent / "risk_config.json")   # next to the data dir
    args = risk_train.add_arguments(__import__("argparse").ArgumentParser()).parse_args([
        "--risk_data_dir", str(data_dir), "--experiment_name", "unit", "--hidden_layer_activation", "selu",
        "--fcnet_hiddens", "[64, 32]", "--batch_size", "250", "--shuffle", "--risky_state_rebalancing_fraction", "0.4",
        "--epochs", "12", "--lr", "0.003", "--logdir", str(tmp_path / "results"), "--seed", "5", "--device", "cpu"])
    log_dir = risk_train.run(args, log=lambda *_: None)
    assert os.path.relpath(log_dir, tmp_path / "results").split(os.sep)[:2] == ["state_action_risk", "unit"]
    cfg = json.load(open(os.path.join(log_dir, "risk_config.json")))
    assert cfg["observation_size"] == 23 and cfg["action_size"] == 7 and cfg["risk_check_next_state"] == 2
    assert json.load(open(os.path.join(log_dir, "arguments.json")))["fcnet_hiddens"] == [64, 32]
    hist = json.load(open(os.path.join(log_dir, "history.json")))
    assert len(hist) == 12 and hist[-1]["loss"] < hist[0]["loss"]
    assert hist[-1]["val_accuracy"] > 0.9 and hist[-1]["val_auc"] > 0.95       # the synthetic rule is learnable
    # the exported weights reproduce the trained model in a NumPy float32 forward pass (the oracle's MLP)
    w = np.load(os.path.join(log_dir, "risk_network.npz"))
    assert str(w["risk/hidden_layer_activation"]) == "selu"
    xv, yv, ss, asz = risk_train.load_risk_dir(str(data_dir / "test"))
    assert ss == 23 and asz == 7
    h = xv
    for i in range(2):
        h = mlp.selu(h @ w["risk/dense_{}/kernel".format(i)] + w["risk/dense_{}/bias".format(i)])
    p = 1.0 / (1.0 + np.exp(-(h @ w["risk/dense_2/kernel"] + w["risk/dense_2/bias"])))
    assert abs(float(((p.reshape(-1) > 0.5) == (yv > 0.5)).mean()) - hist[-1]["val_accuracy"]) < 1e-3
    with pytest.raises(ValueError, match="does not match"):
        json.dump({"observation_size": 99}, open(data_dir.parent / "risk_config.json", "w"))
        risk_train.run(args, log=lambda *_: None)

"""Training of the risk network on the GPU box (the reference's train_risk_network.py): a binary
classifier risk(state[, action]) in [0, 1] on the data set the env generates (`SafeMotionsVecEnv.risk_ground_truth`,
wire format in risk_data.py).  PyTorch does the optimisation (plumbing: Adam, autograd); what this module mirrors is
the reference's interface and semantics:

* data set layout <risk_data_dir>/{train,test}/*.csv, `risk_config.json` next to it (train_risk_network.py:187-236);
* batch composition (RiskDataGenerator, train_risk_network.py:21-149): without --shuffle consecutive batches of the
  truncated data set; with --shuffle every batch is a fresh random subset of the whole set; with
  --risky_state_rebalancing_fraction f every batch holds int(f * batch_size) risky rows and the rest unrisky ones
  (requires --shuffle; not enough rows of a class raise ValueError with the reference's wording); a batch size above
  the data set shrinks to it;
* model (train_risk_network.py:262-273): Dense(h_i, activation) [+ Dropout] ..., Dense(1, sigmoid | linear), optional
  l2 kernel / bias regularisers, max-norm kernel constraint on the hidden layers, class weight of the risky class;
* loss binary cross-entropy, Adam(lr) (train_risk_network.py:275-277); metrics accuracy, precision / recall at the
  reference's thresholds, tp / fp / tn / fn, ROC-AUC and PR-AUC (Keras-style, 200 thresholds);
* output directory <logdir>/<state_action_risk|state_risk>/<experiment_name>/<timestamp>/ with `arguments.json`, the
  updated `risk_config.json` (observation_size / action_size), `history.json` and `risk_network.npz` -- the weights in
  the layout `SafeMotionsVecEnv.load_networks` reads (risk/dense_<i>/kernel [in, out], bias), so that the directory can
  be handed to the env as `risk_config_dir` (the reference stores a Keras SavedModel there).

The device kernels run one to three selu or swish hidden layers of width 16, 32, 48, 64, 128, 192, 256 or 512, the last
one at most 256 (smenv_mlp_load); other shapes and activations train here but cannot be loaded into the step loop.

CLI:  python -m safemotionsrisk_b200.risk_train --risk_data_dir DIR --experiment_name NAME [the reference's flags]
"""
import argparse
import datetime
import glob
import json
import os
import shutil

import numpy as np

from . import risk_data

PR_THRESHOLDS = (0.01, 0.02, 0.03, 0.04, 0.05, 0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)
ACTIVATIONS = ("relu", "selu", "tanh", "sigmoid", "elu", "gelu", "swish", "leaky_relu")


class RiskBatches:
    """The batches of one epoch, composed like the reference's RiskDataGenerator (train_risk_network.py:21-149)."""

    def __init__(self, x, y, batch_size, risky_state_rebalancing_fraction=None, shuffle=False, rng=None):
        if risky_state_rebalancing_fraction is not None and not shuffle:
            raise ValueError("risky_state_rebalancing_fraction requires shuffle to be True")
        x, y = np.asarray(x, dtype=np.float32), np.asarray(y, dtype=np.float32).reshape(-1)
        if x.shape[0] == 0:
            raise FileNotFoundError("Could not find valid data.")
        self.batch_size = int(batch_size)
        self.len = x.shape[0] // self.batch_size
        if self.len == 0:   # batch size bigger than the data set
            self.batch_size, self.len = x.shape[0], 1
        if not shuffle:     # entries that do not fill a batch are dropped
            x, y = x[:self.len * self.batch_size], y[:self.len * self.batch_size]
        self.x, self.y, self.shuffle = x, y, shuffle
        self.rng = rng if rng is not None else np.random.default_rng()
        self.fraction = risky_state_rebalancing_fraction
        if self.fraction is not None:
            self.n_risky = int(self.fraction * self.batch_size)
            self.n_safe = self.batch_size - self.n_risky
            self.idx_risky, self.idx_safe = np.where(y == 1.0)[0], np.where(y != 1.0)[0]
            if len(self.idx_risky) < self.n_risky:
                raise ValueError("Not enough risky datapoints for the selected batch_size and rebalancing_fraction "
                                 "(required {}, available {}.)".format(self.n_risky, len(self.idx_risky)))
            if len(self.idx_safe) < self.n_safe:
                raise ValueError("Not enough unrisky datapoints for the selected batch_size and rebalancing_fraction "
                                 "(required {}, available {}).".format(self.n_safe, len(self.idx_safe)))

    def __len__(self):
        return self.len

    def batch_indices(self, index):
        if not self.shuffle:
            return np.arange(index * self.batch_size, (index + 1) * self.batch_size)
        if self.fraction is not None:   # unrisky rows first, then the risky ones (the reference's concatenation order)
            return np.concatenate([np.sort(self.rng.choice(self.idx_safe, self.n_safe, replace=False)),
                                   np.sort(self.rng.choice(self.idx_risky, self.n_risky, replace=False))])
        return np.sort(self.rng.choice(self.x.shape[0], self.batch_size, replace=False))

    def __iter__(self):
        for i in range(self.len):
            idx = self.batch_indices(i)
            yield self.x[idx], self.y[idx]


def load_risk_dir(directory):
    """All CSV files of <directory> in name order -> (x [N, state (+ action)], y [N], state_size, action_size | None)."""
    xs, ys, state_size, action_size = [], [], None, None
    for path in sorted(glob.glob(os.path.join(directory, "*.csv"))):
        state, action, risk = risk_data.read_risk_csv(path)
        state_size = state.shape[1]
        if action is not None:
            action_size = action.shape[1]
            state = np.concatenate([state, action], axis=1)
        xs.append(state)
        ys.append(risk)
    if not xs:
        raise FileNotFoundError("Could not find valid data files in {}.".format(directory))
    return np.concatenate(xs).astype(np.float32), np.concatenate(ys).astype(np.float32), state_size, action_size


def build_model(n_in, fcnet_hiddens, hidden_layer_activation="relu", last_layer_activation="sigmoid", dropout=None):
    import torch
    acts = {"relu": torch.nn.ReLU, "selu": torch.nn.SELU, "tanh": torch.nn.Tanh, "sigmoid": torch.nn.Sigmoid,
            "elu": torch.nn.ELU, "gelu": torch.nn.GELU, "swish": torch.nn.SiLU,
            "leaky_relu": lambda: torch.nn.LeakyReLU(0.3)}   # Keras' default slope
    if hidden_layer_activation not in acts:
        raise ValueError("hidden_layer_activation must be one of {}".format(ACTIVATIONS))
    if last_layer_activation not in ("linear", "sigmoid"):
        raise ValueError("last_layer_activation must be linear or sigmoid")
    layers, width = [], n_in
    for h in fcnet_hiddens:
        lin = torch.nn.Linear(width, int(h))
        torch.nn.init.xavier_uniform_(lin.weight)   # Keras' glorot_uniform
        torch.nn.init.zeros_(lin.bias)
        layers += [lin, acts[hidden_layer_activation]()]
        if dropout is not None:
            layers.append(torch.nn.Dropout(float(dropout)))
        width = int(h)
    out = torch.nn.Linear(width, 1)
    torch.nn.init.xavier_uniform_(out.weight)
    torch.nn.init.zeros_(out.bias)
    layers.append(out)
    if last_layer_activation == "sigmoid":
        layers.append(torch.nn.Sigmoid())
    return torch.nn.Sequential(*layers)


def binary_metrics(p, y):
    """The reference's metric set (train_risk_network.py:278-317) from predictions p and labels y (NumPy)."""
    p, y = np.asarray(p, dtype=np.float64).reshape(-1), np.asarray(y, dtype=np.float64).reshape(-1) > 0.5
    out = {"accuracy": float(((p > 0.5) == y).mean())}
    for t in PR_THRESHOLDS:
        pred = p > t
        tp, fp, fn = float((pred & y).sum()), float((pred & ~y).sum()), float((~pred & y).sum())
        out["precision_{}".format(t)] = tp / (tp + fp) if tp + fp > 0 else 0.0
        out["recall_{}".format(t)] = tp / (tp + fn) if tp + fn > 0 else 0.0
    pred = p > 0.5
    out.update(tp=float((pred & y).sum()), fp=float((pred & ~y).sum()), tn=float((~pred & ~y).sum()),
               fn=float((~pred & y).sum()))
    # Keras AUC: 200 thresholds, trapezoidal ROC, "careful interpolation" PR replaced by the trapezoid (summary metric)
    ths = np.concatenate([[-1e-7], (np.arange(198) + 1.0) / 199.0, [1.0 + 1e-7]])
    tpr, fpr, prec = [], [], []
    npos, nneg = max(float(y.sum()), 1.0), max(float((~y).sum()), 1.0)
    for t in ths:
        pr = p > t
        tp, fp = float((pr & y).sum()), float((pr & ~y).sum())
        tpr.append(tp / npos); fpr.append(fp / nneg); prec.append(tp / (tp + fp) if tp + fp > 0 else 1.0)
    tpr, fpr, prec = np.array(tpr), np.array(fpr), np.array(prec)
    out["auc"] = float(-np.trapezoid(tpr, fpr))
    out["prc"] = float(-np.trapezoid(prec, tpr))
    return out


def export_weights(model, path, hidden_layer_activation="selu", last_layer_activation="sigmoid"):
    """risk/dense_<i>/{kernel [in, out], bias}: the layout of assets/networks_<scene>.npz (tools/export_networks.py),
    plus the two activation names."""
    import torch
    arrays, i = {"risk/hidden_layer_activation": np.array(hidden_layer_activation),
                 "risk/last_layer_activation": np.array(last_layer_activation)}, 0
    for m in model:
        if isinstance(m, torch.nn.Linear):
            arrays["risk/dense_{}/kernel".format(i)] = m.weight.detach().cpu().numpy().T.astype(np.float32).copy()
            arrays["risk/dense_{}/bias".format(i)] = m.bias.detach().cpu().numpy().astype(np.float32).copy()
            i += 1
    np.savez(path, **arrays)
    return arrays


def train(x_train, y_train, x_test=None, y_test=None, fcnet_hiddens=(128, 256, 256), hidden_layer_activation="relu",
          last_layer_activation="sigmoid", batch_size=1000, shuffle=False, risky_state_rebalancing_fraction=None,
          risky_state_class_weight=None, epochs=100, lr=0.03, dropout=None, kernel_constraint=None,
          kernel_regularizer=None, bias_regularizer=None, seed=None, device=None, log=None):
    """Trains the classifier; returns (model, history) with history[epoch] = {"loss", train metrics, "val_*"}."""
    import torch
    if seed is not None:
        torch.manual_seed(int(seed))
    rng = np.random.default_rng(seed)
    dev = torch.device(device) if device is not None else torch.device("cuda" if torch.cuda.is_available() else "cpu")
    batches = RiskBatches(x_train, y_train, batch_size, risky_state_rebalancing_fraction, shuffle, rng)
    model = build_model(batches.x.shape[1], fcnet_hiddens, hidden_layer_activation, last_layer_activation, dropout).to(dev)
    opt = torch.optim.Adam(model.parameters(), lr=float(lr), eps=1e-7)   # Keras' epsilon
    xt = torch.as_tensor(batches.x, device=dev)
    yt = torch.as_tensor(batches.y, device=dev)
    linears = [m for m in model if isinstance(m, torch.nn.Linear)]
    history = []
    for epoch in range(int(epochs)):
        model.train()
        total, seen = 0.0, 0
        for i in range(len(batches)):
            idx = torch.as_tensor(batches.batch_indices(i), device=dev)
            xb, yb = xt[idx], yt[idx]
            out = model(xb).reshape(-1)
            # Keras' binary_crossentropy clips the model output to [1e-7, 1 - 1e-7] (a linear head is treated as a
            # probability as well)
            per = torch.nn.functional.binary_cross_entropy(out.clamp(1e-7, 1 - 1e-7), yb, reduction="none")
            if risky_state_class_weight is not None:
                per = per * torch.where(yb > 0.5, torch.full_like(yb, float(risky_state_class_weight)), torch.ones_like(yb))
            loss = per.mean()
            if kernel_regularizer is not None:
                loss = loss + float(kernel_regularizer) * sum((m.weight ** 2).sum() for m in linears)
            if bias_regularizer is not None:
                loss = loss + float(bias_regularizer) * sum((m.bias ** 2).sum() for m in linears)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
            if kernel_constraint is not None:   # MaxNorm on the hidden layers: columns of the Keras kernel = rows here
                with torch.no_grad():
                    for m in linears[:-1]:
                        nrm = m.weight.norm(dim=1, keepdim=True)
                        m.weight.mul_(torch.clamp(nrm, max=float(kernel_constraint)) / (nrm + 1e-7))
            total += float(loss.item()) * xb.shape[0]
            seen += xb.shape[0]
        model.eval()
        with torch.no_grad():
            rec = {"epoch": epoch, "loss": total / max(seen, 1)}
            rec.update(binary_metrics(model(xt).reshape(-1).cpu().numpy(), batches.y))
            if x_test is not None and len(x_test):
                pv = model(torch.as_tensor(np.asarray(x_test, dtype=np.float32), device=dev)).reshape(-1).cpu().numpy()
                rec.update({"val_" + k: v for k, v in binary_metrics(pv, y_test).items()})
        history.append(rec)
        if log:
            log("epoch {:3d}  loss {:.4f}  acc {:.4f}  auc {:.4f}{}".format(
                epoch, rec["loss"], rec["accuracy"], rec["auc"],
                "  val_acc {:.4f}  val_auc {:.4f}".format(rec["val_accuracy"], rec["val_auc"]) if "val_auc" in rec else ""))
    return model, history


def add_arguments(parser):
    """The flags of train_risk_network.py:153-175."""
    parser.add_argument("--risk_data_dir", type=str, required=True, default=None)
    parser.add_argument("--experiment_name", type=str, required=True, default=None)
    parser.add_argument("--hidden_layer_activation", default="relu", choices=list(ACTIVATIONS))
    parser.add_argument("--last_layer_activation", default="sigmoid", choices=["linear", "sigmoid"])
    parser.add_argument("--fcnet_hiddens", type=json.loads, default=[128, 256, 256])
    parser.add_argument("--precision_threshold", type=float, default=0.5)
    parser.add_argument("--recall_threshold", type=float, default=0.5)
    parser.add_argument("--risky_state_class_weight", type=float, default=None)
    parser.add_argument("--risky_state_rebalancing_fraction", type=float, default=None)
    parser.add_argument("--batch_size", type=int, default=1000)
    parser.add_argument("--shuffle", action="store_true", default=False)
    parser.add_argument("--epochs", type=int, default=100)
    parser.add_argument("--epochs_per_checkpoint", type=int, default=None)
    parser.add_argument("--lr", type=float, default=0.03)
    parser.add_argument("--dropout", type=float, default=None)
    parser.add_argument("--kernel_constraint", type=float, default=None)
    parser.add_argument("--kernel_regularizer", type=float, default=None)
    parser.add_argument("--bias_regularizer", type=float, default=None)
    parser.add_argument("--logdir", type=str, default=None)
    parser.add_argument("--seed", type=int, default=None)
    parser.add_argument("--device", type=str, default=None, help="torch device (default: cuda if available)")
    return parser


def run(args, log=print):
    """One training run with the directory conventions of train_risk_network.py:177-236; returns the log directory."""
    x, y, state_size, action_size = load_risk_dir(os.path.join(args.risk_data_dir, "train"))
    try:
        xv, yv, _, _ = load_risk_dir(os.path.join(args.risk_data_dir, "test"))
    except FileNotFoundError:
        if os.path.isdir(os.path.join(args.risk_data_dir, "test")):
            raise
        xv, yv = None, None
    log_dir = os.path.join(os.path.expanduser("~"), "risk_results") if args.logdir is None else args.logdir
    log_dir = os.path.join(log_dir, "state_action_risk" if action_size is not None else "state_risk", args.experiment_name,
                           datetime.datetime.now().strftime("%Y%m%dT%H%M%S"))
    os.makedirs(log_dir, exist_ok=True)
    cfg_src = os.path.join(os.path.dirname(os.path.normpath(args.risk_data_dir)), "risk_config.json")
    if os.path.isfile(cfg_src):
        dst = os.path.join(log_dir, "risk_config.json")
        shutil.copy(cfg_src, dst)
        with open(dst) as f:
            config = json.load(f)
        if "observation_size" in config:
            if state_size != config["observation_size"]:
                raise ValueError("The observation size of the risk data ({}) does not match with the observations size "
                                 "specified in risk_config.json ({}).".format(state_size, config["observation_size"]))
        else:
            config["observation_size"] = state_size
        config["action_size"] = action_size
        with open(dst, "w") as f:
            f.write(json.dumps(config, sort_keys=True))
    with open(os.path.join(log_dir, "arguments.json"), "w") as f:
        f.write(json.dumps(vars(args), sort_keys=True))
    model, history = train(x, y, xv, yv, fcnet_hiddens=args.fcnet_hiddens, hidden_layer_activation=args.hidden_layer_activation,
                           last_layer_activation=args.last_layer_activation, batch_size=args.batch_size,
                           shuffle=args.shuffle, risky_state_rebalancing_fraction=args.risky_state_rebalancing_fraction,
                           risky_state_class_weight=args.risky_state_class_weight, epochs=args.epochs, lr=args.lr,
                           dropout=args.dropout, kernel_constraint=args.kernel_constraint,
                           kernel_regularizer=args.kernel_regularizer, bias_regularizer=args.bias_regularizer,
                           seed=args.seed, device=getattr(args, "device", None), log=log)
    export_weights(model, os.path.join(log_dir, "risk_network.npz"), args.hidden_layer_activation, args.last_layer_activation)
    with open(os.path.join(log_dir, "history.json"), "w") as f:
        f.write(json.dumps(history))
    return log_dir


def main(argv=None):
    args = add_arguments(argparse.ArgumentParser(description=__doc__.split("\n")[0])).parse_args(argv)
    print(run(args))
    return 0


if __name__ == "__main__":
    raise SystemExit(main())

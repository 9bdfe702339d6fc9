"""Wire format of the risk data set: the CSV files the reference's env writes while it generates risk ground truth
(safe_motions_base.py:1413-1461) and train_risk_network.py reads back (train_risk_network.py:43-55):
a pandas DataFrame dump with an unnamed index column and the columns `state` (Python list literal of the risk
observation), `action` (list literal, state-action risk only) and `risk` (0.0 / 1.0), file name
episodes_<first>_to_<last>_risk_<mean>_prediction_<mean>_pid_<pid>.csv in <dir>/state_action_risk or <dir>/state_risk.

`SafeMotionsVecEnv.risk_ground_truth` produces the rows on the device; these helpers only serialise them."""
import csv
import json
import os
from ast import literal_eval

import numpy as np


def _row_literal(row):
    return "[" + ", ".join(repr(float(x)) for x in row) + "]"


def write_risk_csv(directory, state, risk, action=None, first_episode=0, last_episode=None, pid=None,
                   risk_prediction=None, risk_config=None):
    """Writes one file of the data set and returns its path.  state [N, D], action [N, A] or None, risk [N]."""
    state = np.asarray(state, dtype=np.float32)
    risk = np.asarray(risk, dtype=np.float64).reshape(-1)
    if state.shape[0] != risk.shape[0]:
        raise ValueError("state and risk differ in length")
    if action is not None:
        action = np.asarray(action, dtype=np.float32)
        if action.shape[0] != risk.shape[0]:
            raise ValueError("action and risk differ in length")
    sub = "state_action_risk" if action is not None else "state_risk"
    out_dir = os.path.join(directory, sub)
    os.makedirs(out_dir, exist_ok=True)
    last_episode = first_episode + len(risk) - 1 if last_episode is None else last_episode
    mean_pred = float(np.mean(risk_prediction)) if risk_prediction is not None and len(risk) else 0.0
    name = "episodes_{}_to_{}_risk_{:.2f}_prediction_{:.2f}_pid_{}.csv".format(
        first_episode, last_episode, float(np.mean(risk)) if len(risk) else 0.0, mean_pred,
        os.getpid() if pid is None else pid)
    path = os.path.join(out_dir, name)
    cols = ["state"] + (["action"] if action is not None else []) + ["risk"] + \
           (["risk_prediction"] if risk_prediction is not None else [])
    with open(path, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([""] + cols)   # DataFrame.to_csv: unnamed index column first
        for i in range(len(risk)):
            row = [i, _row_literal(state[i])]
            if action is not None:
                row.append(_row_literal(action[i]))
            row.append(repr(float(risk[i])))
            if risk_prediction is not None:
                row.append(repr(float(risk_prediction[i])))
            w.writerow(row)
    if risk_config is not None:
        cfg_path = os.path.join(directory, "risk_config.json")
        if not os.path.isfile(cfg_path):   # written once, like safe_motions_base.py:1449-1456
            with open(cfg_path, "w") as f:
                f.write(json.dumps({k: v for k, v in risk_config.items() if k != "config"}, sort_keys=True))
    return path


def read_risk_csv(path):
    """(state, action or None, risk) of one file, parsed the way train_risk_network.py:43-55 does."""
    with open(path, newline="") as f:
        rows = list(csv.reader(f))
    header = rows[0]
    col = {name: i for i, name in enumerate(header)}
    state = np.asarray([literal_eval(r[col["state"]]) for r in rows[1:]], dtype=np.float32)
    action = np.asarray([literal_eval(r[col["action"]]) for r in rows[1:]], dtype=np.float32) if "action" in col else None
    risk = np.asarray([float(r[col["risk"]]) for r in rows[1:]], dtype=np.float64)
    return state, action, risk

"""Risk data set wire format (safe_motions_base.py:1413-1461 writer, train_risk_network.py:43-55 reader)."""
import os
from ast import literal_eval

import numpy as np
import pandas as pd

from safemotionsrisk_b200 import risk_data


def test_csv_reads_back_with_the_reference_reader(tmp_path):
    rng = np.random.default_rng(0)
    state = rng.uniform(-1, 1, (50, 23)).astype(np.float32)
    action = rng.uniform(-1, 1, (50, 7)).astype(np.float32)
    risk = (rng.uniform(size=50) < 0.3).astype(np.float64)
    path = risk_data.write_risk_csv(str(tmp_path), state, risk, action=action, first_episode=100, pid=7,
                                    risk_config={"observation_size": 23, "action_size": 7, "config": {"x": 1}})
    assert os.path.basename(os.path.dirname(path)) == "state_action_risk"
    assert os.path.basename(path) == "episodes_100_to_149_risk_{:.2f}_prediction_0.00_pid_7.csv".format(risk.mean())
    # the reference's reader: pandas with literal_eval converters
    df = pd.read_csv(path, converters={"state": literal_eval, "action": literal_eval})
    assert np.array_equal(np.asarray(list(df["state"]), dtype=np.float32), state)
    assert np.array_equal(np.asarray(list(df["action"]), dtype=np.float32), action)
    assert np.array_equal(np.asarray(list(df["risk"])), risk)
    # and the same layout as DataFrame.to_csv of the reference's dict of lists
    ref = pd.DataFrame({"state": [list(map(float, s)) for s in state], "action": [list(map(float, a)) for a in action],
                        "risk": list(risk)})
    ref_path = os.path.join(str(tmp_path), "ref.csv")
    ref.to_csv(ref_path)
    assert open(ref_path).read().splitlines()[0] == open(path).read().splitlines()[0]
    s2, a2, r2 = risk_data.read_risk_csv(ref_path)
    assert np.array_equal(s2, state) and np.array_equal(a2, action) and np.array_equal(r2, risk)
    assert os.path.isfile(os.path.join(str(tmp_path), "risk_config.json"))
    assert "config" not in open(os.path.join(str(tmp_path), "risk_config.json")).read().replace("risk_config", "")


def test_state_risk_files_have_no_action_column(tmp_path):
    state = np.zeros((3, 23), dtype=np.float32)
    path = risk_data.write_risk_csv(str(tmp_path), state, [0.0, 1.0, 0.0])
    assert os.path.basename(os.path.dirname(path)) == "state_risk"
    s, a, r = risk_data.read_risk_csv(path)
    assert a is None and s.shape == (3, 23) and list(r) == [0.0, 1.0, 0.0]

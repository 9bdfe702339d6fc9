"""Packaged network weights (tools/export_networks.py) and the NumPy reference forward pass (oracle/mlp.py)."""
import os

import numpy as np
import pytest

from oracle import mlp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASSETS = os.path.join(ROOT, "safemotionsrisk_b200", "assets")


@pytest.mark.parametrize("scene,obs", [("space", 23), ("ball", 27)])
def test_packaged_weights_have_the_checkpoint_shapes(scene, obs):
    """Observation widths 23 / 27 are those of risk_config.json and of the checkpoints' fc_1 kernels."""
    w = np.load(os.path.join(ASSETS, "networks_{}.npz".format(scene)))
    assert int(w["risk/observation_size"]) == obs and int(w["risk/action_size"]) == 7
    assert [w["risk/dense_{}/kernel".format(i)].shape for i in range(4)] == [(obs + 7, 512), (512, 256), (256, 128),
                                                                              (128, 1)]
    assert [w["backup/{}/kernel".format(n)].shape for n in ("fc_1", "fc_2", "fc_out")] == [(obs, 256), (256, 128),
                                                                                            (128, 14)]
    for k in w.files:
        assert np.all(np.isfinite(w[k]))


def test_reference_forward_is_well_behaved():
    w = np.load(os.path.join(ASSETS, "networks_space.npz"))
    rng = np.random.default_rng(0)
    obs = rng.uniform(-1, 1, (512, 23)).astype(np.float32)
    act = rng.uniform(-1, 1, (512, 7)).astype(np.float32)
    risk = mlp.risk_forward(w, obs, act)
    pol = mlp.backup_forward(w, obs)
    assert risk.shape == (512,) and np.all((risk >= 0) & (risk <= 1))
    assert pol.shape == (512, 7) and np.all(np.abs(pol) <= 1)
    assert risk.std() > 1e-3 and pol.std() > 1e-2           # trained weights, not constants
    gated, r, risky = mlp.gate(w, obs, act, 0.065)
    assert np.array_equal(gated[~risky], act[~risky]) and np.allclose(gated[risky], pol[risky])


def test_activation_closed_forms():
    x = np.array([-2.0, -0.5, 0.0, 0.5, 2.0], dtype=np.float32)
    assert np.allclose(mlp.selu(x)[2:], 1.0507009873554805 * x[2:])
    assert np.allclose(mlp.selu(x)[:2], 1.0507009873554805 * 1.6732632423543772 * (np.exp(x[:2]) - 1), rtol=1e-6)
    assert np.allclose(mlp.swish(x), x / (1 + np.exp(-x)))

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")
    config.addinivalue_line("markers", "needs_pybullet: needs the reference's pybullet / klimits / gym (skipped without)")


@pytest.fixture(scope="session")
def space_scene():
    from safemotionsrisk_b200 import space_backup_config
    from safemotionsrisk_b200.scene import Scene
    return Scene(space_backup_config())


@pytest.fixture(scope="session")
def ball_scene():
    from safemotionsrisk_b200 import ball_backup_config
    from safemotionsrisk_b200.scene import Scene
    return Scene(ball_backup_config())


@pytest.fixture(scope="session")
def space_bm_scene():
    from safemotionsrisk_b200 import space_backup_config
    from safemotionsrisk_b200.scene import Scene
    return Scene(space_backup_config(ball_machine_mode=True))

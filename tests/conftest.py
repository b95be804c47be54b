import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def golden_car():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_car.npz"))


@pytest.fixture(scope="session")
def golden_sph():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_sph.npz"))


@pytest.fixture(scope="session")
def golden_cyl():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_cyl.npz"))


@pytest.fixture(scope="session")
def golden_oct():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_oct.npz"))


@pytest.fixture(scope="session")
def golden_amr():
    import numpy as np
    return np.load(os.path.join(ROOT, "tests", "golden", "bitlevel_amr.npz"))

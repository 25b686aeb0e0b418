import json
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def kat_dir(kats, tmp_path_factory):
    """The reference's fixture files, materialised from the committed golden JSON."""
    import base64
    d = tmp_path_factory.mktemp("refdata")
    for name, text in kats["files"].items():
        (d / name).write_text(text)
    for name, b in kats["binary_files_b64"].items():
        (d / name).write_bytes(base64.b64decode(b))
    return d

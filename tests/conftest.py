import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def _built_libraries():
    """build the oracle checkers (and, when nvcc is present and the library is stale, the CUDA library) once"""
    from oracle import build_oracle
    build_oracle.build_port()
    try:
        build_oracle.build_ref()
    except Exception as e:  # the reference tree / compiler may be absent on the GPU box: the prebuilt .so is used
        print('oracle/_ref not rebuilt:', e)
    from spitfire_b200 import build as gb_build
    if os.path.exists(gb_build.NVCC):
        gb_build.build(verbose=False)
    yield

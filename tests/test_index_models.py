"""CPU replays of kernel index arithmetic that could not be run on a GPU when it was written (tools/row32_index_model.py:
the 4 x 32-tile split cost-volume kernel, opt-in PWC_CV_SPLIT=row32) against the oracle."""
import importlib.util
import os

import pytest

_TOOL = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools", "row32_index_model.py")


@pytest.mark.parametrize("shape", [(1, 8, 64, 32), (1, 6, 37, 32)])
def test_row32_epilogue_index_model_matches_oracle(shape):
    spec = importlib.util.spec_from_file_location("row32_index_model", _TOOL)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.run(shape)          # asserts max error < 1e-5 and no write outside the 81-channel slot

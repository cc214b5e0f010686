"""Child test of tests/test_emu_kernels.py (needs HB200_EMU_TEST=1): __graft_entry__.smoke() itself."""
import os
import sys

import pytest

pytestmark = pytest.mark.skipif(not os.environ.get("HB200_EMU_TEST"), reason="emulation child only")


def test_smoke_entry_point():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import __graft_entry__ as g
    g.smoke()

"""A short run of the randomised differential test (tools/fuzz_parity.py): every entry point against the oracle on random
series, query lengths, thresholds and interval structures.  The longer runs are recorded in DESIGN.md section 5."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu


def test_random_configurations_match_the_oracle(oracle):
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import fuzz_parity
    bad, exempt = fuzz_parity.run(iters=100, seed=31, verbose=False)
    assert bad == 0

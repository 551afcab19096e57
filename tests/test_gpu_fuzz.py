"""Randomised scenarios against the oracle (tools/emu_fuzz.py): receivers, odd buffer lengths, entry path, input kind, options,
Mode A/C, ICAO flip period, sc16 input, Beast stream.  The same cases run on the emulated kernels in the tier without a GPU."""
import sys
from pathlib import Path

import pytest

sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tools"))

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("first", [0, 25])
def test_fuzz_cases_match_the_oracle(cuda, first):
    import emu_fuzz
    failures = []
    for k in range(first, first + 25):
        params, problems, _ = emu_fuzz.run_case(k, 31, False)
        if problems:
            failures.append(f"{params}\n    " + "\n    ".join(problems[:4]))
    assert not failures, "\n".join(failures)

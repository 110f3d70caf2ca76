"""The reference's OWN unit tests, unmodified, run against this package's modules.

transcoder/colours_test.py, opcodes_test.py, symbol_table_test.py and frame_grabber_test.py
import their subject by
bare module name (``import colours``); here those names are bound to iivision_b200's modules
of the same name and the test files are loaded from the reference tree where they lie.  These
are the suites whose subjects need no CUDA device; screen_test.py and video_test.py exercise
entry points that run on the GPU, where the reference tree is not available -- their cases
are restated in tests/test_gpu_scorer.py and tests/test_gpu_facade.py with the same literals.
Skipped when /root/reference is absent (the GPU box)."""

import importlib
import importlib.util
import os
import sys
import unittest

import pytest

REF = "/root/reference/transcoder"
pytestmark = [pytest.mark.reference,
              pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present")]

ALIASES = ("colours", "palette", "opcodes", "symbol_table", "machine", "video_mode",
           "frame_grabber")


def _run_reference_suite(filename: str):
    saved = {name: sys.modules.get(name) for name in ALIASES}
    try:
        for name in ALIASES:
            sys.modules[name] = importlib.import_module("iivision_b200." + name)
        spec = importlib.util.spec_from_file_location(
            "reference_" + filename[:-3], os.path.join(REF, filename))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        suite = unittest.defaultTestLoader.loadTestsFromModule(mod)
        result = unittest.TestResult()
        suite.run(result)
        return suite.countTestCases(), result
    finally:
        for name, old in saved.items():
            if old is None:
                sys.modules.pop(name, None)
            else:
                sys.modules[name] = old


@pytest.mark.parametrize("filename,n_tests", [("colours_test.py", 5), ("opcodes_test.py", 1),
                                              ("symbol_table_test.py", 1),
                                              ("frame_grabber_test.py", 1)])
def test_reference_suite_passes_on_our_modules(filename, n_tests):
    count, result = _run_reference_suite(filename)
    assert count == n_tests
    problems = [(str(t), tb) for t, tb in result.failures + result.errors]
    assert not problems, problems[0][1]
    assert result.testsRun == n_tests

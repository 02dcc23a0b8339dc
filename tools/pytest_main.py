"""pytest entry for tools/ab_lib.py: python tools/ab_lib.py tools/ab/libe2b_<name>.so tools/pytest_main.py tests/... -m gpu -q"""
import sys
import pytest
sys.exit(pytest.main(sys.argv[1:]))

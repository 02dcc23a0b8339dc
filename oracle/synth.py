"""TEST INFRASTRUCTURE -- the deterministic synthetic conditions and random-init weights live in /synthetic.py at the repo root
(so that bench.py's GPU arm imports nothing from oracle/); this module re-exports them under the name the tests and the
golden-vector scripts use."""
from synthetic import *          # noqa: F401,F403
from synthetic import _gen       # noqa: F401

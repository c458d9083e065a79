"""The synthetic input generators live in the package (refactored_orb_slam2_b200/synthetic.py) so that bench.py's own arm
needs nothing from oracle/; the oracle, the tests and the tools keep importing them under their historical name."""
from refactored_orb_slam2_b200.synthetic import *  # noqa: F401,F403
from refactored_orb_slam2_b200.synthetic import _flip_bits, _gauss, _warp_rows  # noqa: F401

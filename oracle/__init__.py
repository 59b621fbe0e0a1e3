"""CPU oracle of the reference's hot path (matching / triangulation / residuals).

TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this package; the product (sfm_opencv_b200) never does.
"""
from . import geometry, matching, synth  # noqa: F401

"""CPU ORACLE (test infrastructure).  pyref's protocol code with the C++ oracle underneath the two super-linear steps
(MSM and the S polynomial), so the Python restatement of the HyperPlonk driver can run at a few thousand rows."""
import numpy as np

from . import coracle as co
from . import pyref as py


class FastKZG(py.KZG):
    """pcs/src/kzg.rs KZG with the SRS generated and the MSM evaluated by the C++ oracle (same points, same sums)."""

    def __init__(self, max_degree, g1, tau, threads=4):
        self.max_degree, self.g1, self.tau, self.threads = max_degree, g1, tau % py.FR, threads
        self.srs = co.srs_generate(co.g1_to_bytes(g1), co.fr1(tau), max_degree + 1, threads=threads)
        self.g1_points = None

    def commit(self, poly):
        assert len(poly) <= self.max_degree + 1, "Polynomial degree exceeds max degree"
        s = co.to_mont(poly) if len(poly) else np.zeros((0, 32), np.uint8)
        return co.g1_from_bytes(co.msm(self.srs, s, mode=1, threads=self.threads))


def install_fast_s_polynomial():
    """Route pyref.compute_s_polynomial through the C++ oracle's NTT version (cross-checked against the O(n^2) Python
    one in tests/test_oracle.py)."""
    def fast(p1, p2):
        return co.from_mont(co.compute_s_polynomial(co.to_mont(p1) if len(p1) else np.zeros((0, 32), np.uint8),
                                                    co.to_mont(p2) if len(p2) else np.zeros((0, 32), np.uint8)))
    py.compute_s_polynomial = fast

"""pytest configuration: registers the `gpu` marker and puts the repo root and the
package directory (its name has hyphens, so it is a container, not an importable
name) on sys.path."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_DIR = os.path.join(ROOT, "options-pricing-engine-black-scholes-binomial-monte-carlo_b200")
for p in (ROOT, PKG_DIR):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")

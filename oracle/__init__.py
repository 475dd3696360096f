"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU checkers for the g_tessla hot path: `oracle/_ref/` holds the UNMODIFIED reference
compiled from /root/reference (built by oracle/Makefile, prebuilt .so travels to the GPU
box), `liboracle_port.so` is this repo's own plain-C restatement.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (g_tessla_b200/) never does.
"""
from .loader import RefOracle, PortOracle, ref_available, build_oracles  # noqa: F401

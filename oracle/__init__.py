"""TEST INFRASTRUCTURE: the CPU parity oracle (see oracle.h).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (ogl_b200/) never does."""
from .oracle import *  # noqa: F401,F403

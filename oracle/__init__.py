"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
import this package.  The product package (marl-dmfb_b200/) must never do so.
"""
from .oracle import *  # noqa: F401,F403

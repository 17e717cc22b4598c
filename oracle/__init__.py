"""TEST INFRASTRUCTURE — CPU oracles for the ruart_b200 hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (ruart_b200/) never does.
"""

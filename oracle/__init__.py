"""CPU oracle: plain-C restatement of the reference hot path. TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package. gamut_b200/ never does.
"""

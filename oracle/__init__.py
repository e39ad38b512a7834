"""ORACLE — CPU restatement of the reference's Paraformer::Forward path.

Test infrastructure only: nothing outside tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs may import this package, and the product path (asr-2pass_b200/) never does.
"""
